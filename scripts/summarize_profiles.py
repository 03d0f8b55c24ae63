#!/usr/bin/env python
"""Turn the raw ncu outputs of a gpurun call into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py --round r01 \
        --launches gpurun_out/launches_r01_final.csv --report gpurun_out/prof_r01_final.ncu-rep

* <round>_launches_final.csv / _summary.txt: the `--metrics gpu__time_duration.sum` launch list of
  `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e`, per-kernel totals and each kernel's share of
  ONE timed step (kernels of the stage micro-benchmarks that run after the timed region are listed but not shared).
* <round>_ncu_final_summary.txt: selected raw-page metrics of the `--set full` capture, per launch.
* traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, read by bench.py for `roofline.traffic`.
"""
import argparse
import collections
import csv
import io
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEP_KERNELS = ("geos_tables_kernel", "abi_calibrate_lane_kernel<0>", "sunz_fast_kernel<0, 1>", "nn_source_points_kernel<1>",
                "nn_target_tables_kernel", "nn_seg_cmax_kernel", "nn_fixed_query_rect_kernel<1>")
METRICS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
BENCH_CMD = "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"


def kname(name):
    """'void nn_fixed_query_rect_kernel<(bool)1>(GeosQueryArg, ...)' -> 'nn_fixed_query_rect_kernel<1>';
    library templates (cub::...) are cut at their first '<'."""
    n = name[5:] if name.startswith("void ") else name
    lt, par = n.find("<"), n.find("(")
    if 0 <= lt < par or (lt >= 0 and par < 0):
        if n.startswith("cub::") or n.startswith("at::"):
            return n[:lt]
        return n[:n.index(">", lt) + 1].replace("(bool)", "")
    return n[:par] if par >= 0 else n


def launches(path, out_prefix):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows:
        base = kname(r[k])
        tot[base] = tot.get(base, 0.0) + float(r[v]) / 1e6
        cnt[base] += 1
    total = sum(tot.values())
    lines = [f"ncu --metrics gpu__time_duration.sum --clock-control none ; {BENCH_CMD} (C5, 1 x B200)",
             "per-launch times are cold-cache and serialised: compare SHARES, not absolutes.",
             "nn_fixed_query_rect_kernel<1> is the fused search+gather of the timed step (plug-in path); <0> (index "
             "output), the gather, the fused", "calibrate+sunz kernel, bil_sample and fornav only run in the stage "
             "micro-benchmarks after the timed region.", ""]
    for name, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
        lines.append(f"{ms:10.3f} ms total {cnt[name]:3d} launches {ms / cnt[name]:9.3f} ms/launch {100 * ms / total:5.1f}%  {name[:70]}")
    step = {n: tot[n] / cnt[n] * (2 if n.startswith("geos_tables") else 1) for n in tot   # sunz + grid build
            if any(s in n for s in STEP_KERNELS) and not n.startswith("nn_fixed_query_rect_kernel<0>")}
    st = sum(step.values())
    lines += ["", "share of ONE timed step (per-launch times of the kernels a step launches):"]
    for name, ms in sorted(step.items(), key=lambda kv: -kv[1]):
        lines.append(f"   {100 * ms / st:5.1f}%  {ms:8.3f} ms  {name[:70]}")
    shutil.copyfile(path, out_prefix + "_launches_final.csv")
    open(out_prefix + "_launches_final_summary.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def report(paths, out_prefix, capture_args):
    lines = [f"ncu --set full --clock-control none --import-source on {capture_args}",
             f"{BENCH_CMD} ; C5 full scale, 1 x B200; raw-page metrics per launch", ""]
    traffic = collections.defaultdict(list)
    all_rows = []
    for path in paths:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        all_rows += [(hdr, units, r) for r in rows[2:]]
    for hdr, units, r in all_rows:
        k = hdr.index("Kernel Name")
        lines.append(r[k][:110])
        rd = wr = None
        for m in METRICS:
            if m not in hdr:
                continue
            j = hdr.index(m)
            val = r[j]
            try:
                val = f"{float(val.replace(',', '')):.6f}" if "." in val or "e" in val.lower() else val
            except ValueError:
                pass
            lines.append(f"    {m:70s} {val:>22s} {units[j]}")
            if m == "dram__bytes_read.sum":
                rd = (float(r[j].replace(",", "")), units[j])
            if m == "dram__bytes_write.sum":
                wr = (float(r[j].replace(",", "")), units[j])
        if rd and wr:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            base = kname(r[k])
            if base.startswith("nn_fixed_query_rect_kernel<0>"):
                continue  # index-output variant of the stage micro-benchmarks, not part of the timed step
            traffic[base].append(rd[0] * scale[rd[1]] + wr[0] * scale[wr[1]])
    open(out_prefix + "_ncu_final_summary.txt", "w").write("\n".join(lines) + "\n")
    tj = {"source": f"profiles/{os.path.basename(out_prefix)}_ncu_final_summary.txt (ncu --set full, dram__bytes_read.sum + "
                    "dram__bytes_write.sum per launch, C5 on one B200)",
          "dram_bytes_per_launch": {n: sum(v) / len(v) for n, v in sorted(traffic.items())}}
    json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(tj, indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r01")
    ap.add_argument("--launches")
    ap.add_argument("--report", nargs="+")
    ap.add_argument("--capture-args", default="-k regex:'nn_fixed_query_rect|sunz_fast|abi_calibrate_lane|nn_source_points|nn_gather_bulk' ...")
    a = ap.parse_args()
    prefix = os.path.join(ROOT, "profiles", a.round)
    if a.launches:
        launches(a.launches, prefix)
    if a.report:
        report(a.report, prefix, a.capture_args)


if __name__ == "__main__":
    main()
