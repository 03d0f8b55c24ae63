#!/usr/bin/env python
"""Per-kernel timings of the C-ABI entry points at the BASELINE.json sizes (CUDA events, L2-exceeding buffers).

    python scripts/kernel_bench.py [--out profiles/rNN_kernel_bench.json]

Each line: kernel, workload, ms, algorithmic bytes, GB/s and the fraction of the measured HBM
copy bandwidth (MEASURED_PEAKS.json).  These are the numbers DESIGN.md section 5 quotes.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from satpy_b200 import _lib, demo  # noqa: E402
from satpy_b200.modifiers import sunz_correct  # noqa: E402
from satpy_b200.readers import calibrate_abi  # noqa: E402
from satpy_b200.resample import B200BilinearResampler, B200EWAResampler, B200NearestResampler  # noqa: E402
from satpy_b200 import AreaDefinition, SwathDefinition  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, steps=5, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    pk = peak()
    rows = []

    def rec(kernel, workload, ms, nbytes, note=""):
        gbs = nbytes / ms / 1e6 if nbytes else 0.0
        r = {"kernel": kernel, "workload": workload, "ms": round(ms, 4), "algorithmic_bytes": int(nbytes),
             "gbs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / pk, 4), "note": note}
        rows.append(r)
        print(json.dumps(r), flush=True)

    # ---- calibration, sun zenith (C5 source: 10848^2)
    src = demo.make_area("goes_east_abi_f_1km")
    n = src.size
    counts = torch.from_numpy(demo.abi_counts(src.shape, 7, "C02")).cuda()
    out = torch.empty(src.shape, dtype=torch.float32, device="cuda")
    for band, calib in (("C02", "radiance"), ("C02", "reflectance"), ("C13", "brightness_temperature")):
        cal = demo.abi_calibration(band)
        ms = timeit(lambda: calibrate_abi(counts, cal, calib, out=out))
        rec("abi_calibrate_vec_kernel", f"ABI {band} {calib} 10848^2", ms, 6 * n)
    refl = torch.rand(src.shape, device="cuda") * 100
    ms = timeit(lambda: sunz_correct(refl, area=src, start_time=demo.START_TIME, out=out))
    rec("sunz_correct_kernel", "SunZenithCorrector 10848^2 geos, angles derived in-kernel", ms, 8 * n,
        "bound by float64 transcendentals of the inverse projection, not by HBM")
    sza = torch.rand(src.shape, device="cuda") * 100
    ms = timeit(lambda: sunz_correct(refl, sza=sza, out=out))
    rec("sunz_correct_sza_kernel", "SunZenithCorrector 10848^2, SZA dataset provided", ms, 12 * n)
    from satpy_b200.pipeline import ABIResamplePipeline
    dst_small = demo.make_area("global_eqc_500m", 0.01)
    pipe = ABIResamplePipeline(src, dst_small, 2000.0)
    cal = demo.abi_calibration("C02")
    ms = timeit(lambda: pipe.calibrate_sunz(counts, cal, demo.START_TIME))
    rec("abi_vis_sunz_kernel", "fused calibrate + sunz 10848^2", ms, 6 * n, "bound by float64 transcendentals")
    del pipe, sza, refl

    # ---- nearest: search + gather, config 5 (full) and config 1
    dst = demo.make_area("global_eqc_500m")
    r = B200NearestResampler(src, dst)
    ms = timeit(lambda: (r._index_caches.clear(), r.precompute(radius_of_influence=2000.0)), steps=2, warmup=1)
    n_valid = r._current["n_valid"]
    rec("nn_grid_create + nn_query (3 index arrays)", "C5 search 10848^2 -> 80150x40075, r=2 km", ms, 9 * dst.size)
    data = out
    res = torch.empty(dst.shape, dtype=torch.float32, device="cuda")
    gi = r._current["gather_index"]
    fill = (C.c_char * 4).from_buffer_copy(np.float32(np.nan).tobytes())
    lib = _lib.load()

    def gather():
        _lib.check(lib.b200_nn_gather(data.data_ptr(), res.data_ptr(), gi.data_ptr(), gi.numel(), 1, src.size, gi.numel(),
                                      4, C.cast(fill, C.c_void_p), _lib.stream_ptr(torch)))
    ms = timeit(gather)
    rec("nn_gather_kernel<u32>", "C5 gather float32 3212 Mpix", ms, 8 * dst.size + 4 * n_valid)
    del r, gi, res
    torch.cuda.empty_cache()

    s1, d1 = demo.make_area("msg_seviri_fes_3km"), demo.make_area("euro_laea_1km")
    r1 = B200NearestResampler(s1, d1)
    ms = timeit(lambda: (r1._index_caches.clear(), r1.precompute(radius_of_influence=5000.0)), steps=3, warmup=1)
    rec("nn_grid_create + nn_query (general target)", "C1 search 3712^2 -> 5000x4500 LAEA, r=5 km", ms, 9 * d1.size)
    bt = torch.from_numpy(demo.seviri_bt(s1.shape)).cuda()
    ms = timeit(lambda: r1.compute(bt))
    rec("nn_gather_kernel<u32>", "C1 gather 22.5 Mpix", ms, 8 * d1.size + 4 * r1._current["n_valid"])
    ms = timeit(lambda: r1.resample_fused(bt, 5000.0), steps=3, warmup=1)
    rec("nn_grid_create + nn_fixed_query_kernel (fused)", "C1 cold resample, fused", ms, 4 * d1.size + 4 * r1._current["n_valid"])
    del r1

    # ---- config 2: three bands, one search, gather x3
    s2, d2 = demo.make_area("goes_east_abi_f_2km"), demo.make_area("goes_east_eqc_2km")
    r2 = B200NearestResampler(s2, d2)
    ms = timeit(lambda: (r2._index_caches.clear(), r2.precompute(radius_of_influence=5000.0)), steps=3, warmup=1)
    rec("nn_grid_create + nn_query", "C2 search 5424^2 -> 9050^2 eqc, r=5 km", ms, 9 * d2.size)
    rgb = torch.rand((3,) + s2.shape, device="cuda")
    ms = timeit(lambda: r2.compute(rgb))
    rec("nn_gather_kernel<u32> x3 bands", "C2 gather 3 x 81.9 Mpix", ms, (4 + 3 * 4) * d2.size + 3 * 4 * r2._current["n_valid"])
    del r2, rgb

    # ---- config 3: AHI 11000^2 (1 km) -> 1 km LCC 6000^2, bilinear, 32 neighbours, radius 50 km
    s3, d3 = demo.make_area("himawari_ahi_fes_1km"), demo.make_area("east_asia_lcc_1km")
    rb = B200BilinearResampler(s3, d3)
    ms = timeit(lambda: (setattr(rb, "_info", None), rb.precompute(radius_of_influence=50000.0)), steps=2, warmup=1)
    rec("nn_grid_create + b200_bil_info", "C3 bilinear info: AHI 11000^2 -> 6000^2 LCC, 32 neighbours, r=50 km", ms,
        32 * d3.size, "k-nearest + quadrant corners + (s, t); float64 projection / latency bound")
    ahi = torch.rand(s3.shape, device="cuda") * 300
    ms = timeit(lambda: rb.compute(ahi))
    rec("bil_sample_f32_kernel", "C3 sample 36 Mpix", ms, 36 * d3.size)
    del rb, ahi

    # ---- config 4: ten VIIRS M-band granules' worth of swath (7680 x 3200 at the real 750 m sampling) to a 750 m
    # polar-stereographic grid: EWA, and nearest neighbour through the general (bucket) search engine
    lon, lat, rps = demo.viirs_like_swath(nscans=480, rows_per_scan=16, ncols=3200, lat0=28.0, lat_span=51.8,
                                          lon_half_width=10.79)
    swath = SwathDefinition(torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda())
    d4 = AreaDefinition("ps", "ps", "ps", {"proj": "stere", "lat_0": 90.0, "lat_ts": 60.0, "lon_0": 0.0, "ellps": "WGS84"},
                        6400, 9200, (-3.95e6, -7.1e6, 0.85e6, -0.2e6))
    re = B200EWAResampler(swath, d4)
    ms = timeit(lambda: (setattr(re, "_cr", None), re.precompute()), steps=3, warmup=1)
    rec("ll2cr_kernel<float>", "C4 swath 7680x3200 -> (col, row)", ms, 16 * swath.size)
    sw = torch.rand(swath.shape, device="cuda") * 100
    ms = timeit(lambda: re.compute(sw, rows_per_scan=rps), steps=3, warmup=1)
    rec("fornav (params + accumulate + finalize)", f"C4 EWA 24.6 Mpix swath -> {d4.width}x{d4.height} grid, "
        f"{re.valid_counts[0]} cells filled", ms, 12 * swath.size + 12 * d4.size, "float32 atomics into L2")
    del re
    for radius in (2000.0, 5000.0):
        rn = B200NearestResampler(swath, d4)
        ms = timeit(lambda: (rn._index_caches.clear(), rn.precompute(radius_of_influence=radius)), steps=3, warmup=1)
        hits = float((rn._current["gather_index"] >= 0).float().mean())
        rec("nn_grid_create + nn_query (bucket engine)", f"C4 nearest: 24.6 Mpix swath -> {d4.width}x{d4.height} grid, "
            f"r={radius:.0f} m, {hits:.3f} of the grid hit", ms, 24 * swath.size + 4 * d4.size,
            "xyz build + counting sort + query; index array out")
        ms = timeit(lambda: rn.compute(sw, fill_value=float("nan")), steps=3, warmup=1)
        rec("nn_gather_kernel<float>", "C4 gather", ms, 8 * d4.size + 4 * swath.size)
        del rn
    if args.out:
        with open(args.out, "w") as fh:
            json.dump({"peak_hbm_gbs": pk, "rows": rows}, fh, indent=1)


if __name__ == "__main__":
    main()
