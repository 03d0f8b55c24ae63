#!/usr/bin/env python
"""Device time of the sun-zenith kernels on the ABI 1 km full disk at several times of day (CUDA events).

    python scripts/sunz_bench.py         # on a B200

Prints ms and GB/s on the SURVEY 8d algorithmic bytes (8 B/pixel standalone, 6 B/pixel fused) and, once, the largest
relative deviation of the correction from 1 / cos(sza) of b200_cos_sza (the reference's rounded-angle flow)."""
import datetime as dt
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satpy_b200 import demo  # noqa: E402
from satpy_b200.modifiers import cos_sza, sunz_correct  # noqa: E402
from satpy_b200.pipeline import ABIResamplePipeline  # noqa: E402


def timeit(fn, steps=10, warmup=3):
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
    pk = 6554.9
    try:
        pk = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    src = demo.make_area("goes_east_abi_f_1km")
    n = src.size
    refl = torch.rand(src.shape, device="cuda") * 100
    out = torch.empty_like(refl)
    counts = torch.from_numpy(demo.abi_counts(src.shape, 7, "C02")).cuda()
    cal = demo.abi_calibration("C02")
    pipe = ABIResamplePipeline(src, demo.make_area("global_eqc_500m", 0.01), 2000.0)
    for when in (demo.START_TIME, dt.datetime(2024, 3, 20, 11, 30), dt.datetime(2024, 3, 20, 17, 30),
                 dt.datetime(2024, 3, 20, 22, 30)):
        ms = timeit(lambda: sunz_correct(refl, area=src, start_time=when, out=out))
        ms2 = timeit(lambda: pipe.calibrate_sunz(counts, cal, when))
        cz = cos_sza(src, when, device=True)
        ones = torch.ones_like(refl)
        got = sunz_correct(ones, area=src, start_time=when)
        sel = cz > 0.0349 + 1e-3
        rel = ((got[sel].double() - 1.0 / cz[sel]).abs() * cz[sel]).max().item()
        print(json.dumps({"when": when.isoformat(), "sunz_correct_ms": round(ms, 4), "sunz_correct_gbs": round(8 * n / ms / 1e6, 1),
                          "sunz_frac": round(8 * n / ms / 1e6 / pk, 3), "fused_ms": round(ms2, 4),
                          "fused_gbs": round(6 * n / ms2 / 1e6, 1), "fused_frac": round(6 * n / ms2 / 1e6 / pk, 3),
                          "max_rel_dev_of_correction": rel}), flush=True)


if __name__ == "__main__":
    main()
