#!/usr/bin/env python
"""Host-output timing of the plug-in flow on config 5 under different band / span settings (one B200):
    B200_XFER_TRACE=1 python scripts/e2e_probe.py
Prints ms per call of calibrate_abi(numpy) -> SunZenithCorrector -> B200NearestResampler.resample(numpy out)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from satpy_b200 import DataArrayLite, demo
from satpy_b200.modifiers import SunZenithCorrector
from satpy_b200.readers import calibrate_abi
from satpy_b200.resample import B200NearestResampler

src, dst = demo.make_area("goes_east_abi_f_1km"), demo.make_area("global_eqc_500m")
cal = demo.abi_calibration("C02")
counts = demo.abi_counts(src.shape, 7, "C02")
sunz = SunZenithCorrector(name="sunz_corrected", modifiers=("sunz_corrected",))
attrs = {"start_time": demo.START_TIME, "name": "C02", "calibration": "reflectance", "units": "%",
         "standard_name": "toa_bidirectional_reflectance", "area": src}
out = torch.empty(dst.shape, dtype=torch.float32, device="cuda")


def step():
    refl = calibrate_abi(counts, cal, "reflectance", device=True)
    corrected = sunz([DataArrayLite(refl, ("y", "x"), dict(attrs))])
    rs = B200NearestResampler(src, dst)
    res = rs.resample(corrected, radius_of_influence=2000.0, fill_value=float("nan"), out=out, to_host=True)
    return res.data, rs.last_d2h_bytes


CASES = ((24, 416),) if (os.environ.get('B200_FILL_THREADS') or os.environ.get('B200_D2H_STREAMS')) else ((24, 208), (24, 416), (24, 1 << 20), (12, 208), (48, 208))
for chunks, span_rows in CASES:
    B200NearestResampler.ROW_CHUNKS, B200NearestResampler.SPAN_ROWS = chunks, span_rows
    for _ in range(2):
        r, nbytes = step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        r, nbytes = step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    print(f"fill threads {os.environ.get('B200_FILL_THREADS', 'default')} ROW_CHUNKS {chunks:3d} SPAN_ROWS {span_rows:8d}: {ms:7.1f} ms per call, {nbytes / 1e9:.2f} GB over PCIe "
          f"({nbytes / ms / 1e6:.1f} GB/s), finite pixels {int(np.isfinite(r[::97]).sum())}", flush=True)
    del r
