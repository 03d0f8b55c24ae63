#!/usr/bin/env python
"""How far the angle-free cos(sza) of the correction kernels (csrc/sunz.cu, b200_cos_sza_pixel<ALG>) is from the
reference's rounded-angle flow, on the full ABI disk at four times of day: the correction computed by the kernel
against 1 / cos(sza) from b200_cos_sza (which always reproduces the reference's flow).  DESIGN.md section 5 quotes the
result (max 1.1e-6 relative, mean 6e-8; up to 21 % of the disk in the exact band).

    python scripts/sunz_alg_check.py        # on a B200
"""
import datetime as dt
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satpy_b200 import demo  # noqa: E402
from satpy_b200.modifiers import cos_sza, sunz_correct  # noqa: E402

area = demo.make_area("goes_east_abi_f_1km")
data = torch.ones(area.shape, dtype=torch.float32, device="cuda")
for hour in (6, 11, 17, 22):
    t = dt.datetime(2024, 3, 20, hour, 30)
    cz = cos_sza(area, t, device=True)
    out = sunz_correct(data, area, t)
    lim = np.cos(np.deg2rad(88.0))
    sel = cz > lim
    ref = (1.0 / cz[sel]).to(torch.float32)
    got = out[sel]
    rel = ((got - ref).abs() / ref.abs())
    band = ((cz > np.cos(np.deg2rad(95.0)) - 1e-3) & (cz <= 0.2)).float().sum() / torch.isfinite(cz).float().sum()
    alg = (cz > 0.2)
    rel_alg = ((out[alg] - (1.0 / cz[alg]).to(torch.float32)).abs() * cz[alg].to(torch.float32))
    print(f"{hour:02d}:30  lit pixels {int(sel.sum())}  max rel (sza<88) {float(rel.max()):.3e}  max rel on the angle-free part {float(rel_alg.max()):.3e}  mean {float(rel_alg.mean()):.2e}  exact-band share of disc {float(band):.3f}")
