#!/usr/bin/env python
"""The hot path through the plug-in API, on synthetic data (no satpy install needed; with satpy installed call
``satpy_b200.register()`` and use ``Scene.resample(area, resampler="b200_nearest")`` instead).

    python examples/abi_resample_demo.py [--scale 0.1]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from satpy_b200 import DataArrayLite, demo  # noqa: E402
from satpy_b200.composites import SimulatedGreen  # noqa: E402
from satpy_b200.modifiers import SunZenithCorrector  # noqa: E402
from satpy_b200.readers import calibrate_abi  # noqa: E402
from satpy_b200.resample import B200NearestResampler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.1, help="fraction of the full 5424^2 -> 9050^2 config-2 size")
    args = ap.parse_args()
    src = demo.make_area("goes_east_abi_f_2km", args.scale)
    dst = demo.make_area("goes_east_eqc_2km", args.scale)
    sunz = SunZenithCorrector(name="sunz_corrected")            # same class name / keywords as satpy.modifiers
    resampler = B200NearestResampler(src, dst)                  # cls(source_geo_def, target_geo_def), as satpy builds it
    t0 = time.perf_counter()
    bands = []
    for seed, name in ((2, "C01"), (3, "C02"), (4, "C03")):
        counts = demo.abi_counts(src.shape, seed, name)                                   # what calibration="counts" loads
        refl = calibrate_abi(counts, demo.abi_calibration(name), "reflectance")           # reader arithmetic
        ds = DataArrayLite(refl, ("y", "x"), {"name": name, "area": src, "start_time": demo.START_TIME})
        corrected = sunz([ds])                                                            # YAML modifier sunz_corrected
        bands.append(resampler.resample(corrected, radius_of_influence=5000.0 / args.scale))  # search once, gather thrice
    green = SimulatedGreen("green")(bands)
    dt = time.perf_counter() - t0
    v = green.values
    print(f"{len(bands)} bands {src.shape} -> {dst.shape} in {dt * 1e3:.0f} ms (host arrays in and out); "
          f"green: {np.isfinite(v).mean():.1%} of the target covered, mean {np.nanmean(v):.2f} %")


if __name__ == "__main__":
    main()
