"""SEVIRI calibration on the device ("next" row of SURVEY.md section 8f).

Host-side mirror of ``SEVIRICalibrationHandler.calibrate`` / ``SEVIRICalibrationAlgorithm``
(satpy/readers/core/seviri.py:574-633,660-686): counts -> radiance (``where(> 0) * gain + offset``, clipped at 0) ->
reflectance (``pi L 100 / F`` times the squared sun-earth distance factor) or brightness temperature from effective
(``(T - beta) / alpha``) or spectral (``BTFIT`` quadratic) radiance.  Constants ``C1``, ``C2`` as seviri.py:215-216.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from .. import _lib
from ..modifiers.astronomy import days_since_j2000
from ._common import new_f32_like, to_device

C1 = 1.19104273e-5
C2 = 1.43877523


def sun_earth_distance_correction(utc_time) -> float:
    """pyorbital ``astronomy.sun_earth_distance_correction`` (satpy/readers/core/utils.py:487-497 calls it)."""
    return 1 - 0.0167 * math.cos(2 * math.pi * (days_since_j2000(utc_time) - 3) / 365.256363)


@dataclass
class SEVIRICalibration:
    gain: float
    offset: float
    solar_irradiance: float | None = None     # CALIB[platform][channel]["F"]
    scan_time: object = None                  # for the sun-earth distance correction
    wavenumber: float | None = None           # CALIB[...]["VC"]
    alpha: float | None = None
    beta: float | None = None
    btfit: tuple | None = None                # BTFIT[channel] = (A, B, C)

    def struct(self, calibration, radiance_type="effective"):
        s = _lib.SeviriCalStruct()
        s.gain = float(np.float32(self.gain))
        s.offset = float(np.float32(self.offset))
        s.pi_f = float(np.float32(np.pi))
        if calibration == "radiance":
            s.mode = 0
        elif calibration == "reflectance":
            if self.solar_irradiance is None or self.scan_time is None:
                raise ValueError("reflectance calibration needs solar_irradiance and scan_time")
            s.mode = 1
            s.solar_irradiance = float(np.float32(self.solar_irradiance))
            d = sun_earth_distance_correction(self.scan_time)
            s.sun_earth_dist2 = float(np.float32(d * d))
        elif calibration == "brightness_temperature":
            if self.wavenumber is None:
                raise ValueError("brightness_temperature calibration needs the channel wavenumber")
            nu = float(self.wavenumber)
            s.c1 = float(np.float32(C1))
            s.nu3 = float(np.float32(nu ** 3))
            s.c2nu = float(np.float32(C2 * nu))
            if radiance_type == "effective":
                s.mode = 2
                s.alpha, s.beta = float(np.float32(self.alpha)), float(np.float32(self.beta))
            elif radiance_type == "spectral":
                s.mode = 3
                s.fit_a, s.fit_b, s.fit_c = (float(np.float32(v)) for v in self.btfit)
            else:  # IRCalibrationType.not_processed, seviri.py:603-605
                raise ValueError("No calibration coefficients available (PlannedChanProcessing is zero)")
        else:
            raise ValueError(f"Invalid calibration {calibration}")
        return s


def calibrate_seviri(counts, cal: SEVIRICalibration, calibration="radiance", radiance_type="effective", out=None):
    """counts (uint16 or float32) -> float32 radiance / reflectance [%] / brightness temperature [K]."""
    if calibration == "counts":
        return counts
    t, was_numpy, name = to_device(counts, ("uint16", "int16", "float32"))
    torch = _lib.require_cuda()
    if out is None:
        out = new_f32_like(t)
    s = cal.struct(calibration, radiance_type)
    _lib.check(_lib.load().b200_seviri_calibrate(t.data_ptr(), 1 if name == "float32" else 0, out.data_ptr(), t.numel(),
                                                 C.byref(s), _lib.stream_ptr(torch)))
    return out.cpu().numpy() if was_numpy else out
