"""AHI HSD calibration on the device ("next" row of SURVEY.md section 8f).

Host-side mirror of ``AHIHSDFileHandler.calibrate`` (satpy/readers/ahi_hsd.py:675-750): the same calibration names and
the header entries the reference reads (block5 / calibration blocks).  ``bt_in_float64`` reproduces numpy's dtype flow
for the brightness temperature: header values read from a real file are numpy float64 scalars (float64 arithmetic),
the reference's own test fixture passes Python floats (float32 arithmetic, satpy/tests/reader_tests/test_ahi_hsd.py:455-506).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .. import _lib
from ._common import new_f32_like, to_device

_MODES = {"radiance": 0, "reflectance": 1, "brightness_temperature": 2}


@dataclass
class AHICalibration:
    gain_count2rad_conversion: float
    offset_count2rad_conversion: float
    coeff_rad2albedo_conversion: float = 0.0
    central_wave_length: float | None = None      # micrometres, as in block5
    speed_of_light: float = 299792458.0
    planck_constant: float = 6.62606957e-34
    boltzmann_constant: float = 1.3806488e-23
    c0_rad2tb_conversion: float = 0.0
    c1_rad2tb_conversion: float = 1.0
    c2_rad2tb_conversion: float = 0.0
    count_value_outside_scan_pixels: int | None = None
    count_value_error_pixels: int | None = None
    bt_in_float64: bool = True

    def struct(self, calibration):
        if calibration not in _MODES:
            raise ValueError(f"Unknown calibration '{calibration}'")
        s = _lib.AhiCalStruct()
        s.mode = _MODES[calibration]
        s.has_invalid = 0 if self.count_value_outside_scan_pixels is None else 1
        s.count_outside_scan = int(self.count_value_outside_scan_pixels or 0)
        s.count_error = int(self.count_value_error_pixels if self.count_value_error_pixels is not None
                            else s.count_outside_scan)
        s.bt_in_f64 = 1 if self.bt_in_float64 else 0
        s.gain = float(np.float32(self.gain_count2rad_conversion))
        s.offset = float(np.float32(self.offset_count2rad_conversion))
        s.coeff = float(np.float32(self.coeff_rad2albedo_conversion))
        if calibration == "brightness_temperature":
            if self.central_wave_length is None:
                raise ValueError("brightness_temperature calibration needs central_wave_length")
            cwl = float(self.central_wave_length) * 1e-6
            h, c, k = float(self.planck_constant), float(self.speed_of_light), float(self.boltzmann_constant)
            s.a = (h * c) / (k * cwl)
            s.num = 2 * h * c ** 2
            s.cwl5 = cwl ** 5
            s.c0, s.c1, s.c2 = (float(self.c0_rad2tb_conversion), float(self.c1_rad2tb_conversion),
                                float(self.c2_rad2tb_conversion))
        return s


def calibrate_ahi(counts, cal: AHICalibration, calibration="radiance", out=None):
    """uint16 counts -> float32 radiance / reflectance [%] / brightness temperature [K]."""
    if calibration == "counts":
        return counts
    t, was_numpy, _ = to_device(counts, ("uint16", "int16"))
    torch = _lib.require_cuda()
    if out is None:
        out = new_f32_like(t)
    s = cal.struct(calibration)
    _lib.check(_lib.load().b200_ahi_calibrate(t.data_ptr(), out.data_ptr(), t.numel(), C.byref(s),
                                              _lib.stream_ptr(torch)))
    return out.cpu().numpy() if was_numpy else out
