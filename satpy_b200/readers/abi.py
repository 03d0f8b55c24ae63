"""ABI L1b calibration on the device.

Host-side mirror of ``NC_ABI_L1B.get_dataset`` (satpy/readers/abi_l1b.py:45-72):
the same calibration names (``counts`` / ``radiance`` / ``reflectance`` /
``brightness_temperature``), the same inputs (the ``Rad`` variable's
``scale_factor`` / ``add_offset`` / ``_FillValue`` / ``_Unsigned`` attributes
and the file's ``esun`` / ``earth_sun_distance_anomaly_in_AU`` /
``planck_*`` scalars), the same float32 constant rounding
(satpy/readers/core/abi.py:172-173, satpy/readers/abi_l1b.py:133-173) and the
same ``readers.clip_negative_radiances`` switch (satpy/_config.py:54-56).

satpy has no plug point for calibration (it is inline in the file handler), so
the integration is: load with ``calibration="counts"`` and hand the counts and
these attributes to ``calibrate_abi`` -- see INTEGRATION.md.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from .. import _lib

_MODES = {"radiance": 0, "reflectance": 1, "brightness_temperature": 2}


@dataclass
class ABICalibration:
    """Calibration constants of one ABI band, named as in the L1b file."""

    scale_factor: float
    add_offset: float
    fill_value: int | None = None
    unsigned: bool = False
    esun: float | None = None
    earth_sun_distance_anomaly_in_AU: float | None = None
    planck_fk1: float | None = None
    planck_fk2: float | None = None
    planck_bc1: float | None = None
    planck_bc2: float | None = None
    clip_negative_radiances: bool = False

    def struct(self, calibration: str) -> _lib.AbiCalStruct:
        if calibration not in _MODES:
            raise ValueError(f"Unknown calibration '{calibration}'")  # abi_l1b.py:60-62 raises ValueError too
        s = _lib.AbiCalStruct()
        s.mode = _MODES[calibration]
        s.is_unsigned = 1 if self.unsigned else 0
        s.has_fill = 0 if self.fill_value is None else 1
        # the stored bit pattern as a signed 16-bit value (what the int16 array holds)
        s.fill = 0 if self.fill_value is None else int(np.asarray(self.fill_value).astype(np.int64)
                                                       .astype(np.uint16).view(np.int16))
        s.clip_negative_radiances = 1 if self.clip_negative_radiances else 0
        s.scale_factor = float(np.float32(self.scale_factor))
        s.add_offset = float(np.float32(self.add_offset))
        s.refl_factor = 0.0
        s.min_rad = 0.0
        s.fk1 = s.fk2 = s.bc1 = 0.0
        s.bc2 = 1.0
        if calibration == "reflectance":
            if self.esun is None or self.earth_sun_distance_anomaly_in_AU is None:
                raise ValueError("reflectance calibration needs esun and earth_sun_distance_anomaly_in_AU")
            esd = float(self.earth_sun_distance_anomaly_in_AU)
            s.refl_factor = float(np.float32(math.pi * esd * esd / float(self.esun)))
        elif calibration == "brightness_temperature":
            if None in (self.planck_fk1, self.planck_fk2, self.planck_bc1, self.planck_bc2):
                raise ValueError("brightness_temperature calibration needs planck_fk1/fk2/bc1/bc2")
            # python-float constants meet float32 arrays: numpy (NEP 50) rounds them to float32
            s.fk1 = float(np.float32(self.planck_fk1))
            s.fk2 = float(np.float32(self.planck_fk2))
            s.bc1 = float(np.float32(self.planck_bc1))
            s.bc2 = float(np.float32(self.planck_bc2))
            if self.clip_negative_radiances:
                # _get_minimum_radiance, abi_l1b.py:147-155
                # (evaluated in the dtype the attributes come in, float32 for netCDF attributes)
                count_zero = -self.add_offset / self.scale_factor
                s.min_rad = float(np.float32(np.ceil(count_zero) * self.scale_factor + self.add_offset))
        return s


def calibrate_abi(counts, cal: ABICalibration, calibration: str = "radiance", out=None, device=None):
    """int16 counts -> float32 radiance / reflectance [%] / brightness temperature [K].

    ``counts``: numpy array or torch tensor (int16; uint16 is reinterpreted).  A
    numpy input gives a numpy result, a CUDA tensor gives a CUDA tensor; ``device=True`` keeps the result on the
    device whatever came in (the next stage -- modifier, resampler -- takes it from there: nothing crosses PCIe
    between the stages, as nothing is computed between the lazy stages of the reference before ``compute()``).
    """
    torch = _lib.require_cuda()
    if calibration == "counts":
        return counts
    was_numpy = not isinstance(counts, torch.Tensor)
    if was_numpy:
        a = np.ascontiguousarray(counts)
        if a.dtype == np.uint16:
            a = a.view(np.int16)
        if a.dtype != np.int16:
            raise TypeError(f"ABI counts must be int16/uint16, got {a.dtype}")
        from .. import _xfer
        import os as _os, time as _time
        _t0 = _time.perf_counter()
        t = _xfer.to_device(a)
        if _os.environ.get("B200_XFER_TRACE") and _os.environ.get("RANK", "0") == "0":
            import sys as _sys
            torch.cuda.synchronize()
            print(f"[xfer] H2D {a.nbytes / 1e6:.0f} MB in {1e3 * (_time.perf_counter() - _t0):.1f} ms", file=_sys.stderr)
    else:
        t = counts.contiguous()
        if t.dtype != torch.int16:
            raise TypeError(f"ABI counts must be int16, got {t.dtype}")
        if not t.is_cuda:
            t = t.to("cuda")
    if out is None:
        out = torch.empty(t.shape, dtype=torch.float32, device="cuda")
    s = cal.struct(calibration)
    _lib.check(_lib.load().b200_abi_calibrate(t.data_ptr(), out.data_ptr(), t.numel(), C.byref(s),
                                              _lib.stream_ptr(torch)))
    if device is None:
        device = not was_numpy
    if device:
        return out
    from .. import _xfer
    return _xfer.to_host(out)
