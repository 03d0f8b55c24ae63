"""Oracle reader calibration (numpy, the reference's dtype flow) -- TEST INFRASTRUCTURE ONLY.

ABI L1b:
  * ``NC_ABI_BASE._adjust_data``      satpy/readers/core/abi.py:138-174
  * ``NC_ABI_L1B._vis_calibrate``     satpy/readers/abi_l1b.py:133-145 (+ ``*100`` :67-69)
  * ``NC_ABI_L1B._get_minimum_radiance`` satpy/readers/abi_l1b.py:147-155
  * ``NC_ABI_L1B._ir_calibrate``      satpy/readers/abi_l1b.py:157-173

Pinned by satpy/tests/reader_tests/test_abi_l1b.py:48-70,112-120,207-225,391-443.
"""
from __future__ import annotations

import numpy as np


def abi_adjust_data(counts: np.ndarray, scale_factor, add_offset, fill_value, unsigned: bool = False):
    """counts (int16) -> radiance (float32), satpy/readers/core/abi.py:138-174."""
    data = counts
    fill = None if fill_value is None else np.asarray(fill_value, dtype=counts.dtype)
    if unsigned:
        data = data.astype("u%s" % data.dtype.itemsize)
        if fill is not None:
            fill = fill.astype("u%s" % fill.dtype.itemsize)
    if fill is not None:
        # xarray ``where(cond, np.float32(nan))``: int16/uint16 promote with float32 -> float32
        data = np.where(data != fill, data.astype(np.float32), np.float32(np.nan))
    if scale_factor != 1:
        data = data * np.float32(scale_factor) + np.float32(add_offset)
    return data


def abi_vis_calibrate(rad: np.ndarray, esun: float, esd: float) -> np.ndarray:
    """radiance -> reflectance in percent, satpy/readers/abi_l1b.py:133-145,67-69."""
    factor = np.pi * esd * esd / esun
    res = rad * np.float32(factor)
    res = res * 100  # ``res *= 100`` on a float32 array stays float32
    return res.astype(np.float32, copy=False)


def abi_minimum_radiance(scale_factor, add_offset):
    """satpy/readers/abi_l1b.py:147-155."""
    count_zero_rad = - add_offset / scale_factor
    count_pos = np.ceil(count_zero_rad)
    return count_pos * scale_factor + add_offset


def abi_ir_calibrate(rad: np.ndarray, fk1: float, fk2: float, bc1: float, bc2: float,
                     clip_negative_radiances: bool = False, scale_factor=None, add_offset=None) -> np.ndarray:
    """radiance -> brightness temperature, satpy/readers/abi_l1b.py:157-173 (python-float constants, f32 result)."""
    data = rad
    if clip_negative_radiances:
        min_rad = abi_minimum_radiance(scale_factor, add_offset)
        data = data.clip(min=data.dtype.type(min_rad))
    with np.errstate(all="ignore"):
        res = (float(fk2) / np.log(float(fk1) / data + 1) - float(bc1)) / float(bc2)
    return res


def abi_calibrate(counts, calibration, scale_factor, add_offset, fill_value, unsigned=False, **consts):
    """``NC_ABI_L1B.get_dataset`` arithmetic for one calibration level (satpy/readers/abi_l1b.py:45-72)."""
    if calibration == "counts":
        return counts
    rad = abi_adjust_data(counts, scale_factor, add_offset, fill_value, unsigned)
    if calibration == "radiance":
        return rad
    if calibration == "reflectance":
        return abi_vis_calibrate(rad, consts["esun"], consts["esd"])
    if calibration == "brightness_temperature":
        return abi_ir_calibrate(rad, consts["fk1"], consts["fk2"], consts["bc1"], consts["bc2"],
                                consts.get("clip_negative_radiances", False), scale_factor, add_offset)
    raise ValueError(calibration)
