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


# =====================================================================================================================
# "Next" rows of SURVEY.md section 8f: AHI HSD, SEVIRI and VIIRS SDR calibration (same elementwise family as ABI).

def ahi_mask_invalid(counts, count_outside_scan, count_error):
    """satpy/readers/ahi_hsd.py:607-611."""
    invalid = (counts == count_outside_scan) | (counts == count_error)
    return np.where(invalid, np.float32(np.nan), counts)


def ahi_calibrate(counts, calibration, gain, offset, coeff_rad2albedo=None, central_wave_length=None,
                  speed_of_light=None, planck_constant=None, boltzmann_constant=None, c0=None, c1=None, c2=None,
                  count_outside_scan=None, count_error=None):
    """``AHIHSDFileHandler.calibrate`` (satpy/readers/ahi_hsd.py:675-750) with the dtype flow numpy gives the
    constants as passed (Python floats in the reference's test fixture; numpy float64 scalars when they come from a
    real file header, in which case the brightness temperature is evaluated in float64).
    Pinned by satpy/tests/reader_tests/test_ahi_hsd.py:455-506."""
    data = counts
    if count_outside_scan is not None:
        data = ahi_mask_invalid(counts, count_outside_scan, count_error)
    if calibration == "counts":
        return data
    data = data * np.float32(gain) + np.float32(offset)
    if calibration == "radiance":
        return data
    if calibration == "reflectance":
        return (data * np.float32(coeff_rad2albedo) * 100).clip(0)
    with np.errstate(all="ignore"):
        data = np.where(data == 0, np.float32(np.nan), data)
        cwl = central_wave_length * 1e-6
        a__ = (planck_constant * speed_of_light) / (boltzmann_constant * cwl)
        b__ = ((2 * planck_constant * speed_of_light ** 2) / (data * 1.0e6 * cwl ** 5)) + 1
        te_ = a__ / np.log(b__)
        return (c0 + c1 * te_ + c2 * te_ ** 2).clip(0)


SEVIRI_C1 = 1.19104273e-5   # satpy/readers/core/seviri.py:215-216
SEVIRI_C2 = 1.43877523


def seviri_radiance(counts, gain, offset):
    """``SEVIRICalibrationHandler.calibrate`` + ``convert_to_radiance`` (satpy/readers/core/seviri.py:582-585,660-671)."""
    data = np.asarray(counts).astype(np.float32)
    with np.errstate(invalid="ignore"):
        data = np.where(data > 0, data, np.float32(np.nan))
    return (data * np.float32(gain) + np.float32(offset)).clip(0.0, None)


def seviri_tl15(rad, wavenumber):
    """satpy/readers/core/seviri.py:620-623."""
    with np.errstate(all="ignore"):
        return (SEVIRI_C2 * wavenumber) / np.log((1.0 / rad) * SEVIRI_C1 * wavenumber ** 3 + 1.0)


def seviri_erads2bt(rad, wavenumber, alpha, beta):
    """Effective radiance -> BT (satpy/readers/core/seviri.py:587-595)."""
    return (seviri_tl15(rad, wavenumber) - beta) / alpha


def seviri_srads2bt(rad, wavenumber, btfit):
    """Spectral radiance -> BT (satpy/readers/core/seviri.py:612-618)."""
    a__, b__, c__ = btfit
    temp = seviri_tl15(rad, wavenumber)
    return a__ * temp * temp + b__ * temp + c__


def sun_earth_distance_correction(utc_time):
    """pyorbital ``astronomy.sun_earth_distance_correction`` (un-vendored): ``1 - 0.0167 cos(2 pi (jdays2000 - 3) /
    365.256363)``; pinned by VIS008_REFLECTANCE, satpy/tests/reader_tests/test_seviri_l1b_calibration.py:81-100."""
    from .solar import jdays2000
    return 1 - 0.0167 * np.cos(2 * np.pi * (jdays2000(utc_time) - 3) / 365.256363)


def seviri_vis_calibrate(rad, solar_irradiance, utc_time):
    """satpy/readers/core/seviri.py:625-633 + readers/core/utils.py:487-497."""
    reflectance = np.pi * rad * 100.0 / solar_irradiance
    d = sun_earth_distance_correction(utc_time)
    return reflectance * reflectance.dtype.type(d * d)


def viirs_scale(data, factors, rows_per_granule, fill_min_int=65528, fill_max_float=-999.0):
    """VIIRS SDR: ``mask_fill_values`` then per-granule ``data * factor + offset``
    (satpy/readers/core/viirs_atms_sdr.py:168-176,197-214,328-339,360-361).  ``factors``: (n_granules, 2)."""
    data = np.asarray(data)
    if np.issubdtype(data.dtype, np.floating):
        with np.errstate(invalid="ignore"):
            data = np.where(data > np.float32(fill_max_float), data, np.float32(np.nan))
    else:
        data = np.where(data < int(fill_min_int), data, np.float32(np.nan))
    factors = np.asarray(factors, dtype=np.float32).reshape(-1, 2)
    factors = np.where(factors > -999, factors, np.float32(np.nan))
    out = np.empty(data.shape, dtype=np.result_type(data.dtype, np.float32))
    r0 = 0
    for g, n in enumerate(rows_per_granule):
        out[r0:r0 + n] = data[r0:r0 + n] * factors[g, 0] + factors[g, 1]
        r0 += n
    return out
