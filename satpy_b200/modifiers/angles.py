"""Sensor and sun angles on the device ("next" row of SURVEY.md section 8f).

Host-side mirror of ``get_angles``, ``get_satellite_zenith_angle`` and ``compute_relative_azimuth``
(satpy/modifiers/angles.py:336-414).  The satellite position is what ``satpy.utils.get_satpos`` extracts from
``attrs["orbital_parameters"]``: longitude and latitude in degrees, altitude in METRES (angles.py:523-527 divides by 1000).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from ..dataset import is_dataarray
from ..geometry import as_geometry
from .geometry import _sunz_struct


def _satpos(orbital_parameters):
    """``get_satpos`` with the default preference "actual" then "nominal" then "projection" (satpy/utils.py)."""
    for prefix in ("satellite_actual_", "satellite_nominal_", "projection_"):
        try:
            alt = orbital_parameters[prefix + "altitude"] if prefix != "projection_" else orbital_parameters[
                "projection_altitude"]
            return (float(orbital_parameters[prefix + "longitude"]), float(orbital_parameters[prefix + "latitude"]),
                    float(alt))
        except KeyError:
            continue
    raise KeyError("Unable to determine satellite position. Either the reader doesn't provide that information or "
                   "geolocation datasets are missing.")


def get_angles(data_arr=None, area=None, start_time=None, sat_lon=None, sat_lat=None, sat_alt=None, device=False,
               which=("sata", "satz", "suna", "sunz")):
    """Sensor azimuth, sensor zenith, solar azimuth, solar zenith in degrees (float64).

    Either a DataArray-like with ``attrs["area"]``, ``attrs["start_time"]`` and ``attrs["orbital_parameters"]`` or the
    explicit ``area``, ``start_time`` and satellite position (altitude in metres)."""
    torch = _lib.require_cuda()
    if data_arr is not None and is_dataarray(data_arr):
        area = data_arr.attrs["area"]
        start_time = data_arr.attrs["start_time"]
        if sat_lon is None:
            sat_lon, sat_lat, sat_alt = _satpos(data_arr.attrs["orbital_parameters"])
    geo = as_geometry(area)
    g = geo.geo_struct()
    s = _sunz_struct(start_time, 88.0, None, False, 0)
    outs = {k: (torch.empty(geo.shape, dtype=torch.float64, device="cuda") if k in which else None)
            for k in ("sata", "satz", "suna", "sunz")}
    ptr = {k: (v.data_ptr() if v is not None else None) for k, v in outs.items()}
    _lib.check(_lib.load().b200_get_angles(C.byref(g), C.byref(s), float(sat_lon), float(sat_lat), float(sat_alt) / 1000.0,
                                           ptr["sata"], ptr["satz"], ptr["suna"], ptr["sunz"], _lib.stream_ptr(torch)))
    res = tuple(outs[k] for k in which)
    return res if device else tuple(t.cpu().numpy() for t in res)


def get_satellite_zenith_angle(data_arr=None, **kwargs):
    return get_angles(data_arr, which=("satz",), **kwargs)[0]


def compute_relative_azimuth(sat_azi, sun_azi):
    """``min(|sun - sat|, 360 - |sun - sat|)`` in the dtype of the inputs (float32 or float64)."""
    torch = _lib.require_cuda()
    was_numpy = not isinstance(sat_azi, torch.Tensor)
    a = torch.from_numpy(np.ascontiguousarray(sat_azi)) if was_numpy else sat_azi
    b = torch.from_numpy(np.ascontiguousarray(sun_azi)) if not isinstance(sun_azi, torch.Tensor) else sun_azi
    if a.dtype not in (torch.float32, torch.float64) or b.dtype != a.dtype or a.shape != b.shape:
        raise TypeError("azimuth arrays must be float32 or float64 and of the same dtype and shape")
    a, b = a.to("cuda").contiguous(), b.to("cuda").contiguous()
    out = torch.empty_like(a)
    _lib.check(_lib.load().b200_relative_azimuth(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(),
                                                 1 if a.dtype == torch.float32 else 0, _lib.stream_ptr(torch)))
    return out.cpu().numpy() if was_numpy else out
