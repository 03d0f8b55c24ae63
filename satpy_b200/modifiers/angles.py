"""Sensor and sun angles on the device ("next" row of SURVEY.md section 8f).

Host-side mirror of ``get_angles``, ``get_satellite_zenith_angle`` and ``compute_relative_azimuth``
(satpy/modifiers/angles.py:336-414).  The satellite position is what ``satpy.utils.get_satpos`` extracts from
``attrs["orbital_parameters"]``: longitude and latitude in degrees, altitude in METRES (angles.py:523-527 divides by 1000).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import hashlib
import json
import logging
import os

from .. import _lib, _zarr2
from ..dataset import is_dataarray
from ..geometry import AreaDefinition, as_geometry
from .geometry import _sunz_struct

LOG = logging.getLogger(__name__)

# ``satpy.config`` keys of the reference's angle caching (satpy/modifiers/angles.py:53-213, satpy/_config.py): used from
# satpy's own config when satpy is importable, from this dict otherwise
config = {"cache_sensor_angles": False, "cache_dir": None}

_SENSOR_ANGLE_CACHE = {}   # key -> (sensor azimuth, sensor zenith) device tensors (float64)
_CACHE_VERSION = 1
_CACHE_FUNC = "_get_sensor_angles_from_sat_pos"   # the reference's cached function: the file names follow its pattern


def _config_get(key):
    try:
        import satpy
        return satpy.config.get(key, config.get(key))
    except Exception:  # noqa: BLE001 -- satpy not importable
        return config.get(key)


def _sensor_cache_key(geo, sat_lon, sat_lat, sat_alt_m):
    """Hash of the sanitised arguments: the area and the satellite position rounded to one decimal
    (``_sanitize_observer_look_args``, angles.py:306-318) -- the sensor angles of a geostationary scene do not
    depend on the time."""
    area = {"proj": {k: str(v) for k, v in sorted(geo.proj_dict.items())}, "shape": list(geo.shape),
            "extent": list(geo.area_extent)}
    return hashlib.sha1(json.dumps([area, sat_lon, sat_lat, sat_alt_m]).encode("utf8")).hexdigest()


def sensor_angle_cache_clear(cache_dir=None):
    """Forget the device-resident sensor angles and remove their zarr stores (``ZarrCacheHelper.cache_clear``)."""
    import glob
    import shutil
    _SENSOR_ANGLE_CACHE.clear()
    cache_dir = cache_dir or _config_get("cache_dir")
    if cache_dir:
        for d in glob.glob(os.path.join(cache_dir, f"{_CACHE_FUNC}_v*_*_*.zarr")):
            shutil.rmtree(d, ignore_errors=True)


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
               which=("sata", "satz", "suna", "sunz"), cache_sensor_angles=None, cache_dir=None):
    """Sensor azimuth, sensor zenith, solar azimuth, solar zenith in degrees (float64).

    Either a DataArray-like with ``attrs["area"]``, ``attrs["start_time"]`` and ``attrs["orbital_parameters"]`` or the
    explicit ``area``, ``start_time`` and satellite position (altitude in metres).

    ``cache_sensor_angles`` (default: the ``cache_sensor_angles`` config key, angles.py:321-333): the sensor angles of
    an AREA are computed once per (area, satellite position rounded to 0.1) -- with the rounded position, like the
    reference's sanitised arguments -- kept on the device and, when a ``cache_dir`` is configured, persisted as zarr
    stores named ``_get_sensor_angles_from_sat_pos_v1_{0,1}_<hash>.zarr`` (azimuth, zenith); only the sun angles are
    evaluated per scene.  Swaths are never cached (angles.py: ``uncacheable_arg_types``)."""
    torch = _lib.require_cuda()
    if data_arr is not None and is_dataarray(data_arr):
        area = data_arr.attrs["area"]
        start_time = data_arr.attrs["start_time"]
        if sat_lon is None:
            sat_lon, sat_lat, sat_alt = _satpos(data_arr.attrs["orbital_parameters"])
    geo = as_geometry(area)
    if cache_sensor_angles is None:
        cache_sensor_angles = bool(_config_get("cache_sensor_angles"))
    cached = None
    want_sensor = [k for k in which if k in ("sata", "satz")]
    if cache_sensor_angles and want_sensor and isinstance(geo, AreaDefinition):
        sat_lon, sat_lat, sat_alt = float(round(sat_lon, 1)), float(round(sat_lat, 1)), float(round(sat_alt, 1))
        cached = _cached_sensor_angles(torch, geo, sat_lon, sat_lat, sat_alt, cache_dir or _config_get("cache_dir"))
    compute = [k for k in ("sata", "satz", "suna", "sunz") if k in which and not (cached and k in cached)]
    outs = dict(cached or {})
    if compute:
        outs.update(_compute_angles(torch, geo, start_time, sat_lon, sat_lat, sat_alt, compute))
    res = tuple(outs[k] for k in which)
    return res if device else tuple(t.cpu().numpy() for t in res)


def _compute_angles(torch, geo, start_time, sat_lon, sat_lat, sat_alt, which):
    g = geo.geo_struct()
    s = _sunz_struct(start_time, 88.0, None, False, 0)
    outs = {k: (torch.empty(geo.shape, dtype=torch.float64, device="cuda") if k in which else None)
            for k in ("sata", "satz", "suna", "sunz")}
    ptr = {k: (v.data_ptr() if v is not None else None) for k, v in outs.items()}
    _lib.check(_lib.load().b200_get_angles(C.byref(g), C.byref(s), float(sat_lon), float(sat_lat), float(sat_alt) / 1000.0,
                                           ptr["sata"], ptr["satz"], ptr["suna"], ptr["sunz"], _lib.stream_ptr(torch)))
    return {k: v for k, v in outs.items() if v is not None}


def _cached_sensor_angles(torch, geo, sat_lon, sat_lat, sat_alt, cache_dir):
    key = _sensor_cache_key(geo, sat_lon, sat_lat, sat_alt)
    hit = _SENSOR_ANGLE_CACHE.get(key)
    if hit is not None:
        LOG.debug("Reusing device-resident sensor angles")
        return {"sata": hit[0], "satz": hit[1]}
    paths = None
    if cache_dir:
        paths = [os.path.join(cache_dir, f"{_CACHE_FUNC}_v{_CACHE_VERSION}_{i}_{key}.zarr") for i in range(2)]
        if all(os.path.isdir(p) for p in paths):
            try:
                arrs = [torch.from_numpy(_zarr2.read_root_array(p)).to("cuda") for p in paths]
                if all(tuple(a.shape) == geo.shape for a in arrs):
                    _SENSOR_ANGLE_CACHE[key] = (arrs[0], arrs[1])
                    return {"sata": arrs[0], "satz": arrs[1]}
            except IOError:
                pass
    # the sensor angles do not depend on the time: any start_time will do for the two outputs requested
    out = _compute_angles(torch, geo, None, sat_lon, sat_lat, sat_alt, ("sata", "satz"))
    _SENSOR_ANGLE_CACHE[key] = (out["sata"], out["satz"])
    if paths:
        os.makedirs(cache_dir, exist_ok=True)
        for p, k in zip(paths, ("sata", "satz")):
            _zarr2.write_root_array(p, out[k].cpu().numpy())
    return out


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
