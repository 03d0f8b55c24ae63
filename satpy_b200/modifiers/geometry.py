"""``SunZenithCorrector`` / ``EffectiveSolarPathLengthCorrector`` on the device.

Host-side mirror of satpy/modifiers/geometry.py:34-160: same class names, same
constructor keywords (``max_sza=95.0``, ``correction_limit=88.0``), same call
convention ``modifier([vis] or [vis, sza], optional_datasets=[...], **info)``,
same ``sunz_corrected`` short-circuit (:52-54), same ``IncompatibleAreas`` error
when the inputs live on different areas (satpy/tests/test_modifiers.py:178-184),
attrs copied and ``modifiers`` updated like ``apply_modifier_info``
(satpy/composites/core.py:99-125).

The arithmetic -- ``get_cos_sza`` (satpy/modifiers/angles.py:418-469), the
``max_sza`` mask (geometry.py:62-63) and ``_sunzen_corr_cos_ndarray``
(angles.py:555-584) or ``atmospheric_path_length_correction``
(satpy/utils.py:278-316) -- is ONE kernel launch (csrc/sunz.cu); lon/lat never
touch HBM.
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np

from .. import _compat, _lib, _xfer
from ..dataset import is_dataarray, payload, rewrap
from ..geometry import as_geometry
from .astronomy import solar_scalars

logger = logging.getLogger(__name__)

# satpy's own exception type when satpy is importable (its ``except IncompatibleAreas`` clauses then see ours)
IncompatibleAreas = _compat.incompatible_areas()


def _sunz_struct(start_time, limit, max_sza, lonlat_f32, method):
    s = _lib.SunzStruct()
    if start_time is None:
        s.gmst = s.ra = s.dec = 0.0
    else:
        s.gmst, s.ra, s.dec = solar_scalars(start_time)
    s.limit_deg = float(limit)
    s.has_max_sza = 0 if max_sza is None else 1
    s.max_sza_deg = 0.0 if max_sza is None else float(max_sza)
    s.lonlat_f32 = 1 if lonlat_f32 else 0
    s.method = int(method)
    s.reserved = 0
    return s


def _to_cuda_f32(torch, data, allow_f64=False):
    was_numpy = not isinstance(data, torch.Tensor)
    ok = (np.float32, np.float64) if allow_f64 else (np.float32,)
    if was_numpy:
        a = np.ascontiguousarray(data)
        if a.dtype not in ok:
            raise TypeError(f"the device sun-zenith correction works on float32/float64 data, got {a.dtype}")
        return _xfer.to_device(a), True
    if data.dtype not in ((torch.float32, torch.float64) if allow_f64 else (torch.float32,)):
        raise TypeError(f"the device sun-zenith correction works on float32/float64 data, got {data.dtype}")
    return (data if data.is_cuda else data.to("cuda")).contiguous(), False


def cos_sza(area, start_time, lonlat_f32=True, row_slice=None, device=False):
    """cos(solar zenith angle) per pixel, float64, NaN where the area has no lon/lat.

    ``get_cos_sza`` (satpy/modifiers/angles.py:418-431) with the reference's dtype flow:
    lon/lat rounded to float32 when the data is float32 (``lonlat_f32``).
    """
    torch = _lib.require_cuda()
    area = as_geometry(area)
    rows = range(area.height)[row_slice] if row_slice is not None else range(area.height)
    row0, nrows = (rows.start, len(rows)) if len(rows) else (0, 0)
    out = torch.empty((nrows, area.width), dtype=torch.float64, device="cuda")
    g = area.geo_struct(row0, nrows)
    s = _sunz_struct(start_time, 88.0, None, lonlat_f32, 0)
    _lib.check(_lib.load().b200_cos_sza(C.byref(g), C.byref(s), out.data_ptr(), _lib.stream_ptr(torch)))
    return out if device else out.cpu().numpy()


def sunz_correct(data, area=None, start_time=None, correction_limit=88.0, max_sza=95.0, sza=None,
                 method="cos", row0=0, out=None):
    """Apply the sun-zenith correction to ``data`` (``(y, x)`` or ``(bands, y, x)`` float32).

    Either ``area`` + ``start_time`` (angles derived on the device) or ``sza`` (degrees, float32,
    the ``optional_datasets`` branch of satpy/modifiers/geometry.py:64-66).  ``row0``: first area row
    that ``data`` covers (row-tiled multi-GPU use).  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.
    """
    torch = _lib.require_cuda()
    t, was_numpy = _to_cuda_f32(torch, data, allow_f64=True)
    f64 = t.dtype == torch.float64
    if t.dim() not in (2, 3):
        raise ValueError("data must be (y, x) or (bands, y, x)")
    nbands = 1 if t.dim() == 2 else t.shape[0]
    h, w = t.shape[-2:]
    if out is None:
        out = torch.empty_like(t)
    meth = {"cos": 0, "effective_path_length": 1}[method]
    lib = _lib.load()
    if sza is None:
        if area is None or start_time is None:
            raise ValueError("need either (area, start_time) or sza")
        area = as_geometry(area)
        if w != area.width or row0 + h > area.height:
            raise IncompatibleAreas(f"data shape {(h, w)} (row offset {row0}) does not fit area shape {area.shape}")
        g = area.geo_struct(row0, h)
        s = _sunz_struct(start_time, correction_limit, max_sza, not f64, meth)
        fn = lib.b200_sunz_correct_f64 if f64 else lib.b200_sunz_correct
        _lib.check(fn(t.data_ptr(), out.data_ptr(), nbands, h * w, C.byref(g), C.byref(s), _lib.stream_ptr(torch)))
    else:
        z, _ = _to_cuda_f32(torch, sza, allow_f64=True)
        if tuple(z.shape) != (h, w):
            raise IncompatibleAreas(f"sza shape {tuple(z.shape)} != data shape {(h, w)}")
        z = z.to(t.dtype)
        s = _sunz_struct(None, correction_limit, max_sza, not f64, meth)
        fn = lib.b200_sunz_correct_sza_f64 if f64 else lib.b200_sunz_correct_sza
        _lib.check(fn(t.data_ptr(), z.data_ptr(), out.data_ptr(), nbands, h * w, h * w, C.byref(s),
                      _lib.stream_ptr(torch)))
    return _xfer.to_host(out) if was_numpy else out


class _SunZenithCorrectorBase(_compat.modifier_base()):
    """``SunZenithCorrectorBase`` (satpy/modifiers/geometry.py:34-74) without the xarray/dask machinery.  A
    ``satpy.modifiers.ModifierBase`` when satpy is importable (the dependency tree reads ``.id`` and ``.attrs`` of
    every modifier, satpy/node.py:163), the stand-in of satpy_b200._compat otherwise."""

    _method = "cos"

    def __init__(self, max_sza=95.0, **kwargs):
        self.max_sza = max_sza
        self.max_sza_cos = np.cos(np.deg2rad(max_sza)) if max_sza is not None else None
        # ModifierBase/CompositeBase keep every other keyword (name, prerequisites, _satpy_id, ...) in attrs
        super().__init__(**kwargs)

    def _limit(self):
        return 88.0

    @staticmethod
    def _check_areas(arrays):
        areas = [a.attrs.get("area") for a in arrays if is_dataarray(a)]
        if any(a is None for a in areas) and not all(a is None for a in areas):
            raise ValueError("Missing 'area' attribute")
        shapes = {tuple(payload(a).shape[-2:]) for a in arrays}
        if len(shapes) > 1:
            raise IncompatibleAreas("All areas should be of the same size: %s" % sorted(shapes))
        for a in areas[1:]:
            if a is not None and a != areas[0]:
                raise IncompatibleAreas("Areas are different")

    def __call__(self, projectables, optional_datasets=None, **info):
        optional = [d for d in (optional_datasets or []) if d is not None]
        arrays = list(projectables) + list(optional)
        self._check_areas(arrays)
        vis = arrays[0]
        if vis.attrs.get("sunz_corrected"):
            logger.debug("Sun zenith correction already applied")
            return vis
        logger.debug("Applying sun zen correction")
        if len(arrays) < 2:
            res = sunz_correct(payload(vis), area=vis.attrs["area"], start_time=vis.attrs["start_time"],
                               correction_limit=self._limit(), max_sza=self.max_sza, method=self._method)
        else:
            res = sunz_correct(payload(vis), sza=payload(arrays[1]), correction_limit=self._limit(),
                               max_sza=self.max_sza, method=self._method)
        proj = rewrap(vis, res, attrs=dict(vis.attrs))
        self.apply_modifier_info(vis, proj)
        return proj


class SunZenithCorrector(_SunZenithCorrectorBase):
    """Standard ``1 / cos(sunz)`` correction with the log fall-off between ``correction_limit`` and ``max_sza``."""

    def __init__(self, correction_limit=88., **kwargs):
        self.correction_limit = correction_limit
        super().__init__(**kwargs)

    def _limit(self):
        return self.correction_limit


class EffectiveSolarPathLengthCorrector(_SunZenithCorrectorBase):
    """Li & Shibata (2006) effective path length correction (satpy/modifiers/geometry.py:121-163)."""

    _method = "effective_path_length"

    def __init__(self, correction_limit=88., **kwargs):
        self.correction_limit = correction_limit
        super().__init__(**kwargs)

    def _limit(self):
        return self.correction_limit


def sunz_reduce(data, sza, limit=55.0, max_sza=90.0, strength=1.5, out=None):
    """``sunzen_reduction`` (satpy/modifiers/angles.py:587-624): ``data * reduction(sza)``; float32 data and SZA [deg]."""
    torch = _lib.require_cuda()
    t, was_numpy = _to_cuda_f32(torch, data)
    z, _ = _to_cuda_f32(torch, sza)
    if t.dim() not in (2, 3) or tuple(z.shape) != tuple(t.shape[-2:]):
        raise IncompatibleAreas(f"sza shape {tuple(z.shape)} does not match data shape {tuple(t.shape)}")
    nbands = 1 if t.dim() == 2 else t.shape[0]
    n = z.numel()
    if out is None:
        out = torch.empty_like(t)
    _lib.check(_lib.load().b200_sunz_reduce(t.data_ptr(), z.data_ptr(), out.data_ptr(), nbands, n, n, float(limit),
                                            float(max_sza), float(strength), _lib.stream_ptr(torch)))
    return _xfer.to_host(out) if was_numpy else out


class SunZenithReducer(_SunZenithCorrectorBase):
    """Reduce the signal at large sun zenith angles (satpy/modifiers/geometry.py:166-207)."""

    def __init__(self, correction_limit=80., max_sza=90, strength=1.3, **kwargs):
        self.correction_limit = correction_limit
        self.strength = strength
        super().__init__(max_sza=max_sza, **kwargs)
        if self.max_sza is None:
            raise ValueError("`max_sza` must be defined when using the SunZenithReducer.")

    def __call__(self, projectables, optional_datasets=None, **info):
        optional = [d for d in (optional_datasets or []) if d is not None]
        arrays = list(projectables) + list(optional)
        self._check_areas(arrays)
        vis = arrays[0]
        if vis.attrs.get("sunz_corrected"):
            return vis
        torch = _lib.require_cuda()
        if len(arrays) < 2:
            # angle derived from the area inside the kernel (cos(sza) -> mask -> arccos -> degrees -> factor)
            t, was_numpy = _to_cuda_f32(torch, payload(vis))
            area = as_geometry(vis.attrs["area"])
            if t.dim() not in (2, 3) or tuple(t.shape[-2:]) != area.shape:
                raise IncompatibleAreas(f"data shape {tuple(t.shape)} does not fit area shape {area.shape}")
            nbands = 1 if t.dim() == 2 else t.shape[0]
            out = torch.empty_like(t)
            g = area.geo_struct()
            s = _sunz_struct(vis.attrs["start_time"], self.correction_limit, self.max_sza, True, 0)
            _lib.check(_lib.load().b200_sunz_reduce_area(t.data_ptr(), out.data_ptr(), nbands, area.size, C.byref(g),
                                                         C.byref(s), float(self.correction_limit), float(self.max_sza),
                                                         float(self.strength), _lib.stream_ptr(torch)))
            res = _xfer.to_host(out) if was_numpy else out
        else:
            # the solar zenith angle is a dataset: the reference goes through cos() and back (geometry.py:64-66,201);
            # that round trip stays in PyTorch (a rarely used branch), the reduction itself is the kernel
            z = payload(arrays[1])
            z = z if isinstance(z, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(z))
            cz = torch.cos(torch.deg2rad(z.to("cuda")))
            sunz = torch.rad2deg(torch.arccos(cz)).to(torch.float32)
            res = sunz_reduce(payload(vis), sunz, limit=self.correction_limit, max_sza=self.max_sza,
                              strength=self.strength)
        proj = rewrap(vis, res, attrs=dict(vis.attrs))
        self.apply_modifier_info(vis, proj)
        return proj
