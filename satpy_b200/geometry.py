"""Geometry objects of the B200 hot path.

``AreaDefinition`` / ``SwathDefinition`` mirror the part of
``pyresample.geometry`` that satpy touches on this path (SURVEY.md Appendix E:
constructor signature, ``shape``/``width``/``height``, ``area_extent`` of the
OUTER pixel edges, ``get_lonlats``, ``get_proj_vectors``, slicing,
hashability) and know how to describe themselves to the CUDA library as a
``b200_geo_t`` (include/satpy_b200.h).  Real pyresample objects are accepted
anywhere a geometry is expected (``as_geometry``) -- they are translated, not
required.

Only the *constants* of each map projection are derived here, on the host,
once per area (Snyder 1987 / PROJ parameter set-up); the per-pixel transforms
run on the device (satpy_b200/csrc/b200_proj.cuh).  The reference reaches the
same transforms through pyproj via ``AreaDefinition.get_lonlats``
(satpy/modifiers/angles.py:434-441, satpy/resample/kdtree.py:79-87).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib

_ELLPS = {  # (a, rf), rf = 0 -> sphere
    "WGS84": (6378137.0, 298.257223563), "GRS80": (6378137.0, 298.257222101), "WGS72": (6378135.0, 298.26),
    "intl": (6378388.0, 297.0), "bessel": (6377397.155, 299.1528128), "clrk66": (6378206.4, 294.9786982138982),
    "sphere": (6370997.0, 0.0),
}
_DATUM = {"WGS84": "WGS84", "NAD83": "GRS80", "NAD27": "clrk66"}
_HALFPI = math.pi / 2
_EPS10 = 1e-10
_KIND = {"longlat": _lib.PROJ_LONGLAT, "latlong": _lib.PROJ_LONGLAT, "lonlat": _lib.PROJ_LONGLAT,
         "latlon": _lib.PROJ_LONGLAT, "eqc": _lib.PROJ_EQC, "merc": _lib.PROJ_MERC, "geos": _lib.PROJ_GEOS,
         "laea": _lib.PROJ_LAEA, "lcc": _lib.PROJ_LCC, "stere": _lib.PROJ_STERE}


def _parse_proj(projection) -> dict:
    """Accept a PROJ dict or a ``+proj=... +k=v`` string."""
    if isinstance(projection, dict):
        return dict(projection)
    if isinstance(projection, str):
        out = {}
        for tok in projection.replace("+", " ").split():
            if "=" in tok:
                k, v = tok.split("=", 1)
                try:
                    out[k] = float(v)
                except ValueError:
                    out[k] = v
            else:
                out[tok] = True
        return out
    if hasattr(projection, "to_dict"):  # pyproj.CRS
        return dict(projection.to_dict())
    raise TypeError(f"cannot interpret projection {projection!r}")


def _ellipsoid(p: dict) -> tuple[float, float]:
    if "R" in p:
        return float(p["R"]), 0.0
    name = p.get("ellps")
    if name is None and p.get("datum") is not None:
        name = _DATUM[p["datum"]]
    if name is None and "a" not in p:
        name = "WGS84"
    a = es = None
    if name is not None:
        a, rf = _ELLPS[name]
        es = 0.0 if rf == 0 else 2.0 / rf - 1.0 / rf ** 2
    if "a" in p:
        a = float(p["a"])
        es = 0.0 if es is None else es
    if "b" in p:
        b = float(p["b"])
        es = 1.0 - (b * b) / (a * a)
    elif "rf" in p:
        es = 2.0 / float(p["rf"]) - 1.0 / float(p["rf"]) ** 2
    elif "f" in p:
        es = 2.0 * float(p["f"]) - float(p["f"]) ** 2
    elif "es" in p:
        es = float(p["es"])
    return a, es


def _tsfn(phi, e):
    s = math.sin(phi)
    return math.tan(0.5 * (_HALFPI - phi)) / math.pow((1.0 - e * s) / (1.0 + e * s), 0.5 * e)


def _msfn(phi, es):
    s = math.sin(phi)
    return math.cos(phi) / math.sqrt(1.0 - es * s * s)


def _qsfn(sinphi, e, one_es):
    if e < 1e-7:
        return 2.0 * sinphi
    con = e * sinphi
    return one_es * (sinphi / (1.0 - con * con) - (0.5 / e) * math.log((1.0 - con) / (1.0 + con)))


def proj_struct(projection) -> _lib.ProjStruct:
    """Derive the device-side constants of one projection (see ``b200_proj_t``)."""
    p = _parse_proj(projection)
    name = str(p.get("proj", "longlat"))
    if name not in _KIND:
        raise NotImplementedError(f"projection '{name}' is not implemented on the device "
                                  f"(have: {sorted(set(_KIND))})")
    a, es = _ellipsoid(p)
    S = _lib.ProjStruct()
    S.kind = _KIND[name]
    S.mode = 0
    S.a, S.es, S.e, S.one_es = a, es, math.sqrt(es), 1.0 - es
    S.lam0 = math.radians(float(p.get("lon_0", 0.0)))
    S.phi0 = math.radians(float(p.get("lat_0", 0.0)))
    S.x0 = float(p.get("x_0", 0.0))
    S.y0 = float(p.get("y_0", 0.0))
    S.k0 = float(p.get("k_0", p.get("k", 1.0)))
    q = [0.0] * 12
    e = S.e
    if name == "eqc":
        q[0] = math.cos(math.radians(float(p.get("lat_ts", 0.0))))
    elif name == "merc":
        if "lat_ts" in p:
            ts = math.radians(float(p["lat_ts"]))
            S.k0 = _msfn(ts, es) if es != 0.0 else math.cos(ts)
    elif name == "geos":
        h = float(p["h"])
        S.mode = 1 if str(p.get("sweep", "y")) == "x" else 0
        q[0] = h / a
        q[1] = 1.0 + q[0]
        q[2] = q[1] * q[1] - 1.0
        q[3] = math.sqrt(1.0 - es)
        q[4] = 1.0 - es
        q[5] = 1.0 / (1.0 - es)
    elif name == "lcc":
        phi1 = math.radians(float(p["lat_1"]))
        phi2 = math.radians(float(p.get("lat_2", p["lat_1"])))
        if "lat_0" not in p:
            S.phi0 = phi1
        secant = abs(phi1 - phi2) >= _EPS10
        n = math.sin(phi1)
        if es != 0.0:
            m1 = _msfn(phi1, es)
            ml1 = _tsfn(phi1, e)
            if secant:
                n = math.log(m1 / _msfn(phi2, es)) / math.log(ml1 / _tsfn(phi2, e))
            c = m1 * math.pow(ml1, -n) / n
            rho0 = 0.0 if abs(abs(S.phi0) - _HALFPI) < _EPS10 else c * math.pow(_tsfn(S.phi0, e), n)
        else:
            if secant:
                n = math.log(math.cos(phi1) / math.cos(phi2)) / math.log(
                    math.tan(math.pi / 4 + 0.5 * phi2) / math.tan(math.pi / 4 + 0.5 * phi1))
            c = math.cos(phi1) * math.pow(math.tan(math.pi / 4 + 0.5 * phi1), n) / n
            rho0 = 0.0 if abs(abs(S.phi0) - _HALFPI) < _EPS10 else \
                c * math.pow(math.tan(math.pi / 4 + 0.5 * S.phi0), -n)
        q[0], q[1], q[2] = n, c, rho0
    elif name == "stere":
        if abs(abs(S.phi0) - _HALFPI) >= _EPS10:
            raise NotImplementedError("only polar stereographic (lat_0 = +-90) is implemented on the device")
        S.mode = 0 if S.phi0 > 0 else 1
        phits = abs(math.radians(float(p["lat_ts"]))) if "lat_ts" in p else _HALFPI
        if es != 0.0:
            if abs(phits - _HALFPI) < _EPS10:
                akm1 = 2.0 * S.k0 / math.sqrt(math.pow(1 + e, 1 + e) * math.pow(1 - e, 1 - e))
            else:
                s = math.sin(phits)
                akm1 = math.cos(phits) / _tsfn(phits, e) / math.sqrt(1.0 - es * s * s)
        else:
            akm1 = (math.cos(phits) / math.tan(math.pi / 4 - 0.5 * phits)
                    if abs(phits - _HALFPI) >= _EPS10 else 2.0 * S.k0)
        q[0] = akm1
    elif name == "laea":
        t = abs(S.phi0)
        if abs(t - _HALFPI) < _EPS10:
            S.mode = 0 if S.phi0 > 0 else 1
        elif t < _EPS10:
            S.mode = 2
        else:
            S.mode = 3
        if es != 0.0:
            qp = _qsfn(1.0, e, S.one_es)
            rq = math.sqrt(0.5 * qp)
            sinphi = math.sin(S.phi0)
            sinb1 = _qsfn(sinphi, e, S.one_es) / qp
            cosb1 = math.sqrt(max(0.0, 1.0 - sinb1 * sinb1))
            if S.mode == 3:
                dd = math.cos(S.phi0) / (math.sqrt(1.0 - es * sinphi * sinphi) * rq * cosb1)
            else:
                dd = 1.0 / rq if S.mode == 2 else 1.0
            es2, es3 = es * es, es * es * es
            q[0:7] = [qp, rq, sinb1, cosb1, dd, rq * dd, rq / dd]
            q[7] = es / 3.0 + es2 * (31.0 / 180.0) + es3 * (517.0 / 5040.0)
            q[8] = es2 * (23.0 / 360.0) + es3 * (251.0 / 3780.0)
            q[9] = es3 * (761.0 / 45360.0)
        else:
            q[2], q[3] = math.sin(S.phi0), math.cos(S.phi0)
    for i, v in enumerate(q):
        S.p[i] = v
    return S


class BaseDefinition:
    """Common surface of area and swath geometries."""

    shape: tuple[int, int]

    @property
    def size(self):
        return self.shape[0] * self.shape[1]

    @property
    def height(self):
        return self.shape[0]

    @property
    def width(self):
        return self.shape[1]

    def geo_struct(self, row0: int = 0, nrows: int | None = None) -> _lib.GeoStruct:  # pragma: no cover
        raise NotImplementedError

    def get_lonlats(self, row_slice: slice | None = None, device: bool = False):
        """lon/lat (degrees, float64) of the pixel centres, computed on the GPU.

        ``inf`` where the projection has no inverse, like pyproj's default.
        """
        torch = _lib.require_cuda()
        row0, nrows = _rows(row_slice, self.height)
        lons = torch.empty((nrows, self.width), dtype=torch.float64, device="cuda")
        lats = torch.empty_like(lons)
        g = self.geo_struct(row0, nrows)
        _lib.check(_lib.load().b200_area_lonlats(C.byref(g), lons.data_ptr(), lats.data_ptr(), _lib.stream_ptr(torch)))
        if device:
            return lons, lats
        return lons.cpu().numpy(), lats.cpu().numpy()


def _rows(row_slice, height):
    if row_slice is None:
        return 0, height
    start, stop, step = row_slice.indices(height)
    if step != 1:
        raise ValueError("row slices must be contiguous")
    return start, max(0, stop - start)


class AreaDefinition(BaseDefinition):
    """``AreaDefinition(area_id, description, proj_id, projection, width, height, area_extent)``.

    Same constructor order as pyresample's (satpy/tests/test_resample.py:66-73,
    satpy/readers/core/abi.py:271-291); ``area_extent = (x_ll, y_ll, x_ur, y_ur)``
    of the outer pixel edges, row 0 is the northernmost line.
    """

    def __init__(self, area_id, description, proj_id, projection, width, height, area_extent):
        self.area_id = area_id
        self.description = description
        self.proj_id = proj_id
        self.proj_dict = _parse_proj(projection)
        self._width = int(width)
        self._height = int(height)
        self.area_extent = tuple(float(v) for v in area_extent)
        if self._width <= 0 or self._height <= 0:
            raise ValueError("area width and height must be positive")
        self.pixel_size_x = (self.area_extent[2] - self.area_extent[0]) / self._width
        self.pixel_size_y = (self.area_extent[3] - self.area_extent[1]) / self._height
        self.shape = (self._height, self._width)
        self._proj = proj_struct(self.proj_dict)
        self._key = (tuple(sorted((k, str(v)) for k, v in self.proj_dict.items())), self.shape, self.area_extent)

    # pixel (0, 0) centre; pyresample: pixel_upper_left
    @property
    def pixel_upper_left(self):
        return (self.area_extent[0] + self.pixel_size_x / 2.0, self.area_extent[3] - self.pixel_size_y / 2.0)

    def get_proj_vectors(self):
        """1-D x and y projection coordinates of the pixel centres (satpy/coords.py:48-50)."""
        x_ul, y_ul = self.pixel_upper_left
        return (np.arange(self.width) * self.pixel_size_x + x_ul,
                np.arange(self.height) * -self.pixel_size_y + y_ul)

    def geo_struct(self, row0=0, nrows=None):
        g = _lib.GeoStruct()
        g.is_swath = 0
        g.lonlat_is_f32 = 0
        g.area.proj = self._proj
        g.area.width, g.area.height = self.width, self.height
        g.area.pixel_size_x, g.area.pixel_size_y = self.pixel_size_x, self.pixel_size_y
        g.area.x_ul, g.area.y_ul = self.pixel_upper_left
        g.lons = g.lats = None
        g.width, g.height = self.width, self.height
        g.row0 = int(row0)
        g.nrows = self.height - int(row0) if nrows is None else int(nrows)
        return g

    def __getitem__(self, key):
        """``area[slice_y, slice_x]`` -> cropped area (satpy/scene.py:942-947)."""
        ys, xs = key
        y0, y1, sy = ys.indices(self.height)
        x0, x1, sx = xs.indices(self.width)
        if sy != 1 or sx != 1:
            raise ValueError("area slices must be contiguous")
        ext = (self.area_extent[0] + x0 * self.pixel_size_x, self.area_extent[3] - y1 * self.pixel_size_y,
               self.area_extent[0] + x1 * self.pixel_size_x, self.area_extent[3] - y0 * self.pixel_size_y)
        return AreaDefinition(self.area_id, self.description, self.proj_id, self.proj_dict, x1 - x0, y1 - y0, ext)

    def row_window(self, row0: int, nrows: int) -> "AreaDefinition":
        return self[slice(row0, row0 + nrows), slice(0, self.width)]

    def __hash__(self):
        return hash(self._key)

    def __eq__(self, other):
        return isinstance(other, AreaDefinition) and self._key == other._key

    def __repr__(self):
        return f"AreaDefinition({self.area_id!r}, {self.proj_dict}, {self.width}x{self.height}, {self.area_extent})"


class SwathDefinition(BaseDefinition):
    """Explicit lon/lat arrays (degrees), numpy or torch, float32 or float64."""

    def __init__(self, lons, lats):
        if hasattr(lons, "dims") and hasattr(lons, "values"):  # xarray.DataArray (torch tensors have .values() too)
            lons, lats = lons.values, lats.values
        self.lons = lons
        self.lats = lats
        if tuple(lons.shape) != tuple(lats.shape) or len(lons.shape) != 2:
            raise ValueError("swath lons/lats must be 2-D arrays of the same shape")
        self.shape = (int(lons.shape[0]), int(lons.shape[1]))
        self._dev = None

    def _device_arrays(self):
        if self._dev is None:
            torch = _lib.require_cuda()

            def up(a):
                t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
                if t.dtype not in (torch.float32, torch.float64):
                    t = t.to(torch.float64)
                return t.to("cuda").contiguous()
            lo, la = up(self.lons), up(self.lats)
            if lo.dtype != la.dtype:
                lo, la = lo.to(torch.float64), la.to(torch.float64)
            self._dev = (lo, la)
        return self._dev

    def geo_struct(self, row0=0, nrows=None):
        torch = _lib.require_cuda()
        lo, la = self._device_arrays()
        g = _lib.GeoStruct()
        g.is_swath = 1
        g.lonlat_is_f32 = 1 if lo.dtype == torch.float32 else 0
        g.lons, g.lats = lo.data_ptr(), la.data_ptr()
        g.width, g.height = self.width, self.height
        g.row0 = int(row0)
        g.nrows = self.height - int(row0) if nrows is None else int(nrows)
        return g

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other


def as_geometry(obj):
    """Translate a pyresample geometry (duck-typed) into ours; ours pass through."""
    if isinstance(obj, BaseDefinition):
        return obj
    if hasattr(obj, "area_extent") and hasattr(obj, "width"):
        proj = obj.crs.to_dict() if hasattr(obj, "crs") and hasattr(obj.crs, "to_dict") else obj.proj_dict
        return AreaDefinition(getattr(obj, "area_id", "area"), getattr(obj, "description", ""),
                              getattr(obj, "proj_id", ""), proj, obj.width, obj.height, obj.area_extent)
    if hasattr(obj, "lons") and hasattr(obj, "lats"):
        return SwathDefinition(obj.lons, obj.lats)
    raise TypeError(f"not a geometry definition: {obj!r}")
