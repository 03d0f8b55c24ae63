"""A minimal labelled-array container used when xarray is not installed.

satpy moves ``xarray.DataArray`` objects between readers, modifiers and
resamplers.  The plug-ins in this package accept real ``xarray.DataArray``
objects when xarray is importable and otherwise this small stand-in, which
carries exactly what the hot path reads: ``.values`` (numpy or torch),
``.dims``, ``.attrs`` (with ``area``, ``start_time``, ``_FillValue`` ...).
"""
from __future__ import annotations

import numpy as np


class DataArrayLite:
    def __init__(self, data, dims=None, attrs=None, coords=None, name=None):
        self.data = data
        nd = len(data.shape)
        if dims is None:
            dims = ("y", "x") if nd == 2 else (("bands", "y", "x") if nd == 3 else tuple(f"d{i}" for i in range(nd)))
        self.dims = tuple(dims)
        self.attrs = dict(attrs or {})
        # name -> 1-D values along the dimension of that name (``bands``, ``x``, ``y`` ...) or a scalar (``crs``)
        self.coords = dict(coords or {})
        self.name = name

    @property
    def values(self):
        """numpy view of the data; a CUDA tensor is copied through page-locked memory (satpy_b200._xfer)."""
        d = self.data
        if hasattr(d, "detach"):
            if d.is_cuda:
                from ._xfer import to_host
                return to_host(d.detach())
            return d.detach().numpy()
        return np.asarray(d)

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def ndim(self):
        return len(self.data.shape)

    def copy(self, data=None):
        return DataArrayLite(self.data if data is None else data, self.dims, dict(self.attrs), dict(self.coords),
                             self.name)

    def __repr__(self):
        return f"DataArrayLite(shape={self.shape}, dtype={self.dtype}, dims={self.dims})"


def is_dataarray(obj) -> bool:
    return hasattr(obj, "attrs") and hasattr(obj, "dims") and (hasattr(obj, "data") or hasattr(obj, "values"))


def payload(obj):
    """The raw array (numpy or torch) behind a DataArray-like object or a bare array."""
    if is_dataarray(obj):
        d = obj.data
        if hasattr(d, "compute"):  # dask array: materialise once, here
            d = d.compute()
        return d
    return obj


_SPATIAL = ("y", "x", "crs")


def rewrap(template, data, dims=None, attrs=None, new_area=None):
    """New DataArray-like of the same family as ``template`` around ``data``.

    Without ``new_area`` (modifiers): same grid -- every coordinate and the name are kept, like the reference's
    ``proj.copy()`` (satpy/modifiers/geometry.py:60-70).  With ``new_area`` (resamplers): what
    ``_update_resampled_coords`` does (satpy/resample/base.py:45-73, called at kdtree.py:190,293) -- coordinates that
    do not depend on y / x / crs (``bands``, ``time`` ...) are carried over and x / y / crs come from the target area
    (``satpy.coords.add_crs_xy_coords`` when satpy is importable, the area's projection vectors otherwise).
    """
    if not is_dataarray(template):
        return data
    attrs = dict(template.attrs) if attrs is None else attrs
    dims = tuple(template.dims if dims is None else dims)
    if isinstance(template, DataArrayLite):
        coords = dict(template.coords)
        if new_area is not None:
            coords = {k: v for k, v in coords.items() if k not in _SPATIAL}
            coords.update(_area_xy(new_area, tuple(data.shape[-2:])))
        return DataArrayLite(data, dims, attrs, coords, template.name)
    import xarray as xr  # template is a real DataArray
    if hasattr(data, "detach"):
        if data.is_cuda:
            from ._xfer import to_host
            data = to_host(data.detach())
        else:
            data = data.detach().numpy()
    if new_area is None:
        coords = {k: v for k, v in template.coords.items() if all(d in dims for d in v.dims)}
        return xr.DataArray(data, dims=dims, attrs=attrs, coords=coords, name=template.name)
    new = xr.DataArray(data, dims=dims, attrs=attrs, name=template.name)
    try:
        from satpy.resample.base import _update_resampled_coords
        return _update_resampled_coords(template, new, new_area)
    except ImportError:
        keep = {k: v for k, v in template.coords.items()
                if k not in _SPATIAL and not any(d in _SPATIAL for d in v.dims)}
        new = new.assign_coords(**keep)
        xy = _area_xy(new_area, tuple(data.shape[-2:]))
        return new.assign_coords(**{k: ((k,), v) for k, v in xy.items() if k in dims})


def _area_xy(area, shape):
    """``{"x": ..., "y": ...}`` projection coordinates of an area's pixel centres (nothing for swaths or when the
    data covers only a row window of the area)."""
    get = getattr(area, "get_proj_vectors", None)
    if get is None or tuple(getattr(area, "shape", ())) != tuple(shape):
        return {}
    x, y = get()
    return {"x": np.asarray(x), "y": np.asarray(y)}
