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
    def __init__(self, data, dims=None, attrs=None):
        self.data = data
        nd = len(data.shape)
        if dims is None:
            dims = ("y", "x") if nd == 2 else (("bands", "y", "x") if nd == 3 else tuple(f"d{i}" for i in range(nd)))
        self.dims = tuple(dims)
        self.attrs = dict(attrs or {})

    @property
    def values(self):
        d = self.data
        return d.detach().cpu().numpy() if hasattr(d, "detach") else np.asarray(d)

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def ndim(self):
        return len(self.data.shape)

    def copy(self):
        return DataArrayLite(self.data, self.dims, dict(self.attrs))

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


def rewrap(template, data, dims=None, attrs=None):
    """New DataArray-like of the same family as ``template`` around ``data``."""
    if not is_dataarray(template):
        return data
    attrs = dict(template.attrs) if attrs is None else attrs
    dims = template.dims if dims is None else dims
    if isinstance(template, DataArrayLite):
        return DataArrayLite(data, dims, attrs)
    import xarray as xr  # template is a real DataArray
    if hasattr(data, "detach"):
        data = data.detach().cpu().numpy()
    return xr.DataArray(data, dims=dims, attrs=attrs)
