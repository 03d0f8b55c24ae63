"""``b200_ewa``: elliptical weighted averaging resampler on the device.

Host-side mirror of ``pyresample.ewa.DaskEWAResampler``, which satpy registers unchanged as ``ewa``
(satpy/resample/ewa.py:3-11): constructed as ``cls(source_geo_def, target_geo_def)``; ``precompute()`` runs ``ll2cr``
once per geometry pair; ``compute(data, rows_per_scan=, weight_count=10000, weight_min=0.01, weight_distance_max=1.0,
weight_delta_max=10.0, weight_sum_min=-1.0, maximum_weight_mode=False, fill_value=)`` runs ``fornav``.
``rows_per_scan`` defaults to ``data.attrs["rows_per_scan"]`` (set by the reader,
satpy/readers/core/viirs_atms_sdr.py:283), else the whole swath is one scan.
Parity is UNPINNED in the reference (no test exercises ``ewa``); see oracle/ewa.py.
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np

from .. import _compat, _lib, _xfer
from ..dataset import is_dataarray, payload, rewrap
from ..geometry import AreaDefinition, as_geometry

LOG = logging.getLogger(__name__)


def ewa_params(weight_count=10000, weight_min=0.01, weight_distance_max=1.0, weight_delta_max=10.0,
               weight_sum_min=-1.0, maximum_weight_mode=False) -> _lib.EwaParamsStruct:
    p = _lib.EwaParamsStruct()
    p.weight_count = int(weight_count)
    p.maximum_weight_mode = 1 if maximum_weight_mode else 0
    p.weight_min = float(weight_min)
    p.weight_distance_max = float(weight_distance_max)
    p.weight_delta_max = float(weight_delta_max)
    p.weight_sum_min = float(weight_sum_min)
    return p


def max_weight_encode(accums):
    """Value grid of ``maximum_weight_mode`` pass 2 on one rank -> what the ranks MAX-reduce: cells nobody on this rank
    wrote hold -inf (the grid's initial value), a kept sample that is invalid (NaN) becomes +inf so that it survives
    the reduction -- the sample with the largest weight decides, valid or not, as in the sequential loop."""
    import torch
    return torch.where(torch.isnan(accums), torch.full_like(accums, float("inf")), accums)


def max_weight_decode(enc):
    """After the MAX reduction: +-inf -> NaN (an invalid sample won / no sample reached the cell)."""
    import torch
    return torch.where(torch.isinf(enc), torch.full_like(enc, float("nan")), enc)


def scan_band(height: int, rows_per_scan: int, scan_rows=None):
    """Validate a band of swath rows ``(row0, nrows)`` for fornav: it has to start on a scan boundary and hold whole
    scans (ellipse parameters are computed per scan).  ``None``: the whole swath."""
    if scan_rows is None:
        return 0, height
    row0, nrows = int(scan_rows[0]), int(scan_rows[1])
    if row0 < 0 or nrows < 0 or row0 + nrows > height:
        raise ValueError(f"scan_rows {scan_rows} outside the {height}-row swath")
    rps = rows_per_scan if rows_per_scan > 0 else height
    if row0 % rps or nrows % rps:
        raise ValueError(f"scan_rows {scan_rows} does not hold whole scans of {rps} rows")
    return row0, nrows


def scan_bands(height: int, rows_per_scan: int, world_size: int):
    """Split a swath of ``height`` rows into ``world_size`` bands of whole scans, as even as possible (the last ranks
    may get an empty band)."""
    rps = rows_per_scan if rows_per_scan > 0 else height
    nscans = height // rps
    per = -(-nscans // world_size)
    bands = []
    for r in range(world_size):
        s0 = min(nscans, r * per)
        bands.append((s0 * rps, (min(nscans, s0 + per) - s0) * rps))
    return bands


class B200EWAResampler(_compat.resampler_base()):
    """ll2cr + fornav; the accumulation grids live on the device for the duration of ``compute``."""

    def __init__(self, source_geo_def, target_geo_def):
        if _compat.resampler_base() is not object:
            super().__init__(source_geo_def, target_geo_def)
        self.source_geo_def = as_geometry(source_geo_def)
        self.target_geo_def = as_geometry(target_geo_def)
        self._orig_target = target_geo_def
        if not isinstance(self.target_geo_def, AreaDefinition):
            raise NotImplementedError("the EWA target must be an AreaDefinition")
        self._cr = None

    def precompute(self, cache_dir=None, **kwargs):
        """``ll2cr``: project every swath pixel into fractional grid coordinates; returns ``points_in_grid``."""
        del kwargs, cache_dir
        torch = _lib.require_cuda()
        if self._cr is not None:
            return self._cr["points_in_grid"]
        src, dst = self.source_geo_def, self.target_geo_def
        gs, gd = src.geo_struct(), dst.geo_struct()
        f32 = bool(gs.is_swath and gs.lonlat_is_f32)
        dtype = torch.float32 if f32 else torch.float64
        cols = torch.empty(src.shape, dtype=dtype, device="cuda")
        rows = torch.empty(src.shape, dtype=dtype, device="cuda")
        n_in = C.c_int64(0)
        _lib.check(_lib.load().b200_ll2cr(C.byref(gs), C.byref(gd), cols.data_ptr(), rows.data_ptr(), C.byref(n_in),
                                          _lib.stream_ptr(torch)))
        self._cr = {"cols": cols, "rows": rows, "f32": f32, "points_in_grid": int(n_in.value)}
        return self._cr["points_in_grid"]

    def compute(self, data, cache_id=None, rows_per_scan=None, fill_value=None, weight_count=10000, weight_min=0.01,
                weight_distance_max=1.0, weight_delta_max=10.0, weight_sum_min=-1.0, maximum_weight_mode=False,
                scan_rows=None, group=None, **kwargs):
        """``fornav``.  ``scan_rows=(row0, nrows)`` restricts the accumulation to a band of whole scans of the swath
        (``data`` is still the full swath) and ``group`` names a ``torch.distributed`` process group whose ranks each
        accumulate their own band: the (sum of weights, sum of weighted values) grids are summed over the ranks before
        the normalisation, the decomposition pyresample's dask wrapper uses over input chunks."""
        del kwargs, cache_id
        torch = _lib.require_cuda()
        lib = _lib.load()
        if self._cr is None:
            self.precompute()
        if rows_per_scan is None:
            rows_per_scan = data.attrs.get("rows_per_scan", 0) if is_dataarray(data) else 0
        if fill_value is None:
            fill_value = np.nan
        arr = payload(data)
        was_numpy = not isinstance(arr, torch.Tensor)
        t = torch.from_numpy(np.ascontiguousarray(arr)) if was_numpy else arr
        if t.dtype != torch.float32:
            raise TypeError(f"the device EWA resampler works on float32 data, got {t.dtype}")
        t = t.to("cuda", non_blocking=True).contiguous()
        src, dst = self.source_geo_def, self.target_geo_def
        if tuple(t.shape[-2:]) != src.shape:
            raise ValueError(f"data shape {tuple(t.shape)} does not end with the source shape {src.shape}")
        row0, nrows = scan_band(src.height, int(rows_per_scan), scan_rows)
        lead = tuple(t.shape[:-2])
        nbands = int(np.prod(lead)) if lead else 1
        out = torch.empty(lead + dst.shape, dtype=torch.float32, device="cuda")
        if self._cr["points_in_grid"] == 0:   # nothing of the swath reaches the grid: all fill, like the library
            out.fill_(float(fill_value))
        else:
            weights = torch.zeros((nbands,) + dst.shape, dtype=torch.float32, device="cuda")
            accums = torch.zeros((nbands,) + dst.shape, dtype=torch.float32, device="cuda")
            p = ewa_params(weight_count, weight_min, weight_distance_max, weight_delta_max, weight_sum_min,
                           maximum_weight_mode)
            stream = _lib.stream_ptr(torch)
            cr_item = 4 if self._cr["f32"] else 8
            first = row0 * src.width

            def accumulate(mode):
                if nrows <= 0:
                    return
                p.maximum_weight_mode = mode
                _lib.check(lib.b200_fornav_accum(self._cr["cols"].data_ptr() + first * cr_item,
                                                 self._cr["rows"].data_ptr() + first * cr_item,
                                                 1 if self._cr["f32"] else 0, nrows, src.width, int(rows_per_scan),
                                                 t.data_ptr() + first * 4, nbands, src.size, float("nan"), dst.height,
                                                 dst.width, weights.data_ptr(), accums.data_ptr(), C.byref(p), stream))
            if group is None:
                accumulate(1 if maximum_weight_mode else 0)
            elif not maximum_weight_mode:
                import torch.distributed as dist
                accumulate(0)
                dist.all_reduce(weights, op=dist.ReduceOp.SUM, group=group)
                dist.all_reduce(accums, op=dist.ReduceOp.SUM, group=group)
            else:
                # max-by-weight reduction: the maxima of the weights over all ranks first, then every rank writes the
                # samples that reach them, and the value grids are merged (see max_weight_encode)
                import torch.distributed as dist
                accumulate(2)
                dist.all_reduce(weights, op=dist.ReduceOp.MAX, group=group)
                accums.fill_(float("-inf"))
                accumulate(3)
                enc = max_weight_encode(accums)
                dist.all_reduce(enc, op=dist.ReduceOp.MAX, group=group)
                accums = max_weight_decode(enc)
            p.maximum_weight_mode = 1 if maximum_weight_mode else 0
            counts = (C.c_int64 * nbands)()
            _lib.check(lib.b200_fornav_finalize(weights.data_ptr(), accums.data_ptr(), out.data_ptr(), nbands, dst.size,
                                                C.byref(p), float(fill_value), counts, stream))
            self.valid_counts = list(counts)
        res = _xfer.to_host(out) if was_numpy else out
        if not is_dataarray(data):
            return res
        attrs = dict(data.attrs)
        attrs["area"] = self._orig_target
        return rewrap(data, res, dims=tuple(data.dims), attrs=attrs, new_area=self.target_geo_def)

    def resample(self, data, cache_dir=None, mask_area=None, **kwargs):
        del mask_area
        self.precompute(cache_dir=cache_dir)
        return self.compute(data, **kwargs)
