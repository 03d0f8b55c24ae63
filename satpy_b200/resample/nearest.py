"""``b200_nearest``: nearest-neighbour resampler backed by the CUDA bucket-grid search.

Host-side mirror of ``KDTreeResampler`` (satpy/resample/kdtree.py:26-190) and of the
``pyresample.resampler.BaseResampler`` contract it implements:

* constructed as ``cls(source_geo_def, target_geo_def)`` (satpy/resample/base.py:132);
* ``resample(data, **kwargs)`` builds the invalid-data mask for swath sources (or when
  ``mask_area=True``), then ``precompute(mask=, radius_of_influence=, epsilon=, cache_dir=)``
  and ``compute(data, fill_value=)``;
* the neighbour info has the reference's names and meaning (``NN_COORDINATES``,
  satpy/resample/kdtree.py:20-22): ``valid_input_index`` (bool, source shape),
  ``valid_output_index`` (bool, target shape), ``index_array`` (target shape + (1,), positions in
  the COMPRESSED valid-input vector, -1 = no neighbour) -- int32 here instead of int64;
* the result keeps the dtype of the input (satpy/tests/test_resample.py:109-129) and gets
  ``fill_value`` where no neighbour lies within ``radius_of_influence`` (chord distance, metres).

Neighbour info stays resident on the device and is reused for every dataset resampled between the
same two geometries (the analogue of ``self._index_caches``, kdtree.py:58,140).

``resample()`` chooses the work by what is asked for:

* the FIRST float32 dataset between two geometries (no mask, no ``cache_dir``) runs the fused kernel -- search and
  gather in one pass, no index array written (``b200_nn_query_gather_f32``; config 5 would otherwise write 12.85 GB of
  indices it never reads again);
* a SECOND dataset on the same instance (satpy reuses one resampler for every dataset of a source area,
  satpy/scene.py:877,899-901) searches once more, this time keeping the raw gather index on the device, and every
  later dataset is a gather;
* ``precompute()`` called directly, ``cache_dir`` or ``neighbour_info()`` produce the reference's three arrays.

numpy in -> numpy out (page-locked, copied while the next rows are searched), CUDA tensor in -> CUDA tensor out;
``to_host=`` / ``out=`` override that.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import json
import logging
import os

import numpy as np

from .. import _compat, _lib, _xfer, _zarr2
from ..dataset import is_dataarray, payload, rewrap
from ..geometry import AreaDefinition, SwathDefinition, as_geometry

LOG = logging.getLogger(__name__)

SEARCH_METHODS = {"auto": 0, "bucket": 1, "fixed_grid": 2, "fixed_grid_strict": 3}
NN_LIGHT = 0x100   # B200_NN_LIGHT (include/satpy_b200.h)
_PLAN_CACHE = {}   # row bands and reachable-column spans of host results, per (geometries, rows, radius)

NN_COORDINATES = {"valid_input_index": ("y1", "x1"),
                  "valid_output_index": ("y2", "x2"),
                  "index_array": ("y2", "x2", "z2")}

_WGS84_A, _WGS84_F = 6378137.0, 1 / 298.257223563


def _geocentric_spacing(lons, lats):
    """Distances between consecutive points in WGS84 geocentric cartesian metres."""
    e2 = 2 * _WGS84_F - _WGS84_F ** 2
    lo, la = np.deg2rad(lons), np.deg2rad(lats)
    n = _WGS84_A / np.sqrt(1 - e2 * np.sin(la) ** 2)
    xyz = np.stack([n * np.cos(la) * np.cos(lo), n * np.cos(la) * np.sin(lo), n * (1 - e2) * np.sin(la)], -1)
    with np.errstate(invalid="ignore"):
        return np.linalg.norm(np.diff(xyz, axis=0), axis=-1)


def geocentric_resolution(geo) -> float | None:
    """pyresample's ``geocentric_resolution``: pixel spacing along the middle row/column (areas) or
    between the first two columns of the middle row (swaths), in metres."""
    if isinstance(geo, SwathDefinition):
        row = geo.height // 2
        lo = np.asarray(payload_to_numpy(geo.lons[row, :2]), dtype=np.float64)
        la = np.asarray(payload_to_numpy(geo.lats[row, :2]), dtype=np.float64)
        d = _geocentric_spacing(lo, la)
    else:
        mid_row = geo[slice(geo.height // 2, geo.height // 2 + 1), slice(0, geo.width)]
        mid_col = geo[slice(0, geo.height), slice(geo.width // 2, geo.width // 2 + 1)]
        lo_r, la_r = mid_row.get_lonlats()
        lo_c, la_c = mid_col.get_lonlats()
        d = np.concatenate([_geocentric_spacing(lo_r[0], la_r[0]), _geocentric_spacing(lo_c[:, 0], la_c[:, 0])])
    d = d[np.isfinite(d)]
    return float(d.max()) if d.size else None


def payload_to_numpy(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


class _Grid:
    """Owns one ``b200_nn_grid_t`` handle."""

    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        if self.handle:
            try:
                _lib.load().b200_nn_grid_destroy(self.handle)
            except Exception:  # interpreter shutdown
                pass
            self.handle = None


class B200NearestResampler(_compat.resampler_base()):
    """Nearest-neighbour resampling on one B200."""

    ROW_CHUNKS = 24  # host results: rows are produced in this many bands, each copied out while the next is computed
    SPAN_ROWS = 416  # host results: rows per reachable-column span (one device->host copy each; 4 per band of config 5)
    trace = None     # set to a list to collect (stage, start event, end event) of the fused path (bench.py)

    def __init__(self, source_geo_def, target_geo_def):
        if _compat.resampler_base() is not object:
            super().__init__(source_geo_def, target_geo_def)
        self.source_geo_def = as_geometry(source_geo_def)
        self.target_geo_def = as_geometry(target_geo_def)
        self._orig_target = target_geo_def
        self._index_caches = {}
        self._cur = None
        self._fused_uses = {}
        self._last_request = None   # (radius, row_window, method) of the last fused call
        self.last_d2h_bytes = 0

    @property
    def _current(self):
        """The neighbour info of the last request.  A fused ``resample`` leaves none behind; it is computed when
        somebody looks (tests, ``neighbour_info()``, ``save_neighbour_info``)."""
        if self._cur is None and self._last_request is not None:
            radius, rw, method = self._last_request
            self.precompute(radius_of_influence=radius, row_window=rw, method=method)
        return self._cur

    @_current.setter
    def _current(self, value):
        self._cur = value

    # ------------------------------------------------------------------ precompute
    def _default_radius(self):
        res = [geocentric_resolution(g) for g in (self.source_geo_def, self.target_geo_def)]
        res = [r for r in res if r is not None]
        return max(res) if res else 10000.0

    def _mask_ptr(self, torch, mask):
        if mask is None:
            return None, None
        m = payload(mask)
        m = m if isinstance(m, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(m))
        if tuple(m.shape) != self.source_geo_def.shape:
            raise ValueError(f"mask shape {tuple(m.shape)} != source shape {self.source_geo_def.shape}")
        m = m.to("cuda").to(torch.uint8).contiguous()
        return m, m.data_ptr()

    def _build_grid(self, torch, radius, method, mask_ptr=None, want_count=True, light=False):
        """``(grid, valid_input, n_valid or None)``: the search structure over the valid source points.
        ``light``: float32 points only (``B200_NN_LIGHT``) -- for the fused search+gather, which never asks for
        compressed indices or the count."""
        lib = _lib.load()
        src = self.source_geo_def
        valid_input = torch.empty(src.shape, dtype=torch.uint8, device="cuda")
        n_valid = C.c_int64(0)
        handle = C.c_void_p(None)
        gs = src.geo_struct()
        _lib.check(lib.b200_nn_grid_create_ex(C.byref(gs), mask_ptr, radius, valid_input.data_ptr(),
                                              C.byref(n_valid) if want_count else None, C.byref(handle),
                                              SEARCH_METHODS[method] | (NN_LIGHT if light and not want_count else 0),
                                              _lib.stream_ptr(torch)))
        return _Grid(handle), valid_input, (int(n_valid.value) if want_count else None)

    def precompute(self, mask=None, radius_of_influence=None, epsilon=0, cache_dir=None, row_window=None,
                   method="auto", gather_only=False, **kwargs):
        """Build the search structure over the valid source points and search every target pixel.

        ``row_window=(row0, nrows)`` restricts the search to a band of target rows (row-tiled
        multi-GPU use); the default is the whole target.  ``epsilon``: the search is always exact; a non-zero value is
        accepted (an exact answer satisfies every epsilon) and logged.  The neighbour info is cached in device
        memory on this instance; with ``cache_dir`` it is also written to / read from ``nn_lut-<hash>.zarr`` in the
        reference's layout (kdtree.py:125-177).  ``method``: ``"auto"`` (fixed-grid search for geostationary area
        sources, bucket grid otherwise), ``"bucket"``, ``"fixed_grid"`` or ``"fixed_grid_strict"``; the engines are
        tested to return identical indices (``fixed_grid`` settles radii of up to ~2 source pixels with fixed 2x2 / 4x4
        stencils under a rigorous distance bound and larger ones with a loop that rules lattice points out by a local
        model of the grid with measured margins; ``fixed_grid_strict`` examines every point of the rigorous search
        disc).
        ``gather_only``: keep just the raw gather index (what ``compute`` needs), not the reference's three arrays.
        """
        del kwargs
        torch = _lib.require_cuda()
        lib = _lib.load()
        if radius_of_influence is None:
            radius_of_influence = self._default_radius()
        radius = float(radius_of_influence)
        if epsilon:
            LOG.debug("epsilon=%s ignored: the device search is exact", epsilon)
        if mask is not None and cache_dir:
            LOG.warning("Mask and cache_dir both provided to nearest resampler. Cached parameters are affected by "
                        "masked pixels. Will not cache results.")   # kdtree.py:69-73
            cache_dir = None
        src, dst = self.source_geo_def, self.target_geo_def
        row0, nrows = (0, dst.height) if row_window is None else (int(row_window[0]), int(row_window[1]))
        key = (radius, row0, nrows, method)
        if mask is None and key in self._index_caches:
            cached = self._index_caches[key]
            if gather_only or "index_array" in cached:
                self._cur = cached
                LOG.debug("Reusing device-resident neighbour info")
                return key
        if cache_dir and row_window is None and not gather_only:
            try:
                self.load_cached_neighbour_info(cache_dir, radius, epsilon)
                LOG.debug("Read pre-computed kd-tree parameters")
                self._index_caches[key] = self._cur
                return key
            except IOError:
                LOG.debug("Computing kd-tree parameters")
        LOG.debug("Computing neighbour info on the device")
        _m, mask_ptr = self._mask_ptr(torch, mask)
        stream = _lib.stream_ptr(torch)
        grid, valid_input, n_valid = self._build_grid(torch, radius, method, mask_ptr, want_count=not gather_only)
        used = C.c_int32(0)
        _lib.check(lib.b200_nn_grid_method(grid.handle, C.byref(used)))
        gather_index = torch.empty((nrows, dst.width), dtype=torch.int32, device="cuda")
        gd = dst.geo_struct(row0, nrows)
        info = {"valid_input_index": valid_input, "gather_index": gather_index, "n_valid": n_valid, "radius": radius,
                "row_window": (row0, nrows), "requested_method": method,
                "method": {1: "bucket", 2: "fixed_grid"}[used.value] + ("_strict" if method == "fixed_grid_strict" else "")}
        if gather_only:
            _lib.check(lib.b200_nn_query(grid.handle, C.byref(gd), radius, None, gather_index.data_ptr(), None, stream))
        else:
            index_array = torch.empty((nrows, dst.width), dtype=torch.int32, device="cuda")
            valid_output = torch.empty((nrows, dst.width), dtype=torch.uint8, device="cuda")
            _lib.check(lib.b200_nn_query(grid.handle, C.byref(gd), radius, index_array.data_ptr(),
                                         gather_index.data_ptr(), valid_output.data_ptr(), stream))
            info["index_array"] = index_array
            info["valid_output_index"] = valid_output
        del grid  # the index arrays are the complete persistent state (kdtree.py:179-182)
        if mask is None:
            self._index_caches[key] = info
        self._cur = info
        if cache_dir and row_window is None and not gather_only:
            self.save_neighbour_info(cache_dir, radius, epsilon)
        return key

    # --------------------------------------------------------------- cache on disk
    def _cache_filename(self, cache_dir, radius, epsilon=0):
        """``nn_lut-<hash>.zarr`` (satpy/resample/kdtree.py:116-123,141-147): inside satpy the hash is pyresample's
        own (``BaseResampler._create_cache_filename``), so that both resamplers name the file alike; stand-alone
        it is a sha1 over the two geometries and the search keywords."""
        create = getattr(super(), "_create_cache_filename", None)
        if create is not None:
            try:
                return create(cache_dir, prefix="nn_lut-", mask=None, radius_of_influence=radius, epsilon=epsilon)
            except Exception:  # noqa: BLE001 -- geometry objects pyresample cannot hash: fall through
                pass

        def describe(geo):
            if isinstance(geo, AreaDefinition):
                return {"proj": {k: str(v) for k, v in sorted(geo.proj_dict.items())}, "shape": geo.shape,
                        "extent": geo.area_extent}
            lo = payload_to_numpy(geo.lons)
            la = payload_to_numpy(geo.lats)
            return {"swath": hashlib.sha1(np.ascontiguousarray(lo).tobytes() + np.ascontiguousarray(la).tobytes()).hexdigest()}
        key = json.dumps({"src": describe(self.source_geo_def), "dst": describe(self.target_geo_def),
                          "radius_of_influence": radius, "epsilon": epsilon, "neighbours": 1}, sort_keys=True)
        return os.path.join(cache_dir, "nn_lut-" + hashlib.sha1(key.encode()).hexdigest() + ".zarr")

    def save_neighbour_info(self, cache_dir, radius_of_influence, epsilon=0):
        """Persist the three index arrays (``save_neighbour_info``, satpy/resample/kdtree.py:158-177) as a zarr-v2
        store with the reference's names, dims (``NN_COORDINATES``) and dtypes -- ``index_array`` is widened to the
        reference's int64, misses become ``n_valid`` as pykdtree reports them."""
        if not cache_dir:
            return None
        os.makedirs(cache_dir, exist_ok=True)
        filename = self._cache_filename(cache_dir, float(radius_of_influence), epsilon)
        info = self.neighbour_info()
        idx = info["index_array"].astype(np.int64)
        n_valid = int(info["valid_input_index"].sum())
        idx[idx < 0] = n_valid
        _zarr2.write_store(filename, {"valid_input_index": (NN_COORDINATES["valid_input_index"], info["valid_input_index"]),
                                      "valid_output_index": (NN_COORDINATES["valid_output_index"], info["valid_output_index"]),
                                      "index_array": (NN_COORDINATES["index_array"], idx)})
        LOG.info("Saving kd_tree neighbour info to %s", filename)
        return filename

    def load_cached_neighbour_info(self, cache_dir, radius_of_influence, epsilon=0):
        """Re-install index arrays saved earlier (``load_neighbour_info``, kdtree.py:125-156); ``IOError`` when the
        cache has no entry for these geometries and keywords, like the reference."""
        if not cache_dir:
            raise IOError
        filename = self._cache_filename(cache_dir, float(radius_of_influence), epsilon)
        if not os.path.isdir(filename):
            raise IOError(f"no cached neighbour info {filename}")
        arrays = {name: _zarr2.read_array(filename, name) for name in NN_COORDINATES}
        self.load_neighbour_info(arrays["valid_input_index"].astype(bool), arrays["index_array"],
                                 arrays["valid_output_index"].astype(bool))
        self._index_caches[(float(radius_of_influence), 0, self.target_geo_def.height, "auto")] = self._cur
        return filename

    def neighbour_info(self, as_numpy=True):
        """The reference's three cached arrays (names/dims as ``NN_COORDINATES``)."""
        cur = self._current
        if cur is None:
            raise RuntimeError("precompute() has not been called")
        if "index_array" not in cur:   # searched for gathering only: search again, keeping everything
            self._index_caches = {k: v for k, v in self._index_caches.items() if v is not cur}
            self.precompute(radius_of_influence=cur["radius"], row_window=cur["row_window"],
                            method=cur["requested_method"])
            cur = self._cur
        out = {"valid_input_index": cur["valid_input_index"].bool(),
               "valid_output_index": cur["valid_output_index"].bool(),
               "index_array": cur["index_array"].unsqueeze(-1)}
        if as_numpy:
            out = {k: v.cpu().numpy() for k, v in out.items()}
        return out

    def load_neighbour_info(self, valid_input_index, index_array, valid_output_index=None):
        """Install cached index arrays (compressed indices, reference layout: int64, misses = n_valid or -1) without
        searching."""
        torch = _lib.require_cuda()
        src, dst = self.source_geo_def, self.target_geo_def
        vi = torch.as_tensor(np.ascontiguousarray(valid_input_index)).to("cuda").to(torch.uint8).contiguous()
        ia = torch.as_tensor(np.ascontiguousarray(index_array)).to("cuda")
        if ia.dim() == 3:
            ia = ia[..., 0]
        n_valid = int(vi.sum().item())
        ia = torch.where((ia < 0) | (ia >= n_valid), torch.full_like(ia, -1), ia).to(torch.int32).contiguous()
        if tuple(vi.shape) != src.shape or tuple(ia.shape) != dst.shape:
            raise ValueError("cached index arrays do not match the geometries")
        gi = torch.empty_like(ia)
        _lib.check(_lib.load().b200_nn_compressed_to_raw(vi.data_ptr(), vi.numel(), ia.data_ptr(), gi.data_ptr(),
                                                         ia.numel(), _lib.stream_ptr(torch)))
        vo = (torch.ones(dst.shape, dtype=torch.uint8, device="cuda") if valid_output_index is None else
              torch.as_tensor(np.ascontiguousarray(valid_output_index)).to("cuda").to(torch.uint8))
        self._cur = {"valid_input_index": vi, "valid_output_index": vo, "index_array": ia, "gather_index": gi,
                         "n_valid": n_valid, "radius": None, "row_window": (0, dst.height), "method": "loaded"}

    # --------------------------------------------------------------------- compute
    @staticmethod
    def _as_device(torch, arr):
        """``(cuda tensor, was_numpy)``; bool -> uint8 (the kernels move bytes)."""
        was_numpy = not isinstance(arr, torch.Tensor)
        t = _xfer.to_device(arr) if (was_numpy or not arr.is_cuda) else arr
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
        return t.contiguous(), was_numpy

    def _finish(self, data, out_dev, host_arr, dims=None):
        res = host_arr if host_arr is not None else out_dev
        if not is_dataarray(data):
            return res
        attrs = dict(data.attrs)
        attrs["area"] = self._orig_target
        return rewrap(data, res, dims=tuple(data.dims) if dims is None else dims, attrs=attrs,
                      new_area=self.target_geo_def)

    def _host_plan(self, torch, out, row0, nrows, radius, fill_value):
        """Page-locked result + the plan for filling it: row bands, the column span of each band that the source can
        reach (satpy_b200.parallel.block_column_span) -- only that crosses PCIe; the rest of each band is the fill
        value, written by host threads while the device works."""
        from .. import parallel
        host_t, host_arr = _xfer.host_empty(out.shape, out.dtype)
        width = out.shape[-1]
        spans_wanted = radius is not None and out.dtype == torch.float32
        try:
            key = (hash(self.source_geo_def), hash(self.target_geo_def), row0, nrows, radius if spans_wanted else None,
                   self.ROW_CHUNKS, self.SPAN_ROWS)
            bands = _PLAN_CACHE.get(key)
        except TypeError:
            key, bands = None, None
        if bands is None:
            bands = []
            for r0, n in self._row_bands(nrows, True):
                # the span is taken per SUB-band of SPAN_ROWS rows: a band of config 5 is 7.5 degrees of latitude, over
                # which the disk's outline moves by up to 25 degrees of longitude
                subs = []
                for s0 in range(0, n, self.SPAN_ROWS):
                    sn = min(self.SPAN_ROWS, n - s0)
                    span = (0, width)
                    if spans_wanted:
                        span = parallel.block_column_span(self.source_geo_def, self.target_geo_def, row0 + r0 + s0, sn, radius)
                    subs.append((r0 + s0, sn, span))
                bands.append((r0, n, subs))
            if key is not None:
                if len(_PLAN_CACHE) > 64:
                    _PLAN_CACHE.clear()
                _PLAN_CACHE[key] = bands
        h3 = host_t.reshape(-1, nrows, width)
        total = nrows * width
        spanned = sum(sn * max(0, sp[1] - sp[0]) for _, _, subs in bands for _, sn, sp in subs)
        if not _xfer.spans_pay_off(total, spanned):
            bands = [(r0, n, [(r0, n, (0, width))]) for r0, n, _ in bands]
            spanned = total
        fut = _xfer.fill_async(h3, [sub for _, _, subs in bands for sub in subs if sub[2] != (0, width)], fill_value)
        # bytes that actually cross PCIe for this result (bench.py reports them)
        self.last_d2h_bytes = int(spanned * h3.shape[0] * out.element_size())
        return host_t, host_arr, h3, bands, fut

    @staticmethod
    def _band_to_host(torch, cs, h3, out3, r0, n, subs):
        """Queue the device->host copies of one finished row band on the copy stream: per sub-band of rows, the columns
        the source can reach (``subs`` from ``_host_plan``: ``[(row0, nrows, (c_lo, c_hi)), ...]``)."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        cs.wait_event(ev)
        nb, rows, width = out3.shape
        esize = out3.element_size()
        lib = _lib.load()
        hbase, dbase = h3.data_ptr(), out3.data_ptr()
        for s0, sn, (c_lo, c_hi) in subs:
            if c_hi <= c_lo:
                continue
            for b in range(nb):
                off = ((b * rows + s0) * width + c_lo) * esize
                _lib.check(lib.b200_copy_2d(hbase + off, width * esize, dbase + off, width * esize, (c_hi - c_lo) * esize, sn,
                                            int(cs.next().cuda_stream)))

    def _row_bands(self, nrows, to_host):
        width = self.target_geo_def.width
        # a band is worth ~256 MB of result (each costs three launches and a span computation on the host): 24 for the
        # 12.85 GB of config 5, 2 for an eighth of it.  Bands start on multiples of four rows, i.e. on 16-byte
        # boundaries of the index / output arrays whatever the width (the first release only cut widths divisible by
        # four into bands: config 5's 80150 columns went out as ONE band, copied after the last kernel, under the
        # bounding box of the whole disk)
        n = 1
        if to_host:
            n = max(1, min(self.ROW_CHUNKS, (nrows * width * 4) >> 28))
        rows = -(-nrows // max(1, n))
        rows = max(4, -(-rows // 4) * 4)
        return [(r0, min(rows, nrows - r0)) for r0 in range(0, nrows, rows)]

    def compute(self, data, weight_funcs=None, fill_value=np.nan, with_uncert=False, to_host=None, out=None, **kwargs):
        """``get_sample_from_neighbour_info``: gather ``data`` through the cached indices.

        ``to_host``: return a numpy array (page-locked; default: when ``data`` is numpy) -- the rows are gathered in
        ``ROW_CHUNKS`` bands and each band is copied out while the next one is gathered.  ``out``: CUDA tensor to
        write the result into."""
        del kwargs, weight_funcs, with_uncert
        torch = _lib.require_cuda()
        if self._cur is None and self._last_request is not None:   # the last request ran fused: search now, keep the index
            radius, rw, method = self._last_request
            self.precompute(radius_of_influence=radius, row_window=rw, method=method, gather_only=True)
        if self._cur is None:
            raise RuntimeError("precompute() has not been called")
        t, was_numpy = self._as_device(torch, payload(data))
        to_host = was_numpy if to_host is None else bool(to_host)
        src = self.source_geo_def
        if tuple(t.shape[-2:]) != src.shape:
            raise ValueError(f"data shape {tuple(t.shape)} does not end with the source shape {src.shape}")
        lead = tuple(t.shape[:-2])
        nbands = int(np.prod(lead)) if lead else 1
        gi = self._cur["gather_index"]
        nrows, width = gi.shape
        if out is None:
            out = torch.empty(lead + (nrows, width), dtype=t.dtype, device="cuda")
        elif tuple(out.shape) != lead + (nrows, width) or out.dtype != t.dtype or not out.is_contiguous():
            raise ValueError("out must be a contiguous CUDA tensor of the result's shape and dtype")
        elem = t.element_size()
        np_dtype = np.dtype(str(t.dtype).replace("torch.", ""))
        if fill_value is None:
            fill_value = np.nan if np_dtype.kind == "f" else 0
        with np.errstate(invalid="ignore"):
            fill = np.asarray(fill_value).astype(np_dtype)
        fill_buf = (C.c_char * elem).from_buffer_copy(fill.tobytes())
        lib = _lib.load()
        host_arr = None
        bands = [(r0, n, None) for r0, n in self._row_bands(nrows, False)]
        if to_host:
            cur = self._cur
            host_t, host_arr, h3, bands, fut = self._host_plan(torch, out, cur["row_window"][0], nrows,
                                                               cur.get("radius") if np_dtype == np.float32 else None,
                                                               float(fill) if np_dtype.kind == "f" else 0)
            cs = _xfer.CopyStreams()
        out3 = out.reshape(nbands, nrows, width)
        for r0, n, span in bands:
            _lib.check(lib.b200_nn_gather(t.data_ptr(), out3[0, r0].data_ptr(), gi[r0].data_ptr(), n * width, nbands,
                                          src.size, nrows * width, elem, C.cast(fill_buf, C.c_void_p),
                                          _lib.stream_ptr(torch)))
            if to_host:
                self._band_to_host(torch, cs, h3, out3, r0, n, span)
        if to_host:
            cs.finish(out)
            fut.result()
        return self._finish(data, out, host_arr)

    # -------------------------------------------------------------------- resample
    def resample(self, data, cache_dir=None, mask_area=None, to_host=None, out=None, fused=None, **kwargs):
        """``BaseResampler.resample``: mask (swaths by default) -> precompute -> compute, with the fused single-pass
        kernel for the first float32 dataset (module docstring).  ``fused=True/False`` forces / forbids it."""
        torch = _lib.require_cuda()
        if mask_area is None and isinstance(self.source_geo_def, SwathDefinition):
            mask_area = True
        arr = payload(data)
        if mask_area:
            t = arr if isinstance(arr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(arr))
            t = t.to("cuda")
            if t.dtype.is_floating_point:
                m = torch.isnan(t)
            else:
                attrs = data.attrs if is_dataarray(data) else {}
                fv = attrs.get("_FillValue", torch.iinfo(t.dtype).max)
                m = t == int(fv)
            while m.dim() > 2:
                m = m.all(dim=0)
            kwargs["mask"] = m
            arr = t
        pre_kw = {k: kwargs[k] for k in ("mask", "radius_of_influence", "epsilon", "row_window", "method")
                  if k in kwargs}
        comp_kw = {k: v for k, v in kwargs.items() if k not in pre_kw}
        dtype_name = str(getattr(arr, "dtype", ""))
        is_f32 = dtype_name in ("float32", "torch.float32")
        radius = pre_kw.get("radius_of_influence")
        radius = float(self._default_radius() if radius is None else radius)
        dst = self.target_geo_def
        rw = pre_kw.get("row_window")
        row0, nrows = (0, dst.height) if rw is None else (int(rw[0]), int(rw[1]))
        key = (radius, row0, nrows, pre_kw.get("method", "auto"))
        can_fuse = is_f32 and "mask" not in pre_kw and not cache_dir and key not in self._index_caches
        if fused is None:
            fused = can_fuse and self._fused_uses.get(key, 0) == 0
        elif fused and not can_fuse:
            raise ValueError("fused resampling needs float32 data, no mask, no cache_dir and no cached neighbour info")
        if fused:
            self._fused_uses[key] = self._fused_uses.get(key, 0) + 1
            self._cur, self._last_request = None, (radius, (row0, nrows), pre_kw.get("method", "auto"))
            return self._resample_fused(data, arr, radius, (row0, nrows), pre_kw.get("method", "auto"),
                                        comp_kw.get("fill_value", np.nan), to_host, out)
        full = bool(cache_dir) or "mask" in pre_kw
        self.precompute(cache_dir=cache_dir, gather_only=not full, **pre_kw)
        return self.compute(data, to_host=to_host, out=out, **comp_kw)

    def _resample_fused(self, data, arr, radius, row_window, method, fill_value, to_host, out):
        """float32 only: search and gather in one kernel, no index array materialised (SURVEY.md section 8d,
        "fused query+gather"); host results are produced in row bands, each copied out while the next is searched."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        src, dst = self.source_geo_def, self.target_geo_def
        import os as _os, time as _time
        _tr = _os.environ.get("B200_XFER_TRACE") and _os.environ.get("RANK", "0") == "0"
        _t0 = _time.perf_counter()
        t, was_numpy = self._as_device(torch, arr)
        if t.dtype != torch.float32:
            raise TypeError("fused resampling works on float32 data")
        if tuple(t.shape[-2:]) != src.shape:
            raise ValueError(f"data shape {tuple(t.shape)} does not end with the source shape {src.shape}")
        to_host = was_numpy if to_host is None else bool(to_host)
        lead = tuple(t.shape[:-2])
        nbands = int(np.prod(lead)) if lead else 1
        row0, nrows = row_window
        width = dst.width
        if fill_value is None:
            fill_value = np.nan
        if out is None:
            out = torch.empty(lead + (nrows, width), dtype=torch.float32, device="cuda")
        elif tuple(out.shape) != lead + (nrows, width) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 CUDA tensor of the result's shape")
        if nrows == 0:
            return self._finish(data, out, _xfer.host_empty(out.shape, out.dtype)[1] if to_host else None)
        tr = self.trace
        if tr is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        grid, _valid_input, _ = self._build_grid(torch, radius, method, None, want_count=False, light=True)
        if tr is not None:
            ev[1].record()
        host_arr = None
        bands = [(r0, n, None) for r0, n in self._row_bands(nrows, False)]
        _t1 = _time.perf_counter()
        if to_host:
            host_t, host_arr, h3, bands, fut = self._host_plan(torch, out, row0, nrows, radius, float(fill_value))
            cs = _xfer.CopyStreams()
        _t2 = _time.perf_counter()
        out3 = out.reshape(nbands, nrows, width)
        stream = _lib.stream_ptr(torch)
        for r0, n, span in bands:
            gd = dst.geo_struct(row0 + r0, n)
            _lib.check(lib.b200_nn_query_gather_f32(grid.handle, C.byref(gd), radius, t.data_ptr(),
                                                    out3[0, r0].data_ptr(), nbands, src.size, nrows * width,
                                                    float(fill_value), stream))
            if to_host:
                self._band_to_host(torch, cs, h3, out3, r0, n, span)
        if tr is not None:
            ev[2].record()
            tr.append(("build", ev[0], ev[1]))
            tr.append(("query_gather", ev[1], ev[2]))
        del grid   # stream-ordered: freed after the queries queued above
        _t3 = _time.perf_counter()
        if to_host:
            cs.finish(out)
            _t4 = _time.perf_counter()
            fut.result()
            if _tr:
                import sys as _sys
                print(f"[xfer] rows {nrows}: grid {1e3 * (_t1 - _t0):.1f} ms, host plan {1e3 * (_t2 - _t1):.1f} ms, launches "
                      f"{1e3 * (_t3 - _t2):.1f} ms, wait copies {1e3 * (_t4 - _t3):.1f} ms, wait fill "
                      f"{1e3 * (_time.perf_counter() - _t4):.1f} ms, bands {len(bands)}", file=_sys.stderr)
        return self._finish(data, out, host_arr)

    def resample_fused(self, data, radius_of_influence, fill_value=np.nan, row_window=None, out=None,
                       method="auto"):
        """Always the fused kernel (kept for callers of the first release; ``resample`` picks it by itself)."""
        dst = self.target_geo_def
        rw = (0, dst.height) if row_window is None else (int(row_window[0]), int(row_window[1]))
        res = self._resample_fused(None, payload(data), float(radius_of_influence), rw, method, fill_value, None, out)
        return res
