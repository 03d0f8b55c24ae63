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
"""
from __future__ import annotations

import ctypes as C
import hashlib
import json
import logging
import os

import numpy as np

from .. import _lib
from ..dataset import is_dataarray, payload, rewrap
from ..geometry import AreaDefinition, SwathDefinition, as_geometry

LOG = logging.getLogger(__name__)

SEARCH_METHODS = {"auto": 0, "bucket": 1, "fixed_grid": 2, "fixed_grid_strict": 3}

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


class B200NearestResampler:
    """Nearest-neighbour resampling on one B200."""

    def __init__(self, source_geo_def, target_geo_def):
        self.source_geo_def = as_geometry(source_geo_def)
        self.target_geo_def = as_geometry(target_geo_def)
        self._orig_target = target_geo_def
        self._index_caches = {}
        self._current = None

    # ------------------------------------------------------------------ precompute
    def _default_radius(self):
        res = [geocentric_resolution(g) for g in (self.source_geo_def, self.target_geo_def)]
        res = [r for r in res if r is not None]
        return max(res) if res else 10000.0

    def precompute(self, mask=None, radius_of_influence=None, epsilon=0, cache_dir=None, row_window=None,
                   method="auto", **kwargs):
        """Build the bucket grid over the valid source points and search every target pixel.

        ``row_window=(row0, nrows)`` restricts the search to a band of target rows (row-tiled
        multi-GPU use); the default is the whole target.  ``epsilon`` is accepted for signature
        compatibility: the search is always exact (epsilon = 0).  The neighbour info is cached in device
        memory on this instance; with ``cache_dir`` it is also written to / read from ``nn_lut-<hash>.npz``
        (the reference's zarr cache, kdtree.py:125-177, with the same array names).  ``method``: ``"auto"`` (fixed-grid
        search for geostationary area sources, bucket grid otherwise), ``"bucket"`` or ``"fixed_grid"``;
        both engines return identical indices.
        """
        del kwargs
        torch = _lib.require_cuda()
        lib = _lib.load()
        if radius_of_influence is None:
            radius_of_influence = self._default_radius()
        radius = float(radius_of_influence)
        if mask is not None and cache_dir:
            LOG.warning("Mask and cache_dir both provided to nearest resampler. Cached parameters are affected by "
                        "masked pixels. Will not cache results.")   # kdtree.py:69-73
            cache_dir = None
        src, dst = self.source_geo_def, self.target_geo_def
        row0, nrows = (0, dst.height) if row_window is None else (int(row_window[0]), int(row_window[1]))
        key = (radius, row0, nrows, method)
        if mask is None and key in self._index_caches:
            self._current = self._index_caches[key]
            LOG.debug("Reusing device-resident neighbour info")
            return key
        if cache_dir and row_window is None:
            try:
                self.load_cached_neighbour_info(cache_dir, radius, epsilon)
                LOG.debug("Read pre-computed kd-tree parameters")
                self._index_caches[key] = self._current
                return key
            except IOError:
                LOG.debug("Computing kd-tree parameters")
        LOG.debug("Computing bucket-grid neighbour info")
        mask_ptr = None
        if mask is not None:
            m = payload(mask)
            m = m if isinstance(m, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(m))
            if tuple(m.shape) != src.shape:
                raise ValueError(f"mask shape {tuple(m.shape)} != source shape {src.shape}")
            m = m.to("cuda").to(torch.uint8).contiguous()
            mask_ptr = m.data_ptr()
        stream = _lib.stream_ptr(torch)
        valid_input = torch.empty(src.shape, dtype=torch.uint8, device="cuda")
        n_valid = C.c_int64(0)
        handle = C.c_void_p(None)
        gs = src.geo_struct()
        _lib.check(lib.b200_nn_grid_create_ex(C.byref(gs), mask_ptr, radius, valid_input.data_ptr(),
                                              C.byref(n_valid), C.byref(handle), SEARCH_METHODS[method], stream))
        grid = _Grid(handle)
        used = C.c_int32(0)
        _lib.check(lib.b200_nn_grid_method(grid.handle, C.byref(used)))
        gather_index = torch.empty((nrows, dst.width), dtype=torch.int32, device="cuda")
        index_array = torch.empty((nrows, dst.width), dtype=torch.int32, device="cuda")
        valid_output = torch.empty((nrows, dst.width), dtype=torch.uint8, device="cuda")
        gd = dst.geo_struct(row0, nrows)
        _lib.check(lib.b200_nn_query(grid.handle, C.byref(gd), radius, index_array.data_ptr(),
                                     gather_index.data_ptr(), valid_output.data_ptr(), stream))
        info = {"valid_input_index": valid_input, "valid_output_index": valid_output, "index_array": index_array,
                "gather_index": gather_index, "n_valid": int(n_valid.value), "radius": radius,
                "row_window": (row0, nrows), "method": {1: "bucket", 2: "fixed_grid"}[used.value] + ("_strict" if method == "fixed_grid_strict" else "")}
        del grid  # the index arrays are the complete persistent state (kdtree.py:179-182)
        if mask is None:
            self._index_caches[key] = info
        self._current = info
        if cache_dir and row_window is None:
            self.save_neighbour_info(cache_dir, radius, epsilon)
        return key

    # --------------------------------------------------------------- cache on disk
    def _cache_filename(self, cache_dir, radius, epsilon=0):
        """``nn_lut-<hash>.npz``: the reference hashes the two geometries and the search keywords into
        ``nn_lut-<hash>.zarr`` (satpy/resample/kdtree.py:116-123); zarr is not available here, the array names and
        dims are the reference's (``NN_COORDINATES``)."""
        def describe(geo):
            if isinstance(geo, AreaDefinition):
                return {"proj": {k: str(v) for k, v in sorted(geo.proj_dict.items())}, "shape": geo.shape,
                        "extent": geo.area_extent}
            lo = payload_to_numpy(geo.lons)
            la = payload_to_numpy(geo.lats)
            return {"swath": hashlib.sha1(np.ascontiguousarray(lo).tobytes() + np.ascontiguousarray(la).tobytes()).hexdigest()}
        key = json.dumps({"src": describe(self.source_geo_def), "dst": describe(self.target_geo_def),
                          "radius_of_influence": radius, "epsilon": epsilon, "neighbours": 1}, sort_keys=True)
        return os.path.join(cache_dir, "nn_lut-" + hashlib.sha1(key.encode()).hexdigest() + ".npz")

    def save_neighbour_info(self, cache_dir, radius_of_influence, epsilon=0):
        """Persist the three index arrays (``save_neighbour_info``, satpy/resample/kdtree.py:158-177)."""
        if not cache_dir:
            return None
        os.makedirs(cache_dir, exist_ok=True)
        filename = self._cache_filename(cache_dir, float(radius_of_influence), epsilon)
        info = self.neighbour_info()
        np.savez_compressed(filename, **info)
        LOG.info("Saving kd_tree neighbour info to %s", filename)
        return filename

    def load_cached_neighbour_info(self, cache_dir, radius_of_influence, epsilon=0):
        """Re-install index arrays saved earlier (``load_neighbour_info``, kdtree.py:125-156); ``IOError`` when the
        cache has no entry for these geometries and keywords, like the reference."""
        if not cache_dir:
            raise IOError
        filename = self._cache_filename(cache_dir, float(radius_of_influence), epsilon)
        if not os.path.exists(filename):
            raise IOError(f"no cached neighbour info {filename}")
        with np.load(filename) as fid:
            self.load_neighbour_info(fid["valid_input_index"], fid["index_array"], fid["valid_output_index"])
        self._index_caches[(float(radius_of_influence), 0, self.target_geo_def.height, "auto")] = self._current
        return filename

    def neighbour_info(self, as_numpy=True):
        """The reference's three cached arrays (names/dims as ``NN_COORDINATES``)."""
        if self._current is None:
            raise RuntimeError("precompute() has not been called")
        cur = self._current
        out = {"valid_input_index": cur["valid_input_index"].bool(),
               "valid_output_index": cur["valid_output_index"].bool(),
               "index_array": cur["index_array"].unsqueeze(-1)}
        if as_numpy:
            out = {k: v.cpu().numpy() for k, v in out.items()}
        return out

    def load_neighbour_info(self, valid_input_index, index_array, valid_output_index=None):
        """Install cached index arrays (compressed indices, reference layout) without searching."""
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
        self._current = {"valid_input_index": vi, "valid_output_index": vo, "index_array": ia, "gather_index": gi,
                         "n_valid": n_valid, "radius": None, "row_window": (0, dst.height)}

    # --------------------------------------------------------------------- compute
    def compute(self, data, weight_funcs=None, fill_value=np.nan, with_uncert=False, **kwargs):
        """``get_sample_from_neighbour_info``: gather ``data`` through the cached indices."""
        del kwargs, weight_funcs, with_uncert
        torch = _lib.require_cuda()
        if self._current is None:
            raise RuntimeError("precompute() has not been called")
        arr = payload(data)
        was_numpy = not isinstance(arr, torch.Tensor)
        t = torch.from_numpy(np.ascontiguousarray(arr)) if was_numpy else arr
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
        t = t.to("cuda", non_blocking=True).contiguous()
        src = self.source_geo_def
        if tuple(t.shape[-2:]) != src.shape:
            raise ValueError(f"data shape {tuple(t.shape)} does not end with the source shape {src.shape}")
        lead = tuple(t.shape[:-2])
        nbands = int(np.prod(lead)) if lead else 1
        gi = self._current["gather_index"]
        n_out = gi.numel()
        out = torch.empty(lead + tuple(gi.shape), dtype=t.dtype, device="cuda")
        elem = t.element_size()
        np_dtype = np.dtype(str(t.dtype).replace("torch.", ""))
        if fill_value is None:
            fill_value = np.nan if np_dtype.kind == "f" else 0
        with np.errstate(invalid="ignore"):
            fill = np.asarray(fill_value).astype(np_dtype)
        fill_buf = (C.c_char * elem).from_buffer_copy(fill.tobytes())
        _lib.check(_lib.load().b200_nn_gather(t.data_ptr(), out.data_ptr(), gi.data_ptr(), n_out, nbands,
                                              src.size, n_out, elem, C.cast(fill_buf, C.c_void_p),
                                              _lib.stream_ptr(torch)))
        res = out.cpu().numpy() if was_numpy else out
        if not is_dataarray(data):
            return res
        dims = tuple(data.dims)
        attrs = dict(data.attrs)
        attrs["area"] = self._orig_target
        return rewrap(data, res, dims=dims, attrs=attrs)

    # -------------------------------------------------------------------- resample
    def resample(self, data, cache_dir=None, mask_area=None, **kwargs):
        """``BaseResampler.resample``: mask (swaths by default) -> precompute -> compute."""
        torch = _lib.require_cuda()
        if mask_area is None and isinstance(self.source_geo_def, SwathDefinition):
            mask_area = True
        if mask_area:
            arr = payload(data)
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
        pre_kw = {k: kwargs[k] for k in ("mask", "radius_of_influence", "epsilon", "row_window", "method")
                  if k in kwargs}
        self.precompute(cache_dir=cache_dir, **pre_kw)
        comp_kw = {k: v for k, v in kwargs.items() if k not in pre_kw}
        return self.compute(data, **comp_kw)

    def resample_fused(self, data, radius_of_influence, fill_value=np.nan, row_window=None, out=None,
                       method="auto"):
        """float32 only: search and gather in one kernel, no index array materialised.

        For one-off resampling of a single dataset ("fused query+gather").
        """
        torch = _lib.require_cuda()
        lib = _lib.load()
        src, dst = self.source_geo_def, self.target_geo_def
        arr = payload(data)
        was_numpy = not isinstance(arr, torch.Tensor)
        t = torch.from_numpy(np.ascontiguousarray(arr)) if was_numpy else arr
        if t.dtype != torch.float32:
            raise TypeError("resample_fused works on float32 data")
        t = t.to("cuda", non_blocking=True).contiguous()
        lead = tuple(t.shape[:-2])
        nbands = int(np.prod(lead)) if lead else 1
        row0, nrows = (0, dst.height) if row_window is None else (int(row_window[0]), int(row_window[1]))
        stream = _lib.stream_ptr(torch)
        valid_input = torch.empty(src.shape, dtype=torch.uint8, device="cuda")
        n_valid = C.c_int64(0)
        handle = C.c_void_p(None)
        gs = src.geo_struct()
        _lib.check(lib.b200_nn_grid_create_ex(C.byref(gs), None, float(radius_of_influence), valid_input.data_ptr(),
                                              C.byref(n_valid), C.byref(handle), SEARCH_METHODS[method], stream))
        grid = _Grid(handle)
        if out is None:
            out = torch.empty(lead + (nrows, dst.width), dtype=torch.float32, device="cuda")
        gd = dst.geo_struct(row0, nrows)
        _lib.check(lib.b200_nn_query_gather_f32(grid.handle, C.byref(gd), float(radius_of_influence), t.data_ptr(),
                                                out.data_ptr(), nbands, src.size, nrows * dst.width,
                                                float(fill_value), stream))
        torch.cuda.current_stream().synchronize()
        del grid
        return out.cpu().numpy() if was_numpy else out
