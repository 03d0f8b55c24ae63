"""``b200_bilinear``: bilinear resampler on the device.

Host-side mirror of ``BilinearResampler`` (satpy/resample/kdtree.py:193-293): constructed as
``cls(source_geo_def, target_geo_def)``; ``precompute(mask=None, radius_of_influence=50000, epsilon=0,
reduce_data=True, cache_dir=False)`` computes the bilinear coefficients once (``neighbours=32``, kdtree.py:231) and
keeps them on the instance; ``compute(data, fill_value=None)`` defaults the fill value to ``data.attrs["_FillValue"]``
(kdtree.py:285-286) and returns an array of the target shape (kdtree.py:287-291).

The coefficients keep pyresample's names: ``bilinear_t``, ``bilinear_s``, ``index_array`` (N, 4: upper-left,
upper-right, lower-left, lower-right corner; compressed valid-input indices), ``valid_input_index``.
Parity is UNPINNED in the reference (its tests mock the library); see oracle/bilinear.py.
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
from ..geometry import AreaDefinition, as_geometry
from .nearest import SEARCH_METHODS, _Grid

LOG = logging.getLogger(__name__)


class B200BilinearResampler(_compat.resampler_base()):
    """Bilinear interpolation between the four nearest source pixels around each target pixel centre."""

    def __init__(self, source_geo_def, target_geo_def):
        if _compat.resampler_base() is not object:
            super().__init__(source_geo_def, target_geo_def)
        self.source_geo_def = as_geometry(source_geo_def)
        self.target_geo_def = as_geometry(target_geo_def)
        self._orig_target = target_geo_def
        self._info = None

    def precompute(self, mask=None, radius_of_influence=50000, epsilon=0, reduce_data=True, cache_dir=False,
                   neighbours=32, row_window=None, **kwargs):
        """``get_bil_info``.  ``mask`` is deleted exactly as the reference does (``del mask``, kdtree.py:224);
        ``epsilon`` and ``reduce_data`` are accepted for signature compatibility (the search is exact and always reads
        only the source pixels around each target).  The coefficients are cached on this instance and, with
        ``cache_dir``, loaded from / saved to ``bil_lut-<hash>.zarr`` (``load_bil_info`` / ``save_bil_info``,
        kdtree.py:244-279)."""
        del kwargs, mask, epsilon, reduce_data
        torch = _lib.require_cuda()
        lib = _lib.load()
        src, dst = self.source_geo_def, self.target_geo_def
        if not isinstance(dst, AreaDefinition):
            raise NotImplementedError("the bilinear target must be an AreaDefinition")
        radius = float(radius_of_influence)
        row0, nrows = (0, dst.height) if row_window is None else (int(row_window[0]), int(row_window[1]))
        key = (radius, int(neighbours), row0, nrows)
        if self._info is not None and self._info["key"] == key:
            LOG.debug("Reusing device-resident bilinear parameters")
            return key
        if cache_dir and row_window is None:
            try:
                self.load_bil_info(cache_dir, radius_of_influence=radius, neighbours=int(neighbours))
                LOG.debug("Loaded bilinear parameters")
                return key
            except IOError:
                pass
        LOG.debug("Computing bilinear parameters")
        stream = _lib.stream_ptr(torch)
        valid_input = torch.empty(src.shape, dtype=torch.uint8, device="cuda")
        n_valid = C.c_int64(0)
        handle = C.c_void_p(None)
        gs = src.geo_struct()
        # the fixed-grid engine is required (NotImplementedError for non-geostationary sources)
        _lib.check(lib.b200_nn_grid_create_ex(C.byref(gs), None, radius, valid_input.data_ptr(), C.byref(n_valid),
                                              C.byref(handle), SEARCH_METHODS["fixed_grid"], stream))
        grid = _Grid(handle)
        n = nrows * dst.width
        index_array = torch.empty((n, 4), dtype=torch.int32, device="cuda")
        gather_index = torch.empty((n, 4), dtype=torch.int32, device="cuda")
        t = torch.empty(n, dtype=torch.float64, device="cuda")
        s = torch.empty(n, dtype=torch.float64, device="cuda")
        gd = dst.geo_struct(row0, nrows)
        _lib.check(lib.b200_bil_info(grid.handle, valid_input.data_ptr(), C.byref(gd), radius, int(neighbours),
                                     index_array.data_ptr(), gather_index.data_ptr(), t.data_ptr(), s.data_ptr(),
                                     stream))
        del grid
        self._info = {"key": key, "bilinear_t": t, "bilinear_s": s, "index_array": index_array,
                      "gather_index": gather_index, "valid_input_index": valid_input, "n_valid": int(n_valid.value),
                      "out_shape": (nrows, dst.width)}
        if cache_dir and row_window is None:
            LOG.debug("Saving bilinear parameters.")
            self.save_bil_info(cache_dir, radius_of_influence=radius, neighbours=int(neighbours))
        return key

    # --------------------------------------------------------------- cache on disk
    def _cache_filename(self, cache_dir, radius_of_influence, neighbours):
        """``bil_lut-<hash>.zarr`` (kdtree.py:244-279): pyresample's own naming inside satpy, a sha1 over the two
        geometries and the keywords stand-alone."""
        create = getattr(super(), "_create_cache_filename", None)
        if create is not None:
            try:
                return create(cache_dir, prefix="bil_lut-", radius_of_influence=radius_of_influence,
                              neighbours=neighbours, epsilon=0)
            except Exception:  # noqa: BLE001
                pass
        src, dst = self.source_geo_def, self.target_geo_def
        key = json.dumps({"src": [sorted((k, str(v)) for k, v in src.proj_dict.items()), src.shape, src.area_extent],
                          "dst": [sorted((k, str(v)) for k, v in dst.proj_dict.items()), dst.shape, dst.area_extent],
                          "radius_of_influence": radius_of_influence, "neighbours": neighbours}, sort_keys=True)
        return os.path.join(cache_dir, "bil_lut-" + hashlib.sha1(key.encode()).hexdigest() + ".zarr")

    def save_bil_info(self, cache_dir, radius_of_influence=50000.0, neighbours=32):
        """Persist the coefficients as a zarr-v2 store: ``bilinear_t`` / ``bilinear_s`` (float64, N), ``index_array``
        (int64, N x 4: the four corners, compressed valid-input indices) and ``valid_input_index`` (bool, source
        shape) -- the quantities pyresample's ``save_resampling_info`` stores (its own array names for the corner
        indices could not be checked: pyresample is not installable here)."""
        if not cache_dir or self._info is None:
            return None
        os.makedirs(cache_dir, exist_ok=True)
        filename = self._cache_filename(cache_dir, float(radius_of_influence), int(neighbours))
        info = self.bil_info()
        _zarr2.write_store(filename, {"bilinear_t": (("x1",), info["bilinear_t"]), "bilinear_s": (("x1",), info["bilinear_s"]),
                                      "index_array": (("x1", "n"), info["index_array"].astype(np.int64)),
                                      "valid_input_index": (("y2", "x2"), info["valid_input_index"])})
        LOG.info("Saving BIL neighbour info to %s", filename)
        return filename

    def load_bil_info(self, cache_dir, radius_of_influence=50000.0, neighbours=32):
        """Re-install coefficients saved earlier; ``IOError`` when the cache has no entry (kdtree.py:244-258)."""
        if not cache_dir:
            raise IOError
        torch = _lib.require_cuda()
        filename = self._cache_filename(cache_dir, float(radius_of_influence), int(neighbours))
        if not os.path.isdir(filename):
            raise IOError(f"no cached bilinear info {filename}")
        src, dst = self.source_geo_def, self.target_geo_def
        t = torch.from_numpy(_zarr2.read_array(filename, "bilinear_t")).to("cuda")
        s_ = torch.from_numpy(_zarr2.read_array(filename, "bilinear_s")).to("cuda")
        idx = torch.from_numpy(_zarr2.read_array(filename, "index_array")).to("cuda").to(torch.int32).contiguous()
        vin = torch.from_numpy(_zarr2.read_array(filename, "valid_input_index").astype(np.uint8)).to("cuda")
        if tuple(vin.shape) != src.shape or t.numel() != dst.size or tuple(idx.shape) != (dst.size, 4):
            raise IOError("cached bilinear info does not match the geometries")
        # compressed -> raw source indices for the gather (the kernel works on raw pixel numbers)
        gi = torch.empty_like(idx)
        _lib.check(_lib.load().b200_nn_compressed_to_raw(vin.data_ptr(), vin.numel(), idx.data_ptr(), gi.data_ptr(),
                                                         idx.numel(), _lib.stream_ptr(torch)))
        self._info = {"key": (float(radius_of_influence), int(neighbours), 0, dst.height), "bilinear_t": t,
                      "bilinear_s": s_, "index_array": idx, "gather_index": gi, "valid_input_index": vin,
                      "n_valid": int(vin.sum().item()), "out_shape": dst.shape}
        return filename

    def bil_info(self, as_numpy=True):
        if self._info is None:
            raise RuntimeError("precompute() has not been called")
        out = {k: self._info[k] for k in ("bilinear_t", "bilinear_s", "index_array")}
        out["valid_input_index"] = self._info["valid_input_index"].bool()
        return {k: v.cpu().numpy() for k, v in out.items()} if as_numpy else out

    def compute(self, data, fill_value=None, **kwargs):
        """``get_sample_from_bil_info``: 4-tap blend of float32 ``data`` ((y, x) or (bands, y, x))."""
        del kwargs
        torch = _lib.require_cuda()
        if self._info is None:
            raise RuntimeError("precompute() has not been called")
        if fill_value is None and is_dataarray(data):
            fill_value = data.attrs.get("_FillValue")
        if fill_value is None:
            fill_value = np.nan
        arr = payload(data)
        was_numpy = not isinstance(arr, torch.Tensor)
        tdata = torch.from_numpy(np.ascontiguousarray(arr)) if was_numpy else arr
        if tdata.dtype != torch.float32:
            raise TypeError(f"the device bilinear resampler works on float32 data, got {tdata.dtype}")
        tdata = tdata.to("cuda", non_blocking=True).contiguous()
        src = self.source_geo_def
        if tuple(tdata.shape[-2:]) != src.shape:
            raise ValueError(f"data shape {tuple(tdata.shape)} does not end with the source shape {src.shape}")
        lead = tuple(tdata.shape[:-2])
        nbands = int(np.prod(lead)) if lead else 1
        info = self._info
        n_out = info["bilinear_t"].numel()
        out = torch.empty(lead + info["out_shape"], dtype=torch.float32, device="cuda")
        _lib.check(_lib.load().b200_bil_sample_f32(tdata.data_ptr(), out.data_ptr(), info["gather_index"].data_ptr(),
                                                   info["bilinear_t"].data_ptr(), info["bilinear_s"].data_ptr(), n_out,
                                                   nbands, src.size, n_out, float(fill_value), _lib.stream_ptr(torch)))
        res = _xfer.to_host(out) if was_numpy else out
        if not is_dataarray(data):
            return res
        attrs = dict(data.attrs)
        attrs["area"] = self._orig_target
        return rewrap(data, res, dims=tuple(data.dims), attrs=attrs, new_area=self.target_geo_def)

    def resample(self, data, cache_dir=None, mask_area=None, **kwargs):
        del mask_area
        pre = {k: kwargs[k] for k in ("radius_of_influence", "epsilon", "reduce_data", "neighbours", "row_window")
               if k in kwargs}
        self.precompute(cache_dir=cache_dir, **pre)
        return self.compute(data, **{k: v for k, v in kwargs.items() if k not in pre})
