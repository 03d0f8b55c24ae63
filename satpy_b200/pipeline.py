"""The BASELINE.json hot path end to end on one GPU: ABI counts -> reflectance -> sun-zenith
correction -> nearest-neighbour resampling (configs 2 and 5).

This is the chain ``Scene.load([C02], modifiers=("sunz_corrected",))`` + ``Scene.resample(area,
resampler="nearest")`` evaluates lazily in the reference (SURVEY.md sections 3.1-3.3):
``NC_ABI_L1B.get_dataset`` (satpy/readers/abi_l1b.py:45-72) -> ``SunZenithCorrector.__call__``
(satpy/modifiers/geometry.py:48-71) -> ``KDTreeResampler.precompute/compute``
(satpy/resample/kdtree.py:60-95,184-190).  Here it is three kernel families on one stream:
fused calibrate+sunz (6 B/pixel), bucket-grid build, fused query+gather (or query once + gather).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .geometry import as_geometry
from .modifiers.geometry import _sunz_struct
from .resample.nearest import SEARCH_METHODS, _Grid


class ABIResamplePipeline:
    """Reusable device state for resampling a stream of ABI images between two fixed geometries."""

    def __init__(self, src_area, dst_area, radius_of_influence, row_window=None, src_row_window=None,
                 method="auto"):
        """``row_window``: band of TARGET rows this instance produces; ``src_row_window``: band of SOURCE
        rows it is fed (counts passed to ``run`` then cover only those rows) -- see satpy_b200.parallel."""
        self.torch = _lib.require_cuda()
        self.lib = _lib.load()
        self.src = as_geometry(src_area)
        self.dst = as_geometry(dst_area)
        self.radius = float(radius_of_influence)
        self.method = SEARCH_METHODS[method]
        self.row0, self.nrows = (0, self.dst.height) if row_window is None else (int(row_window[0]), int(row_window[1]))
        self.src_row0, self.src_nrows = ((0, self.src.height) if src_row_window is None else
                                         (int(src_row_window[0]), int(src_row_window[1])))
        torch = self.torch
        self.src_shape = (self.src_nrows, self.src.width)
        self.refl = torch.empty(self.src_shape, dtype=torch.float32, device="cuda")
        self.valid_input = torch.empty(self.src_shape, dtype=torch.uint8, device="cuda")
        self.grid = None
        self.gather_index = None
        self.n_valid = None
        self.launches = 0
        self.trace = None  # set to a list to collect (stage name, start event, end event) per stage

    def _stage(self, name):
        return _Stage(self, name)

    @property
    def out_shape(self):
        return (self.nrows, self.dst.width)

    # ------------------------------------------------------------------ stages
    def calibrate_sunz(self, counts_dev, cal, start_time, correction_limit=88.0, max_sza=95.0):
        """counts (int16, device) -> sun-zenith-corrected reflectance [%] in ``self.refl``; one launch."""
        torch = self.torch
        s_cal = (_lib.AbiCalStruct * 1)(cal.struct("reflectance"))
        g = self.src.geo_struct(self.src_row0, self.src_nrows)
        sz = _sunz_struct(start_time, correction_limit, max_sza, True, 0)
        cptr = (C.c_void_p * 1)(counts_dev.data_ptr())
        optr = (C.c_void_p * 1)(self.refl.data_ptr())
        _lib.check(self.lib.b200_abi_vis_sunz(cptr, optr, s_cal, 1, C.byref(g), C.byref(sz), _lib.stream_ptr(torch)))
        self.launches += 2  # geos_tables_kernel + abi_vis_sunz_kernel
        return self.refl

    def build_grid(self):
        """Bucket grid over the valid source points (4 launches + 2 cub scans)."""
        torch = self.torch
        self.grid = None
        n_valid = C.c_int64(0)
        handle = C.c_void_p(None)
        gs = self.src.geo_struct(self.src_row0, self.src_nrows)
        _lib.check(self.lib.b200_nn_grid_create_ex(C.byref(gs), None, self.radius, self.valid_input.data_ptr(),
                                                   C.byref(n_valid), C.byref(handle), self.method,
                                                   _lib.stream_ptr(torch)))
        self.grid = _Grid(handle)
        self.n_valid = int(n_valid.value)
        self.launches += 2  # geos_tables_kernel + nn_source_points_kernel (the cub scan kernels are not counted)
        return self.grid

    def query_gather(self, data, out, fill=float("nan"), row_window=None):
        """Fused search + gather of float32 ``data`` into ``out``; one launch.  ``row_window``: a sub-band of this
        instance's target rows (``out`` then holds just those rows)."""
        torch = self.torch
        r0, n = (self.row0, self.nrows) if row_window is None else (int(row_window[0]), int(row_window[1]))
        gd = self.dst.geo_struct(r0, n)
        _lib.check(self.lib.b200_nn_query_gather_f32(self.grid.handle, C.byref(gd), self.radius, data.data_ptr(),
                                                     out.data_ptr(), 1, self.refl.numel(), out.numel(), fill,
                                                     _lib.stream_ptr(torch)))
        self.launches += 3  # nn_target_tables_kernel + nn_seg_cmax_kernel + nn_fixed_query_rect_kernel
        return out

    def query(self):
        """Search once and keep the gather index on the device (the cached-neighbour-info mode)."""
        torch = self.torch
        self.gather_index = torch.empty(self.out_shape, dtype=torch.int32, device="cuda")
        gd = self.dst.geo_struct(self.row0, self.nrows)
        _lib.check(self.lib.b200_nn_query(self.grid.handle, C.byref(gd), self.radius, None,
                                          self.gather_index.data_ptr(), None, _lib.stream_ptr(torch)))
        self.launches += 1
        return self.gather_index

    def gather(self, data, out, fill=float("nan")):
        torch = self.torch
        fill_buf = (C.c_char * 4).from_buffer_copy(np.float32(fill).tobytes())
        _lib.check(self.lib.b200_nn_gather(data.data_ptr(), out.data_ptr(), self.gather_index.data_ptr(),
                                           out.numel(), 1, self.refl.numel(), out.numel(), 4,
                                           C.cast(fill_buf, C.c_void_p), _lib.stream_ptr(torch)))
        self.launches += 1
        return out

    # --------------------------------------------------------------------- step
    def run(self, counts_dev, cal, start_time, out=None, reuse_neighbours=False):
        """One pass of the hot path.  ``reuse_neighbours``: keep the index arrays of the previous pass
        (what ``cache_dir`` / ``resamplers_cache`` give the reference for fixed geostationary grids)."""
        torch = self.torch
        if out is None:
            out = torch.empty(self.out_shape, dtype=torch.float32, device="cuda")
        if self.nrows == 0:
            return out
        if self.src_nrows == 0:  # nothing of the source can reach this tile
            out.fill_(float("nan"))
            return out
        if tuple(counts_dev.shape) != self.src_shape:
            raise ValueError(f"counts shape {tuple(counts_dev.shape)} != source window shape {self.src_shape}")
        with self._stage("calibrate_sunz"):
            self.calibrate_sunz(counts_dev, cal, start_time)
        if reuse_neighbours:
            if self.gather_index is None:
                self.build_grid()
                self.query()
                self.grid = None
            with self._stage("gather"):
                self.gather(self.refl, out)
            return out
        with self._stage("build"):
            self.build_grid()
        with self._stage("query_gather"):
            self.query_gather(self.refl, out)
        return out

    def run_host(self, counts_host, cal, start_time, out_host, out_dev=None, chunks=8):
        """``run`` with HOST buffers (pinned torch tensors or numpy arrays): counts are copied to the device, the image
        is copied back into ``out_host``; returns after the last copy has completed.  The target rows are searched in
        ``chunks`` bands and each band starts its device->host copy (second stream) while the next one is searched."""
        torch = self.torch
        c = counts_host if isinstance(counts_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(counts_host))
        o = out_host if isinstance(out_host, torch.Tensor) else torch.from_numpy(out_host)
        if self.nrows == 0:
            return out_host
        if out_dev is None:
            out_dev = torch.empty(self.out_shape, dtype=torch.float32, device="cuda")
        if self.src_nrows == 0:
            out_dev.fill_(float("nan"))
            o.copy_(out_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out_host
        if getattr(self, "_counts_dev", None) is None:
            self._counts_dev = torch.empty(self.src_shape, dtype=torch.int16, device="cuda")
            self._copy_stream = torch.cuda.Stream()
        self._counts_dev.copy_(c.view(torch.int16) if c.dtype == torch.uint16 else c, non_blocking=True)
        self.calibrate_sunz(self._counts_dev, cal, start_time)
        self.build_grid()
        main = torch.cuda.current_stream()
        rows = -(-self.nrows // max(1, chunks))
        for r0 in range(0, self.nrows, rows):
            n = min(rows, self.nrows - r0)
            self.query_gather(self.refl, out_dev[r0:r0 + n], row_window=(self.row0 + r0, n))
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(ev)
                o[r0:r0 + n].copy_(out_dev[r0:r0 + n], non_blocking=True)
        self._copy_stream.synchronize()
        return out_host


class _Stage:
    """Brackets one pipeline stage with CUDA events on the current stream when tracing is on."""

    def __init__(self, pipe, name):
        self.pipe, self.name = pipe, name

    def __enter__(self):
        if self.pipe.trace is not None:
            torch = self.pipe.torch
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.pipe.trace is not None:
            self.e1.record()
            self.pipe.trace.append((self.name, self.e0, self.e1))
        return False


def abi_sunz_nearest(counts, cal, src_area, dst_area, start_time, radius_of_influence, row_window=None):
    """Convenience wrapper: host or device counts in, device float32 image out."""
    torch = _lib.require_cuda()
    t = counts if isinstance(counts, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(counts))
    if t.dtype == torch.uint16:
        t = t.view(torch.int16)
    t = t.to("cuda", non_blocking=True).contiguous()
    pipe = ABIResamplePipeline(src_area, dst_area, radius_of_influence, row_window)
    out = pipe.run(t, cal, start_time)
    torch.cuda.current_stream().synchronize()
    return out
