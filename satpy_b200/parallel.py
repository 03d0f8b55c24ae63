"""Row-tile decomposition of a resampling job over the GPUs of one box.

The reference has no distributed compute; its analogue of this module is dask chunking of the target
(``CHUNK_SIZE``, satpy/resample/base.py:32) plus ``Scene._reduce_data`` cropping the source to what a
target needs (``get_area_slices``, satpy/scene.py:930-954).  Here the target area is cut into equal
contiguous row tiles, one per rank; each rank works on the band of SOURCE rows that can reach its tile
(tile latitude range +- the search radius) and the image is assembled on every rank either by NCCL all-gathers
(``all_gather_rows``) or, preferably, by pushing finished rows straight into the peers' copies of the image over
NVLink with the copy engines (``PeerImage``).  No index array or source data is exchanged.
"""
from __future__ import annotations

import math

from . import _lib
from .geometry import as_geometry

R_EARTH = 6370997.0


def row_tiles(height: int, world_size: int):
    """Equal tiles of ``ceil(height / world_size)`` rows; returns ``(tile_rows, [(row0, nrows), ...])``.

    The last tiles may be short or empty; the all-gather buffer is padded to ``world_size * tile_rows`` rows.
    """
    tile = -(-height // world_size)
    tiles = []
    for r in range(world_size):
        r0 = min(height, r * tile)
        tiles.append((r0, min(height, r0 + tile) - r0))
    return tile, tiles


def _lat_range_of_rows(geo, row0, nrows, chunk_rows=None):
    """(min, max) latitude over a band of rows, on the device, in row chunks to bound memory."""
    torch = _lib.require_cuda()
    if chunk_rows is None:
        chunk_rows = max(1, (64 << 20) // max(1, geo.width))
    lo, hi = math.inf, -math.inf
    for r in range(row0, row0 + nrows, chunk_rows):
        n = min(chunk_rows, row0 + nrows - r)
        _, lats = geo.get_lonlats(slice(r, r + n), device=True)
        ok = torch.isfinite(lats)
        if ok.any():
            v = lats[ok]
            lo, hi = min(lo, float(v.min())), max(hi, float(v.max()))
    return lo, hi


def _row_lat_extent(geo, chunk_rows=None):
    """Per-row (min, max) latitude of a geometry (NaN rows where nothing is on the earth); cached on the object."""
    cached = getattr(geo, "_row_lat_extent", None)
    if cached is not None:
        return cached
    torch = _lib.require_cuda()
    if chunk_rows is None:
        chunk_rows = max(1, (64 << 20) // max(1, geo.width))
    mins, maxs = [], []
    for r in range(0, geo.height, chunk_rows):
        n = min(chunk_rows, geo.height - r)
        _, lats = geo.get_lonlats(slice(r, r + n), device=True)
        ok = torch.isfinite(lats) & (lats.abs() <= 90.0)
        mins.append(torch.where(ok, lats, torch.full_like(lats, math.inf)).amin(dim=1))
        maxs.append(torch.where(ok, lats, torch.full_like(lats, -math.inf)).amax(dim=1))
    out = (torch.cat(mins).cpu(), torch.cat(maxs).cpu())
    try:
        geo._row_lat_extent = out
    except AttributeError:
        pass
    return out


def source_row_window(src, dst, row0: int, nrows: int, radius: float):
    """Band of source rows ``(src_row0, src_nrows)`` that can hold a neighbour of target rows [row0, row0+nrows).

    Conservative: a source row is kept when its latitude extent intersects the tile's latitude range widened
    by the search radius (as an arc on the R = 6370997 m sphere, plus 1 %).  ``(0, 0)`` when nothing can match.
    """
    src, dst = as_geometry(src), as_geometry(dst)
    if nrows <= 0:
        return 0, 0
    lo, hi = _lat_range_of_rows(dst, row0, nrows)
    if lo > hi:
        return 0, 0
    margin = math.degrees(2.0 * math.asin(min(1.0, radius / (2.0 * R_EARTH)))) * 1.01 + 1e-6
    lo, hi = lo - margin, hi + margin
    rmin, rmax = _row_lat_extent(src)
    hit = (rmax >= lo) & (rmin <= hi)
    idx = hit.nonzero().flatten()
    if idx.numel() == 0:
        return 0, 0
    first, last = int(idx[0]), int(idx[-1])
    return first, last - first + 1


_SPAN_CACHE: dict = {}


def block_column_span(src, dst, row0: int, nrows: int, radius: float, align: int = 32):
    """Cached front of ``_block_column_span`` (a span depends on the two geometries, the rows and the radius only; the
    resamplers ask for the same bands scene after scene)."""
    src, dst = as_geometry(src), as_geometry(dst)
    try:
        key = (hash(src), hash(dst), int(row0), int(nrows), float(radius), int(align))
    except TypeError:
        return _block_column_span(src, dst, row0, nrows, radius, align)
    hit = _SPAN_CACHE.get(key)
    if hit is None:
        if len(_SPAN_CACHE) > 4096:
            _SPAN_CACHE.clear()
        hit = _SPAN_CACHE[key] = _block_column_span(src, dst, row0, nrows, radius, align)
    return hit


def _block_column_span(src, dst, row0: int, nrows: int, radius: float, align: int = 32):
    """Columns ``[c_lo, c_hi)`` of target rows [row0, row0 + nrows) that a geostationary source can reach at all;
    everything outside is fill value whatever the data (the far side of the earth as seen from the satellite).

    The test is the one the search kernel applies before it looks at any source pixel (``nn_fixed_position``,
    csrc/nn_geos.cu): with (rc, rs) the geocentric radius vector of the target latitude, a target is on the far side
    when ``rc * cos(lon - lon_0) < 1 / radius_g - 1.5 * radius / a``; here in float64 with 1e-4 of slack, over the row
    of the block nearest the equator (largest rc).  Only for rectilinear targets (lon depends on the column only:
    eqc, merc, longlat) and geostationary area sources; any other pair, or a span that wraps around the antimeridian,
    returns the full width.  ``align``: the span is widened to multiples of ``align`` columns (128-byte rows for the
    copy engines)."""
    src, dst = as_geometry(src), as_geometry(dst)
    width = dst.width
    full = (0, width)
    if nrows <= 0:
        return (0, 0)
    sp, dp = getattr(src, "proj_dict", None), getattr(dst, "proj_dict", None)
    if not sp or not dp or sp.get("proj") != "geos" or dp.get("proj") not in ("eqc", "merc", "longlat", "latlong", "latlon", "lonlat"):
        return full
    import numpy as np
    from .geometry import _ellipsoid
    a, es = _ellipsoid(sp)
    rg = 1.0 + float(sp["h"]) / a
    lon0 = math.radians(float(sp.get("lon_0", 0.0)))
    x, y = dst.get_proj_vectors()
    y = y[row0:row0 + nrows]
    kind = dp.get("proj")
    if kind == "eqc":        # PROJ eqc: x = a lam cos(lat_ts) + x_0, y = a (phi - lat_0) + y_0
        da, _ = _ellipsoid(dp)
        lons = np.degrees((x - float(dp.get("x_0", 0.0))) / (da * math.cos(math.radians(float(dp.get("lat_ts", 0.0)))))) \
            + float(dp.get("lon_0", 0.0))
        lats = np.degrees((y - float(dp.get("y_0", 0.0))) / da) + float(dp.get("lat_0", 0.0))
    elif kind in ("longlat", "latlong", "latlon", "lonlat"):
        lons, lats = x.astype(np.float64), y.astype(np.float64)
    else:                    # merc: the latitude needs the iterative inverse -- take it from the device
        _, lats = dst[slice(row0, row0 + nrows), slice(0, 1)].get_lonlats()
        lons, _ = dst[slice(row0, row0 + 1), slice(0, width)].get_lonlats()
    lats, lons = np.asarray(lats, dtype=np.float64).ravel(), np.asarray(lons, dtype=np.float64).ravel()
    ok = np.isfinite(lats) & (np.abs(lats) <= 90.0)
    if not ok.any() or not np.isfinite(lons).all():
        return full
    phi = np.deg2rad(lats[ok])
    rp2 = 1.0 - es
    pg = np.arctan(rp2 * np.tan(phi))
    r = math.sqrt(rp2) / np.sqrt(rp2 * np.cos(pg) ** 2 + np.sin(pg) ** 2)
    rc = float((r * np.cos(pg)).max())
    vis_min = 1.0 / rg - 1.5 * radius / a - 1e-5 - 1e-4
    thr = vis_min / rc
    if thr > 1.0:
        return (0, 0)
    vis = np.cos(np.deg2rad(lons) - lon0) >= thr
    idx = np.flatnonzero(vis)
    if idx.size == 0:
        return (0, 0)
    if idx.size != idx[-1] - idx[0] + 1:     # two pieces (the disk straddles the antimeridian of the target grid)
        return full
    c_lo = (int(idx[0]) // align) * align
    c_hi = min(width, -(-(int(idx[-1]) + 1) // align) * align)
    return (c_lo, c_hi)


def all_gather_rows(full_image, tile_rows: int, rank: int, group=None):
    """One NCCL all-gather, in place: rank r owns rows [r*tile_rows, (r+1)*tile_rows) of ``full_image``
    (shape ``(world_size * tile_rows, width)``); afterwards every rank holds every tile."""
    import torch.distributed as dist
    mine = full_image[rank * tile_rows:(rank + 1) * tile_rows]
    dist.all_gather_into_tensor(full_image, mine, group=group)
    return full_image


def exchange_rows_async(full_image, tile_rows: int, rank: int, world_size: int, sub_row0: int, sub_nrows: int,
                        group=None):
    """Start the exchange of ONE sub-tile (rows [sub_row0, sub_row0 + sub_nrows) of every rank's tile) as a group of
    NCCL point-to-point operations: this rank's rows go to every peer and every peer's rows land directly in their
    final place in ``full_image`` (no staging buffer).  Lets a rank ship the sub-tiles it has finished while it still
    computes the next one; the union over all sub-tiles is the same all-gather.  Returns the work handles."""
    import torch.distributed as dist
    if world_size == 1 or sub_nrows <= 0:
        return []
    ops = []
    mine = full_image[rank * tile_rows + sub_row0: rank * tile_rows + sub_row0 + sub_nrows]
    for peer in range(world_size):
        if peer == rank:
            continue
        theirs = full_image[peer * tile_rows + sub_row0: peer * tile_rows + sub_row0 + sub_nrows]
        ops.append(dist.P2POp(dist.isend, mine, peer, group))
        ops.append(dist.P2POp(dist.irecv, theirs, peer, group))
    return dist.batch_isend_irecv(ops)


def round_peer(rank: int, k: int, world_size: int) -> int:
    """Destination of ``rank`` in round ``k`` (1 <= k < world_size) of an exchange: a rotation, so that each round is a
    permutation without fixed points -- every rank sends one copy and receives one copy, every link is used once."""
    if not 1 <= k < world_size:
        raise ValueError(f"round {k} outside 1..{world_size - 1}")
    return (rank + k) % world_size


class PeerImage:
    """The assembled image as one SYMMETRIC allocation per rank (``torch.distributed._symmetric_memory``): every rank
    maps every peer's copy, and ``push_rows`` copies rows this rank has finished into all of them with the copy
    engines over NVLink / NVSwitch.  No SM and no NCCL kernel is involved, so the exchange overlaps the search of the
    next row block without competing for the SMs the search kernel fills, and small pieces cost nothing extra
    (measured: 775 GB/s per direction on 2 and on 4 x B200 for pieces from 0.8 GB down to 0.1 GB; NCCL all-gathers of
    such pieces next to the running search kernel ran at half their stand-alone bandwidth).

    ``push_rows`` is COLLECTIVE: every rank calls it the same number of times per scene, with pieces of (nearly) the
    same size.  Call k of every rank forms one exchange of world - 1 rounds; in round j rank r writes to rank
    (r + j) mod world, so that every link carries exactly one copy in each direction, and a device-side barrier on the
    exchange stream starts each round on all ranks together.  Without that lock step the ranks drift (row blocks cost
    differently), several of them write to the same peer while other links idle, and the exchange took 21-25 ms
    instead of 13 ms for config 5 on 4 GPUs.

    ``finish()`` waits for this rank's copies and passes a device-side barrier: after it, ``local`` holds every
    rank's rows.  The image stays valid until a peer's next ``push_rows``; callers that consume it while the next
    scene is being computed need a second instance (double buffering).
    """

    BARRIER_TIMEOUT_MS = 30000  # a rank that died must not leave the others spinning

    def __init__(self, shape, dtype=None, group=None):
        torch = _lib.require_cuda()
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.torch = torch
        self.group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        dtype = torch.float32 if dtype is None else dtype
        self.local = symm_mem.empty(tuple(shape), dtype=dtype, device=torch.device("cuda", torch.cuda.current_device()))
        self.handle = symm_mem.rendezvous(self.local, self.group)
        self.peers = [self.local if r == self.rank else self.handle.get_buffer(r, tuple(shape), dtype)
                      for r in range(self.world)]
        # High priority: the device-side barrier that opens every round is a (tiny) kernel, and the block scheduler
        # hands free CTA slots to the oldest kernel of equal priority first -- behind a running search kernel with
        # thousands of CTAs still to dispatch, a barrier waited 0.5 - 2 ms and the rounds trailed the computation
        # (trace: 12 pushes of 0.47 ms spread over 10.5 ms).  With priority the barrier takes the next slot that frees.
        self.stream = torch.cuda.Stream(priority=-1)
        # the fill (one bandwidth-bound launch over everything outside the spans) runs at normal priority: issued after
        # the last block's search it takes the SM slots that kernel's tail leaves and overlaps the last pushes
        self.fill_stream = torch.cuda.Stream()
        # "burst" mode (B200_EXCHANGE_MODE=burst; the default from 4 ranks on): one barrier per push, then the copies
        # to all peers at once on their own streams -- every rank sends 1/(N-1) of its egress to every peer and receives
        # the same from each, so the switch is loaded evenly without a barrier per round (28 barriers of ~0.1 ms per
        # step at N = 8 against pushes of 0.23 ms).  "rounds": the lock-step rotation described above.
        import os
        mode = os.environ.get("B200_EXCHANGE_MODE", "auto")
        # N = 4, final round-2 library: burst 9.10 ms per step, rounds 9.44 ms (both verified; different boxes of the pool)
        self.burst = mode == "burst" or (mode == "auto" and self.world >= 4)
        self.peer_streams = [torch.cuda.Stream(priority=-1) for _ in range(self.world - 1)] if self.burst else []
        self.trace = None  # set to a list to collect (peer, row0, nbytes, start event, end event) per copy
        self.bytes_pushed = 0   # payload this rank sent since the counter was last reset (bench.py reports it)

    def fill_outside(self, blocks, value=float("nan")):
        """``blocks``: ``[(row0, nrows, (c_lo, c_hi)), ...]`` -- write ``value`` into the columns outside ``[c_lo, c_hi)``
        of those rows of the LOCAL image (what no peer will send, see ``push_rows(col_span=)``), on a side stream that
        ``finish()`` joins."""
        torch = self.torch
        lib = _lib.load()
        width = self.local.shape[-1]
        key = (tuple(blocks), float(value) if value == value else "nan")
        if getattr(self, "_rects_key", None) != key:      # the rectangle table of a fixed decomposition is built once
            import ctypes as C
            rects = []
            for row0, nrows, (c_lo, c_hi) in blocks:
                if nrows <= 0:
                    continue
                if c_hi <= c_lo:
                    rects += [row0, nrows, 0, width]
                    continue
                if c_lo > 0:
                    rects += [row0, nrows, 0, c_lo]
                if c_hi < width:
                    rects += [row0, nrows, c_hi, width - c_hi]
            self._rects = (C.c_int64 * max(1, len(rects)))(*rects)
            self._n_rects = len(rects) // 4
            self._rects_key = key
        st = self.fill_stream
        st.wait_stream(torch.cuda.current_stream())
        if self._n_rects:
            _lib.check(lib.b200_fill_f32_rects(self.local.data_ptr(), width, self._rects, self._n_rects, value,
                                               int(st.cuda_stream)))

    def push_rows(self, row0: int, nrows: int, after=None, col_span=None):
        """Copy rows [row0, row0 + nrows) of ``local`` into every peer's image (``nrows`` may be 0: the rank still
        takes part in the rounds); ``after``: CUDA event recorded when those rows are complete (default: everything
        queued on the current stream so far).  ``col_span=(c_lo, c_hi)``: send only those columns (2-D copies on the
        copy engines, ``b200_copy_2d``); the receivers fill the rest themselves (``fill_outside``)."""
        torch = self.torch
        if self.world == 1:
            return
        if after is None:
            after = torch.cuda.Event()
            after.record(torch.cuda.current_stream())
        st = self.stream
        st.wait_event(after)
        mine = self.local[row0:row0 + nrows]
        width = self.local.shape[-1]
        esize = self.local.element_size()
        c_lo, c_hi = (0, width) if col_span is None else (int(col_span[0]), int(col_span[1]))
        empty = nrows <= 0 or c_hi <= c_lo
        nbytes = 0 if empty else nrows * (c_hi - c_lo) * esize
        lib = _lib.load()
        if self.burst:
            with torch.cuda.stream(st):
                self.handle.barrier(channel=1, timeout_ms=self.BARRIER_TIMEOUT_MS)
                start = torch.cuda.Event()
                start.record(st)
            if not empty:
                for k in range(1, self.world):
                    peer = round_peer(self.rank, k, self.world)
                    ps = self.peer_streams[k - 1]
                    ps.wait_event(start)
                    _lib.check(lib.b200_copy_2d(self.peers[peer][row0, c_lo:].data_ptr(), width * esize,
                                                mine[0, c_lo:].data_ptr(), width * esize, (c_hi - c_lo) * esize, nrows,
                                                int(ps.cuda_stream)))
                    self.bytes_pushed += nbytes
                    st.wait_stream(ps)     # the next push (and finish) follow these copies
            return
        with torch.cuda.stream(st):
            for k in range(1, self.world):
                self.handle.barrier(channel=1, timeout_ms=self.BARRIER_TIMEOUT_MS)
                if empty:
                    continue
                peer = round_peer(self.rank, k, self.world)
                if self.trace is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st)
                if c_lo == 0 and c_hi == width:
                    self.peers[peer][row0:row0 + nrows].copy_(mine, non_blocking=True)
                else:
                    _lib.check(lib.b200_copy_2d(self.peers[peer][row0, c_lo:].data_ptr(), width * esize,
                                                mine[0, c_lo:].data_ptr(), width * esize, (c_hi - c_lo) * esize, nrows,
                                                int(st.cuda_stream)))
                self.bytes_pushed += nbytes
                if self.trace is not None:
                    e1.record(st)
                    self.trace.append((peer, row0, nbytes, e0, e1))

    def finish(self):
        """Stream-ordered on the current stream: own copies done, then all ranks' copies done."""
        main = self.torch.cuda.current_stream()
        main.wait_stream(self.stream)
        main.wait_stream(self.fill_stream)
        if self.world > 1:
            self.handle.barrier(channel=0, timeout_ms=self.BARRIER_TIMEOUT_MS)
        return self.local
