"""Row-tile decomposition of a resampling job over the GPUs of one box.

The reference has no distributed compute; its analogue of this module is dask chunking of the target
(``CHUNK_SIZE``, satpy/resample/base.py:32) plus ``Scene._reduce_data`` cropping the source to what a
target needs (``get_area_slices``, satpy/scene.py:930-954).  Here the target area is cut into equal
contiguous row tiles, one per rank; each rank works on the band of SOURCE rows that can reach its tile
(tile latitude range +- the search radius) and one NCCL all-gather assembles the image on every rank.
No index array or source data is exchanged.
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
