"""Oracle end-to-end pipeline (numpy / scipy) -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the chain the benchmark times on the device (configs 2 and 5):
``NC_ABI_L1B.get_dataset`` (satpy/readers/abi_l1b.py:45-72) -> ``SunZenithCorrector.__call__``
(satpy/modifiers/geometry.py:48-71) -> ``KDTreeResampler.precompute`` / ``.compute``
(satpy/resample/kdtree.py:60-95,184-190).  Used by tests, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; never by ``satpy_b200``.
"""
from __future__ import annotations

import os
import time
import warnings
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import calibration, nearest, solar
from .projections import Area


_ROW_EXTENT_CACHE: dict = {}


def make_area(params) -> Area:
    proj, w, h, ext = params
    return Area(proj, w, h, ext)


def abi_reflectance(counts, band):
    return calibration.abi_calibrate(counts, "reflectance", band["scale_factor"], band["add_offset"],
                                     band["fill_value"], band.get("unsigned", False), esun=band["esun"],
                                     esd=band["earth_sun_distance_anomaly_in_AU"])


def abi_sunz(counts, band, src: Area, start_time, row_slice=None):
    """calibrate + sun-zenith correct (float32)."""
    refl = abi_reflectance(counts, band)
    return solar.sun_zenith_corrector(refl, src, start_time, row_slice=row_slice)


def _row_blocks(n_rows, width, itemsize=4, chunk_bytes=32 << 20):
    """Row blocks of about 32 MiB, the reference benchmarks' dask chunk size (benchmarks/abi_l1b_benchmarks.py:52)."""
    rows = max(1, int(chunk_bytes // (width * itemsize)))
    return [slice(r, min(n_rows, r + rows)) for r in range(0, n_rows, rows)]


def abi_sunz_threaded(counts, band, src: Area, start_time, workers=None):
    """Same as ``abi_sunz`` over ~32 MiB row blocks on a thread pool (stand-in for dask's threaded scheduler)."""
    workers = workers or len(os.sched_getaffinity(0))
    out = np.empty(counts.shape, dtype=np.float32)
    blocks = _row_blocks(counts.shape[0], counts.shape[1])

    def work(sl):
        out[sl] = abi_sunz(counts[sl], band, src, start_time, row_slice=sl)
    with ThreadPoolExecutor(max_workers=workers) as ex:
        list(ex.map(work, blocks))
    return out


def abi_sunz_nearest(counts, band, src_params, dst_params, start_time, radius, dst_rows: slice | None = None,
                     src_rows: slice | None = None, workers=-1, timings: dict | None = None,
                     info: dict | None = None):
    """The whole chain; ``dst_rows`` restricts the target to a row window and ``src_rows`` says which source rows
    ``counts`` covers (bounded CPU samples: a target tile with the band of source rows that can reach it).
    ``info``: a dict that receives the intermediate arrays (neighbour info, lon/lat, corrected source data) so that
    a checker can compare indices as well as values (oracle/compare.py)."""
    src, dst = make_area(src_params), make_area(dst_params)
    t0 = time.perf_counter()
    if src_rows is None:
        data = abi_sunz_threaded(counts, band, src, start_time)
    else:
        data = abi_sunz(counts, band, src, start_time, row_slice=src_rows)
    t1 = time.perf_counter()
    slon, slat = src.get_lonlats(src_rows)
    dlon, dlat = dst.get_lonlats(dst_rows)
    t2 = time.perf_counter()
    valid_in, valid_out, index = nearest.get_neighbour_info(slon, slat, dlon, dlat, radius, workers=workers)
    t3 = time.perf_counter()
    out = nearest.get_sample_from_neighbour_info(data, valid_in, index, np.float32(np.nan))
    t4 = time.perf_counter()
    if timings is not None:
        for k, v in (("calibrate_sunz", t1 - t0), ("lonlats", t2 - t1), ("search", t3 - t2), ("gather", t4 - t3)):
            timings[k] = timings.get(k, 0.0) + v
    if info is not None:
        info.update(valid_input_index=valid_in, valid_output_index=valid_out, index_array=index, src_lons=slon,
                    src_lats=slat, dst_lons=dlon, dst_lats=dlat, data=data,
                    src_row0=0 if src_rows is None else (src_rows.start or 0), src_width=src.width)
    return out


def source_rows_for_tile(src: Area, dst: Area, dst_rows: slice, radius, col_step=8, pad_rows=8) -> slice:
    """Band of source rows that can hold neighbours of a target row tile (the CPU stand-in for
    ``get_area_slices``, satpy/scene.py:942-947): latitude extents from every ``col_step``-th column."""
    _, dlat = dst.proj.inverse(*np.meshgrid(dst.proj_vectors()[0][::max(1, dst.width // 512)],
                                            dst.proj_vectors(rows=np.arange(dst.height)[dst_rows])[1]))
    ok = np.isfinite(dlat)
    if not ok.any():
        return slice(0, 0)
    margin = np.degrees(radius / nearest.R_EARTH) * 1.05 + 1e-6
    lo, hi = dlat[ok].min() - margin, dlat[ok].max() + margin
    key = (tuple(sorted((k, str(v)) for k, v in src.proj_dict.items())), src.width, src.height, src.area_extent, col_step)
    if key not in _ROW_EXTENT_CACHE:
        x, y = src.proj_vectors()
        _, slat = src.proj.inverse(*np.meshgrid(x[::col_step], y))
        with np.errstate(invalid="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            slat = np.where(np.isfinite(slat), slat, np.nan)
            _ROW_EXTENT_CACHE[key] = (np.nanmin(slat, axis=1), np.nanmax(slat, axis=1))
    rmin, rmax = _ROW_EXTENT_CACHE[key]
    with np.errstate(invalid="ignore"):
        hit = np.flatnonzero((rmax >= lo) & (rmin <= hi))
    if hit.size == 0:
        return slice(0, 0)
    return slice(max(0, hit[0] - pad_rows), min(src.height, hit[-1] + 1 + pad_rows))


def sample_tiles(src: Area, dst: Area, radius, rows_frac=1.0 / 256.0, at=(1 / 8, 3 / 8, 5 / 8, 7 / 8), n_rows=None):
    """Bounded sample of a large resampling job for the CPU side: target row tiles of ``rows_frac`` of the height
    (or ``n_rows`` rows) starting at the fractions ``at`` of the height, each with the band of source rows that can
    reach it.  Returns ``[(dst_rows, src_rows), ...]`` (slices).  ``bench.py`` times the oracle on these tiles and
    the parity tests / the bench's parity leg compare the device's rows of the same tiles with the result."""
    n = max(1, int(round(dst.height * rows_frac))) if n_rows is None else int(n_rows)
    tiles = []
    for frac in at:
        r0 = max(0, min(dst.height - n, int(dst.height * frac)))
        rows = slice(r0, min(dst.height, r0 + n))
        tiles.append((rows, source_rows_for_tile(src, dst, rows, radius)))
    return tiles
