"""Checker helpers shared by the parity tests and ``bench.py``'s parity leg -- TEST INFRASTRUCTURE ONLY.

They compare what the CUDA path produced for a target row tile with what the oracle
(``oracle/reference_pipeline.py``: the restated ``KDTreeResampler`` chain, satpy/resample/kdtree.py:60-95,184-190)
produced for the same tile.  Index arrays are compared as RAW source pixel numbers (row * width + col of the whole
source grid) because the oracle sees only the band of source rows that can reach the tile, whose compressed
numbering differs from the device's.  Exactly equidistant sources are the one documented freedom (pykdtree's choice
depends on its tree layout): a mismatch counts as a *tie* only when the two candidates'
float64 chord distances to the target agree within ``tol_m``.
"""
from __future__ import annotations

import numpy as np

from .nearest import lonlat2xyz


def raw_indices(valid_input_index, index_array, src_row0: int, src_width: int) -> np.ndarray:
    """Compressed indices of a source row band (misses = n_valid or -1) -> raw pixel numbers of the whole source
    grid (-1 = miss), shape of ``index_array`` without its trailing neighbour axis."""
    idx = np.asarray(index_array)
    if idx.ndim == 3:
        idx = idx[..., 0]
    pos = np.flatnonzero(np.asarray(valid_input_index).ravel())
    n_valid = pos.size
    miss = (idx < 0) | (idx >= n_valid)
    raw = np.full(idx.shape, -1, dtype=np.int64)
    if n_valid:
        raw[~miss] = pos[idx[~miss]] + int(src_row0) * int(src_width)
    return raw


def compare_raw_indices(raw_gpu, raw_ref, src_lons, src_lats, dst_lons, dst_lats, src_row0: int, src_width: int,
                        tol_m: float = 1e-6) -> dict:
    """``{"n": pixels, "mismatch": differing pixels, "tie": of those equidistant, "non_tie": the rest}``.

    ``src_lons/src_lats``: lon/lat of the source row band starting at ``src_row0``; ``dst_lons/dst_lats``: of the tile."""
    a = np.asarray(raw_gpu, dtype=np.int64).ravel()
    b = np.asarray(raw_ref, dtype=np.int64).ravel()
    if a.shape != b.shape:
        raise ValueError(f"index tiles differ in shape: {a.shape} vs {b.shape}")
    diff = np.flatnonzero(a != b)
    res = {"n": int(a.size), "mismatch": int(diff.size), "tie": 0, "non_tie": 0, "hits": int((b >= 0).sum())}
    if diff.size == 0:
        return res
    both = (a[diff] >= 0) & (b[diff] >= 0)
    res["non_tie"] = int((~both).sum())
    d = diff[both]
    if d.size:
        off = int(src_row0) * int(src_width)
        band = np.asarray(src_lons).size
        ia, ib = a[d] - off, b[d] - off
        inside = (ia >= 0) & (ia < band)     # a device pick outside the oracle's source band cannot be checked: bad
        res["non_tie"] += int((~inside).sum())
        d, ia, ib = d[inside], ia[inside], ib[inside]
        sl, st = np.asarray(src_lons, dtype=np.float64).ravel(), np.asarray(src_lats, dtype=np.float64).ravel()
        q = lonlat2xyz(np.asarray(dst_lons, dtype=np.float64).ravel()[d], np.asarray(dst_lats, dtype=np.float64).ravel()[d])
        da = np.sqrt(((lonlat2xyz(sl[ia], st[ia]) - q) ** 2).sum(-1))
        db = np.sqrt(((lonlat2xyz(sl[ib], st[ib]) - q) ** 2).sum(-1))
        tie = np.abs(da - db) <= tol_m
        res["tie"] = int(tie.sum())
        res["non_tie"] += int((~tie).sum())
    return res


def compare_values(out_gpu, out_ref, rtol: float = 1e-5, floor: float = 1e-3, exclude=None) -> dict:
    """NaN pattern and relative error of two float images of one tile.  ``exclude``: boolean mask of pixels to leave
    out (targets whose neighbour is an equidistant tie resolved differently: the two sides then sample different
    source pixels)."""
    g, r = np.asarray(out_gpu), np.asarray(out_ref)
    keep = np.ones(g.shape, bool) if exclude is None else ~np.asarray(exclude, bool).reshape(g.shape)
    nan_mismatch = int(((np.isnan(g) != np.isnan(r)) & keep).sum())
    both = ~np.isnan(g) & ~np.isnan(r) & keep
    if both.any():
        err = np.abs(g[both].astype(np.float64) - r[both]) / np.maximum(np.abs(r[both]), floor)
        max_rel, n_bad = float(err.max()), int((err > rtol).sum())
    else:
        max_rel, n_bad = 0.0, 0
    return {"n": int(keep.sum()), "valid": int(both.sum()), "nan_mismatch": nan_mismatch, "max_rel_err": max_rel,
            "n_over_rtol": n_bad, "rtol": rtol}
