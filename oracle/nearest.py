"""Oracle nearest-neighbour resampling (numpy + scipy cKDTree) -- TEST INFRASTRUCTURE ONLY.

Restates ``pyresample.kd_tree.XArrayResamplerNN.get_neighbour_info`` /
``get_sample_from_neighbour_info`` as satpy drives them:
  * ``KDTreeResampler.precompute``  satpy/resample/kdtree.py:60-95  (neighbours=1, epsilon, radius)
  * ``KDTreeResampler.compute``     satpy/resample/kdtree.py:184-190
  * cached array names / dims       satpy/resample/kdtree.py:20-22
  * fill value choice               satpy/resample/base.py:170-174,201

pyresample (>=1.24, pyproject.toml:19) and pykdtree (unpinned, pyproject.toml:16)
are un-vendored dependencies; their published algorithm:
spherical cartesian coordinates with R = 6370997.0 m, a kd-tree over the VALID
source points only, Euclidean (chord) distance against ``radius_of_influence``
(pykdtree accepts a neighbour when ``dist^2 < radius^2``), misses -> index N_valid,
which the gather turns into ``fill_value``.

Pinned by the only real (un-mocked) nearest case in the reference:
satpy/tests/test_resample.py:109-129.

``scipy.spatial.cKDTree`` stands in for pykdtree (same miss convention).  It is
exact for eps=0 but its choice between EXACTLY equidistant sources may differ
from pykdtree's, so callers compare indices with ``compare_indices`` which
accepts a mismatch only when the two candidates' squared distances agree to
a few ulp ("equidistant ties documented").
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

R_EARTH = 6370997.0  # pyresample.geometry / kd_tree module constant


def lonlat2xyz(lons, lats):
    """pyresample ``_get_cartesian_coords``: R*(cos lat cos lon, cos lat sin lon, sin lat), float64."""
    lons = np.deg2rad(np.asarray(lons, dtype=np.float64))
    lats = np.deg2rad(np.asarray(lats, dtype=np.float64))
    x = R_EARTH * np.cos(lats) * np.cos(lons)
    y = R_EARTH * np.cos(lats) * np.sin(lons)
    z = R_EARTH * np.sin(lats)
    return np.stack([x.ravel(), y.ravel(), z.ravel()], axis=-1)


def valid_lonlat(lons, lats):
    """pyresample kd_tree ``_get_valid_input_index``: -180<=lon<=180 & -90<=lat<=90 (NaN/inf fail)."""
    with np.errstate(invalid="ignore"):
        return (lons >= -180.0) & (lons <= 180.0) & (lats >= -90.0) & (lats <= 90.0)


def swath_default_radius(lons, lats):
    """pyresample ``SwathDefinition.geocentric_resolution``-style default used by the 2x2 golden case.

    Distance between the first two columns of the middle row in WGS84 geocentric
    cartesian metres; only used to reproduce
    satpy/tests/test_resample.py:109-129 with the default radius.
    """
    a, f = 6378137.0, 1 / 298.257223563
    e2 = 2 * f - f * f
    row = lons.shape[0] // 2
    lo = np.deg2rad(np.asarray(lons[row, :2], dtype=np.float64))
    la = np.deg2rad(np.asarray(lats[row, :2], dtype=np.float64))
    n = a / np.sqrt(1 - e2 * np.sin(la) ** 2)
    xyz = np.stack([n * np.cos(la) * np.cos(lo), n * np.cos(la) * np.sin(lo), n * (1 - e2) * np.sin(la)], -1)
    return float(np.linalg.norm(xyz[1] - xyz[0]))


def get_neighbour_info(src_lons, src_lats, dst_lons, dst_lats, radius_of_influence, mask=None,
                       epsilon=0.0, neighbours=1, workers=-1, brute_force=False):
    """Return ``(valid_input_index, valid_output_index, index_array)``.

    ``index_array`` has shape ``dst.shape + (neighbours,)`` (int64), indexes the
    COMPRESSED valid-input vector, and holds ``n_valid`` (the pykdtree miss
    convention) for misses and invalid outputs -- ``gather`` maps that to fill.
    """
    src_lons = np.asarray(src_lons, dtype=np.float64)
    src_lats = np.asarray(src_lats, dtype=np.float64)
    dst_lons = np.asarray(dst_lons, dtype=np.float64)
    dst_lats = np.asarray(dst_lats, dtype=np.float64)
    valid_in = valid_lonlat(src_lons, src_lats)
    if mask is not None:
        valid_in &= ~np.asarray(mask, dtype=bool)
    valid_out = valid_lonlat(dst_lons, dst_lats)
    n_valid = int(valid_in.sum())
    out_shape = dst_lons.shape + (neighbours,)
    index = np.full((dst_lons.size, neighbours), n_valid, dtype=np.int64)
    if n_valid and valid_out.any():
        src_xyz = lonlat2xyz(src_lons[valid_in], src_lats[valid_in])
        dst_xyz = lonlat2xyz(dst_lons[valid_out], dst_lats[valid_out])
        if brute_force:
            idx = _brute(src_xyz, dst_xyz, radius_of_influence, neighbours)
        else:
            tree = cKDTree(src_xyz, leafsize=16)
            _, idx = tree.query(dst_xyz, k=neighbours, eps=epsilon,
                                distance_upper_bound=radius_of_influence, workers=workers)
            idx = np.asarray(idx, dtype=np.int64).reshape(-1, neighbours)
        index[valid_out.ravel()] = idx
    return valid_in, valid_out, index.reshape(out_shape)


def _brute(src_xyz, dst_xyz, radius, k):
    """O(N*M) exact float64 search; ties -> lowest compressed index (the CUDA path's documented rule)."""
    n = src_xyz.shape[0]
    out = np.full((dst_xyz.shape[0], k), n, dtype=np.int64)
    r2 = float(radius) ** 2
    for i in range(dst_xyz.shape[0]):
        d = src_xyz - dst_xyz[i]
        d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
        order = np.lexsort((np.arange(n), d2))[:k]
        ok = d2[order] < r2
        out[i, :ok.sum()] = order[ok]
    return out


def get_sample_from_neighbour_info(data, valid_input_index, index_array, fill_value=np.nan):
    """``data.ravel()[valid_input_index][index_array]`` with misses -> fill (leading dims broadcast)."""
    data = np.asarray(data)
    n_valid = int(valid_input_index.sum())
    idx = index_array[..., 0]
    miss = (idx >= n_valid) | (idx < 0)
    safe = np.where(miss, 0, idx)
    lead = data.shape[:-2]
    flat = data.reshape(lead + (-1,))[..., valid_input_index.ravel()]
    if n_valid == 0:
        out = np.empty(lead + idx.shape, dtype=data.dtype)
        out[...] = fill_value
        return out
    out = flat[..., safe]
    out[..., miss] = fill_value
    return out


def to_minus_one(index_array, n_valid):
    """The stored ``index_array`` convention used by this repo's C-ABI: misses are -1."""
    return np.where(index_array >= n_valid, -1, index_array)


def compare_indices(idx_a, idx_b, src_lons, src_lats, dst_lons, dst_lats, valid_input_index, tol_m=1e-6):
    """Compare two compressed index arrays (misses = -1).  Returns ``(n_mismatch, n_tie, n_bad)``.

    A mismatch is a *tie* when both candidates are hits whose float64 chord
    distances to the target differ by <= ``tol_m`` metres (default one micron:
    coordinates near 6.4e6 m carry ~1e-9 m of rounding, and the device's
    sin/cos differ from numpy's in the last ulp); anything else is bad.
    """
    a = np.asarray(idx_a).reshape(-1)
    b = np.asarray(idx_b).reshape(-1)
    diff = np.flatnonzero(a != b)
    if diff.size == 0:
        return 0, 0, 0
    both_hit = (a[diff] >= 0) & (b[diff] >= 0)
    n_bad = int((~both_hit).sum())
    d = diff[both_hit]
    if d.size:
        sl = np.asarray(src_lons, dtype=np.float64)[valid_input_index]
        st = np.asarray(src_lats, dtype=np.float64)[valid_input_index]
        q = lonlat2xyz(np.asarray(dst_lons, dtype=np.float64).reshape(-1)[d],
                       np.asarray(dst_lats, dtype=np.float64).reshape(-1)[d])
        pa = lonlat2xyz(sl[a[d]], st[a[d]])
        pb = lonlat2xyz(sl[b[d]], st[b[d]])
        da = ((pa - q) ** 2).sum(-1)
        db = ((pb - q) ** 2).sum(-1)
        tie = np.abs(np.sqrt(da) - np.sqrt(db)) <= tol_m
        n_bad += int((~tie).sum())
        n_tie = int(tie.sum())
    else:
        n_tie = 0
    return int(diff.size), n_tie, n_bad
