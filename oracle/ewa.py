"""Oracle elliptical-weighted-averaging resampling (numpy) -- TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED.**

Restates ``pyresample.ewa`` -- ``ll2cr`` (``_ll2cr.pyx``) and ``fornav`` (``_fornav.pyx`` + ``_fornav_templates.cpp``,
ms2gt lineage; pyresample >= 1.24 is an un-vendored dependency, pyproject.toml:19) -- which satpy exposes unchanged as
the ``ewa`` resampler (satpy/resample/ewa.py:3-11); ``rows_per_scan`` comes from the reader
(satpy/readers/core/viirs_atms_sdr.py:247-255,283).  No test in the reference exercises ``ewa``, so nothing pins these
numbers: the module follows the library's published algorithm, float32 weights, ellipse
parameters and accumulators (``typedef float weight_type / ewa_param_type / accum_type``) and the C++ loop's
incremental evaluation of ``q`` (``q += dq; dq += ddq``) so that the weight-table indices are reproduced exactly.

``DaskEWAResampler`` runs the same ``fornav`` for every (input scan chunk, output chunk) pair and sums the weight and
accumulation grids before normalising; that sum equals one full-grid pass, which is what is restated here.

Known-answer anchors (no reference vector exists): tests/test_oracle_golden.py::test_fornav_oracle_on_a_one_to_one_swath_is_a_fixed_weighted_average / test_ll2cr_oracle_is_the_affine_map_of_the_target_area.
"""
from __future__ import annotations

import numpy as np

EPSILON = np.float32(1e-8)
F32 = np.float32


def ll2cr(lons, lats, dst):
    """Swath lon/lat -> fractional (cols, rows) of the target area; NaN where not projectable (``fill=np.nan``).

    ``col 0.0 / row 0.0`` is the centre of the top-left cell.  Returns ``(points_in_grid, cols, rows)`` in the dtype
    of the lon/lat arrays (``cr_dtype``)."""
    dt = np.float32 if np.asarray(lons).dtype == np.float32 else np.float64
    cw = dst.pixel_size_x
    ch = -abs(dst.pixel_size_y)
    ox = dst.area_extent[0] + cw / 2.0
    oy = dst.area_extent[3] + ch / 2.0
    x, y = dst.proj.forward(np.asarray(lons, dtype=np.float64), np.asarray(lats, dtype=np.float64))
    x, y = x.astype(dt), y.astype(dt)   # pyproj returns the input dtype; the arithmetic below happens in cr_dtype
    with np.errstate(invalid="ignore"):
        bad = (x >= 1e30) | np.isnan(x) | (y >= 1e30) | np.isnan(y)
        cols = ((x - dt(ox)) / dt(cw)).astype(dt)
        rows = ((y - dt(oy)) / dt(ch)).astype(dt)
        cols[bad] = np.nan
        rows[bad] = np.nan
        inside = (cols >= -1) & (cols <= dst.width + 1) & (rows >= -1) & (rows <= dst.height + 1)
    return int(inside.sum()), cols, rows


def weight_table(weight_count=10000, weight_min=0.01, weight_distance_max=1.0):
    """``initialize_weight``: ``wtab[i] = exp(-alpha * qmax * i / (count - 1))`` in float32."""
    qmax = F32(weight_distance_max) * F32(weight_distance_max)
    alpha = F32(-np.log(F32(weight_min)) / qmax)
    idx = np.arange(weight_count, dtype=np.float32)
    wtab = np.exp(-alpha * qmax * idx / F32(weight_count - 1)).astype(np.float32)
    qfactor = F32(F32(weight_count) / qmax)
    return wtab, qmax, qfactor


def ewa_parameters(u, v, qmax, distance_max=1.0, delta_max=10.0):
    """``compute_ewa_parameters`` for one scan (``u``/``v`` = col/row images of shape (rows_per_scan, ncols))."""
    u = u.astype(np.float32, copy=False)
    v = v.astype(np.float32, copy=False)
    nrows, ncols = u.shape
    mid, rowsm1 = nrows // 2, nrows - 1
    dmax, delta_max = F32(distance_max), F32(delta_max)
    a = np.zeros(ncols, np.float32)
    b = np.zeros(ncols, np.float32)
    c = np.zeros(ncols, np.float32)
    udel = np.full(ncols, dmax, np.float32)
    vdel = np.full(ncols, dmax, np.float32)
    with np.errstate(all="ignore"):
        ux = (u[mid, 2:] - u[mid, :-2]) / F32(2.0) * dmax
        vx = (v[mid, 2:] - v[mid, :-2]) / F32(2.0) * dmax
        uy = (u[rowsm1, 1:-1] - u[0, 1:-1]) / F32(rowsm1) * dmax
        vy = (v[rowsm1, 1:-1] - v[0, 1:-1]) / F32(rowsm1) * dmax
        nan = np.isnan(ux) | np.isnan(vx) | np.isnan(uy) | np.isnan(vy)
        fs = ux * vy - uy * vx
        fs = fs * fs
        fs = np.where(fs < EPSILON, EPSILON, fs)
        fs = qmax / fs
        ai = (vx * vx + vy * vy) * fs
        bi = F32(-2.0) * (ux * vx + uy * vy) * fs
        ci = (ux * ux + uy * uy) * fs
        d = F32(4.0) * ai * ci - bi * bi
        d = np.where(d < EPSILON, EPSILON, d)
        d = (F32(4.0) * qmax) / d
        ud = np.minimum(np.sqrt(ci * d), delta_max)
        vd = np.minimum(np.sqrt(ai * d), delta_max)
    a[1:-1] = np.where(nan, 0, ai)
    b[1:-1] = np.where(nan, 0, bi)
    c[1:-1] = np.where(nan, 0, ci)
    udel[1:-1] = np.where(nan, dmax, ud)
    vdel[1:-1] = np.where(nan, dmax, vd)
    for arr in (a, b, c, udel, vdel):
        arr[0], arr[-1] = arr[1], arr[-2]
    return a, b, c, udel, vdel


def fornav(cols, rows, dst_shape, data, rows_per_scan=0, fill=np.nan, weight_count=10000, weight_min=0.01,
           weight_distance_max=1.0, weight_delta_max=10.0, weight_sum_min=-1.0, maximum_weight_mode=False,
           return_grids=False):
    """``fornav``: splat every swath pixel into the grid cells inside its ellipse, then normalise.

    ``data``: (rows, cols) or (bands, rows, cols).  Returns ``(valid_counts, out)`` like the library (out in the
    data dtype, ``fill`` where the weight sum is below ``weight_sum_min``)."""
    data = np.asarray(data)
    bands = data.reshape((-1,) + data.shape[-2:])
    nrows, ncols = cols.shape
    if rows_per_scan <= 0:
        rows_per_scan = nrows
    if nrows % rows_per_scan:
        raise ValueError("swath rows must be a multiple of rows_per_scan")
    gh, gw = dst_shape
    wtab, qmax, qfactor = weight_table(weight_count, weight_min, weight_distance_max)
    u_all = cols.astype(np.float32, copy=False)
    v_all = rows.astype(np.float32, copy=False)
    grid_w = np.zeros((len(bands), gh * gw), np.float32)
    grid_a = np.zeros((len(bands), gh * gw), np.float32)
    with np.errstate(all="ignore"):
        for r0 in range(0, nrows, rows_per_scan):
            u, v = u_all[r0:r0 + rows_per_scan], v_all[r0:r0 + rows_per_scan]
            a, b, c, udel, vdel = ewa_parameters(u, v, qmax, weight_distance_max, weight_delta_max)
            A, B, C = (np.broadcast_to(p, u.shape) for p in (a, b, c))
            UD, VD = np.broadcast_to(udel, u.shape), np.broadcast_to(vdel, u.shape)
            keep = ~((u < -UD) | (v < -VD) | np.isnan(u) | np.isnan(v))
            pr, pc = np.nonzero(keep)
            if pr.size == 0:
                continue
            u0, v0 = u[pr, pc], v[pr, pc]
            pa, pb, pcc = A[pr, pc], B[pr, pc], C[pr, pc]
            pud, pvd = UD[pr, pc], VD[pr, pc]
            iu1 = np.trunc(u0 - pud).astype(np.int64)
            iu2 = np.trunc(u0 + pud).astype(np.int64)
            iv1 = np.trunc(v0 - pvd).astype(np.int64)
            iv2 = np.trunc(v0 + pvd).astype(np.int64)
            iu1 = np.maximum(iu1, 0)
            iu2 = np.minimum(iu2, gw - 1)
            iv1 = np.maximum(iv1, 0)
            iv2 = np.minimum(iv2, gh - 1)
            ok = (iu1 < gw) & (iu2 >= 0) & (iv1 < gh) & (iv2 >= 0)
            if not ok.any():
                continue
            sel = np.flatnonzero(ok)
            u0, v0, pa, pb, pcc = u0[sel], v0[sel], pa[sel], pb[sel], pcc[sel]
            iu1, iu2, iv1, iv2 = iu1[sel], iu2[sel], iv1[sel], iv2[sel]
            vals = bands[:, r0 + pr[sel], pc[sel]]
            ddq = F32(2.0) * pa
            uu = (iu1 - u0).astype(np.float32)
            a2up1 = pa * (F32(2.0) * uu + F32(1.0))
            bu = pb * uu
            au2 = pa * uu * uu
            nv = int((iv2 - iv1).max()) + 1
            nu = int((iu2 - iu1).max()) + 1
            for j in range(nv):
                iv = iv1 + j
                vrow = iv <= iv2
                vv = (iv - v0).astype(np.float32)
                dq = a2up1 + pb * vv
                q = (pcc * vv + bu) * vv + au2
                for k in range(nu):
                    iu = iu1 + k
                    m = vrow & (iu <= iu2) & (q >= 0) & (q < qmax)
                    if m.any():
                        iw = np.minimum((q[m] * qfactor).astype(np.int64), weight_count - 1)
                        w = wtab[iw]
                        cell = iv[m] * gw + iu[m]
                        for ch in range(len(bands)):
                            val = vals[ch][m]
                            if maximum_weight_mode:
                                # sequential semantics: keep the sample with the largest weight seen so far
                                order = np.argsort(-w, kind="stable")
                                cs, ws, vs = cell[order], w[order], val[order]
                                first = np.unique(cs, return_index=True)[1]
                                cs, ws, vs = cs[first], ws[first], vs[first]
                                upd = ws > grid_w[ch, cs]
                                grid_w[ch, cs[upd]] = ws[upd]
                                bad = np.isnan(vs) | (vs == fill)
                                grid_a[ch, cs[upd]] = np.where(bad, np.nan, vs)[upd].astype(np.float32)
                            else:
                                good = ~(np.isnan(val) | (val == fill))
                                np.add.at(grid_w[ch], cell[good], w[good])
                                np.add.at(grid_a[ch], cell[good], (val[good].astype(np.float32) * w[good]))
                    q = q + dq
                    dq = dq + ddq
    wsm = F32(weight_sum_min)
    if wsm <= 0:
        wsm = F32(weight_min)   # documented: "if weight_sum_min <= 0 it is set to weight_min"
    with np.errstate(all="ignore"):
        bad = (grid_w < wsm) | np.isnan(grid_a)
        res = grid_a if maximum_weight_mode else grid_a / grid_w
        out = np.where(bad, fill, res).astype(data.dtype)
    out = out.reshape(data.shape[:-2] + (gh, gw))
    counts = (~bad).sum(axis=1)
    if return_grids:
        return counts, out, grid_w.reshape(-1, gh, gw), grid_a.reshape(-1, gh, gw)
    return counts, out


def fornav_accumulate(cols, rows, dst_shape, data, **kwargs):
    """The two accumulation grids (sum of weights, sum of weight * value) of ``fornav`` before normalisation, for
    2-D ``data``: what one input chunk contributes in pyresample's dask wrapper, which SUMS these pairs over the input
    chunks and normalises once.  Test infrastructure like the rest of this module."""
    _, _, grid_w, grid_a = fornav(cols, rows, dst_shape, data, return_grids=True, **kwargs)
    return grid_w[0], grid_a[0]
