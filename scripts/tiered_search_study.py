#!/usr/bin/env python
"""CPU prototype (numpy) of the TIERED fixed-grid search of csrc/nn_geos.cu (nn_fixed_search), checked against
scipy's cKDTree on row tiles of config 5 (ABI 1 km full disk -> 0.5 km global grid, r = 2 km):

  tier 1   2 x 2 enclosing block, sufficient when  bd / (0.975 ps) * 1.0001 < (1 + edge - 0.02) * kf
  tier 2   4 x 4 window,          sufficient when                      ... < (2 + edge - 0.02) * kf
  kf       satellite-distance factor min(1.10, 0.9999 + 0.36 (rho_T^2 / h^2 - 1))   (scripts/rho_factor_check.py)
  ranking  float32 distances to float32 points, keys = distance bits with the candidate's slot in the low 4 bits;
           winner accepted when the runner-up is > 2.02 m behind and the winner > 1.02 m from the radius.

What it establishes (the GPU tests repeat it on every pixel against the strict general loop): every target the tiers
settle has the kd-tree's answer, and with kf no target of config 5 falls through to the general loop.

    python scripts/tiered_search_study.py [rows per tile]
"""
import os
import sys
import time

import numpy as np
from scipy.spatial import cKDTree

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_pipeline as rp          # noqa: E402
from oracle.nearest import lonlat2xyz                # noqa: E402
from oracle.projections import Area as OArea         # noqa: E402
from satpy_b200 import demo                          # noqa: E402

R = 2000.0
C1, KB, CAP = 0.36, 0.9999, 1.10
POS_ERR = 0.02


def run_tile(frac, nrows=16, src_name="goes_east_abi_f_1km", dst_name="global_eqc_500m", lon0_deg=-75.0, verbose=True):
    """One tile of `nrows` target rows starting at fraction `frac` of the target's height.  Returns a dict with the
    shares of the tiers and the number of settled targets whose answer differs from the kd-tree's."""
    src = OArea(*demo.area_params(src_name, 1.0))
    dst = OArea(*demo.area_params(dst_name, 1.0))
    P = src.proj
    a = P.a
    rg = P.radius_g
    hm = a * (rg - 1.0)
    rp2 = 1.0 - P.es
    rpp = np.sqrt(rp2)
    lam0 = np.radians(lon0_deg)
    psx, psy = src.pixel_size_x, src.pixel_size_y
    x_ul = src.area_extent[0] + psx / 2
    y_ul = src.area_extent[3] - psy / 2
    inv_step = np.float32(1.0 / (0.975 * min(psx, psy)))
    r0 = int(dst.height * frac)
    rows = slice(r0, r0 + nrows)
    srows = rp.source_rows_for_tile(src, dst, rows, R)
    if srows.stop - srows.start == 0:
        return None
    slon, slat = src.get_lonlats(srows)
    ok = np.isfinite(slon) & np.isfinite(slat)
    H, W = slon.shape
    xyz = lonlat2xyz(np.where(ok, slon, 0).ravel(), np.where(ok, slat, 0).ravel()).reshape(H, W, 3)
    P32 = xyz.astype(np.float32)
    P32[~ok] = np.inf
    xs, ys = dst.proj_vectors(rows=np.arange(dst.height)[rows])
    tlon, tlat = dst.proj.inverse(*np.meshgrid(xs, ys))
    T = lonlat2xyz(tlon.ravel(), tlat.ravel())
    lam, phi = np.radians(tlon.ravel()), np.radians(tlat.ravel())
    rel = (lam - lam0 + np.pi) % (2 * np.pi) - np.pi
    pg = np.arctan(rp2 * np.tan(phi))
    rr = rpp / np.hypot(rpp * np.cos(pg), np.sin(pg))
    vx, vy, vz = rr * np.cos(pg) * np.cos(rel), rr * np.cos(pg) * np.sin(rel), rr * np.sin(pg)
    vis = vx >= 1.0 / rg - 1.5 * R / a - 1e-5
    tmp = rg - vx
    ax, ay = np.arctan(vy / np.hypot(vz, tmp)), np.arctan(vz / tmp)        # sweep x (ABI)
    fc = ax * hm / psx - x_ul / psx
    fr = -ay * hm / psy + y_ul / psy - srows.start
    s = (tmp * tmp + vy * vy + vz * vz) * a * a / (hm * hm) - 1.0
    kf = np.minimum(CAP, KB + C1 * s)
    idx = np.flatnonzero(vis)
    fc, fr, kf = fc[idx].astype(np.float32), fr[idx].astype(np.float32), kf[idx].astype(np.float32)
    T32 = T[idx].astype(np.float32)
    c0, r0_ = np.floor(fc).astype(np.int64), np.floor(fr).astype(np.int64)
    interior = (c0 >= 1) & (c0 <= W - 3) & (r0_ >= 1) & (r0_ <= H - 3)
    n = idx.size
    res = np.full(n, -2, dtype=np.int64)          # -2: not settled by the tiers (general loop / border)
    ii = np.flatnonzero(interior)
    c0i, r0i, Ti = c0[ii], r0_[ii], T32[ii]

    def keys(dr, dc):
        d = (P32[r0i + dr, c0i + dc] - Ti) ** 2
        d2 = (d[:, 0] + d[:, 1] + d[:, 2]).astype(np.float32)
        return (d2.view(np.uint32) & np.uint32(0xFFFFFFF0)) | np.uint32((dr + 1) * 4 + (dc + 1))

    def track(k, m1, m2):
        return np.minimum(m1, k), np.minimum(m2, np.maximum(k, m1))
    INF = np.uint32(0x7f800000)
    m1 = np.full(ii.size, 0xFFFFFFFF, np.uint32)
    m2 = m1.copy()
    for dr in (0, 1):
        for dc in (0, 1):
            m1, m2 = track(keys(dr, dc), m1, m2)
    with np.errstate(invalid="ignore"):
        bd = np.minimum(np.float32(R), np.sqrt((m1 & np.uint32(0xFFFFFFF0)).view(np.float32)) + np.float32(1.01))
    bd = np.where(np.isfinite(bd), bd, np.float32(R))
    fx, fy = fc[ii] - c0i.astype(np.float32), fr[ii] - r0i.astype(np.float32)
    edge = np.minimum(np.minimum(fx, 1 - fx), np.minimum(fy, 1 - fy))
    wq = bd * inv_step * np.float32(1.0001)
    lim = (edge + (1.0 - POS_ERR)) * kf[ii]
    t1 = wq < lim
    t2 = (~t1) & (wq < lim + kf[ii])
    t3 = ~(t1 | t2)
    t1_plain = wq < edge + (1.0 - POS_ERR)                       # without the satellite-distance factor
    t2_plain = (~t1_plain) & (wq < edge + (2.0 - POS_ERR))
    m1b, m2b = m1.copy(), m2.copy()
    for dr in (-1, 0, 1, 2):
        for dc in (-1, 0, 1, 2):
            if dr in (0, 1) and dc in (0, 1):
                continue
            m1b, m2b = track(keys(dr, dc), m1b, m2b)
    M1, M2 = np.where(t2, m1b, m1), np.where(t2, m2b, m2)
    s1 = np.sqrt((M1 & np.uint32(0xFFFFFFF0)).view(np.float32))
    slot = (M1 & np.uint32(15)).astype(np.int64)
    win = (r0i + (slot >> 2) - 1) * W + c0i + (slot & 3) - 1
    novalid = M1 >= INF
    t = s1 + np.float32(2.02)
    with np.errstate(invalid="ignore"):
        clear_w = (M2 >= INF) | ((M2 & np.uint32(0xFFFFFFF0)).view(np.float32) > t * t * np.float32(1.000001))
        clear_r = np.abs(s1 - np.float32(R)) > np.float32(1.02)
    out = np.where(novalid, -1, np.where(s1 < R, win, -1))
    out = np.where(~novalid & ~(clear_w & clear_r), -3, out)     # -3: near-tie (refined / exact ranking)
    out = np.where(t3, -2, out)
    res[ii] = out
    vsrc = np.flatnonzero(ok.ravel())
    tree = cKDTree(xyz.reshape(-1, 3)[vsrc])
    dd, jj = tree.query(T[idx], k=1, distance_upper_bound=R)
    truth = np.where(np.isfinite(dd), vsrc[np.minimum(jj, vsrc.size - 1)], -1)
    settled = res > -2
    hidden = np.flatnonzero(~vis)
    hidden_hits = 0
    if hidden.size:
        d_h, _ = tree.query(T[hidden[:: max(1, hidden.size // 100000)]], k=1, distance_upper_bound=R)
        hidden_hits = int(np.isfinite(d_h).sum())
    info = {"frac": frac, "visible": int(n), "interior": float(interior.mean()) if n else 0.0,
            "tier1": float(t1.mean()), "tier2": float(t2.mean()), "general": float(t3.mean()),
            "general_without_kf": float((~(t1_plain | t2_plain)).mean()), "near_tie": float(np.mean(out == -3)),
            "miss": float(np.mean(out == -1)), "settled": float(settled.mean()),
            "mismatches": int((settled & (res != truth)).sum()), "hidden_hits": hidden_hits}
    if verbose:
        print(info)
    return info


if __name__ == "__main__":
    rows_per_tile = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    for f in (0.0505, 0.055, 0.07, 0.125, 0.30, 0.5):
        t0 = time.time()
        run_tile(f, rows_per_tile)
        print(f"   {time.time() - t0:.1f} s")
