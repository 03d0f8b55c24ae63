"""CPU study (numpy, no GPU) for the next round: a RIGOROUS lattice-model pruning from discrete second differences.
For a window of the ABI 1 km grid at the diagonal limb, per target: candidates of the rigorous disc (what the kernel
walks today where the model is rejected) vs candidates that survive  modelDist(i,j) - E(i,j) < bd,
E(i,j) = i(i-1)/2 Mcc + j(j-1)/2 Mrr + |ij| Mcr  (a bound on the deviation from the local affine lattice that holds
for ANY lattice: it is a sum of discrete second differences along the path)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satpy_b200 import demo
from oracle.projections import Area as OArea
from oracle.nearest import lonlat2xyz

src = OArea(*demo.area_params("goes_east_abi_f_1km", 1.0))
x, y = src.proj_vectors()
R = 2000.0
rng = np.random.default_rng(0)

def window(r0, c0, n=260):
    X, Y = np.meshgrid(x[c0:c0 + n], y[r0:r0 + n])
    lon, lat = src.proj.inverse(X, Y)
    ok = np.isfinite(lon) & np.isfinite(lat)
    P = lonlat2xyz(np.where(ok, lon, 0).ravel(), np.where(ok, lat, 0).ravel()).reshape(n, n, 3)
    P[~ok] = np.nan
    return P, lon, lat, ok

def study(name, r0, c0, nwin=260):
    P, lon, lat, ok = window(r0, c0, nwin)
    n = P.shape[0]
    # discrete second differences (norms)
    dcc = np.linalg.norm(P[:, 2:] - 2 * P[:, 1:-1] + P[:, :-2], axis=-1)
    drr = np.linalg.norm(P[2:] - 2 * P[1:-1] + P[:-2], axis=-1)
    dcr = np.linalg.norm(P[1:, 1:] - P[1:, :-1] - P[:-1, 1:] + P[:-1, :-1], axis=-1)
    K = 4
    tot_disc = tot_rig = tot_need = 0
    n2pct = 0
    ntar = 0
    nmiss = 0
    for _ in range(4000):
        r = rng.integers(K + 3, n - K - 4); c = rng.integers(K + 3, n - K - 4)
        blk = P[r:r + 2, c:c + 2]
        if not np.isfinite(blk).all() or not np.isfinite(P[r - K - 1:r + K + 3, c - K - 1:c + K + 3]).all():
            continue
        fx, fy = rng.random(2)
        # a target inside the cell: bilinear point of the 3-D block, pushed back to the sphere radius
        T = (blk[0, 0] * (1 - fx) * (1 - fy) + blk[0, 1] * fx * (1 - fy) + blk[1, 0] * (1 - fx) * fy + blk[1, 1] * fx * fy)
        T = T / np.linalg.norm(T) * 6370997.0
        d4 = np.linalg.norm(blk.reshape(4, 3) - T, axis=1)
        bd = min(R, d4.min() + 1.0)
        w = bd / (0.975 * 1002.0) * 1.0001 + 0.02
        if w < 1.0:
            continue
        ntar += 1
        if d4.min() >= R: nmiss += 1
        dev = np.linalg.norm(blk[1, 1] - (blk[0, 0] + (blk[0, 1] - blk[0, 0]) + (blk[1, 0] - blk[0, 0])))
        gmin = min(np.sum((blk[0, 1] - blk[0, 0]) ** 2), np.sum((blk[1, 0] - blk[0, 0]) ** 2))
        if dev * dev > 4e-4 * gmin: n2pct += 1
        ec = P[r, c + 1] - P[r, c]; er = P[r + 1, c] - P[r, c]
        Mcc = dcc[r - K:r + K + 2, c - K - 1:c + K + 1].max(); Mrr = drr[r - K - 1:r + K + 1, c - K:c + K + 2].max()
        Mcr = dcr[r - K:r + K + 1, c - K:c + K + 1].max()
        fc, fr = c + fx, r + fy
        for j in range(int(np.ceil(fr - w)), int(np.floor(fr + w)) + 1):
            for i in range(int(np.ceil(fc - np.sqrt(max(0, w * w - (fr - j) ** 2)))), int(np.floor(fc + np.sqrt(max(0, w * w - (fr - j) ** 2)))) + 1):
                if (j in (r, r + 1)) and (i in (c, c + 1)):
                    continue
                tot_disc += 1
                a, b = i - c, j - r
                model = P[r, c] + a * ec + b * er
                E = abs(a * (a - 1)) / 2 * Mcc + abs(b * (b - 1)) / 2 * Mrr + abs(a * b) * Mcr
                true_d = np.linalg.norm(P[j, i] - T)
                assert np.linalg.norm(P[j, i] - model) <= E + 1e-6, (a, b, np.linalg.norm(P[j, i] - model), E)
                if np.linalg.norm(model - T) - E < bd:
                    tot_rig += 1
                if true_d < bd:
                    tot_need += 1
    print(f"{name}: 2% model check fails for {n2pct}/{ntar};", end=" ")
    print(f"targets needing the disc {ntar} (misses {nmiss}); per target: rigorous disc {tot_disc / max(ntar,1):.2f} candidates, "
          f"second-difference bound leaves {tot_rig / max(ntar,1):.2f}, truly closer than best {tot_need / max(ntar,1):.2f}; "
          f"Mcc {np.nanmedian(dcc):.1f} Mrr {np.nanmedian(drr):.1f} Mcr {np.nanmedian(dcr):.1f} m; step c {np.nanmedian(np.linalg.norm(P[:,1:]-P[:,:-1],axis=-1)):.0f} r {np.nanmedian(np.linalg.norm(P[1:]-P[:-1],axis=-1)):.0f} m")

# find windows: diagonal limb (NW), west limb (row centre), moderate (zenith ~60)
def first_valid_col(r):
    X = x; Y = np.full_like(x, y[r])
    lon, lat = src.proj.inverse(X, Y)
    v = np.flatnonzero(np.isfinite(lon) & np.isfinite(lat))
    return v[0]
def study_edge(name, r_mid, inset, n=60):
    """a narrow window whose first valid column is `inset` pixels from the limb on its middle row"""
    c = first_valid_col(r_mid) + inset
    study(name, r_mid - n // 2, c - 6, n)

for inset in (12, 20, 30):
    study_edge(f"diagonal limb (NW) {inset} px inside", 1730, inset)
for inset in (12, 20, 30):
    study_edge(f"west limb {inset} px inside", 5424, inset)
