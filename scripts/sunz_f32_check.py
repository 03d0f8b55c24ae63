#!/usr/bin/env python
"""CPU emulation (numpy float32) of the float32 sun-zenith fast path of csrc/sunz.cu (``sunz_fast_corr``) against the
oracle's rounded-angle float64 flow (oracle/solar.py, restating satpy/modifiers/angles.py:418-469,555-584), on the
full ABI 1 km disk (every ``STEP``-th pixel) and the SEVIRI disk (sweep y) at several times of day.

The fast path: q = B_r (D_r - T_c), D_r = A_r / B_r, from float32 hi/lo pairs (4 instructions), then float32:
    s = sqrt(q); X = 1 + rg s; Y = Yr cy; Z = Zr cz; num = sd Z + cd_ca X - cd_sa Y; n2 = X^2 + Y^2 + Z^2
    cz = num * rsqrt(n2),  1 / cos(sza) = rcp(cz),  used where cz > alg_hi
The hardware's approximate sqrt / rsqrt / rcp are emulated as the correctly rounded value +- up to 2 ulp (random).
Prints the largest relative deviation of the correction from the oracle's above ``alg_hi`` -- it has to stay below
1e-5 (north_star) with margin; DESIGN.md quotes the result.

    python scripts/sunz_f32_check.py [alg_hi]
"""
import datetime as dt
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import solar as osolar  # noqa: E402
from oracle.projections import Area as OArea  # noqa: E402
from satpy_b200 import demo  # noqa: E402

F = np.float32
STEP = 5
rng = np.random.default_rng(0)


def approx(x64):
    """a hardware approximation: correctly rounded float32 value moved by up to 2 ulp"""
    x = x64.astype(F)
    k = rng.integers(-2, 3, size=x.shape)
    return (x.view(np.int32) + k.astype(np.int32)).view(F)


def fast_corr(area, when, rows, cols):
    proj = area.proj_dict
    a, es = area.proj.a, area.proj.es
    h = float(proj["h"])
    sweep_x = str(proj.get("sweep", "y")) == "x"
    hh = h / a
    rg = 1.0 + hh
    c = rg * rg - 1.0
    inv2 = 1.0 / (1.0 - es)
    x, y = area.proj_vectors(rows=rows, cols=cols)
    tx = np.tan(x / (a * hh))[None, :]
    ty = np.tan(y / (a * hh))[:, None]
    hz, hx = np.hypot(1.0, ty), np.hypot(1.0, tx)
    e_r = ty * ty * inv2
    T_c = tx * tx
    if sweep_x:
        A_r, B_r = 1.0 - c * e_r, c * hz * hz
        Yr, Zr, cy, cz = (c * hz).astype(F), (c * inv2 * ty).astype(F), tx.astype(F), np.ones_like(tx).astype(F)
    else:
        A_r, B_r = 1.0 - c * e_r, c * (1.0 + e_r)
        Yr, Zr, cy, cz = np.full_like(ty, c).astype(F), (c * inv2 * ty).astype(F), tx.astype(F), hx.astype(F)
    # q = B_r (D_r - T_c), D_r = A_r / B_r, with D and T as float32 hi + lo pairs
    D_r = A_r / B_r
    Dh, Th = D_r.astype(F), T_c.astype(F)
    Dl, Tl = (D_r - Dh).astype(F), (T_c - Th).astype(F)
    qf = (B_r.astype(F) * ((Dh - Th).astype(F) + (Dl - Tl).astype(F)).astype(F)).astype(F)
    q64 = A_r - B_r * T_c
    with np.errstate(invalid="ignore", divide="ignore"):
        relq = np.abs(qf - q64) / np.abs(q64)
    assert np.nanmax(np.where(np.abs(q64) > 1e-9, relq, 0)) < 1e-6, np.nanmax(relq)
    on = qf >= 0
    qf = np.where(on, qf, F(0))
    s = approx(np.sqrt(qf.astype(np.float64)))
    gmst, ra, dec = osolar.solar_scalars(when)
    ang = gmst - ra + np.deg2rad(float(proj["lon_0"]))
    sd, cd_ca, cd_sa = F(np.sin(dec)), F(np.cos(dec) * np.cos(ang)), F(np.cos(dec) * np.sin(ang))
    X = (F(1.0) + F(rg) * s).astype(F)
    Y = (Yr * cy).astype(F)
    Z = (Zr * cz).astype(F)
    num = (sd * Z).astype(F)
    num = (cd_ca * X + num).astype(F)        # FFMA: one rounding each
    num = (-cd_sa * Y + num).astype(F)
    n2 = (Z * Z).astype(F)
    n2 = (Y.astype(np.float64) * Y + n2).astype(F)
    n2 = (X.astype(np.float64) * X + n2).astype(F)
    rs = approx(1.0 / np.sqrt(n2.astype(np.float64)))
    czf = (num * rs).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        corr = approx(1.0 / czf.astype(np.float64))
    return on, czf, np.ones_like(czf), corr


def main():
    alg_hi = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
    for name, times in (("goes_east_abi_f_1km", [dt.datetime(2024, 3, 20, hh, 30) for hh in (6, 11, 17, 22)] +
                         [demo.START_TIME]),
                        ("msg_seviri_fes_3km", [dt.datetime(2020, 6, 21, 5, 0), dt.datetime(2020, 12, 21, 12, 0)])):
        area = OArea(*demo.area_params(name))
        step = STEP if area.width > 5000 else 1
        rows, cols = np.arange(0, area.height, step), np.arange(0, area.width, step)
        sub = OArea(area.proj_dict, len(cols), len(rows),
                    (area.area_extent[0] + (0.5 - step / 2) * area.pixel_size_x, 0, 0, 0))  # placeholder (unused)
        x, y = area.proj_vectors(rows=rows, cols=cols)
        lon, lat = area.proj.inverse(*np.meshgrid(x, y))
        with np.errstate(invalid="ignore"):
            lon = np.where(lon >= 1e30, np.nan, lon).astype(F)
            lat = np.where(lat >= 1e30, np.nan, lat).astype(F)
        for when in times:
            with np.errstate(invalid="ignore"):
                cz_ref = osolar.cos_zen(when, lon, lat)          # the reference's dtype flow (float32 lon/lat)
            on, num, nrm, corr = fast_corr(area, when, rows, cols)
            assert (on == np.isfinite(cz_ref)).mean() > 0.9999
            with np.errstate(invalid="ignore"):
                use = on & (num > F(alg_hi) * nrm) & np.isfinite(cz_ref)
                ref = (1.0 / cz_ref).astype(F)
                rel = np.abs(corr[use].astype(np.float64) - ref[use]) / ref[use]
                dcz = np.abs(num[use].astype(np.float64) / nrm[use] - cz_ref[use])
                lit = np.isfinite(cz_ref) & (cz_ref > np.cos(np.deg2rad(95.0)) - 1e-3)
                band = lit & ~use
            print(f"{name} {when:%Y-%m-%d %H:%M} alg_hi {alg_hi}: fast-path pixels {use.sum()} "
                  f"max rel err of 1/cos {rel.max():.2e} (mean {rel.mean():.1e}), max |d cos| {dcz.max():.2e}; "
                  f"exact-band share of the disk {band.sum() / on.sum():.3f}")
            assert rel.max() < 8e-6


if __name__ == "__main__":
    main()
