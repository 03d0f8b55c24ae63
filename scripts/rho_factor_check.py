#!/usr/bin/env python
"""Numerical check of the satellite-distance factor of the tiered search (csrc/nn_geos.cu, nn_fixed_search):

    kf = min(K_CAP, K_B + K_A * (rho_T^2 / h^2 - 1))   must satisfy   h * kf <= rho(candidate)

for every ellipsoid point seen within `disc_px` pixels of the target's view direction -- then two directions alpha apart
are at least 2 h kf sin(alpha / 2) apart on the ground, and a grid distance g is at least 0.975 kf g pixel_size metres.
Checked over all view angles up to the horizon, seven azimuths (equatorial ... polar), visible targets and targets just
behind the limb (their range is that of the far intersection), for pixel angles up to 250 urad.

    python scripts/rho_factor_check.py
"""
import numpy as np

A, B, H = 6378137.0, 6356752.314, 35786023.0      # GRS80, GOES-R perspective height
K_A, K_B, K_CAP = 0.36, 0.9999, 1.10               # the kernel's constants (nn_fixed_grid_query_impl)


def rho(theta, psi, far=False):
    """Range from the satellite to the ellipsoid along the direction `theta` off nadir at azimuth `psi`
    (0 = towards the equator's limb, pi/2 = towards a pole); NaN where the ray misses the earth."""
    D = A + H
    u = np.array([-np.cos(theta), np.sin(theta) * np.cos(psi), np.sin(theta) * np.sin(psi)])
    us = u * np.array([1.0, 1.0, A / B])[:, None]
    a2 = (us * us).sum(0)
    b2 = 2.0 * D * us[0]
    disc = b2 * b2 - 4.0 * a2 * (D * D - A * A)
    ok = disc >= 0
    sq = np.sqrt(np.where(ok, disc, 0.0))
    t = (-b2 + sq) / (2 * a2) if far else (-b2 - sq) / (2 * a2)
    return np.where(ok, t, np.nan)


def kfac(r):
    return np.minimum(K_CAP, K_B + K_A * ((r / H) ** 2 - 1.0))


def worst_ratio(px_angle, disc_px=3.0, n_theta=200001):
    """min over directions of rho(candidate) / (h * kf(target)); must be >= 1."""
    worst = np.inf
    delta = disc_px * px_angle
    th = np.linspace(0.0, 0.1525, n_theta)
    for psi in np.linspace(0.0, np.pi / 2, 7):
        ps = np.full_like(th, psi)
        r_near, r_far = rho(th, ps), rho(th, ps, far=True)
        vis = np.isfinite(r_near)
        rmin = np.full_like(th, np.inf)     # smallest range within `delta` of the direction: inward and sideways
        for dth, side in ((-delta, 0.0), (-0.7 * delta, 0.7), (-0.7 * delta, -0.7), (0.0, 1.0), (0.0, -1.0)):
            with np.errstate(divide="ignore", invalid="ignore"):
                psi2 = psi + np.where(th > 0, side * delta / np.maximum(th, 1e-9), 0.0)
            r2 = rho(np.abs(th + dth), psi2)
            rmin = np.where(np.isfinite(r2), np.minimum(rmin, r2), rmin)
        # hidden targets: up to 300 km of path behind the limb (1.5 radii of influence is far less)
        for r_t, sel in ((r_near, vis), (r_far, vis & ((r_far - r_near) < 300e3))):
            m = sel & np.isfinite(rmin) & np.isfinite(r_t)
            worst = min(worst, float(((rmin[m] / H) / kfac(r_t[m])).min()))
    return worst


if __name__ == "__main__":
    for px in (14e-6, 56e-6, 112e-6, 250e-6):
        print(f"pixel angle {px * 1e6:5.0f} urad: min rho_candidate / (h kf) = {worst_ratio(px):.6f}")
