"""Per-scene solar scalars (host, float64) for the sun-zenith kernels.

The reference obtains cos(SZA) from ``pyorbital.astronomy.cos_zen``
(satpy/modifiers/angles.py:34-36,467-469).  Everything in that formula that does
not depend on the pixel -- Greenwich mean sidereal time, the sun's right
ascension and declination at ``start_time`` -- is evaluated here once and passed
to the device in ``b200_sunz_t``; the per-pixel part runs in
satpy_b200/csrc/sunz.cu.  Formulas: Meeus low-precision solar coordinates and the
GMST polynomial in the form pyorbital publishes them.
"""
from __future__ import annotations

import datetime as dt
import math

_J2000 = dt.datetime(2000, 1, 1, 12, 0)


def _as_datetime(t) -> dt.datetime:
    if isinstance(t, dt.datetime):
        return t.replace(tzinfo=None) if t.tzinfo is None else t.astimezone(dt.timezone.utc).replace(tzinfo=None)
    import numpy as np
    return np.datetime64(t, "us").astype(dt.datetime)


def days_since_j2000(utc_time) -> float:
    d = _as_datetime(utc_time) - _J2000
    return d.days + (d.seconds + d.microseconds * 1e-6) / 86400.0


def gmst(utc_time) -> float:
    """Greenwich mean sidereal time [rad]; the T^3 coefficient is pyorbital's literal ``6.2 * 10e-6``."""
    t = days_since_j2000(utc_time) / 36525.0
    theta = 67310.54841 + t * (876600 * 3600 + 8640184.812866 + t * (0.093104 - t * 6.2 * 10e-6))
    return math.radians(theta / 240.0) % (2 * math.pi)


def sun_ra_dec(utc_time) -> tuple[float, float]:
    """Right ascension and declination of the sun [rad]."""
    t = days_since_j2000(utc_time) / 36525.0
    m_a = math.radians(357.52910 + 35999.05030 * t - 0.0001559 * t * t - 0.00000048 * t * t * t)
    l_0 = 280.46645 + 36000.76983 * t + 0.0003032 * t * t
    d_l = ((1.914600 - 0.004817 * t - 0.000014 * t * t) * math.sin(m_a)
           + (0.019993 - 0.000101 * t) * math.sin(2 * m_a) + 0.000290 * math.sin(3 * m_a))
    eclon = math.radians(l_0 + d_l)
    eps = math.radians(23.0 + 26.0 / 60.0 + 21.448 / 3600.0
                       - (46.8150 * t + 0.00059 * t * t - 0.001813 * t * t * t) / 3600)
    x = math.cos(eclon)
    y = math.cos(eps) * math.sin(eclon)
    z = math.sin(eps) * math.sin(eclon)
    r = math.sqrt(1.0 - z * z)
    return 2 * math.atan2(y, x + r), math.atan2(z, r)


def solar_scalars(utc_time) -> tuple[float, float, float]:
    """``(gmst, ra, dec)`` in radians."""
    ra, dec = sun_ra_dec(utc_time)
    return gmst(utc_time), ra, dec
