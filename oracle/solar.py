"""Oracle solar geometry + sun-zenith correction (numpy) -- TEST INFRASTRUCTURE ONLY.

Restates, with the reference's own dtype flow:
  * pyorbital ``astronomy.cos_zen`` / ``sun_ra_dec`` / ``gmst`` -- the library is
    an un-vendored dependency (pyproject.toml:17, unpinned); call sites
    satpy/modifiers/angles.py:34-36,467-469.  Published algorithm: Meeus,
    "Astronomical Algorithms" low-precision solar coordinates + the IAU-82 GMST
    polynomial as written in pyorbital.
  * ``get_cos_sza``            satpy/modifiers/angles.py:418-431
  * ``_get_valid_lonlats``     satpy/modifiers/angles.py:434-441
  * ``SunZenithCorrectorBase.__call__`` masking  satpy/modifiers/geometry.py:48-71
  * ``_sunzen_corr_cos_ndarray``   satpy/modifiers/angles.py:555-584
  * ``_sunzen_reduction_ndarray``  satpy/modifiers/angles.py:594-624
  * ``atmospheric_path_length_correction``  satpy/utils.py:274-316

Pinned by satpy/tests/test_modifiers.py:34-110,116-176,198-220 (see
tests/test_oracle_golden.py).
"""
from __future__ import annotations

import datetime as dt

import numpy as np


def jdays2000(utc_time) -> np.float64:
    """Days since 2000-01-01T12:00 (pyorbital ``jdays2000``)."""
    t = np.datetime64(utc_time, "ns")
    return np.float64((t - np.datetime64("2000-01-01T12:00", "ns")) / np.timedelta64(1, "D"))


def gmst(utc_time) -> np.float64:
    """Greenwich mean sidereal time in radians (pyorbital ``gmst``; 6.2 * 10e-6 is literal)."""
    ut1 = jdays2000(utc_time) / 36525.0
    theta = 67310.54841 + ut1 * (876600 * 3600 + 8640184.812866 + ut1 * (0.093104 - ut1 * 6.2 * 10e-6))
    return np.float64(np.deg2rad(theta / 240.0) % (2 * np.pi))


def sun_ecliptic_longitude(utc_time) -> np.float64:
    jdate = jdays2000(utc_time) / 36525.0
    m_a = np.deg2rad(357.52910 + 35999.05030 * jdate - 0.0001559 * jdate * jdate
                     - 0.00000048 * jdate * jdate * jdate)
    l_0 = 280.46645 + 36000.76983 * jdate + 0.0003032 * jdate * jdate
    d_l = ((1.914600 - 0.004817 * jdate - 0.000014 * jdate * jdate) * np.sin(m_a)
           + (0.019993 - 0.000101 * jdate) * np.sin(2 * m_a) + 0.000290 * np.sin(3 * m_a))
    return np.float64(np.deg2rad(l_0 + d_l))


def sun_ra_dec(utc_time) -> tuple[np.float64, np.float64]:
    jdate = jdays2000(utc_time) / 36525.0
    eps = np.deg2rad(23.0 + 26.0 / 60.0 + 21.448 / 3600.0
                     - (46.8150 * jdate + 0.00059 * jdate * jdate - 0.001813 * jdate * jdate * jdate) / 3600)
    eclon = sun_ecliptic_longitude(utc_time)
    x__ = np.cos(eclon)
    y__ = np.cos(eps) * np.sin(eclon)
    z__ = np.sin(eps) * np.sin(eclon)
    r__ = np.sqrt(1.0 - z__ * z__)
    declination = np.arctan2(z__, r__)
    right_ascension = 2 * np.arctan2(y__, (x__ + r__))
    return np.float64(right_ascension), np.float64(declination)


def solar_scalars(utc_time) -> tuple[float, float, float]:
    """``(gmst, ra, dec)`` in radians: everything per-scene the per-pixel formula needs."""
    ra, dec = sun_ra_dec(utc_time)
    return float(gmst(utc_time)), float(ra), float(dec)


def cos_zen(utc_time, lon, lat):
    """pyorbital ``cos_zen``: dtype flow is numpy's (f32 lon/lat -> f32 deg2rad/sin/cos, f64 after)."""
    lon = np.deg2rad(lon)
    lat = np.deg2rad(lat)
    r_a, dec = sun_ra_dec(utc_time)
    h__ = (gmst(utc_time) + lon) - r_a
    return np.sin(lat) * np.sin(dec) + np.cos(lat) * np.cos(dec) * np.cos(h__)


def get_cos_sza(area, start_time, data_dtype=np.float32, row_slice=None):
    """satpy/modifiers/angles.py:418-441: lon/lat (f64) -> NaN where >=1e30 -> cast to data dtype -> cos_zen."""
    lons, lats = area.get_lonlats(row_slice)
    with np.errstate(invalid="ignore"):
        lons = np.where(lons >= 1e30, np.nan, lons)
        lats = np.where(lats >= 1e30, np.nan, lats)
    if lons.dtype != np.dtype(data_dtype) and np.issubdtype(data_dtype, np.floating):
        lons = lons.astype(data_dtype)
        lats = lats.astype(data_dtype)
    with np.errstate(invalid="ignore"):
        return cos_zen(start_time, lons, lats)


def sunzen_corr_cos(data: np.ndarray, cos_zen: np.ndarray, limit: float = 88., max_sza=95.) -> np.ndarray:
    """satpy/modifiers/angles.py:555-584, expression for expression."""
    limit_rad = np.deg2rad(limit)
    limit_cos = np.cos(limit_rad)
    max_sza_rad = np.deg2rad(max_sza) if max_sza is not None else max_sza
    with np.errstate(all="ignore"):
        corr = (1. / cos_zen).astype(data.dtype, copy=False)
        if max_sza is not None:
            grad_factor = (np.arccos(cos_zen) - limit_rad) / (max_sza_rad - limit_rad)
            grad_factor = 1. - np.log(grad_factor + 1) / np.log(2)
            grad_factor = grad_factor.clip(0.)
        else:
            grad_factor = 1.
        corr = np.where(cos_zen > limit_cos, corr,
                        np.asarray(grad_factor / limit_cos).astype(data.dtype, copy=False))
        corr[np.isnan(cos_zen)] = 0
        return data * corr


def sun_zenith_corrector(data, area, start_time, correction_limit=88., max_sza=95., sza_deg=None,
                         row_slice=None):
    """``SunZenithCorrector.__call__`` (satpy/modifiers/geometry.py:48-71,114-118)."""
    if sza_deg is None:
        coszen = get_cos_sza(area, start_time, data.dtype, row_slice)
        if max_sza is not None:
            with np.errstate(invalid="ignore"):
                coszen = np.where(coszen >= np.cos(np.deg2rad(max_sza)), coszen, np.nan)
    else:
        coszen = np.cos(np.deg2rad(sza_deg))
    return sunzen_corr_cos(data, coszen, limit=correction_limit, max_sza=max_sza)


def _li_shibata(cos_zen):
    return 24.35 / (2. * cos_zen + np.sqrt(498.5225 * cos_zen ** 2 + 1))


def atmospheric_path_length_correction(data, cos_zen, limit=88., max_sza=95.):
    """satpy/utils.py:278-316 (Li & Shibata 2006)."""
    limit_rad = np.deg2rad(limit)
    limit_cos = np.cos(limit_rad)
    max_sza_rad = np.deg2rad(max_sza) if max_sza is not None else max_sza
    with np.errstate(all="ignore"):
        corr = _li_shibata(cos_zen)
        corr_lim = _li_shibata(limit_cos)
        if max_sza is not None:
            grad_factor = (np.arccos(cos_zen) - limit_rad) / (max_sza_rad - limit_rad)
            grad_factor = 1. - np.log(grad_factor + 1) / np.log(2)
            grad_factor = grad_factor.clip(0.)
        else:
            grad_factor = 1.
        corr = np.where(cos_zen > limit_cos, corr, grad_factor * corr_lim)
        corr = np.where(~np.isnan(cos_zen), corr, 0)
        return data * corr


def sunzen_reduction(data, sunz, limit=55., max_sza=90., strength=1.5):
    """satpy/modifiers/angles.py:594-624."""
    with np.errstate(all="ignore"):
        reduction_factor = (sunz - limit) / (max_sza - limit)
        reduction_factor = reduction_factor.clip(0., 1.)
        reduction_factor = 1. - np.log2(reduction_factor + 1)
        reduction_factor = reduction_factor ** strength / (
            reduction_factor ** strength + (1 - reduction_factor) ** strength)
        corr = np.where(sunz < limit, 1.0, reduction_factor)
        corr[np.isnan(sunz)] = 0
        return data * corr


def as_datetime(t) -> dt.datetime:
    if isinstance(t, dt.datetime):
        return t
    return np.datetime64(t, "us").astype(dt.datetime)


# =====================================================================================================================
# Sensor / sun angles ("next" row of SURVEY.md section 8f): satpy/modifiers/angles.py:336-402,443-531.
# pyorbital (un-vendored, unpinned) supplies get_observer_look / observer_position / get_alt_az; restated from the
# library's published code (Celestrak "Orbital Coordinate Systems" parts II and III).  UNPINNED except
# compute_relative_azimuth (satpy/tests/modifier_tests/test_angles.py:361-382).
_F = 1 / 298.257223563   # WGS-84 flattening (pyorbital.astronomy.F)
_A = 6378.137            # WGS-84 equatorial radius [km] (pyorbital.astronomy.A)


def observer_position(utc_time, lon, lat, alt):
    """ECI position [km] of a point at geodetic lon/lat [deg] and altitude [km] (pyorbital ``observer_position``)."""
    lon = np.deg2rad(lon)
    lat = np.deg2rad(lat)
    theta = (gmst(utc_time) + lon) % (2 * np.pi)
    c = 1 / np.sqrt(1 + _F * (_F - 2) * np.sin(lat) ** 2)
    sq = c * (1 - _F) ** 2
    achcp = (_A * c + alt) * np.cos(lat)
    return achcp * np.cos(theta), achcp * np.sin(theta), (_A * sq + alt) * np.sin(lat)


def get_observer_look(sat_lon, sat_lat, sat_alt, utc_time, lon, lat, alt):
    """(azimuth, elevation) [deg] of the satellite seen from the ground (pyorbital ``orbital.get_observer_look``)."""
    pos_x, pos_y, pos_z = observer_position(utc_time, sat_lon, sat_lat, sat_alt)
    opos_x, opos_y, opos_z = observer_position(utc_time, lon, lat, alt)
    lon = np.deg2rad(lon)
    lat = np.deg2rad(lat)
    theta = (gmst(utc_time) + lon) % (2 * np.pi)
    rx, ry, rz = pos_x - opos_x, pos_y - opos_y, pos_z - opos_z
    sin_lat, cos_lat, sin_theta, cos_theta = np.sin(lat), np.cos(lat), np.sin(theta), np.cos(theta)
    top_s = sin_lat * cos_theta * rx + sin_lat * sin_theta * ry - cos_lat * rz
    top_e = -sin_theta * rx + cos_theta * ry
    top_z = cos_lat * cos_theta * rx + cos_lat * sin_theta * ry + sin_lat * rz
    az_ = np.mod(np.arctan2(-top_e, top_s) + np.pi, 2 * np.pi)
    rg_ = np.sqrt(rx * rx + ry * ry + rz * rz)
    with np.errstate(invalid="ignore"):
        el_ = np.arcsin((top_z / rg_).clip(max=1))
    return np.rad2deg(az_), np.rad2deg(el_)


def get_alt_az(utc_time, lon, lat):
    """Sun altitude and azimuth [rad] (pyorbital ``astronomy.get_alt_az``)."""
    lon = np.deg2rad(lon)
    lat = np.deg2rad(lat)
    ra_, dec = sun_ra_dec(utc_time)
    h__ = (gmst(utc_time) + lon) - ra_
    alt_ = np.arcsin(np.sin(lat) * np.sin(dec) + np.cos(lat) * np.cos(dec) * np.cos(h__))
    az_ = np.arctan2(-np.sin(h__), (np.cos(lat) * np.tan(dec) - np.sin(lat) * np.cos(h__)))
    return alt_, az_


def get_angles(area, start_time, sat_lon, sat_lat, sat_alt_m):
    """``get_angles`` (satpy/modifiers/angles.py:377-402): sensor azimuth / zenith, solar azimuth / zenith [deg], float64."""
    lons, lats = area.get_lonlats()
    with np.errstate(invalid="ignore"):
        lons = np.where(lons >= 1e30, np.nan, lons)
        lats = np.where(lats >= 1e30, np.nan, lats)
        sata, satel = get_observer_look(sat_lon, sat_lat, sat_alt_m / 1000.0, start_time, lons, lats, 0)
        suna = np.rad2deg(get_alt_az(start_time, lons, lats)[1]) % 360.
        sunz = np.rad2deg(np.arccos(cos_zen(start_time, lons, lats)))
    return sata, 90 - satel, suna, sunz


def compute_relative_azimuth(sat_azi, sun_azi):
    """satpy/modifiers/angles.py:371-374."""
    ssadiff = np.absolute(sun_azi - sat_azi)
    return np.minimum(ssadiff, sun_azi.dtype.type(360.0) - ssadiff)
