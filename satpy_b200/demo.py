"""Synthetic scenes of the BASELINE.json configurations (no files, no network).

The role ``satpy/demo/`` plays for the reference (ready-made inputs for examples, tests and
benchmarks), restricted to deterministic synthetic data: area definitions copied by VALUE from
satpy/etc/areas.yaml, realistic ABI calibration constants, and seeded count arrays.  Everything is
a plain dict / numpy array so that the oracle (tests, bench CPU arm) and the device path are fed
the same numbers.
"""
from __future__ import annotations

import datetime as dt
import math

import numpy as np

# ---------------------------------------------------------------------------- areas
# (proj dict, width, height, area_extent) -- outer pixel edges, as in satpy/etc/areas.yaml
_A_WGS84 = 6378137.0

AREAS = {
    # satpy/etc/areas.yaml:15-31
    "msg_seviri_fes_3km": ({"proj": "geos", "lon_0": 0.0, "a": 6378169.0, "b": 6356583.8, "h": 35785831.0},
                           3712, 3712, (-5570248.686685662, -5567248.28340708, 5567248.28340708, 5570248.686685662)),
    # satpy/etc/areas.yaml:551-568
    "goes_east_abi_f_2km": ({"proj": "geos", "sweep": "x", "lon_0": -75.0, "h": 35786023.0, "ellps": "GRS80"},
                            5424, 5424, (-5434894.885056, -5434894.885056, 5434894.885056, 5434894.885056)),
    # satpy/etc/areas.yaml:531-548
    "goes_east_abi_f_1km": ({"proj": "geos", "sweep": "x", "lon_0": -75.0, "h": 35786023.0, "ellps": "GRS80"},
                            10848, 10848, (-5434894.885056, -5434894.885056, 5434894.885056, 5434894.885056)),
    # satpy/etc/areas.yaml:287-302 (config 3 source); lon_0 = 140.7: the visible disk crosses the antimeridian
    "himawari_ahi_fes_1km": ({"proj": "geos", "lon_0": 140.7, "a": 6378137.0, "rf": 298.257024882273, "h": 35785863.0},
                             11000, 11000, (-5500000.0355, -5500000.0355, 5500000.0355, 5500000.0355)),
    # 1 km Lambert conformal conic over East Asia (config 3 target, builder-fixed: SURVEY.md section 8d)
    "east_asia_lcc_1km": ({"proj": "lcc", "lat_1": 30.0, "lat_2": 60.0, "lat_0": 38.0, "lon_0": 126.0, "ellps": "WGS84"},
                          6000, 6000, (-3.0e6, -3.0e6, 3.0e6, 3.0e6)),
    # "1 km EuroLAEA" (config 1): ETRS89-LAEA (EPSG:3035) parameters, 5000 x 4500 pixels of 1 km
    "euro_laea_1km": ({"proj": "laea", "lat_0": 52.0, "lon_0": 10.0, "x_0": 4321000.0, "y_0": 3210000.0,
                       "ellps": "GRS80"}, 5000, 4500, (2500000.0, 1400000.0, 7500000.0, 5900000.0)),
    # 2 km equirectangular over the GOES-East disk bounding box (config 2): lon -75 +- 81.3, lat +- 81.3
    "goes_east_eqc_2km": ({"proj": "eqc", "lon_0": 0.0, "ellps": "WGS84"}, 9050, 9050,
                          (math.radians(-75.0) * _A_WGS84 - 9050e3, -9050e3,
                           math.radians(-75.0) * _A_WGS84 + 9050e3, 9050e3)),
    # 0.5 km global equirectangular grid (config 5): 80150 x 40075
    "global_eqc_500m": ({"proj": "eqc", "lon_0": 0.0, "ellps": "WGS84"}, 80150, 40075,
                        (-math.pi * _A_WGS84, -math.pi * _A_WGS84 / 2, math.pi * _A_WGS84, math.pi * _A_WGS84 / 2)),
}


def area_params(name: str, scale: float = 1.0):
    """``(proj, width, height, extent)``; ``scale < 1`` keeps the extent and shrinks the pixel count
    (coarser pixels) so that the oracle can run the same geometry quickly."""
    proj, w, h, ext = AREAS[name]
    if scale != 1.0:
        w, h = max(2, int(round(w * scale))), max(2, int(round(h * scale)))
    return dict(proj), w, h, tuple(ext)


def make_area(name: str, scale: float = 1.0):
    from .geometry import AreaDefinition
    proj, w, h, ext = area_params(name, scale)
    return AreaDefinition(name, name, name, proj, w, h, ext)


# ------------------------------------------------------------------- ABI calibration constants
# Representative GOES-16 L1b header values (builder-fixed and recorded here).
ABI_BANDS = {
    "C01": dict(scale_factor=0.8121064, add_offset=-25.936647, fill_value=1023, unsigned=True, max_count=1022,
                esun=2017.1648, earth_sun_distance_anomaly_in_AU=0.99447),
    "C02": dict(scale_factor=0.15868431, add_offset=-20.289911, fill_value=4095, unsigned=True, max_count=4094,
                esun=1631.3351, earth_sun_distance_anomaly_in_AU=0.99447),
    "C03": dict(scale_factor=0.37691253, add_offset=-20.289911, fill_value=1023, unsigned=True, max_count=1022,
                esun=957.0699, earth_sun_distance_anomaly_in_AU=0.99447),
    "C13": dict(scale_factor=0.04572892, add_offset=-1.6443, fill_value=4095, unsigned=True, max_count=4094,
                planck_fk1=10803.3, planck_fk2=1392.74, planck_bc1=0.07550, planck_bc2=0.99975),
}
START_TIME = dt.datetime(2019, 3, 14, 18, 0, 0)  # configs 2 and 5


def abi_counts(shape, seed: int, band: str = "C02", fill_fraction: float = 0.005) -> np.ndarray:
    """Seeded int16 count image: uniform over the band's count range, ``fill_fraction`` set to _FillValue."""
    rng = np.random.default_rng(seed)
    b = ABI_BANDS[band]
    counts = rng.integers(0, b["max_count"] + 1, size=shape, dtype=np.int16)
    if fill_fraction > 0:
        n = int(counts.size * fill_fraction)
        idx = rng.integers(0, counts.size, size=n)
        counts.reshape(-1)[idx] = np.int16(b["fill_value"])
    return counts


def abi_calibration(band: str, **overrides):
    from .readers.abi import ABICalibration
    kw = {k: v for k, v in ABI_BANDS[band].items() if k != "max_count"}
    kw.update(overrides)
    return ABICalibration(**kw)


def seviri_bt(shape, seed: int = 1) -> np.ndarray:
    """Config 1 input: float32 brightness temperatures uniform in [200, 320) K."""
    rng = np.random.default_rng(seed)
    return rng.uniform(200.0, 320.0, size=shape).astype(np.float32)


def viirs_like_swath(nscans=12, rows_per_scan=16, ncols=400, dtype=np.float32, bowtie=0.0, lat0=70.0, lat_span=6.0,
                     lon_half_width=9.0):
    """Config 4 stand-in: a smooth high-latitude swath of ``nscans * rows_per_scan`` rows (M-band geometry in
    miniature: 16 rows per scan); ``bowtie > 0`` makes consecutive scans overlap towards the swath edges.
    ``lat0=28, lat_span=51.8, lon_half_width=10.79`` with 480 scans x 3200 columns gives the real 750 m M-band
    sampling of ten granules (scripts/kernel_bench.py)."""
    nrows = nscans * rows_per_scan
    scan = (np.arange(nrows) // rows_per_scan)[:, None].astype(np.float64)
    within = ((np.arange(nrows) % rows_per_scan)[:, None] - (rows_per_scan - 1) / 2.0) / rows_per_scan
    xt = np.linspace(-1.0, 1.0, ncols)[None, :]
    at = (scan + 0.5 + within * (1.0 + bowtie * xt * xt)) / nscans
    lat = lat0 + lat_span * at + 0.25 * xt
    lon = -20.0 + lon_half_width * xt / np.cos(np.deg2rad(lat)) + 4.0 * at
    return lon.astype(dtype), lat.astype(dtype), rows_per_scan
