"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Run with ``-m gpu`` on a B200.

Bars: bit-exact for integer / index work (equidistant ties counted and bounded), <= 1e-5 relative for
float32 values (north_star), exact equality for the reference's golden vectors where numpy is exact.
"""
import datetime as dt

import numpy as np
import pytest

from oracle import calibration as ocal
from oracle import nearest as onn
from oracle import solar as osolar
from oracle.projections import Area as OArea
from oracle.projections import Swath as OSwath
from satpy_b200 import AreaDefinition, DataArrayLite, SwathDefinition, demo
from satpy_b200.modifiers import EffectiveSolarPathLengthCorrector, IncompatibleAreas, SunZenithCorrector, cos_sza
from satpy_b200.modifiers import sunz_correct
from satpy_b200.readers import ABICalibration, calibrate_abi
from satpy_b200.resample import B200NearestResampler

pytestmark = pytest.mark.gpu

REL = 1e-5  # north_star: resampled and calibrated float32 values within 1e-5 relative


def both_areas(name, scale=1.0):
    params = demo.area_params(name, scale)
    return demo.make_area(name, scale), OArea(*params)


# ---------------------------------------------------------------------------------- calibration
def test_abi_ir_golden_vector():
    """satpy/tests/reader_tests/test_abi_l1b.py:207-225,391-414."""
    shape = (750, 1250)
    rad = (np.arange(shape[0] * shape[1]).reshape(shape) + 1.0) * 50.0
    rad[0, 0] = -0.0001
    counts = ((rad + 1.3) / 0.5).astype(np.int16)
    fill = np.int16(np.floor(((9 + 1) * 50.0 + 1.3) / 0.5))
    for clip, first in ((False, np.nan), (True, 134.68753)):
        cal = ABICalibration(scale_factor=0.5, add_offset=-1.3, fill_value=fill, planck_fk1=13432.1,
                             planck_fk2=1497.61, planck_bc1=0.09102, planck_bc2=0.99971,
                             clip_negative_radiances=clip)
        res = calibrate_abi(counts, cal, "brightness_temperature")
        assert res.dtype == np.float32
        expected = np.array([first, 304.97037, 332.22778, 354.6147, 374.08688, 391.58655, 407.64786, 422.60635,
                             436.68802, np.nan])
        np.testing.assert_allclose(res[0, :10], expected, equal_nan=True, atol=1e-4)
        ref = ocal.abi_calibrate(counts, "brightness_temperature", 0.5, -1.3, fill, fk1=13432.1, fk2=1497.61,
                                 bc1=0.09102, bc2=0.99971, clip_negative_radiances=clip)
        np.testing.assert_allclose(res, ref, rtol=REL, equal_nan=True)


def test_abi_vis_golden_vector():
    """satpy/tests/reader_tests/test_abi_l1b.py:48-70,423-443."""
    shape = (1500, 2500)
    rad = (np.arange(shape[0] * shape[1]).reshape(shape) + 1.0) * 50.0
    counts = ((rad + 1.0) / 0.5).astype(np.int16)
    cal = ABICalibration(scale_factor=0.5, add_offset=-1.0, fill_value=1002, esun=2017,
                         earth_sun_distance_anomaly_in_AU=0.99)
    res = calibrate_abi(counts, cal, "reflectance")
    expected = np.array([7.632808, 15.265616, 22.898426, 30.531233, 38.164043, 45.796852, 53.429657, 61.062466,
                         68.695274, np.nan])
    np.testing.assert_allclose(res[0, :10], expected, equal_nan=True)
    ref = ocal.abi_calibrate(counts, "reflectance", 0.5, -1.0, 1002, esun=2017, esd=0.99)
    np.testing.assert_array_equal(res, ref)  # float32 mul/add in the reference's order: bit-exact
    rad_dev = calibrate_abi(counts, cal, "radiance")
    np.testing.assert_array_equal(rad_dev, ocal.abi_calibrate(counts, "radiance", 0.5, -1.0, 1002))


@pytest.mark.parametrize("band,calibration", [("C01", "reflectance"), ("C02", "reflectance"), ("C02", "radiance"),
                                              ("C13", "brightness_temperature")])
@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 4099, 1 << 20])
def test_abi_calibrate_sizes(band, calibration, n):
    """Ragged sizes exercise the vector body and the scalar tail; unsigned counts with fill values."""
    counts = demo.abi_counts((n,), seed=11 + n % 7, band=band, fill_fraction=0.01)
    b = demo.ABI_BANDS[band]
    cal = demo.abi_calibration(band)
    res = calibrate_abi(counts, cal, calibration)
    kw = {k: b[k] for k in ("esun",) if k in b}
    if "earth_sun_distance_anomaly_in_AU" in b:
        kw["esd"] = b["earth_sun_distance_anomaly_in_AU"]
    for k in ("fk1", "fk2", "bc1", "bc2"):
        if "planck_" + k in b:
            kw[k] = b["planck_" + k]
    ref = ocal.abi_calibrate(counts, calibration, b["scale_factor"], b["add_offset"], b["fill_value"], b["unsigned"],
                             **kw)
    assert res.shape == ref.shape and res.dtype == np.float32
    if calibration == "brightness_temperature":
        np.testing.assert_allclose(res, ref, rtol=REL, equal_nan=True)
    else:
        np.testing.assert_array_equal(res, ref)


def test_abi_calibrate_errors():
    cal = demo.abi_calibration("C02")
    with pytest.raises(ValueError):
        calibrate_abi(np.zeros(4, np.int16), cal, "albedo")
    with pytest.raises(TypeError):
        calibrate_abi(np.zeros(4, np.float32), cal, "radiance")
    with pytest.raises(ValueError):
        calibrate_abi(np.zeros(4, np.int16), demo.abi_calibration("C13"), "reflectance")


# ------------------------------------------------------------------------------------ geometry
PROJ_CASES = [
    ("geos_y", {"proj": "geos", "lon_0": 0.0, "a": 6378169.0, "b": 6356583.8, "h": 35785831.0}, 120, 120,
     (-5570248.68, -5567248.28, 5567248.28, 5570248.68)),
    ("geos_x", {"proj": "geos", "sweep": "x", "lon_0": -75.0, "h": 35786023.0, "ellps": "GRS80"}, 130, 110,
     (-5434894.885056, -5434894.885056, 5434894.885056, 5434894.885056)),
    ("eqc", {"proj": "eqc", "lon_0": 10.0, "ellps": "WGS84"}, 100, 80, (-2.0e7, -1.0e7, 2.0e7, 1.0e7)),
    ("merc", {"proj": "merc"}, 90, 70, (-1.5e7, -1.2e7, 1.5e7, 1.2e7)),
    ("merc_ts_sphere", {"proj": "merc", "lat_ts": 30.0, "R": 6371000.0}, 64, 64, (-1.0e7, -8.0e6, 1.0e7, 8.0e6)),
    ("laea_ob", {"proj": "laea", "lat_0": 52.0, "lon_0": 10.0, "x_0": 4321000.0, "y_0": 3210000.0, "ellps": "GRS80"},
     100, 90, (2500000.0, 1400000.0, 7500000.0, 5900000.0)),
    ("laea_n", {"proj": "laea", "lat_0": 90.0, "lon_0": 0.0, "ellps": "WGS84"}, 80, 80, (-5e6, -5e6, 5e6, 5e6)),
    ("laea_s_sphere", {"proj": "laea", "lat_0": -90.0, "lon_0": 20.0, "R": 6371228.0}, 80, 80, (-5e6, -5e6, 5e6, 5e6)),
    ("laea_eq", {"proj": "laea", "lat_0": 0.0, "lon_0": -30.0, "ellps": "WGS84"}, 80, 80, (-6e6, -6e6, 6e6, 6e6)),
    ("lcc", {"proj": "lcc", "lat_1": 30.0, "lat_2": 60.0, "lat_0": 38.0, "lon_0": 126.0, "ellps": "WGS84"}, 100, 100,
     (-3.0e6, -3.0e6, 3.0e6, 3.0e6)),
    ("lcc_1sp", {"proj": "lcc", "lat_1": 25.0, "lat_0": 25.0, "lon_0": -95.0, "datum": "WGS84"}, 60, 60,
     (-2.0e6, -2.0e6, 2.0e6, 2.0e6)),
    ("stere_n", {"proj": "stere", "lat_0": 90.0, "lat_ts": 60.0, "lon_0": 0.0, "ellps": "WGS84"}, 100, 100,
     (-4.0e6, -4.0e6, 4.0e6, 4.0e6)),
    ("stere_s", {"proj": "stere", "lat_0": -90.0, "lat_ts": -71.0, "lon_0": 0.0, "ellps": "WGS84"}, 100, 100,
     (-4.0e6, -4.0e6, 4.0e6, 4.0e6)),
    ("longlat", {"proj": "longlat", "datum": "WGS84"}, 72, 36, (-180.0, -90.0, 180.0, 90.0)),
]


@pytest.mark.parametrize("name,proj,w,h,ext", PROJ_CASES, ids=[c[0] for c in PROJ_CASES])
def test_area_lonlats_match_oracle(name, proj, w, h, ext):
    """Device inverse projections vs the oracle (float64): < 1e-9 degrees, same off-earth pattern."""
    area = AreaDefinition(name, name, name, proj, w, h, ext)
    lons, lats = area.get_lonlats()
    olons, olats = OArea(proj, w, h, ext).get_lonlats()
    assert lons.shape == (h, w)
    fin = np.isfinite(olons)
    np.testing.assert_array_equal(np.isfinite(lons), fin)
    dlon = np.abs(lons[fin] - olons[fin])
    dlon = np.minimum(dlon, 360.0 - dlon)
    assert dlon.max() < 1e-9
    assert np.abs(lats[fin] - olats[fin]).max() < 1e-9
    # row windows address the same pixels
    lons_w, lats_w = area.get_lonlats(slice(3, 11))
    np.testing.assert_array_equal(lons_w, lons[3:11])
    np.testing.assert_array_equal(lats_w, lats[3:11])


# ---------------------------------------------------------------------------------- sun zenith
def _sunz_area():
    return AreaDefinition("test", "test", "test", {"proj": "merc"}, 2, 2, (-2000, -2000, 2000, 2000))


def test_cos_sza_golden():
    """cos(SZA) fixture of satpy/tests/test_modifiers.py:104."""
    cz = cos_sza(_sunz_area(), dt.datetime(2018, 1, 1, 18), lonlat_f32=False)
    np.testing.assert_allclose(cz, [[0.0149581333, 0.0146694376], [0.0150812684, 0.0147925727]], rtol=2e-8)


def test_sunz_corrector_golden_not_provided():
    """satpy/tests/test_modifiers.py:116-146 (float32 branch)."""
    attrs = {"area": _sunz_area(), "start_time": dt.datetime(2018, 1, 1, 18), "modifiers": tuple(), "name": "test_vis"}
    ds = DataArrayLite(np.ones((2, 2), dtype=np.float32), ("y", "x"), attrs)
    comp = SunZenithCorrector(name="sza_test", modifiers=tuple())
    res = comp((ds,), test_attr="test")
    assert res.dtype == np.float32
    np.testing.assert_allclose(res.values, np.array([[22.401667, 22.31777], [22.437503, 22.353533]]), rtol=1e-6)
    # apply_modifier_info only fills keys the data does not have yet (satpy/composites/core.py:99-125)
    assert res.attrs["area"] is attrs["area"] and res.attrs["name"] == "test_vis"
    assert res.attrs["modifiers"] == tuple()
    comp = SunZenithCorrector(name="sza_test", modifiers=tuple(), correction_limit=90)
    res = comp((ds,), test_attr="test")
    np.testing.assert_allclose(res.values, np.array([[66.853262, 68.168939], [66.30742, 67.601493]]), rtol=1e-5)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_sunz_corrector_golden_both_dtypes(dtype):
    """The reference runs these cases for float32 AND float64 data (satpy/tests/test_modifiers.py:116-176)."""
    area = _sunz_area()
    attrs = {"area": area, "start_time": dt.datetime(2018, 1, 1, 18), "modifiers": tuple(), "name": "test_vis"}
    ds = DataArrayLite(np.ones((2, 2), dtype=dtype), ("y", "x"), attrs)
    sza = DataArrayLite(np.rad2deg(np.arccos(np.array([[0.0149581333, 0.0146694376], [0.0150812684, 0.0147925727]])))
                        .astype(dtype), ("y", "x"), {"area": area})
    for args in ((ds,), (ds, sza)):
        res = SunZenithCorrector(name="sza_test", modifiers=tuple())(args, test_attr="test")
        assert res.dtype == dtype and res.values.dtype == dtype
        np.testing.assert_allclose(res.values, np.array([[22.401667, 22.31777], [22.437503, 22.353533]], dtype=dtype),
                                   rtol=2e-6)
        res = SunZenithCorrector(name="sza_test", modifiers=tuple(), correction_limit=90)(args, test_attr="test")
        np.testing.assert_allclose(res.values, np.array([[66.853262, 68.168939], [66.30742, 67.601493]], dtype=dtype),
                                   rtol=1e-5)


def test_sunz_correct_float64_matches_oracle():
    area, oarea = both_areas("goes_east_abi_f_2km", 500 / 5424)
    when = dt.datetime(2019, 3, 14, 23, 30)
    data = np.random.default_rng(5).uniform(0, 120, size=area.shape)
    res = sunz_correct(data, area=area, start_time=when)
    ref = osolar.sun_zenith_corrector(data, oarea, when)
    assert res.dtype == np.float64
    np.testing.assert_array_equal(np.isnan(res), np.isnan(ref))
    np.testing.assert_allclose(res, ref, rtol=1e-9, atol=1e-9)   # float64 throughout: no float32 trig in the way


def test_sunz_corrector_golden_provided():
    """satpy/tests/test_modifiers.py:148-176: SZA given as an optional dataset."""
    area = _sunz_area()
    attrs = {"area": area, "start_time": dt.datetime(2018, 1, 1, 18), "modifiers": tuple(), "name": "test_vis"}
    ds = DataArrayLite(np.ones((2, 2), dtype=np.float32), ("y", "x"), attrs)
    sza = np.rad2deg(np.arccos(np.array([[0.0149581333, 0.0146694376], [0.0150812684, 0.0147925727]])))
    sza = DataArrayLite(sza.astype(np.float32), ("y", "x"), {"area": area})
    res = SunZenithCorrector(name="sza_test", modifiers=tuple())((ds, sza), test_attr="test")
    np.testing.assert_allclose(res.values, np.array([[22.401667, 22.31777], [22.437503, 22.353533]], np.float32),
                               rtol=1e-5)
    res = SunZenithCorrector(name="sza_test", modifiers=tuple(), correction_limit=90)((ds, sza))
    np.testing.assert_allclose(res.values, np.array([[66.853262, 68.168939], [66.30742, 67.601493]]), rtol=1e-5)
    # already corrected -> returned untouched (satpy/modifiers/geometry.py:52-54)
    ds2 = DataArrayLite(np.ones((2, 2), dtype=np.float32), ("y", "x"), dict(attrs, sunz_corrected=True))
    assert SunZenithCorrector(name="x")((ds2,)) is ds2


def test_sunz_incompatible_areas():
    """satpy/tests/test_modifiers.py:178-184."""
    big = AreaDefinition("test", "test", "test", {"proj": "merc"}, 4, 4, (-2000, -2000, 2000, 2000))
    ds2 = DataArrayLite(np.ones((4, 4), dtype=np.float32), ("y", "x"),
                        {"area": big, "start_time": dt.datetime(2018, 1, 1, 18)})
    sza = DataArrayLite(np.ones((2, 2), dtype=np.float32), ("y", "x"), {"area": _sunz_area()})
    with pytest.raises(IncompatibleAreas):
        SunZenithCorrector(name="sza_test", modifiers=tuple())((ds2, sza))


def _assert_sunz_close(res, ref, data=None, cz=None, max_sza=None, limit=88.0):
    """1e-5 relative (north_star).  In the last 0.25 degree before ``max_sza`` the log fall-off drives the corrected
    value to zero and the reference itself is ill-conditioned there: numpy's float32 sin/cos are not correctly
    rounded (1 ulp differences on 24 % of the pixels), which alone moves the reference by up to 3e-4 relative in
    that sliver.  There the bar is absolute: 1e-5 of the correction at the limit, |data| / cos(limit)."""
    assert res.shape == ref.shape
    np.testing.assert_array_equal(np.isnan(res), np.isnan(ref))
    ok = ~np.isnan(ref)
    err = np.abs(res.astype(np.float64) - ref)
    tol = REL * np.abs(ref) + 1e-6
    if cz is not None and max_sza is not None:
        with np.errstate(invalid="ignore"):
            sliver = cz < np.cos(np.deg2rad(max_sza - 0.25))
        tol = np.where(sliver, REL * np.abs(data) / np.cos(np.deg2rad(limit)) + 1e-6, tol)
    bad = ok & (err > tol)
    # how many pixels only pass because of the absolute bar of the sliver (the tolerance north_star does not grant)
    needed_abs = int((ok & (err > REL * np.abs(ref) + 1e-6) & ~bad).sum())
    assert not bad.any(), (f"{bad.sum()} pixels outside tolerance, worst {np.nanmax(np.where(ok, err / tol, 0)):.2f} x tol; "
                           f"{needed_abs} of {int(ok.sum())} pixels needed the absolute bar of the last 0.25 degree")
    return needed_abs


@pytest.mark.parametrize("areaname,scale,when", [
    ("goes_east_abi_f_2km", 700 / 5424, dt.datetime(2019, 3, 14, 18, 0)),     # day
    ("goes_east_abi_f_2km", 700 / 5424, dt.datetime(2019, 3, 14, 23, 30)),    # terminator on the disk
    ("msg_seviri_fes_3km", 600 / 3712, dt.datetime(2020, 6, 21, 5, 0)),
])
@pytest.mark.parametrize("max_sza,limit", [(95.0, 88.0), (None, 88.0), (90.0, 80.0)])
def test_sunz_correct_matches_oracle(areaname, scale, when, max_sza, limit):
    area, oarea = both_areas(areaname, scale)
    rng = np.random.default_rng(5)
    data = rng.uniform(0, 120, size=area.shape).astype(np.float32)
    data[rng.random(area.shape) < 0.01] = np.nan
    res = sunz_correct(data, area=area, start_time=when, correction_limit=limit, max_sza=max_sza)
    ref = osolar.sun_zenith_corrector(data, oarea, when, correction_limit=limit, max_sza=max_sza)
    cz = osolar.get_cos_sza(oarea, when, np.float32)
    assert res.dtype == np.float32
    needed_abs = _assert_sunz_close(res, ref, data, cz, max_sza, limit)
    print(f"sun-zenith parity: {needed_abs} of {res.size} pixels needed the absolute bar of the last 0.25 degree")
    # multi-band input shares the factor
    res3 = sunz_correct(np.stack([data, data * 2]), area=area, start_time=when, correction_limit=limit,
                        max_sza=max_sza)
    np.testing.assert_array_equal(res3[0], res)
    # a row window gives the same pixels (row-tiled multi-GPU use)
    resw = sunz_correct(data[100:164], area=area, start_time=when, correction_limit=limit, max_sza=max_sza, row0=100)
    np.testing.assert_array_equal(resw, res[100:164])


def test_effective_path_length_matches_oracle():
    area, oarea = both_areas("goes_east_abi_f_2km", 500 / 5424)
    when = dt.datetime(2019, 3, 14, 22, 0)
    data = np.random.default_rng(6).uniform(0, 100, size=area.shape).astype(np.float32)
    ds = DataArrayLite(data, ("y", "x"), {"area": area, "start_time": when, "name": "C02"})
    res = EffectiveSolarPathLengthCorrector(name="eff")((ds,)).values
    cz = osolar.get_cos_sza(oarea, when, np.float32)
    with np.errstate(invalid="ignore"):
        cz = np.where(cz >= np.cos(np.deg2rad(95.0)), cz, np.nan)
    ref = osolar.atmospheric_path_length_correction(data, cz).astype(np.float32)
    _assert_sunz_close(res, ref, data, cz, 95.0, 88.0)


def test_fused_abi_vis_sunz_matches_oracle():
    from satpy_b200.pipeline import ABIResamplePipeline
    import torch
    scale = 900 / 10848
    area, oarea = both_areas("goes_east_abi_f_1km", scale)
    counts = demo.abi_counts(area.shape, seed=7, band="C02")
    pipe = ABIResamplePipeline(area, demo.make_area("global_eqc_500m", 64 / 80150), 5e4)
    res = pipe.calibrate_sunz(torch.from_numpy(counts).cuda(), demo.abi_calibration("C02"), demo.START_TIME)
    res = res.cpu().numpy()
    from oracle import reference_pipeline as rp
    ref = rp.abi_sunz(counts, demo.ABI_BANDS["C02"], oarea, demo.START_TIME)
    refl = rp.abi_reflectance(counts, demo.ABI_BANDS["C02"])
    _assert_sunz_close(res, ref, refl, osolar.get_cos_sza(oarea, demo.START_TIME, np.float32), 95.0, 88.0)


# ------------------------------------------------------------------------------------- nearest
def _check_indices(resampler, o_src, o_dst, radius, mask=None, max_tie_frac=2e-3):
    info = resampler.neighbour_info()
    slon, slat = o_src.get_lonlats()
    dlon, dlat = o_dst.get_lonlats()
    vin, vout, idx = onn.get_neighbour_info(slon, slat, dlon, dlat, radius, mask=mask)
    n_valid = int(vin.sum())
    np.testing.assert_array_equal(info["valid_input_index"], vin)
    np.testing.assert_array_equal(info["valid_output_index"], vout)
    ref = onn.to_minus_one(idx, n_valid)[..., 0]
    got = info["index_array"][..., 0]
    assert got.shape == ref.shape and got.dtype == np.int32
    n_mis, n_tie, n_bad = onn.compare_indices(got, ref, slon, slat, dlon, dlat, vin)
    assert n_bad == 0, f"{n_bad} index mismatches that are not equidistant ties (of {n_mis})"
    assert n_tie <= max_tie_frac * got.size + 8, f"{n_tie} ties of {got.size}"
    return vin, ref, got, n_tie


NN_KEYS = ("valid_input_index", "valid_output_index", "index_array")

NN_CASES = [
    # (source area, scale, target area, scale, radius [m])
    ("msg_seviri_fes_3km", 300 / 3712, "euro_laea_1km", 400 / 5000, 50000.0),     # config 1 geometry
    ("goes_east_abi_f_2km", 400 / 5424, "goes_east_eqc_2km", 500 / 9050, 40000.0),  # config 2 geometry
    ("goes_east_abi_f_1km", 500 / 10848, "global_eqc_500m", 900 / 80150, 30000.0),  # config 5 geometry
    ("goes_east_abi_f_1km", 300 / 10848, "global_eqc_500m", 300 / 80150, 5000.0),   # radius << pixel: many misses
]


@pytest.mark.parametrize("method", ["bucket", "fixed_grid"])
@pytest.mark.parametrize("sname,sscale,dname,dscale,radius", NN_CASES)
def test_nearest_indices_match_oracle(sname, sscale, dname, dscale, radius, method):
    src, o_src = both_areas(sname, sscale)
    dst, o_dst = both_areas(dname, dscale)
    r = B200NearestResampler(src, dst)
    r.precompute(radius_of_influence=radius, method=method)
    assert r._current["method"] == method
    vin, ref, got, n_tie = _check_indices(r, o_src, o_dst, radius)
    # gather parity for several dtypes, including multi-band
    rng = np.random.default_rng(3)
    for dtype, fill in ((np.float32, np.nan), (np.float64, np.nan), (np.int16, -999), (np.uint8, 255)):
        data = (rng.uniform(0, 200, size=(2,) + src.shape)).astype(dtype)
        out = r.compute(data, fill_value=fill)
        assert out.dtype == dtype and out.shape == (2,) + dst.shape
        idx3 = np.where(got < 0, int(vin.sum()), got)[..., None]
        exp = onn.get_sample_from_neighbour_info(data, vin, idx3, fill)
        np.testing.assert_array_equal(out, exp)


@pytest.mark.parametrize("dname,dscale", [("global_eqc_500m", 1500 / 80150), ("euro_laea_1km", 700 / 5000)])
@pytest.mark.parametrize("radius", [3000.0, 12000.0, 60000.0])
def test_nearest_engines_agree_exactly(dname, dscale, radius):
    """The bucket grid and the fixed-grid ring search are two exact searches with one tie rule: identical output,
    from sub-pixel radii (every target a miss or a single candidate) to radii of several source pixels."""
    src, _ = both_areas("goes_east_abi_f_2km", 900 / 5424)   # 12 km pixels
    dst, _ = both_areas(dname, dscale)
    out = {}
    for method in ("bucket", "fixed_grid", "fixed_grid_strict"):
        r = B200NearestResampler(src, dst)
        r.precompute(radius_of_influence=radius, method=method)
        out[method] = r.neighbour_info()
    for k in NN_KEYS:
        np.testing.assert_array_equal(out["bucket"][k], out["fixed_grid"][k])
        np.testing.assert_array_equal(out["bucket"][k], out["fixed_grid_strict"][k])
    hit = (out["bucket"]["index_array"] >= 0).mean()
    assert 0.0 < hit < 1.0


@pytest.mark.parametrize("seed", range(10))
def test_engines_agree_on_random_limb_crossing_cases(seed):
    """The lattice-model pruning of ``fixed_grid`` rests on empirically fitted margins, so it is held to the strict mode
    and to the bucket engine on randomised cases: both sweeps (ABI x, SEVIRI y), source pixels of 8-40 km, radii from 1
    to 50 source pixels, random masks plus a masked rectangle, source row windows, and eqc / laea target patches
    placed across the limb."""
    import math
    import torch
    rng = np.random.default_rng(100 + seed)
    sname, lon0 = (("goes_east_abi_f_2km", -75.0) if seed % 2 == 0 else ("msg_seviri_fes_3km", 0.0))
    nsrc = int(rng.integers(280, 700))
    src = demo.make_area(sname, nsrc / demo.AREAS[sname][1])
    pix = (src.area_extent[2] - src.area_extent[0]) / src.width
    if rng.random() < 0.4:                               # a band of source rows (what a multi-GPU block sees)
        r0 = int(rng.integers(0, src.height // 3))
        src = src.row_window(r0, int(rng.integers(src.height // 3, src.height - r0)))
    radius = float(pix * rng.choice([1.0, 1.7, 3.0, 6.0, 12.0, 25.0, 50.0]))
    clon = lon0 + float(rng.choice([-1, 1])) * float(rng.uniform(45.0, 84.0))     # across the east / west limb
    clat = float(rng.uniform(-75.0, 75.0))
    tpix = pix * float(rng.uniform(0.3, 1.5))
    w, h = int(rng.integers(150, 380)), int(rng.integers(120, 300))
    if seed % 3 == 0:
        proj = {"proj": "laea", "lat_0": clat, "lon_0": clon, "ellps": "WGS84"}
        ext = (-w * tpix / 2, -h * tpix / 2, w * tpix / 2, h * tpix / 2)
    else:
        a = 6378137.0
        x0, y0 = math.radians(clon) * a, math.radians(clat) * a
        proj = {"proj": "eqc", "lon_0": 0.0, "ellps": "WGS84"}
        ext = (x0 - w * tpix / 2, y0 - h * tpix / 2, x0 + w * tpix / 2, y0 + h * tpix / 2)
    dst = AreaDefinition("t", "t", "t", proj, w, h, ext)
    mask = rng.random(src.shape) < 0.05
    mr, mc = int(rng.integers(0, src.height - 20)), int(rng.integers(0, src.width - 20))
    mask[mr:mr + 20, mc:mc + 20] = True
    out = {}
    for method in ("bucket", "fixed_grid", "fixed_grid_strict"):
        r = B200NearestResampler(src, dst)
        r.precompute(radius_of_influence=radius, method=method, mask=mask)
        out[method] = r._current["gather_index"].clone()
        np.testing.assert_array_equal(r._current["valid_input_index"].cpu().numpy().astype(bool) & mask, False)
    assert bool(torch.equal(out["fixed_grid"], out["fixed_grid_strict"])), "pruned != strict"
    assert bool(torch.equal(out["bucket"], out["fixed_grid_strict"])), "bucket != fixed grid"


def test_fixed_grid_pruning_is_exact_on_the_full_disk():
    """The lattice-model pruning of the fixed-grid search against its strict mode over a whole (coarse) full disk seen
    from a fine global grid -- nadir, mid-latitudes and the limb, where pixels stretch to many times their nadir size --
    and against SEVIRI's sweep-y geometry."""
    for sname, sscale, dname, dscale, radius in (("goes_east_abi_f_1km", 2000 / 10848, "global_eqc_500m", 6000 / 80150, 11e3),
                                                 ("msg_seviri_fes_3km", 1500 / 3712, "global_eqc_500m", 5000 / 80150, 16e3)):
        src, _ = both_areas(sname, sscale)
        dst, _ = both_areas(dname, dscale)
        out = {}
        for method in ("fixed_grid", "fixed_grid_strict"):
            r = B200NearestResampler(src, dst)
            r.precompute(radius_of_influence=radius, method=method)
            out[method] = r._current["gather_index"].clone()
        assert bool((out["fixed_grid"] == out["fixed_grid_strict"]).all())
        assert 0.2 < float((out["fixed_grid"] >= 0).float().mean()) < 0.6


def test_fixed_grid_rejects_non_geos_sources():
    src, _ = both_areas("euro_laea_1km", 100 / 5000)
    dst, _ = both_areas("global_eqc_500m", 100 / 80150)
    with pytest.raises(NotImplementedError):
        B200NearestResampler(src, dst).precompute(radius_of_influence=1e5, method="fixed_grid")


def test_nearest_fused_equals_two_step():
    src, _ = both_areas("goes_east_abi_f_1km", 400 / 10848)
    dst, _ = both_areas("global_eqc_500m", 700 / 80150)
    data = np.random.default_rng(1).uniform(0, 100, size=src.shape).astype(np.float32)
    r = B200NearestResampler(src, dst)
    two = r.resample(data, radius_of_influence=40000.0)
    fused = r.resample_fused(data, 40000.0)
    np.testing.assert_array_equal(two, fused)
    # row windows tile the target exactly
    parts = [r.resample_fused(data, 40000.0, row_window=(r0, n)) for r0, n in ((0, 100), (100, 150), (250, dst.height - 250))]
    np.testing.assert_array_equal(np.concatenate(parts), two)
    # cached neighbour info is reused and can be re-installed from the reference-layout arrays
    info = r.neighbour_info()
    r2 = B200NearestResampler(src, dst)
    r2.load_neighbour_info(info["valid_input_index"], info["index_array"].astype(np.int64))
    np.testing.assert_array_equal(r2.compute(data), two)


def test_nearest_swath_golden_2x2():
    """The only un-mocked nearest case of the reference: satpy/tests/test_resample.py:109-129."""
    lons = np.arange(4).reshape((2, 2))
    source = SwathDefinition(lons, lons)
    dest = SwathDefinition(lons + .0001, lons + .0001)
    expected_gap = np.array([[1, 2], [3, 255]])
    data = DataArrayLite(expected_gap, ("y", "x"), {"_FillValue": 255, "area": source})
    res = B200NearestResampler(source, dest).resample(data, fill_value=255)
    assert res.dtype == expected_gap.dtype
    assert np.all(res.values == expected_gap)
    res = B200NearestResampler(source, dest).resample(data, fill_value=255, radius_of_influence=1000000)
    assert np.all(res.values == np.array([[1, 2], [3, 3]]))
    assert res.attrs["area"] is dest


def test_nearest_swath_source_with_mask_and_dateline():
    """Swath source crossing the antimeridian and the pole region, NaN data masked like BaseResampler.resample."""
    rng = np.random.default_rng(9)
    h, w = 220, 180
    lat = np.linspace(60, 89.5, h)[:, None] + rng.uniform(-0.02, 0.02, (h, w))
    lon = ((np.linspace(150, 215, w)[None, :] + rng.uniform(-0.02, 0.02, (h, w)) + 180) % 360) - 180
    data = rng.uniform(0, 1, (h, w)).astype(np.float32)
    data[rng.random((h, w)) < 0.05] = np.nan
    src = SwathDefinition(lon, lat)
    dst, o_dst = both_areas("euro_laea_1km", 1.0)
    dst = AreaDefinition("ps", "ps", "ps", {"proj": "stere", "lat_0": 90.0, "lat_ts": 60.0, "lon_0": 180.0,
                                            "ellps": "WGS84"}, 200, 200, (-2.5e6, -2.5e6, 2.5e6, 2.5e6))
    o_dst = OArea(dst.proj_dict, 200, 200, dst.area_extent)
    radius = 30000.0
    r = B200NearestResampler(src, dst)
    out = r.resample(data, radius_of_influence=radius)
    mask = np.isnan(data)
    vin, ref, got, _ = _check_indices(r, OSwath(lon, lat), o_dst, radius, mask=mask)
    idx3 = np.where(ref < 0, int(vin.sum()), ref)[..., None]
    exp = onn.get_sample_from_neighbour_info(data, vin, idx3, np.float32(np.nan))
    same = (got == ref)
    np.testing.assert_array_equal(out[same], exp[same])
    assert np.isfinite(out).sum() > 1000


def test_nearest_float32_swath_and_empty_results():
    lon = np.linspace(-10, 10, 50, dtype=np.float32)[None, :].repeat(40, 0)
    lat = np.linspace(40, 50, 40, dtype=np.float32)[:, None].repeat(50, 1)
    src = SwathDefinition(lon, lat)
    far = AreaDefinition("far", "far", "far", {"proj": "eqc", "lon_0": 0.0, "ellps": "WGS84"}, 30, 20,
                         (1.0e7, -5.0e6, 1.5e7, -4.0e6))
    r = B200NearestResampler(src, far)
    out = r.resample(np.ones((40, 50), np.float32), radius_of_influence=20000.0, mask_area=False)
    assert np.isnan(out).all()
    assert (r.neighbour_info()["index_array"] == -1).all()
    with pytest.raises(ValueError):
        r.precompute(radius_of_influence=-5.0)
    with pytest.raises(ValueError):
        r.compute(np.ones((3, 3), np.float32))


@pytest.mark.parametrize("radius", [12000.0, 60000.0])
def test_bucket_search_first_cap_holes_and_stray_rows(radius):
    """The bucket engine sizes its cells from a SAMPLE of the source and scans a one-cell cap before the full radius.
    Neither may change a result: targets next to masked holes have their nearest neighbour beyond the first cap, and a
    scan line far outside the sampled latitude range (row 6 is not in the 512-row sample of 600 rows) still has to
    be found through the edge band."""
    h, w = 600, 520
    rng = np.random.default_rng(21)
    lat = np.linspace(44.0, 62.0, h)[:, None] + 0.3 * np.linspace(-1, 1, w)[None, :] ** 2
    lon = np.linspace(-8.0, 22.0, w)[None, :] + 0.02 * np.arange(h)[:, None]
    lat = lat + rng.uniform(-0.002, 0.002, (h, w))
    lat[6, :] = 71.5 + 0.01 * np.sin(np.arange(w))          # a stray scan line 9 degrees north of everything else
    mask = np.zeros((h, w), bool)
    mask[100:140, 200:260] = True                            # holes of ~120 x 200 km and ~40 x 60 km
    mask[400:412, 50:70] = True
    src = SwathDefinition(lon, lat)
    proj = {"proj": "stere", "lat_0": 90.0, "lat_ts": 60.0, "lon_0": 10.0, "ellps": "WGS84"}
    ext = (-1.6e6, -5.0e6, 1.6e6, -1.6e6)
    dst = AreaDefinition("ps", "ps", "ps", proj, 360, 380, ext)
    r = B200NearestResampler(src, dst)
    r.precompute(mask=mask, radius_of_influence=radius)
    assert r._current["method"] == "bucket"
    vin, ref, got, _ = _check_indices(r, OSwath(lon, lat), OArea(proj, 360, 380, ext), radius, mask=mask)
    stray = np.flatnonzero(vin.ravel())[got[got >= 0]] // w == 6
    assert stray.sum() > 20                                  # the stray line was reached
    assert (got < 0).sum() > 1000 and (got >= 0).sum() > 50000


def test_swath_lonlats_may_be_device_tensors():
    """A swath whose lon/lat already live in HBM (what a device reader hands over) gives the result of the same
    swath passed as numpy arrays."""
    import torch
    lon, lat, _ = demo.viirs_like_swath(nscans=10, ncols=300)
    data = np.random.default_rng(3).uniform(0, 1, lon.shape).astype(np.float32)
    dst = AreaDefinition("ps", "ps", "ps", {"proj": "stere", "lat_0": 90.0, "lat_ts": 60.0, "lon_0": 0.0,
                                            "ellps": "WGS84"}, 240, 200, (-1.6e6, -2.4e6, 0.2e6, -0.9e6))
    on_host = B200NearestResampler(SwathDefinition(lon, lat), dst).resample(data, radius_of_influence=8000.0)
    dev = SwathDefinition(torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda())
    assert dev.shape == lon.shape
    on_dev = B200NearestResampler(dev, dst).resample(torch.from_numpy(data).cuda(), radius_of_influence=8000.0)
    np.testing.assert_array_equal(on_dev.cpu().numpy(), on_host)
    assert np.isfinite(on_host).sum() > 500


# ------------------------------------------------------------------------------------ pipeline
def test_pipeline_matches_oracle_and_cached_mode():
    import torch
    from oracle import reference_pipeline as rp
    from satpy_b200.pipeline import ABIResamplePipeline
    s, d = 700 / 10848, 1400 / 80150
    src, dst = demo.make_area("goes_east_abi_f_1km", s), demo.make_area("global_eqc_500m", d)
    counts = demo.abi_counts(src.shape, seed=7, band="C02")
    radius = 2.5 * 1002.0 / s
    pipe = ABIResamplePipeline(src, dst, radius)
    cdev = torch.from_numpy(counts).cuda()
    out = pipe.run(cdev, demo.abi_calibration("C02"), demo.START_TIME).cpu().numpy()
    out2 = pipe.run(cdev, demo.abi_calibration("C02"), demo.START_TIME, reuse_neighbours=True).cpu().numpy()
    np.testing.assert_array_equal(out, out2)
    ref = rp.abi_sunz_nearest(counts, demo.ABI_BANDS["C02"], demo.area_params("goes_east_abi_f_1km", s),
                              demo.area_params("global_eqc_500m", d), demo.START_TIME, radius)
    nan_mismatch = (np.isnan(out) != np.isnan(ref)).mean()
    assert nan_mismatch < 1e-5
    ok = ~np.isnan(out) & ~np.isnan(ref)
    err = np.abs(out[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-3)
    # ties may pick the other of two equidistant pixels: bounded, everything else within 1e-5
    assert (err > REL).mean() < 2e-3


# ------------------------------------------------------------------------------------ bilinear
# PARITY UNPINNED in the reference (satpy/tests/test_resample.py:340-416 mocks pyresample): the oracle restates the
# library's published algorithm (oracle/bilinear.py); these tests hold the device path to that restatement.
def _lcc_target(w=260, h=220):
    proj = {"proj": "lcc", "lat_1": 25.0, "lat_2": 45.0, "lat_0": 35.0, "lon_0": -95.0, "ellps": "WGS84"}
    ext = (-2.6e6, -2.2e6, 2.6e6, 2.2e6)
    return AreaDefinition("lcc", "lcc", "lcc", proj, w, h, ext), OArea(proj, w, h, ext)


@pytest.mark.parametrize("sscale,radius", [(500 / 5424, 100e3), (900 / 5424, 50e3)])
def test_bilinear_matches_oracle(sscale, radius):
    from oracle import bilinear as obil
    from satpy_b200.resample import B200BilinearResampler
    src, o_src = both_areas("goes_east_abi_f_2km", sscale)
    dst, o_dst = _lcc_target()
    r = B200BilinearResampler(src, dst)
    r.precompute(radius_of_influence=radius)
    info = r.bil_info()
    t, s, vin, idx = obil.get_bil_info(o_src, o_dst, radius, 32)
    np.testing.assert_array_equal(info["valid_input_index"], vin)
    # same corners wherever the interpolation is defined (elsewhere the index is a don't-care: the result is fill)
    defined = np.isfinite(t) & np.isfinite(s)
    got_defined = np.isfinite(info["bilinear_t"]) & np.isfinite(info["bilinear_s"])
    assert (defined != got_defined).mean() < 1e-4
    both = defined & got_defined
    assert both.mean() > 0.5
    same_corners = (info["index_array"][both] == idx[both]).all(axis=1)
    assert same_corners.mean() > 0.9999, same_corners.mean()   # equidistant ties may pick the twin pixel
    ok = both.copy()
    ok[both] = same_corners
    # the quadratic (-b +- sqrt(disc)) / 2a cancels badly where opposite sides are almost parallel (a -> 0): the
    # reference's own float64 result carries ~1e-7 of noise there, so (t, s) are held to 2e-6, not to the last bit
    np.testing.assert_allclose(info["bilinear_t"][ok], t[ok], rtol=0, atol=2e-6)
    np.testing.assert_allclose(info["bilinear_s"][ok], s[ok], rtol=0, atol=2e-6)
    assert np.median(np.abs(info["bilinear_t"][ok] - t[ok])) < 1e-12
    data = np.random.default_rng(2).uniform(200, 320, size=(2,) + src.shape).astype(np.float32)
    out = r.compute(data, fill_value=np.nan)
    exp = obil.get_sample_from_bil_info(data, t, s, vin, idx, dst.shape)
    assert out.shape == exp.shape and out.dtype == np.float32
    okimg = ok.reshape(dst.shape)
    np.testing.assert_allclose(out[:, okimg], exp[:, okimg], rtol=REL)
    np.testing.assert_array_equal(np.isnan(out[:, defined.reshape(dst.shape) == got_defined.reshape(dst.shape)]),
                                  np.isnan(exp[:, defined.reshape(dst.shape) == got_defined.reshape(dst.shape)]))


def test_bilinear_reproduces_linear_fields():
    """Size-independent property: a field linear in the target projection's coordinates is reproduced exactly."""
    import torch
    from satpy_b200.resample import B200BilinearResampler
    src, o_src = both_areas("goes_east_abi_f_2km", 1200 / 5424)
    dst, o_dst = _lcc_target(700, 600)
    slon, slat = o_src.get_lonlats()
    ix, iy = o_dst.proj.forward(slon, slat)
    data = np.where(np.isfinite(ix), 1e-5 * ix + 2e-5 * iy + 30.0, np.nan).astype(np.float32)
    ds = DataArrayLite(data, ("y", "x"), {"area": src, "_FillValue": -999.0})
    res = B200BilinearResampler(src, dst).resample(ds, radius_of_influence=50e3)
    assert res.attrs["area"] is dst and res.shape == dst.shape
    x, y = o_dst.proj_vectors()
    exp = 1e-5 * x[None, :] + 2e-5 * y[:, None] + 30.0
    out = res.values
    ok = out != -999.0
    assert ok.mean() > 0.95
    assert np.abs(out[ok] - exp[ok]).max() < 2e-4   # float32 data: 30..130 with 8e-6 resolution


def test_bilinear_needs_geostationary_source_and_float32():
    from satpy_b200.resample import B200BilinearResampler
    src, _ = both_areas("euro_laea_1km", 100 / 5000)
    dst, _ = _lcc_target(50, 40)
    with pytest.raises(NotImplementedError):
        B200BilinearResampler(src, dst).precompute()
    gsrc, _ = both_areas("goes_east_abi_f_2km", 100 / 5424)
    r = B200BilinearResampler(gsrc, dst)
    r.precompute(radius_of_influence=300e3)
    with pytest.raises(TypeError):
        r.compute(np.zeros(gsrc.shape, np.int16))


# ----------------------------------------------------------------------------------------- EWA
# PARITY UNPINNED in the reference (no test exercises "ewa"): held to oracle/ewa.py, within north_star's 1e-4.
def _ps_target(w=300, h=300):
    proj = {"proj": "stere", "lat_0": 90.0, "lat_ts": 60.0, "lon_0": 0.0, "ellps": "WGS84"}
    ext = (-1.2e6, -2.4e6, 0.3e6, -0.9e6)
    return AreaDefinition("ps", "ps", "ps", proj, w, h, ext), OArea(proj, w, h, ext)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_ll2cr_matches_oracle(dtype):
    import torch
    from oracle import ewa as oewa
    from satpy_b200.resample import B200EWAResampler
    lon, lat, _ = demo.viirs_like_swath(dtype=dtype)
    lon[3, 5] = np.nan
    lat[7, 9] = 95.0
    dst, o_dst = _ps_target()
    r = B200EWAResampler(SwathDefinition(lon, lat), dst)
    n_in = r.precompute()
    n_ref, cols, rows = oewa.ll2cr(lon, lat, o_dst)
    assert n_in == n_ref
    got_c, got_r = r._cr["cols"].cpu().numpy(), r._cr["rows"].cpu().numpy()
    assert got_c.dtype == dtype
    np.testing.assert_array_equal(np.isnan(got_c), np.isnan(cols))
    tol = 2e-3 if dtype == np.float32 else 1e-7   # float32: |x| ~ 1e6 m carries 0.06 m = 5e-4 cells of rounding
    np.testing.assert_allclose(got_c, cols, rtol=0, atol=tol, equal_nan=True)
    np.testing.assert_allclose(got_r, rows, rtol=0, atol=tol, equal_nan=True)


@pytest.mark.parametrize("rows_per_scan,bowtie,maxmode", [(16, 0.0, False), (16, 0.6, False), (0, 0.0, False),
                                                          (16, 0.3, True)])
def test_fornav_matches_oracle(rows_per_scan, bowtie, maxmode):
    import torch
    from oracle import ewa as oewa
    from satpy_b200.resample import B200EWAResampler
    lon, lat, rps = demo.viirs_like_swath(bowtie=bowtie)
    dst, o_dst = _ps_target()
    rng = np.random.default_rng(4)
    data = rng.uniform(0, 120, size=(2,) + lon.shape).astype(np.float32)
    data[0][rng.random(lon.shape) < 0.01] = np.nan
    r = B200EWAResampler(SwathDefinition(lon, lat), dst)
    r.precompute()
    # feed the oracle the device's own col/row images so that the comparison isolates fornav
    cols, rows = r._cr["cols"].cpu().numpy(), r._cr["rows"].cpu().numpy()
    out = r.compute(data, rows_per_scan=rows_per_scan, maximum_weight_mode=maxmode)
    cnt, exp = oewa.fornav(cols, rows, dst.shape, data, rows_per_scan=rows_per_scan, maximum_weight_mode=maxmode)
    assert out.shape == exp.shape == (2,) + dst.shape and out.dtype == np.float32
    nan_mismatch = (np.isnan(out) != np.isnan(exp)).mean()
    assert nan_mismatch < 2e-4, nan_mismatch      # cells whose weight sum sits on weight_sum_min
    ok = ~np.isnan(out) & ~np.isnan(exp)
    assert ok.mean() > 0.3
    if maxmode:
        assert (out[ok] != exp[ok]).mean() < 1e-3  # equal maximum weights: either sample is "the" maximum
    else:
        np.testing.assert_allclose(out[ok], exp[ok], rtol=1e-4)   # north_star: EWA within 1e-4
        assert abs(int(r.valid_counts[0]) - int(cnt[0])) <= max(3, 2e-4 * dst.size)


def test_fornav_scan_band_equals_the_oracle_on_that_band():
    """``scan_rows`` (what each rank of a multi-GPU EWA accumulates before the grids are summed): the
    result of a band of whole scans is the oracle's result for that band alone, and the bands of a split add up."""
    import torch
    from oracle import ewa as oewa
    from satpy_b200.resample import B200EWAResampler
    from satpy_b200.resample.ewa import scan_bands
    lon, lat, rps = demo.viirs_like_swath(bowtie=0.3)
    dst, _ = _ps_target()
    rng = np.random.default_rng(14)
    data = rng.uniform(0, 120, size=lon.shape).astype(np.float32)
    r = B200EWAResampler(SwathDefinition(lon, lat), dst)
    r.precompute()
    cols, rows = r._cr["cols"].cpu().numpy(), r._cr["rows"].cpu().numpy()
    bands = scan_bands(lon.shape[0], rps, 3)
    assert sum(n for _, n in bands) == lon.shape[0]
    row0, nrows = bands[1]
    out = r.compute(data, rows_per_scan=rps, scan_rows=(row0, nrows))
    sl = slice(row0, row0 + nrows)
    _, exp = oewa.fornav(cols[sl], rows[sl], dst.shape, data[sl], rows_per_scan=rps)
    assert (np.isnan(out) != np.isnan(exp)).mean() < 2e-4
    ok = ~np.isnan(out) & ~np.isnan(exp)
    assert ok.mean() > 0.05
    np.testing.assert_allclose(out[ok], exp[ok], rtol=1e-4)
    with pytest.raises(ValueError):
        r.compute(data, rows_per_scan=rps, scan_rows=(row0 + 1, nrows))
    with pytest.raises(ValueError):
        r.compute(data, rows_per_scan=rps, scan_rows=(row0, lon.shape[0]))


def test_fornav_properties():
    """Size-independent properties: a constant field is reproduced, NaN input is skipped, an empty overlap is all fill."""
    from satpy_b200.resample import B200EWAResampler
    lon, lat, rps = demo.viirs_like_swath(nscans=20, ncols=600)
    dst, _ = _ps_target(500, 500)
    r = B200EWAResampler(SwathDefinition(lon, lat), dst)
    const = np.full(lon.shape, 7.5, np.float32)
    ds = DataArrayLite(const, ("y", "x"), {"rows_per_scan": rps, "area": r.source_geo_def})
    out = r.resample(ds)
    assert out.attrs["area"] is dst
    v = out.values
    ok = ~np.isnan(v)
    assert ok.mean() > 0.2 and np.abs(v[ok] - 7.5).max() < 1e-5
    holes = const.copy()
    holes[::3] = np.nan
    v2 = r.compute(holes, rows_per_scan=rps)
    ok2 = ~np.isnan(v2)
    assert np.abs(v2[ok2] - 7.5).max() < 1e-5 and ok2.sum() <= ok.sum()
    far, _ = _ps_target(50, 50)
    far = AreaDefinition("far", "far", "far", far.proj_dict, 50, 50, (3.0e6, 3.0e6, 3.5e6, 3.5e6))
    r3 = B200EWAResampler(SwathDefinition(lon, lat), far)
    assert r3.precompute() == 0
    assert np.isnan(r3.compute(const, rows_per_scan=rps)).all()
    with pytest.raises(ValueError):
        r.compute(const, rows_per_scan=7)      # swath rows not a multiple of rows_per_scan
    with pytest.raises(ValueError):
        r.compute(const, rows_per_scan=rps, weight_min=0.0)


def test_gather_into_misaligned_row_band():
    """Row bands of an image whose width is not a multiple of 4 start at addresses that are only 4- or 8-byte aligned
    (the multi-GPU decomposition writes such bands): the gather and the fused search must handle them."""
    import torch
    from satpy_b200.pipeline import ABIResamplePipeline
    s, d = 300 / 10848, 602 / 80150      # target width 602: 602 * 4 B = 8 (mod 16)
    src, dst = demo.make_area("goes_east_abi_f_1km", s), demo.make_area("global_eqc_500m", d)
    assert dst.width % 4 == 2
    counts = torch.from_numpy(demo.abi_counts(src.shape, seed=7, band="C02")).cuda()
    cal = demo.abi_calibration("C02")
    radius = 2.5 * 1002.0 / s
    ref = ABIResamplePipeline(src, dst, radius).run(counts, cal, demo.START_TIME).cpu().numpy()
    full = torch.full((dst.height, dst.width), -7.0, dtype=torch.float32, device="cuda")
    for row0, nrows in ((0, 101), (101, 37), (138, dst.height - 138)):   # 101 * 602 is odd-aligned
        pipe = ABIResamplePipeline(src, dst, radius, row_window=(row0, nrows))
        view = full[row0:row0 + nrows]
        pipe.run(counts, cal, demo.START_TIME, out=view)
        np.testing.assert_array_equal(view.cpu().numpy(), ref[row0:row0 + nrows])
        pipe.query()
        view.fill_(-7.0)
        pipe.gather(pipe.refl, view)
        np.testing.assert_array_equal(view.cpu().numpy(), ref[row0:row0 + nrows])


def test_nearest_cache_dir_round_trip(tmp_path):
    """cache_dir persistence (satpy/resample/kdtree.py:125-177): a second resampler loads the index arrays from disk
    (no search) and gives the same image; a mask disables caching like in the reference."""
    src, _ = both_areas("goes_east_abi_f_2km", 300 / 5424)
    dst, _ = both_areas("goes_east_eqc_2km", 400 / 9050)
    data = np.random.default_rng(2).uniform(0, 1, size=src.shape).astype(np.float32)
    r1 = B200NearestResampler(src, dst)
    out1 = r1.resample(data, radius_of_influence=60e3, cache_dir=str(tmp_path))
    import json
    from satpy_b200 import _zarr2
    files = list(tmp_path.glob("nn_lut-*.zarr"))
    assert len(files) == 1 and files[0].is_dir()
    # the reference's layout: a zarr-v2 group, arrays named and dimensioned as NN_COORDINATES, int64 indices whose
    # misses are n_valid (what pykdtree reports), bool masks (satpy/resample/kdtree.py:20-22,141-177)
    assert json.load(open(files[0] / ".zgroup")) == {"zarr_format": 2}
    assert {p.name for p in files[0].iterdir() if p.is_dir()} == {"valid_input_index", "valid_output_index", "index_array"}
    meta = json.load(open(files[0] / "index_array" / ".zarray"))
    assert meta["shape"] == list(dst.shape) + [1] and meta["dtype"] == "<i8" and meta["zarr_format"] == 2
    assert json.load(open(files[0] / "index_array" / ".zattrs"))["_ARRAY_DIMENSIONS"] == ["y2", "x2", "z2"]
    assert json.load(open(files[0] / "valid_input_index" / ".zarray"))["dtype"] == "|b1"
    idx = _zarr2.read_array(str(files[0]), "index_array")
    vin = _zarr2.read_array(str(files[0]), "valid_input_index")
    info = r1.neighbour_info()
    np.testing.assert_array_equal(vin, info["valid_input_index"])
    np.testing.assert_array_equal(np.where(idx == vin.sum(), -1, idx), info["index_array"])
    r2 = B200NearestResampler(src, dst)
    r2.load_cached_neighbour_info(str(tmp_path), 60e3)
    np.testing.assert_array_equal(r2.compute(data), out1)
    r3 = B200NearestResampler(src, dst)
    np.testing.assert_array_equal(r3.resample(data, radius_of_influence=60e3, cache_dir=str(tmp_path)), out1)
    with pytest.raises(IOError):
        B200NearestResampler(src, dst).load_cached_neighbour_info(str(tmp_path), 61e3)
    B200NearestResampler(src, dst).resample(data, radius_of_influence=61e3, cache_dir=str(tmp_path),
                                            mask=np.zeros(src.shape, bool))
    assert len(list(tmp_path.glob("nn_lut-*.zarr"))) == 1


def test_row_blocks_with_source_windows_reassemble_the_full_image():
    """The multi-GPU decomposition on one GPU: every target row block, fed only the band of source rows that can reach
    it (satpy_b200.parallel.source_row_window, the analogue of get_area_slices), reproduces its rows of the full run."""
    import torch
    from satpy_b200 import parallel
    from satpy_b200.pipeline import ABIResamplePipeline
    s, d = 900 / 10848, 1802 / 80150
    src, dst = demo.make_area("goes_east_abi_f_1km", s), demo.make_area("global_eqc_500m", d)
    counts = demo.abi_counts(src.shape, seed=7, band="C02")
    cdev = torch.from_numpy(counts).cuda()
    cal = demo.abi_calibration("C02")
    radius = 2.2 * 1002.0 / s
    full = ABIResamplePipeline(src, dst, radius).run(cdev, cal, demo.START_TIME).cpu().numpy()
    block_rows, blocks = parallel.row_tiles(dst.height, 8)
    assert sum(n for _, n in blocks) == dst.height
    n_src_rows = 0
    for row0, nrows in blocks:
        s0, sn = parallel.source_row_window(src, dst, row0, nrows, radius)
        n_src_rows += sn
        pipe = ABIResamplePipeline(src, dst, radius, row_window=(row0, nrows), src_row_window=(s0, sn))
        out = pipe.run(cdev[s0:s0 + sn].contiguous(), cal, demo.START_TIME).cpu().numpy()
        np.testing.assert_array_equal(out, full[row0:row0 + nrows])
    assert n_src_rows < 1.6 * src.height          # the windows overlap a little, they do not replicate the source


def test_config3_geometry_bilinear_and_nearest_across_the_antimeridian():
    """Config 3 geometry in miniature: Himawari (lon_0 = 140.7, disk crossing +-180) -> LCC over East Asia; also a
    global grid so that both sides of the antimeridian are hit."""
    from oracle import bilinear as obil
    from satpy_b200.resample import B200BilinearResampler
    src, o_src = both_areas("himawari_ahi_fes_1km", 700 / 11000)
    dst, o_dst = both_areas("east_asia_lcc_1km", 300 / 6000)
    r = B200BilinearResampler(src, dst)
    r.precompute(radius_of_influence=50e3 * 11000 / 700 / 4)
    info = r.bil_info()
    t, s, vin, idx = obil.get_bil_info(o_src, o_dst, 50e3 * 11000 / 700 / 4, 32)
    both = np.isfinite(t) & np.isfinite(s) & np.isfinite(info["bilinear_t"]) & np.isfinite(info["bilinear_s"])
    assert both.mean() > 0.9
    same = (info["index_array"][both] == idx[both]).all(axis=1)
    assert same.mean() > 0.9999
    glob, o_glob = both_areas("global_eqc_500m", 1200 / 80150)
    rn = B200NearestResampler(src, glob)
    rn.precompute(radius_of_influence=30e3)
    vin, ref, got, n_tie = _check_indices(rn, o_src, o_glob, 30e3)
    lon, _ = o_glob.get_lonlats()
    hit = got >= 0
    assert (hit & (lon > 170)).any() and (hit & (lon < -170)).any()   # neighbours found on both sides of +-180


def test_c_abi_edge_cases_and_errors():
    """Empty and degenerate inputs and the documented error codes, straight through the C ABI."""
    import ctypes as C
    import torch
    from satpy_b200 import _lib
    lib = _lib.load()
    stream = _lib.stream_ptr(torch)
    # n = 0 is a no-op everywhere, NULL pointers allowed then
    cal = demo.abi_calibration("C02").struct("radiance")
    assert lib.b200_abi_calibrate(None, None, 0, C.byref(cal), stream) == 0
    fill = (C.c_char * 4)()
    assert lib.b200_nn_gather(None, None, None, 0, 1, 0, 0, 4, C.cast(fill, C.c_void_p), stream) == 0
    # bad arguments -> B200_EINVAL with a message, never a crash
    assert lib.b200_abi_calibrate(None, None, 8, C.byref(cal), stream) == _lib.B200_EINVAL
    assert b"NULL" in lib.b200_last_error()
    assert lib.b200_nn_gather(None, None, None, 4, 1, 0, 0, 3, C.cast(fill, C.c_void_p), stream) == _lib.B200_EINVAL
    # a 1 x 1 target and a 2 x 2 source
    src = AreaDefinition("s", "s", "s", {"proj": "geos", "lon_0": 0.0, "h": 35785831.0, "ellps": "WGS84"}, 2, 2,
                         (-3000.0, -3000.0, 3000.0, 3000.0))
    dst = AreaDefinition("d", "d", "d", {"proj": "eqc", "lon_0": 0.0, "ellps": "WGS84"}, 1, 1, (-1500.0, -1500.0, 1500.0, 1500.0))
    for method in ("bucket", "fixed_grid"):
        r = B200NearestResampler(src, dst)
        out = r.resample(np.array([[1, 2], [3, 4]], np.float32), radius_of_influence=5000.0, method=method)
        assert out.shape == (1, 1) and out[0, 0] in (1, 2, 3, 4)   # four equidistant pixels: lowest index wins
        assert out[0, 0] == 1
    # a radius of more than 4096 source pixels is refused by the fixed-grid engine, served by the bucket engine
    big = B200NearestResampler(src, dst)
    with pytest.raises(NotImplementedError, match="bucket"):
        big.precompute(radius_of_influence=2.0e7, method="fixed_grid")
    big.precompute(radius_of_influence=2.0e7, method="bucket")
    assert big.neighbour_info()["index_array"][0, 0, 0] == 0
    # a swath whose every pixel is invalid: nothing to find, no crash
    nan_swath = SwathDefinition(np.full((4, 5), np.nan), np.full((4, 5), np.nan))
    rs = B200NearestResampler(nan_swath, dst)
    res = rs.resample(np.ones((4, 5), np.float32), radius_of_influence=1e5)
    assert np.isnan(res).all() and rs._current["n_valid"] == 0
