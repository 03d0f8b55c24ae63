""""Next" rows of SURVEY.md section 8f on the device: AHI / SEVIRI / VIIRS calibration and SunZenithReducer, against the
reference's golden vectors and the oracle.  Run with ``-m gpu``."""
import datetime as dt

import numpy as np
import pytest

from oracle import calibration as ocal
from oracle import solar as osolar
from satpy_b200 import AreaDefinition, DataArrayLite, demo
from satpy_b200.modifiers import SunZenithReducer, sunz_reduce
from satpy_b200.readers import AHICalibration, SEVIRICalibration, calibrate_ahi, calibrate_seviri, scale_viirs

from .test_oracle_golden import (SEV_IR108_MET10, _SEV_COUNTS, _SEV_RAD, _SEV_TBS1, _SEV_TBS2, _SEV_VIS_RAD,
                                 _SEV_VIS_REFL)

pytestmark = pytest.mark.gpu

AHI_KW = dict(gain_count2rad_conversion=-0.0037, offset_count2rad_conversion=15.20, coeff_rad2albedo_conversion=0.0019255,
              central_wave_length=10.4073, speed_of_light=299792458.0, planck_constant=6.62606957e-34,
              boltzmann_constant=1.3806488e-23, c0_rad2tb_conversion=-0.116127314574, c1_rad2tb_conversion=1.00099153832,
              c2_rad2tb_conversion=-1.76961091571e-06)
AHI_ORACLE_KW = dict(gain=-0.0037, offset=15.20, coeff_rad2albedo=0.0019255, central_wave_length=10.4073,
                     speed_of_light=299792458.0, planck_constant=6.62606957e-34, boltzmann_constant=1.3806488e-23,
                     c0=-0.116127314574, c1=1.00099153832, c2=-1.76961091571e-06)


def test_ahi_golden_vectors():
    """satpy/tests/reader_tests/test_ahi_hsd.py:455-506."""
    counts = np.array([[0, 1000], [2000, 5000]], dtype=np.uint16)
    cal = AHICalibration(bt_in_float64=False, **AHI_KW)   # the fixture passes Python floats: float32 arithmetic
    assert calibrate_ahi(123, cal, "counts") == 123
    rad = calibrate_ahi(counts, cal, "radiance")
    assert rad.dtype == np.float32 and np.allclose(rad, np.array([[15.2, 11.5], [7.8, -3.3]]))
    bt = calibrate_ahi(counts, cal, "brightness_temperature")
    np.testing.assert_allclose(bt, np.array([[330.978979, 310.524688], [285.845017, np.nan]]), rtol=2e-7)
    refl = calibrate_ahi(counts, cal, "reflectance")
    assert np.allclose(refl, np.array([[2.92676, 2.214325], [1.50189, 0.]]))


@pytest.mark.parametrize("n", [0, 7, 8, 9, 1 << 20])
@pytest.mark.parametrize("f64", [False, True])
def test_ahi_matches_oracle(n, f64):
    rng = np.random.default_rng(n + 3)
    counts = rng.integers(0, 2048, size=n, dtype=np.uint16)
    counts[rng.random(n) < 0.01] = 65535
    counts[rng.random(n) < 0.01] = 65534
    cal = AHICalibration(count_value_outside_scan_pixels=65535, count_value_error_pixels=65534, bt_in_float64=f64, **AHI_KW)
    okw = dict(AHI_ORACLE_KW, count_outside_scan=65535, count_error=65534)
    if f64:  # numpy float64 header scalars, as read from a real file
        okw.update({k: np.float64(v) for k, v in okw.items() if k in ("central_wave_length", "speed_of_light",
                    "planck_constant", "boltzmann_constant", "c0", "c1", "c2")})
    for calib in ("radiance", "reflectance"):
        np.testing.assert_array_equal(calibrate_ahi(counts, cal, calib), ocal.ahi_calibrate(counts, calib, **okw))
    bt = calibrate_ahi(counts, cal, "brightness_temperature")
    ref = ocal.ahi_calibrate(counts, "brightness_temperature", **okw)
    np.testing.assert_allclose(bt, ref.astype(np.float32), rtol=1e-5, equal_nan=True)


def test_seviri_golden_vectors():
    """satpy/tests/reader_tests/test_seviri_l1b_calibration.py:29-149 (Meteosat-10, IR_108 and VIS008)."""
    c = SEV_IR108_MET10
    cal = SEVIRICalibration(gain=0.20503567620766011, offset=-10.456819486590666, wavenumber=c["wavenumber"],
                            alpha=c["alpha"], beta=c["beta"], btfit=c["btfit"])
    rad = calibrate_seviri(_SEV_COUNTS, cal, "radiance")
    assert rad.dtype == np.float32
    np.testing.assert_allclose(rad, _SEV_RAD, rtol=1e-5)
    np.testing.assert_allclose(calibrate_seviri(_SEV_COUNTS, cal, "brightness_temperature", "spectral"), _SEV_TBS1, rtol=1e-5)
    np.testing.assert_allclose(calibrate_seviri(_SEV_COUNTS, cal, "brightness_temperature", "effective"), _SEV_TBS2, rtol=1e-5)
    with pytest.raises(ValueError, match="No calibration coefficients"):
        calibrate_seviri(_SEV_COUNTS, cal, "brightness_temperature", "not_processed")
    # VIS008: feed radiances through gain 1 / offset 0
    vis = SEVIRICalibration(gain=1.0, offset=0.0, solar_irradiance=73.1807, scan_time=dt.datetime(2020, 8, 15, 13, 0, 40))
    np.testing.assert_allclose(calibrate_seviri(_SEV_VIS_RAD, vis, "reflectance"), _SEV_VIS_REFL, rtol=1e-5)


def test_seviri_matches_oracle_uint16():
    rng = np.random.default_rng(8)
    counts = rng.integers(0, 1024, size=(300, 333), dtype=np.uint16)
    c = SEV_IR108_MET10
    cal = SEVIRICalibration(gain=0.2050356762, offset=-10.4568194866, wavenumber=c["wavenumber"], alpha=c["alpha"],
                            beta=c["beta"], btfit=c["btfit"])
    rad = ocal.seviri_radiance(counts, cal.gain, cal.offset)
    np.testing.assert_array_equal(calibrate_seviri(counts, cal, "radiance"), rad)
    np.testing.assert_allclose(calibrate_seviri(counts, cal, "brightness_temperature", "effective"),
                               ocal.seviri_erads2bt(rad, c["wavenumber"], c["alpha"], c["beta"]), rtol=1e-5, equal_nan=True)
    np.testing.assert_allclose(calibrate_seviri(counts, cal, "brightness_temperature", "spectral"),
                               ocal.seviri_srads2bt(rad, c["wavenumber"], c["btfit"]), rtol=1e-5, equal_nan=True)


def test_viirs_scale_matches_oracle():
    rng = np.random.default_rng(12)
    rows_per_gran = [768, 768, 752]
    data = rng.integers(0, 65536, size=(sum(rows_per_gran), 320), dtype=np.uint16)
    factors = np.array([2.0e-3, -0.5, 1.9e-3, -0.4, -999.9, -999.9], np.float32)
    out = scale_viirs(data, factors, rows_per_gran)
    ref = ocal.viirs_scale(data, factors, rows_per_gran)
    assert out.dtype == np.float32
    np.testing.assert_array_equal(out, ref)
    fdata = rng.uniform(-1200, 300, size=(32, 64)).astype(np.float32)
    np.testing.assert_array_equal(scale_viirs(fdata, [1.5, 2.0], [32]), ocal.viirs_scale(fdata, [1.5, 2.0], [32]))
    with pytest.raises(ValueError):
        scale_viirs(data, factors, [768, 768])


def test_sunz_reducer_golden_and_oracle():
    """satpy/tests/test_modifiers.py:186-220."""
    area = AreaDefinition("test", "test", "test", {"proj": "merc"}, 2, 2, (-2000, -2000, 2000, 2000))
    sza = np.rad2deg(np.arccos(np.array([[0.0149581333, 0.0146694376], [0.0150812684, 0.0147925727]]))).astype(np.float32)
    ds = DataArrayLite(np.ones((2, 2), np.float32), ("y", "x"), {"area": area, "start_time": dt.datetime(2018, 1, 1, 18),
                                                                "name": "test_vis", "modifiers": tuple()})
    sz = DataArrayLite(sza, ("y", "x"), {"area": area})
    res = SunZenithReducer(name="sza_reduction_test_default", modifiers=tuple())((ds, sz))
    np.testing.assert_allclose(res.values, np.array([[0.02916261, 0.02839063], [0.02949383, 0.02871911]], np.float32), rtol=2e-5)
    res = SunZenithReducer(name="custom", modifiers=tuple(), correction_limit=70, max_sza=95, strength=3.0)((ds, sz))
    np.testing.assert_allclose(res.values, np.array([[0.01041319, 0.01030033], [0.01046164, 0.01034834]], np.float32), rtol=2e-5)
    with pytest.raises(ValueError, match="`max_sza` must be defined"):
        SunZenithReducer(name="x", max_sza=None)
    rng = np.random.default_rng(1)
    data = rng.uniform(0, 100, (2, 200, 300)).astype(np.float32)
    z = rng.uniform(0, 100, (200, 300)).astype(np.float32)
    z[0, :5] = np.nan
    out = sunz_reduce(data, z, limit=55.0, max_sza=90.0, strength=1.5)
    ref = osolar.sunzen_reduction(data, z, limit=55., max_sza=90., strength=1.5).astype(np.float32)
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6)


def test_native_resampler_replicate_and_aggregate():
    """satpy/resample/native.py:41-159 (cf. satpy/tests/test_resample.py TestNativeResampler: expand, reduce, errors)."""
    from satpy_b200.resample import B200NativeResampler
    rng = np.random.default_rng(5)
    small = AreaDefinition("s", "s", "s", {"proj": "merc"}, 50, 40, (-2000, -2000, 2000, 2000))
    big = AreaDefinition("b", "b", "b", {"proj": "merc"}, 100, 120, (-2000, -2000, 2000, 2000))
    for dtype in (np.float32, np.int16, np.uint8, np.float64):
        data = (rng.uniform(0, 200, (3, 40, 50))).astype(dtype)
        out = B200NativeResampler(small, big).resample(data)
        assert out.dtype == dtype
        np.testing.assert_array_equal(out, np.repeat(np.repeat(data, 3, axis=1), 2, axis=2))
    d2 = rng.uniform(0, 1, (120, 100)).astype(np.float32)
    d2[rng.random(d2.shape) < 0.2] = np.nan
    d2[:3, :2] = np.nan   # an all-NaN block
    ds = DataArrayLite(d2, ("y", "x"), {"area": big})
    res = B200NativeResampler(big, [big, small]).compute(ds, expand=False)
    assert res.attrs["area"] is small
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exp = np.nanmean(d2.reshape(40, 3, 50, 2), axis=(1, 3))
    np.testing.assert_allclose(res.values, exp, rtol=1e-6, equal_nan=True)
    np.testing.assert_array_equal(B200NativeResampler(small, small).resample(d2[:40, :50]), d2[:40, :50])
    odd = AreaDefinition("o", "o", "o", {"proj": "merc"}, 75, 60, (-2000, -2000, 2000, 2000))
    with pytest.raises(ValueError, match="whole number"):
        B200NativeResampler(small, odd).resample(d2[:40, :50])
    with pytest.raises(ValueError, match="more than 2 dimensions"):
        B200NativeResampler(big, small).resample(np.zeros((2, 120, 100), np.float32))


def test_simulated_green():
    """satpy/composites/abi.py:56-61."""
    from satpy_b200.composites import SimulatedGreen
    rng = np.random.default_rng(6)
    c = [rng.uniform(0, 120, (300, 500)).astype(np.float32) for _ in range(3)]
    c[1][5, 5] = np.nan
    area = AreaDefinition("s", "s", "s", {"proj": "merc"}, 500, 300, (-2000, -2000, 2000, 2000))
    ds = [DataArrayLite(a, ("y", "x"), {"area": area, "name": n}) for a, n in zip(c, ("C01", "C02", "C03"))]
    res = SimulatedGreen("green", fractions=(0.465, 0.465, 0.07))(ds)
    exp = c[0] * 0.465 + c[1] * 0.465 + c[2] * 0.07
    assert exp.dtype == np.float32 and res.dtype == np.float32
    np.testing.assert_array_equal(res.values, exp)
    assert res.attrs["name"] == "green" and res.attrs["area"] is area


def test_config2_chain_matches_oracle():
    """BASELINE.json configs[1] in miniature: C01-C03 counts -> reflectance -> sunz (one fused launch for the three
    bands) -> ONE neighbour search -> gather x3 -> SimulatedGreen, against the oracle chain."""
    import ctypes as C
    import torch
    from oracle import nearest as onn
    from oracle import reference_pipeline as rp
    from oracle.projections import Area as OArea
    from satpy_b200 import _lib, demo
    from satpy_b200.composites import SimulatedGreen
    from satpy_b200.modifiers.geometry import _sunz_struct
    from satpy_b200.resample import B200NearestResampler
    s, d = 500 / 5424, 800 / 9050
    src, dst = demo.make_area("goes_east_abi_f_2km", s), demo.make_area("goes_east_eqc_2km", d)
    o_src, o_dst = OArea(*demo.area_params("goes_east_abi_f_2km", s)), OArea(*demo.area_params("goes_east_eqc_2km", d))
    bands = ("C01", "C02", "C03")
    counts = [demo.abi_counts(src.shape, seed=2 + i, band=b) for i, b in enumerate(bands)]
    cdev = [torch.from_numpy(c).cuda() for c in counts]
    refl = torch.empty((3,) + src.shape, dtype=torch.float32, device="cuda")
    cals = (_lib.AbiCalStruct * 3)(*[demo.abi_calibration(b).struct("reflectance") for b in bands])
    g = src.geo_struct()
    sz = _sunz_struct(demo.START_TIME, 88.0, 95.0, True, 0)
    _lib.check(_lib.load().b200_abi_vis_sunz((C.c_void_p * 3)(*[t.data_ptr() for t in cdev]),
                                             (C.c_void_p * 3)(*[refl[i].data_ptr() for i in range(3)]), cals, 3,
                                             C.byref(g), C.byref(sz), _lib.stream_ptr(torch)))
    radius = 2.5 * 2004.0 / s
    r = B200NearestResampler(src, dst)
    out = r.resample(refl, radius_of_influence=radius)          # one search, three gathers
    green = SimulatedGreen("green")([out[0], out[1], out[2]]).cpu().numpy()
    # oracle chain
    ref_bands = np.stack([rp.abi_sunz(c, demo.ABI_BANDS[b], o_src, demo.START_TIME) for c, b in zip(counts, bands)])
    slon, slat = o_src.get_lonlats()
    dlon, dlat = o_dst.get_lonlats()
    vin, vout, idx = onn.get_neighbour_info(slon, slat, dlon, dlat, radius)
    ref = onn.get_sample_from_neighbour_info(ref_bands, vin, idx, np.float32(np.nan))
    ref_green = ref[0] * 0.465 + ref[1] * 0.465 + ref[2] * 0.07
    assert (np.isnan(green) != np.isnan(ref_green)).mean() < 1e-5
    ok = ~np.isnan(green) & ~np.isnan(ref_green)
    err = np.abs(green[ok] - ref_green[ok]) / np.maximum(np.abs(ref_green[ok]), 1e-3)
    assert ok.mean() > 0.3 and (err > 1e-5).mean() < 2e-3   # equidistant ties may pick the twin pixel


def test_get_angles_matches_oracle_and_relative_azimuth_golden():
    """satpy/modifiers/angles.py:336-402 on the device; relative azimuth golden: test_angles.py:361-382."""
    from oracle.projections import Area as OArea
    from satpy_b200 import demo
    from satpy_b200.modifiers.angles import compute_relative_azimuth, get_angles, get_satellite_zenith_angle
    scale = 400 / 5424
    area, oarea = demo.make_area("goes_east_abi_f_2km", scale), OArea(*demo.area_params("goes_east_abi_f_2km", scale))
    when = dt.datetime(2019, 3, 14, 18, 0)
    ds = DataArrayLite(np.zeros(area.shape, np.float32), ("y", "x"),
                       {"area": area, "start_time": when,
                        "orbital_parameters": {"satellite_nominal_longitude": -75.2, "satellite_nominal_latitude": 0.03,
                                               "satellite_nominal_altitude": 35786023.0}})
    sata, satz, suna, sunz = get_angles(ds)
    ref = osolar.get_angles(oarea, when, -75.2, 0.03, 35786023.0)
    for got, exp, name in zip((sata, satz, suna, sunz), ref, ("sata", "satz", "suna", "sunz")):
        assert got.dtype == np.float64
        np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
        ok = ~np.isnan(exp)
        d = np.abs(got[ok] - exp[ok])
        if name in ("sata", "suna"):
            d = np.minimum(d, 360.0 - d)   # azimuth wraps; near the sub-satellite / sub-solar point it is ill-conditioned
            assert np.quantile(d, 0.999) < 1e-6, name
        else:
            assert d.max() < 1e-7, name
    assert np.nanmin(satz) < 1.0 and 85 < np.nanmax(satz) < 92
    np.testing.assert_array_equal(get_satellite_zenith_angle(ds), satz)
    for dtype in (np.float32, np.float64):
        saa = np.array([-120, 40., 0.04, 179.4, 94.2, 12.1], dtype=dtype)
        vaa = np.array([60., 57.7, 175.1, 234.18, 355.4, 12.1], dtype=dtype)
        raa = compute_relative_azimuth(vaa, saa)
        assert raa.dtype == dtype
        np.testing.assert_allclose(np.array([180., 17.7, 175.06, 54.78, 98.8, 0.], dtype=dtype), raa, rtol=2e-7)
        np.testing.assert_array_equal(raa, osolar.compute_relative_azimuth(vaa, saa))


# ------------------------------------------------------------------------------- ratio sharpening
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_ratio_and_self_sharpened_rgb_golden_and_oracle(dtype):
    """RatioSharpenedRGB / SelfSharpenedRGB: the reference's golden vectors (satpy/tests/compositor_tests/
    test_resolution.py:143-215), error behaviour (:96-121), then odd shapes, crop offsets and NaNs vs the oracle."""
    from oracle import composites as oc
    from satpy_b200 import DataArrayLite
    from satpy_b200.composites import RatioSharpenedRGB, SelfSharpenedRGB
    from satpy_b200.modifiers import IncompatibleAreas
    from .test_oracle_golden import SHARPEN_CASES, _sharpen_fixture
    low, ds2, ds3, high = _sharpen_fixture(dtype)
    attrs = {"resolution": 1000, "calibration": "reflectance", "units": "%", "name": "test_vis"}
    mk = lambda a, **kw: DataArrayLite(a, ("y", "x"), dict(attrs, **kw))   # noqa: E731
    for hi, neutral, exp_r, exp_g, exp_b in SHARPEN_CASES:
        comp = RatioSharpenedRGB(name="true_color", high_resolution_band=hi, neutral_resolution_band=neutral)
        res = comp((mk(low), mk(ds2), mk(ds3)), optional_datasets=(mk(high, resolution=500),))
        assert "units" not in res.attrs and res.attrs["name"] == "true_color" and res.attrs["resolution"] == 500
        assert res.dims == ("bands", "y", "x") and list(res.coords["bands"]) == ["R", "G", "B"]
        data = res.values
        assert data.dtype == dtype and data.shape == (3, 2, 2)
        for got, exp in zip(data, (exp_r, exp_g, exp_b)):
            np.testing.assert_allclose(got, np.array(exp), rtol=1e-5)
    data = SelfSharpenedRGB(name="true_color")((mk(low), mk(ds2), mk(ds3))).values
    np.testing.assert_allclose(data[0], [[5.0, 5.0], [5.0, 0]], rtol=1e-5)
    np.testing.assert_allclose(data[1], [[4.0, 4.0], [4.0, 0]], rtol=1e-5)
    np.testing.assert_allclose(data[2], [[16 / 3, 16 / 3], [16 / 3, 0]], rtol=1e-5)
    res = RatioSharpenedRGB(name="true_color", high_resolution_band=None)((mk(low), mk(ds2), mk(ds3)),
                                                                            optional_datasets=(mk(high),))
    assert res.values.shape == (3, 2, 2) and res.values.dtype == dtype
    with pytest.raises(ValueError):
        RatioSharpenedRGB(name="x", high_resolution_band="bad")
    with pytest.raises(ValueError):
        RatioSharpenedRGB(name="x")((mk(low), mk(ds2), mk(ds3), mk(ds3)))
    with pytest.raises(IncompatibleAreas):
        RatioSharpenedRGB(name="x")((mk(low), mk(ds2), mk(ds3)), optional_datasets=(mk(np.ones((4, 4), dtype)),))
    with pytest.raises(ValueError, match="at least one high resolution band"):
        SelfSharpenedRGB(name="x", high_resolution_band=None)((mk(low), mk(ds2), mk(ds3)))
    # bigger, odd-shaped, with NaNs and both crop-offset parities
    rng = np.random.default_rng(11)
    shape = (301, 413)
    bands = [rng.uniform(0, 100, size=shape).astype(dtype) for _ in range(3)]
    bands[0][rng.random(shape) < 0.02] = np.nan
    bands[1][rng.random(shape) < 0.01] = np.nan
    bands[0][rng.random(shape) < 0.01] = 0.0
    hi_band = rng.uniform(0, 100, size=shape).astype(dtype)
    hi_band[rng.random(shape) < 0.02] = np.nan
    tol = 2e-6 if dtype == np.float32 else 1e-13
    for hi, neutral in (("red", None), ("green", "blue")):
        got = RatioSharpenedRGB(name="t", high_resolution_band=hi, neutral_resolution_band=neutral)(
            bands, optional_datasets=(hi_band,))
        exp = oc.ratio_sharpened_rgb(*bands, hi_band, hi, neutral)
        np.testing.assert_allclose(got, exp, rtol=tol, equal_nan=True)

    class _Area:
        def __init__(self, off):
            self.crop_offset = off
    for off in ((0, 0), (1, 0), (0, 1), (3, 5)):
        ds = [DataArrayLite(b, ("y", "x"), {"area": _Area(off)}) for b in bands]
        got = SelfSharpenedRGB(name="t", high_resolution_band="red")(ds).values
        exp = oc.ratio_sharpened_rgb(*bands, None, "red", None, self_sharpen=True, offset=off)
        np.testing.assert_allclose(got, exp, rtol=tol, equal_nan=True)
    unmasked = SelfSharpenedRGB(name="t", common_channel_mask=False)(bands)
    exp = oc.ratio_sharpened_rgb(*bands, None, "red", None, self_sharpen=True, common_channel_mask=False)
    np.testing.assert_allclose(unmasked, exp, rtol=tol, equal_nan=True)


# ------------------------------------------------------------------------------- EWA maximum_weight_mode over ranks
def test_fornav_maximum_weight_mode_split_over_two_emulated_ranks():
    """The max-by-weight reduction of ``B200EWAResampler.compute(group=...)`` with the two ranks emulated on one GPU:
    pass 1 per band (mode 2), element-wise maximum of the weight grids, pass 2 per band (mode 3), MAX of the encoded
    value grids -- equals the single-call ``maximum_weight_mode`` result up to equal-weight ties."""
    import ctypes as C
    import torch
    from satpy_b200 import AreaDefinition, SwathDefinition, _lib
    from satpy_b200.resample import B200EWAResampler
    from satpy_b200.resample.ewa import ewa_params, max_weight_decode, max_weight_encode, scan_bands
    lon, lat, rps = demo.viirs_like_swath(bowtie=0.3)
    proj = {"proj": "stere", "lat_0": 90.0, "lat_ts": 60.0, "lon_0": 0.0, "ellps": "WGS84"}
    dst = AreaDefinition("ps", "ps", "ps", proj, 300, 300, (-1.2e6, -2.4e6, 0.3e6, -0.9e6))
    rng = np.random.default_rng(4)
    data = rng.uniform(0, 120, size=lon.shape).astype(np.float32)
    data[rng.random(lon.shape) < 0.02] = np.nan
    r = B200EWAResampler(SwathDefinition(lon, lat), dst)
    whole = r.compute(data, rows_per_scan=rps, maximum_weight_mode=True)
    lib = _lib.load()
    stream = _lib.stream_ptr(torch)
    t = torch.from_numpy(data).cuda()
    src = r.source_geo_def
    bands = scan_bands(lon.shape[0], rps, 2)
    item = 4 if r._cr["f32"] else 8

    def accum(row0, nrows, mode, weights, accums):
        p = ewa_params(maximum_weight_mode=True)
        p.maximum_weight_mode = mode
        first = row0 * src.width
        _lib.check(lib.b200_fornav_accum(r._cr["cols"].data_ptr() + first * item, r._cr["rows"].data_ptr() + first * item,
                                         1 if r._cr["f32"] else 0, nrows, src.width, rps, t.data_ptr() + first * 4, 1,
                                         src.size, float("nan"), dst.height, dst.width, weights.data_ptr(),
                                         accums.data_ptr(), C.byref(p), stream))
    ws = [torch.zeros(dst.shape, dtype=torch.float32, device="cuda") for _ in bands]
    dummy = torch.zeros(dst.shape, dtype=torch.float32, device="cuda")
    for (row0, nrows), w in zip(bands, ws):
        accum(row0, nrows, 2, w, dummy)
    wmax = torch.maximum(ws[0], ws[1])
    encs = []
    for row0, nrows in bands:
        a = torch.full(dst.shape, float("-inf"), dtype=torch.float32, device="cuda")
        accum(row0, nrows, 3, wmax.clone(), a)
        encs.append(max_weight_encode(a))
    merged = max_weight_decode(torch.maximum(encs[0], encs[1]))
    out = torch.empty(dst.shape, dtype=torch.float32, device="cuda")
    p = ewa_params(maximum_weight_mode=True)
    _lib.check(lib.b200_fornav_finalize(wmax.data_ptr(), merged.data_ptr(), out.data_ptr(), 1, dst.size, C.byref(p),
                                        float("nan"), None, stream))
    got = out.cpu().numpy()
    assert (np.isnan(got) != np.isnan(whole)).mean() < 1e-3
    ok = ~np.isnan(got) & ~np.isnan(whole)
    assert ok.mean() > 0.3 and (got[ok] != whole[ok]).mean() < 2e-3


# ------------------------------------------------------------------------------- sensor-angle caching
def test_sensor_angles_are_cached_on_the_device_and_as_zarr(tmp_path):
    """``cache_sensor_angles`` (satpy/modifiers/angles.py:53-213,321-333): sensor angles computed once per (area,
    rounded satellite position), kept on the device and persisted as zarr stores with the reference's file-name
    pattern; the sun angles are evaluated per scene."""
    import datetime as dt
    import torch
    from satpy_b200.modifiers import angles
    area = demo.make_area("goes_east_abi_f_2km", 400 / 5424)
    t1, t2 = dt.datetime(2019, 3, 14, 18, 0), dt.datetime(2019, 3, 14, 21, 0)
    pos = dict(sat_lon=-75.23, sat_lat=0.04, sat_alt=35786023.7)
    angles.sensor_angle_cache_clear(str(tmp_path))
    plain = angles.get_angles(area=area, start_time=t1, sat_lon=-75.2, sat_lat=0.0, sat_alt=35786023.7, device=True)
    a1 = angles.get_angles(area=area, start_time=t1, device=True, cache_sensor_angles=True, cache_dir=str(tmp_path), **pos)
    for got, exp in zip(a1, plain):       # cached = computed with the position rounded to one decimal
        np.testing.assert_allclose(got.cpu().numpy(), exp.cpu().numpy(), rtol=0, atol=1e-9, equal_nan=True)
    stores = sorted(p.name for p in tmp_path.glob("_get_sensor_angles_from_sat_pos_v1_*_*.zarr"))
    assert len(stores) == 2 and "_v1_0_" in stores[0] and "_v1_1_" in stores[1]
    a2 = angles.get_angles(area=area, start_time=t2, device=True, cache_sensor_angles=True, cache_dir=str(tmp_path), **pos)
    assert a2[0] is a1[0] and a2[1] is a1[1]                       # device-resident
    assert not torch.equal(torch.nan_to_num(a2[3]), torch.nan_to_num(a1[3]))   # the sun moved
    angles._SENSOR_ANGLE_CACHE.clear()                              # a new process: read the stores back
    a3 = angles.get_angles(area=area, start_time=t2, device=True, cache_sensor_angles=True, cache_dir=str(tmp_path), **pos)
    assert a3[0] is not a1[0]
    np.testing.assert_array_equal(a3[0].cpu().numpy(), a1[0].cpu().numpy())
    np.testing.assert_array_equal(a3[1].cpu().numpy(), a1[1].cpu().numpy())
    angles.config["cache_sensor_angles"], angles.config["cache_dir"] = True, str(tmp_path)
    try:
        satz = angles.get_satellite_zenith_angle(area=area, start_time=t1, device=True, **pos)
        assert satz is a3[1]
    finally:
        angles.config["cache_sensor_angles"], angles.config["cache_dir"] = False, None
    angles.sensor_angle_cache_clear(str(tmp_path))
    assert not list(tmp_path.glob("*.zarr")) and not angles._SENSOR_ANGLE_CACHE


def test_bilinear_cache_dir_round_trip(tmp_path):
    """``cache_dir`` for the bilinear coefficients (satpy/resample/kdtree.py:216-279): a second resampler loads
    ``bil_lut-<hash>.zarr`` instead of searching and gives the same image."""
    import json
    from satpy_b200 import AreaDefinition
    from satpy_b200.resample import B200BilinearResampler
    src = demo.make_area("goes_east_abi_f_2km", 300 / 5424)
    proj = {"proj": "lcc", "lat_1": 25.0, "lat_2": 45.0, "lat_0": 35.0, "lon_0": -95.0, "ellps": "WGS84"}
    dst = AreaDefinition("lcc", "lcc", "lcc", proj, 120, 100, (-2.6e6, -2.2e6, 2.6e6, 2.2e6))
    data = np.random.default_rng(3).uniform(200, 320, size=src.shape).astype(np.float32)
    r1 = B200BilinearResampler(src, dst)
    out1 = r1.resample(data, radius_of_influence=100e3, cache_dir=str(tmp_path), fill_value=np.nan)
    stores = list(tmp_path.glob("bil_lut-*.zarr"))
    assert len(stores) == 1
    assert json.load(open(stores[0] / "bilinear_t" / ".zarray"))["dtype"] == "<f8"
    assert json.load(open(stores[0] / "index_array" / ".zarray"))["shape"] == [dst.size, 4]
    r2 = B200BilinearResampler(src, dst)
    r2.load_bil_info(str(tmp_path), radius_of_influence=100e3)
    np.testing.assert_array_equal(r2.compute(data, fill_value=np.nan), out1)
    r3 = B200BilinearResampler(src, dst)
    np.testing.assert_array_equal(r3.resample(data, radius_of_influence=100e3, cache_dir=str(tmp_path),
                                              fill_value=np.nan), out1)
    with pytest.raises(IOError):
        B200BilinearResampler(src, dst).load_bil_info(str(tmp_path), radius_of_influence=101e3)
