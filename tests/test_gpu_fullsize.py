"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run these sizes):
fused == two-step, engines agree, identity resampling is the identity, every output value is the value at its index.
Run with ``-m gpu``; about half a minute on a B200."""
import numpy as np
import pytest

from satpy_b200 import demo
from satpy_b200.resample import B200NearestResampler

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c5():
    import torch
    src, dst = demo.make_area("goes_east_abi_f_1km"), demo.make_area("global_eqc_500m")
    data = torch.rand(src.shape, device="cuda", dtype=torch.float32) * 100.0
    return src, dst, data


def test_c5_fused_equals_search_then_gather_and_values_come_from_their_index(c5):
    """Config 5 at full size (10848^2 -> 80150 x 40075, 3212 Mpix)."""
    import torch
    src, dst, data = c5
    r = B200NearestResampler(src, dst)
    r.precompute(radius_of_influence=2000.0)          # the reference's three arrays
    two_step = r.compute(data)
    info = r._current
    gi = info["gather_index"]
    hit = gi >= 0
    frac = float(hit.float().mean())
    assert 0.25 < frac < 0.45                       # the visible disk on a global grid
    # every hit carries exactly the source value at its index, every miss is NaN
    assert bool(torch.isnan(two_step[~hit]).all())
    assert bool((two_step[hit] == data.reshape(-1)[gi[hit].long()]).all())
    # compressed index = raw index minus the invalid pixels before it
    vin = info["valid_input_index"].reshape(-1)
    assert info["n_valid"] == int(vin.sum())
    sample = torch.randint(0, gi.numel(), (1 << 20,), device="cuda")
    g, c = gi.reshape(-1)[sample], info["index_array"].reshape(-1)[sample]
    ok = g >= 0
    prefix = torch.cumsum(vin.to(torch.int32), 0) - vin.to(torch.int32)
    assert bool((c[ok] == prefix[g[ok].long()]).all()) and bool((c[~ok] == -1).all())
    del prefix, info
    r._index_caches.clear()
    r._current = None
    torch.cuda.empty_cache()
    fused = r.resample_fused(data, 2000.0)
    assert bool(((fused == two_step) | (torch.isnan(fused) & torch.isnan(two_step))).all())


def test_c5_engines_agree_on_full_resolution_row_bands(c5):
    """Bucket grid vs fixed grid (and its strict mode) on three 256-row bands of the full-resolution target."""
    import torch
    src, dst, _ = c5
    for row0 in (5000, 20000, 33000):
        out = {}
        for method in ("bucket", "fixed_grid", "fixed_grid_strict"):
            r = B200NearestResampler(src, dst)
            r.precompute(radius_of_influence=2000.0, row_window=(row0, 256), method=method)
            out[method] = r._current["gather_index"].clone()
            del r
            torch.cuda.empty_cache()
        assert bool((out["bucket"] == out["fixed_grid"]).all())
        assert bool((out["bucket"] == out["fixed_grid_strict"]).all())


def test_identity_resampling_is_the_identity_full_disk():
    """Resampling a full-disk image onto its own grid with a sub-pixel radius returns it unchanged (11.8e7 pixels)."""
    import torch
    src = demo.make_area("goes_east_abi_f_1km")
    data = torch.rand(src.shape, device="cuda", dtype=torch.float32)
    r = B200NearestResampler(src, src)
    out = r.resample(data, radius_of_influence=400.0)
    valid = r._current["valid_input_index"].bool()
    assert bool((out[valid] == data[valid]).all()) and bool(torch.isnan(out[~valid]).all())
    n = src.width
    raw = torch.arange(src.size, device="cuda", dtype=torch.int32).reshape(src.shape)
    assert bool((r._current["gather_index"][valid] == raw[valid]).all())
    assert abs(float(valid.float().mean()) - 0.775) < 0.02    # area of the disk inside the square grid


def test_c2_one_search_serves_three_bands_full_size():
    """Config 2 at full size: 3 x 5424^2 -> 9050^2; the gather of a 3-band stack equals three single-band gathers."""
    import torch
    src, dst = demo.make_area("goes_east_abi_f_2km"), demo.make_area("goes_east_eqc_2km")
    rgb = torch.rand((3,) + src.shape, device="cuda", dtype=torch.float32)
    r = B200NearestResampler(src, dst)
    out = r.resample(rgb, radius_of_influence=5000.0)
    assert tuple(out.shape) == (3,) + dst.shape
    for b in range(3):
        single = r.compute(rgb[b])
        assert bool(((single == out[b]) | (torch.isnan(single) & torch.isnan(out[b]))).all())
    u8 = (rgb[0] * 255).to(torch.uint8)
    o8 = r.compute(u8, fill_value=255)
    gi = r._current["gather_index"]
    assert o8.dtype == torch.uint8 and bool((o8[gi < 0] == 255).all())


def test_c5_tiered_search_equals_strict_mode_on_the_whole_target(c5):
    """The fixed 2x2 / 4x4 stencils with the satellite-distance factor (``fixed_grid``) against the general loop that
    examines every lattice point of the rigorous disc (``fixed_grid_strict``) on EVERY pixel of config 5, band by band."""
    import torch
    src, dst, _ = c5
    hits = 0
    step = 4008
    for row0 in range(0, dst.height, step):
        n = min(step, dst.height - row0)
        out = {}
        for method in ("fixed_grid", "fixed_grid_strict"):
            r = B200NearestResampler(src, dst)
            r.precompute(radius_of_influence=2000.0, row_window=(row0, n), method=method, gather_only=True)
            out[method] = r._current["gather_index"]
            del r
        diff = int((out["fixed_grid"] != out["fixed_grid_strict"]).sum())
        assert diff == 0, f"rows {row0}..{row0 + n}: {diff} pixels differ"
        hits += int((out["fixed_grid"] >= 0).sum())
        del out
    assert 0.25 < hits / dst.size < 0.45


def test_light_search_structure_serves_the_fused_path_only():
    """``B200_NN_LIGHT``: float32 points only.  The fused search+gather and raw gather indices work (and equal the
    full structure's, exact float64 decisions included); what needs the prefix table or the float64 records is
    refused with B200_EUNSUPPORTED."""
    import ctypes as C
    import torch
    from satpy_b200 import _lib
    from satpy_b200.resample.nearest import NN_LIGHT, SEARCH_METHODS
    lib = _lib.load()
    stream = _lib.stream_ptr(torch)
    src, dst = demo.make_area("goes_east_abi_f_2km"), demo.make_area("goes_east_eqc_2km")
    data = torch.rand(src.shape, device="cuda", dtype=torch.float32)
    gs, gd = src.geo_struct(), dst.geo_struct()
    res = {}
    for light in (0, NN_LIGHT):
        vin = torch.empty(src.shape, dtype=torch.uint8, device="cuda")
        h = C.c_void_p(None)
        _lib.check(lib.b200_nn_grid_create_ex(C.byref(gs), None, 5000.0, vin.data_ptr(), None, C.byref(h),
                                              SEARCH_METHODS["fixed_grid"] | light, stream))
        out = torch.empty(dst.shape, dtype=torch.float32, device="cuda")
        gi = torch.empty(dst.shape, dtype=torch.int32, device="cuda")
        _lib.check(lib.b200_nn_query_gather_f32(h, C.byref(gd), 5000.0, data.data_ptr(), out.data_ptr(), 1, 0, 0,
                                                float("nan"), stream))
        _lib.check(lib.b200_nn_query(h, C.byref(gd), 5000.0, None, gi.data_ptr(), None, stream))
        nbytes = C.c_int64(0)
        _lib.check(lib.b200_nn_grid_nbytes(h, C.byref(nbytes)))
        if light:
            ia = torch.empty(dst.shape, dtype=torch.int32, device="cuda")
            assert lib.b200_nn_query(h, C.byref(gd), 5000.0, ia.data_ptr(), None, None, stream) == _lib.B200_EUNSUPPORTED
            nv = C.c_int64(0)
            assert lib.b200_nn_grid_n_valid(h, C.byref(nv)) == _lib.B200_EUNSUPPORTED
        torch.cuda.synchronize()
        res[light] = (out, gi, vin, nbytes.value)
        lib.b200_nn_grid_destroy(h)
    (o0, g0, v0, b0), (o1, g1, v1, b1) = res[0], res[NN_LIGHT]
    assert b1 == src.size * 16 and b0 == src.size * 52
    assert bool((g0 == g1).all()) and bool((v0 == v1).all())
    assert bool(((o0 == o1) | (torch.isnan(o0) & torch.isnan(o1))).all())
    assert 0.3 < float((g0 >= 0).float().mean()) < 0.9
