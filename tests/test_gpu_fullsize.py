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
