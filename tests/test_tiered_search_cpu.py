"""CPU evidence for the two claims the tiered fixed-grid search of csrc/nn_geos.cu rests on (no GPU needed; the GPU
suite checks the kernel itself against its strict mode on every pixel of config 5):

* the satellite-distance factor ``kf`` never exceeds rho / h of any candidate near the target's direction, so the
  sharpened distance bound is a bound (``scripts/rho_factor_check.py``);
* a numpy restatement of the tiers (2x2 block, 4x4 window, float32 ranking keys, acceptance margins) settles its
  targets exactly as a kd-tree does, at the limb and near nadir, and with ``kf`` no target of config 5 needs the general
  loop (``scripts/tiered_search_study.py``).
"""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "scripts", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("px_angle", [14e-6, 56e-6, 250e-6])     # AHI 0.5 km, ABI 1 km, the largest admitted
def test_satellite_distance_factor_is_a_lower_bound(px_angle):
    chk = _load("rho_factor_check")
    assert (chk.K_A, chk.K_B, chk.K_CAP) == (0.36, 0.9999, 1.10)          # the constants of nn_fixed_grid_query_impl
    assert chk.worst_ratio(px_angle, disc_px=3.0, n_theta=60001) >= 1.00005


def test_tiered_search_prototype_matches_the_kd_tree():
    study = _load("tiered_search_study")
    limb = study.run_tile(0.055, nrows=4, verbose=False)        # 80 degrees north: pixels stretched 10x and more
    mid = study.run_tile(0.30, nrows=4, verbose=False)          # 36 degrees north: tier 1 settles 85 %
    for info in (limb, mid):
        assert info["visible"] > 30000
        assert info["mismatches"] == 0 and info["hidden_hits"] == 0
        assert info["general"] == 0.0                            # the 4x4 window always suffices with kf ...
        assert info["settled"] > 0.99
    assert limb["general_without_kf"] > 0.1                      # ... and would not without it
    assert limb["tier2"] > 0.9 and mid["tier1"] > 0.8
