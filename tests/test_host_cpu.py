"""CPU-only checks: the C ABI library loads and exports every declared symbol, ctypes mirrors the C structs,
host-side geometry / calibration logic, the row-tile decomposition and a world_size-2 gloo run of the all-gather."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "satpy_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib_built):
    from satpy_b200 import _lib
    declared = _declared_functions()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib_built, name), f"{name} declared in include/satpy_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "satpy_b200/_lib.py SIGNATURES and the header disagree"
    assert b"sm_100a" in lib_built.b200_version()


def test_library_is_sm100a_only(lib_built):
    from satpy_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_ctypes_structs_match_the_c_header(tmp_path):
    """sizeof/offsetof printed by a C program compiled against the header == the ctypes mirror."""
    from satpy_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "satpy_b200.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu\n", sizeof(b200_proj_t), sizeof(b200_area_t), sizeof(b200_geo_t),
           sizeof(b200_abi_cal_t), sizeof(b200_sunz_t));
    printf("%zu %zu %zu %zu %zu\n", offsetof(b200_proj_t, p), offsetof(b200_area_t, x_ul), offsetof(b200_geo_t, lons),
           offsetof(b200_geo_t, row0), offsetof(b200_abi_cal_t, fk1));
    return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    lines = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    sizes = [int(v) for v in lines[0].split()]
    offs = [int(v) for v in lines[1].split()]
    assert sizes == [C.sizeof(_lib.ProjStruct), C.sizeof(_lib.AreaStruct), C.sizeof(_lib.GeoStruct),
                     C.sizeof(_lib.AbiCalStruct), C.sizeof(_lib.SunzStruct)]
    assert offs == [_lib.ProjStruct.p.offset, _lib.AreaStruct.x_ul.offset, _lib.GeoStruct.lons.offset,
                    _lib.GeoStruct.row0.offset, _lib.AbiCalStruct.fk1.offset]


def test_no_cpu_fallback(lib_built):
    """Without a CUDA device every compute path raises instead of silently running elsewhere."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from satpy_b200 import demo
    from satpy_b200.readers import calibrate_abi
    from satpy_b200.resample import B200NearestResampler
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        calibrate_abi(np.zeros(8, np.int16), demo.abi_calibration("C02"), "radiance")
    src, dst = demo.make_area("goes_east_abi_f_2km", 0.01), demo.make_area("goes_east_eqc_2km", 0.01)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200NearestResampler(src, dst).precompute(radius_of_influence=1e5)
    assert lib_built.b200_device_count(C.byref(C.c_int(0))) in (0, 2)  # B200_OK with 0 devices or B200_ECUDA


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "satpy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"


PROJS = [
    {"proj": "geos", "sweep": "x", "lon_0": -75.0, "h": 35786023.0, "ellps": "GRS80"},
    {"proj": "geos", "lon_0": 0.0, "a": 6378169.0, "b": 6356583.8, "h": 35785831.0},
    {"proj": "eqc", "lon_0": 0.0, "lat_ts": 30.0, "ellps": "WGS84"},
    {"proj": "merc", "lat_ts": 45.0, "ellps": "WGS84"},
    {"proj": "laea", "lat_0": 52.0, "lon_0": 10.0, "x_0": 4321000.0, "y_0": 3210000.0, "ellps": "GRS80"},
    {"proj": "laea", "lat_0": -90.0, "lon_0": 0.0, "ellps": "WGS84"},
    {"proj": "lcc", "lat_1": 30.0, "lat_2": 60.0, "lat_0": 38.0, "lon_0": 126.0, "ellps": "WGS84"},
    {"proj": "stere", "lat_0": 90.0, "lat_ts": 60.0, "lon_0": 0.0, "ellps": "WGS84"},
]


@pytest.mark.parametrize("proj", PROJS, ids=[p["proj"] + str(i) for i, p in enumerate(PROJS)])
def test_host_projection_constants_match_oracle(proj):
    """satpy_b200.geometry derives the device constants on its own (it cannot import the oracle): same numbers."""
    from oracle.projections import Proj
    from satpy_b200.geometry import proj_struct
    S, O = proj_struct(proj), Proj(proj)
    assert S.a == O.a and S.es == pytest.approx(O.es, abs=1e-18) and S.lam0 == O.lam0
    name = proj["proj"]
    if name == "geos":
        assert [S.p[i] for i in range(6)] == pytest.approx(
            [O.radius_g_1, O.radius_g, O.c_geos, O.radius_p, O.radius_p2, O.radius_p_inv2], rel=1e-15)
        assert S.mode == (1 if O.flip_axis else 0)
    elif name == "eqc":
        assert S.p[0] == pytest.approx(O.rc, rel=1e-15)
    elif name == "merc":
        assert S.k0 == pytest.approx(O.k0, rel=1e-15)
    elif name == "lcc":
        assert [S.p[0], S.p[1], S.p[2]] == pytest.approx([O.n, O.c, O.rho0], rel=1e-14)
    elif name == "stere":
        assert S.p[0] == pytest.approx(O.akm1, rel=1e-15)
    elif name == "laea":
        assert [S.p[i] for i in range(7)] == pytest.approx([O.qp, O.rq, O.sinb1, O.cosb1, O.dd, O.xmf, O.ymf], rel=1e-14)


def test_proj_string_and_unsupported_projection():
    from satpy_b200.geometry import AreaDefinition, proj_struct
    a = AreaDefinition("t", "t", "t", "+proj=geos +lon_0=-95.0 +h=35786023.0 +a=6378137.0 +b=6356752.31414 +sweep=x "
                                     "+units=m +no_defs", 10, 12, (-1000., -1500., 1000., 1500.))
    assert a.shape == (12, 10) and a.pixel_size_x == 200.0 and a.pixel_size_y == 250.0
    assert a.proj_dict["sweep"] == "x" and a.geo_struct().area.proj.mode == 1
    sub = a[slice(2, 6), slice(1, 4)]
    assert sub.shape == (4, 3) and sub.area_extent == (-800.0, 0.0, -200.0, 1000.0)
    assert hash(a) == hash(AreaDefinition("x", "", "", a.proj_dict, 10, 12, a.area_extent)) and a != sub
    with pytest.raises(NotImplementedError):
        proj_struct({"proj": "tmerc", "lon_0": 9.0})


def test_abi_calibration_struct_follows_reference_rounding():
    """float32 rounding of the constants exactly where the reference rounds them (abi_l1b.py:133-173, core/abi.py:172-173)."""
    from satpy_b200.readers import ABICalibration
    cal = ABICalibration(scale_factor=0.5, add_offset=-1.3, fill_value=np.int16(1002), planck_fk1=13432.1,
                         planck_fk2=1497.61, planck_bc1=0.09102, planck_bc2=0.99971, clip_negative_radiances=True)
    s = cal.struct("brightness_temperature")
    assert s.mode == 2 and s.has_fill == 1 and s.fill == 1002
    assert s.fk1 == float(np.float32(13432.1)) and s.bc2 == float(np.float32(0.99971))
    assert s.min_rad == pytest.approx(0.2, abs=1e-6)  # ceil(1.3 / 0.5) * 0.5 - 1.3
    vis = ABICalibration(scale_factor=0.5, add_offset=-1.0, fill_value=65535, unsigned=True, esun=2017,
                         earth_sun_distance_anomaly_in_AU=0.99).struct("reflectance")
    assert vis.fill == -1 and vis.is_unsigned == 1  # stored bit pattern of 65535 as int16
    assert vis.refl_factor == float(np.float32(np.pi * 0.99 * 0.99 / 2017))
    with pytest.raises(ValueError):
        cal.struct("reflectance")


def test_modifier_attrs_and_registration_without_satpy():
    import satpy_b200
    from satpy_b200.modifiers import SunZenithCorrector
    from satpy_b200.resample import get_resampler_classes
    m = SunZenithCorrector(name="sunz", prerequisites=[], optional_prerequisites=["solar_zenith_angle"], max_sza=None)
    assert m.max_sza is None and m.max_sza_cos is None and m.correction_limit == 88.0
    assert m.attrs["optional_prerequisites"] == ["solar_zenith_angle"]
    assert set(get_resampler_classes()) >= {"b200_nearest"}
    assert satpy_b200.register() is False  # satpy is not installed in this image
    import yaml

    class _L(yaml.SafeLoader):
        pass
    _L.add_multi_constructor("tag:yaml.org,2002:python/name:", lambda loader, suffix, node: suffix)
    cfg = yaml.load(open(os.path.join(satpy_b200.ETC_DIR, "composites", "visir.yaml")), Loader=_L)
    assert cfg["modifiers"]["sunz_corrected"]["modifier"] == "satpy_b200.modifiers.SunZenithCorrector"


def test_row_tiles():
    from satpy_b200.parallel import row_tiles
    tile, tiles = row_tiles(40075, 8)
    assert tile == 5010 and len(tiles) == 8
    assert tiles[0] == (0, 5010) and tiles[7] == (35070, 5005)
    assert sum(n for _, n in tiles) == 40075
    tile, tiles = row_tiles(5, 8)
    assert tile == 1 and tiles[5:] == [(5, 0), (5, 0), (5, 0)]


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from satpy_b200.parallel import row_tiles, all_gather_rows
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
height, width = 11, 6
tile, tiles = row_tiles(height, world)
full = torch.full((world * tile, width), -1.0)
r0, n = tiles[rank]
full[rank * tile: rank * tile + n] = torch.arange(r0 * width, (r0 + n) * width, dtype=torch.float32).reshape(n, width)
all_gather_rows(full, tile, rank)
expect = torch.arange(height * width, dtype=torch.float32).reshape(height, width)
assert torch.equal(full[:height], expect), (rank, full)
dist.barrier()
if rank == 0:
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_ewa_scan_bands():
    from satpy_b200.resample.ewa import scan_band, scan_bands
    assert scan_bands(7680, 16, 8) == [(i * 960, 960) for i in range(8)]
    bands = scan_bands(16 * 10, 16, 4)
    assert bands == [(0, 48), (48, 48), (96, 48), (144, 16)] and sum(n for _, n in bands) == 160
    assert scan_bands(32, 16, 4) == [(0, 16), (16, 16), (32, 0), (32, 0)]
    assert scan_bands(100, 0, 2) == [(0, 100), (100, 0)]          # rows_per_scan = 0: the swath is one scan
    assert scan_band(160, 16, None) == (0, 160)
    assert scan_band(160, 16, (32, 64)) == (32, 64)
    for bad in ((8, 16), (16, 24), (-16, 16), (144, 32)):
        with pytest.raises(ValueError):
            scan_band(160, 16, bad)
    with pytest.raises(ValueError):
        scan_band(160, 0, (0, 80))


_EWA_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from oracle import ewa as oewa
from oracle.projections import Area
from satpy_b200 import demo
from satpy_b200.resample.ewa import scan_bands
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lon, lat, rps = demo.viirs_like_swath(nscans=6, ncols=120)
proj = {"proj": "stere", "lat_0": 90.0, "lat_ts": 60.0, "lon_0": 0.0, "ellps": "WGS84"}
dst = Area(proj, 60, 60, (-1.2e6, -2.4e6, 0.3e6, -0.9e6))
_, cols, rows = oewa.ll2cr(lon, lat, dst)
data = np.random.default_rng(2).uniform(0, 100, lon.shape).astype(np.float32)
row0, nrows = scan_bands(lon.shape[0], rps, world)[rank]
sl = slice(row0, row0 + nrows)
w, a = oewa.fornav_accumulate(cols[sl], rows[sl], (60, 60), data[sl], rows_per_scan=rps)
tw, ta = torch.from_numpy(w.astype(np.float64)), torch.from_numpy(a.astype(np.float64))
dist.all_reduce(tw); dist.all_reduce(ta)
w_all, a_all = oewa.fornav_accumulate(cols, rows, (60, 60), data, rows_per_scan=rps)
np.testing.assert_allclose(tw.numpy(), w_all, rtol=1e-5, atol=1e-6)
np.testing.assert_allclose(ta.numpy(), a_all, rtol=1e-5, atol=1e-4)
dist.barrier()
if rank == 0:
    print("EWA_GLOO_OK", int((w_all > 0).sum()))
dist.destroy_process_group()
'''


def test_ewa_scan_bands_sum_to_the_whole_world_size_2_gloo(tmp_path):
    """Multi-GPU EWA: ranks accumulate bands of whole scans and SUM their grids; checked with the
    oracle's accumulation on two gloo ranks."""
    script = tmp_path / "worker.py"
    script.write_text(_EWA_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29633", str(script), ROOT],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "EWA_GLOO_OK" in out.stdout


_EWA_MAX_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from oracle import ewa as oewa
from oracle.projections import Area
from satpy_b200 import demo
from satpy_b200.resample.ewa import scan_bands, max_weight_encode, max_weight_decode
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lon, lat, rps = demo.viirs_like_swath(nscans=6, ncols=120, bowtie=0.3)
proj = {"proj": "stere", "lat_0": 90.0, "lat_ts": 60.0, "lon_0": 0.0, "ellps": "WGS84"}
dst = Area(proj, 60, 60, (-1.2e6, -2.4e6, 0.3e6, -0.9e6))
_, cols, rows = oewa.ll2cr(lon, lat, dst)
rng = np.random.default_rng(2)
data = rng.uniform(0, 100, lon.shape).astype(np.float32)
data[rng.random(lon.shape) < 0.05] = np.nan
row0, nrows = scan_bands(lon.shape[0], rps, world)[rank]
sl = slice(row0, row0 + nrows)
# what one rank holds after its two passes: the maxima of ITS weights, then -- given the maxima over all ranks -- the
# samples that reach them (-inf where none of its samples does, NaN for an invalid sample)
_, _, w, a = oewa.fornav(cols[sl], rows[sl], (60, 60), data[sl], rows_per_scan=rps, maximum_weight_mode=True,
                         return_grids=True)
tw = torch.from_numpy(w[0].copy())
dist.all_reduce(tw, op=dist.ReduceOp.MAX)
mine = torch.from_numpy(a[0].copy())
mine = torch.where(torch.from_numpy(w[0]) == tw, mine, torch.full_like(mine, float("-inf")))
mine = torch.where(tw > 0, mine, torch.full_like(mine, float("-inf")))
enc = max_weight_encode(mine)
dist.all_reduce(enc, op=dist.ReduceOp.MAX)
merged = max_weight_decode(enc).numpy()
_, _, w_all, a_all = oewa.fornav(cols, rows, (60, 60), data, rows_per_scan=rps, maximum_weight_mode=True,
                                 return_grids=True)
np.testing.assert_array_equal(tw.numpy(), w_all[0])
hit = w_all[0] > 0
same = (merged[hit] == a_all[0][hit]) | (np.isnan(merged[hit]) & np.isnan(a_all[0][hit]))
assert same.mean() > 0.995, same.mean()          # equal maximum weights in two bands: either sample is "the" maximum
assert np.isnan(merged[~hit]).all()
dist.barrier()
if rank == 0:
    print("EWA_MAX_GLOO_OK", int(hit.sum()), float(same.mean()))
dist.destroy_process_group()
'''


def test_ewa_maximum_weight_mode_reduces_by_weight_world_size_2_gloo(tmp_path):
    """Multi-GPU EWA in ``maximum_weight_mode``: MAX-reduce the weight grids, keep the samples that reach
    the global maxima, MAX-reduce the encoded value grids (``max_weight_encode`` / ``_decode``); checked with the
    oracle's grids on two gloo ranks against the whole-swath result."""
    script = tmp_path / "worker.py"
    script.write_text(_EWA_MAX_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29635", str(script), ROOT],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "EWA_MAX_GLOO_OK" in out.stdout


def test_peer_exchange_rounds_are_collision_free():
    """PeerImage's schedule: in every round the destinations form a permutation without fixed points, and over the
    world_size - 1 rounds every rank reaches every peer exactly once."""
    from satpy_b200.parallel import round_peer
    for world in (2, 3, 4, 8):
        reached = {r: set() for r in range(world)}
        for k in range(1, world):
            dests = [round_peer(r, k, world) for r in range(world)]
            assert sorted(dests) == list(range(world))
            assert all(d != r for r, d in enumerate(dests))
            for r, d in enumerate(dests):
                reached[r].add(d)
        assert all(reached[r] == set(range(world)) - {r} for r in range(world))
    with pytest.raises(ValueError):
        round_peer(0, 0, 4)
    with pytest.raises(ValueError):
        round_peer(0, 4, 4)


def test_all_gather_rows_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", str(script), ROOT],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm), at a debug scale: one JSON line
    with the contract's keys, `impl`, a `cpu_baseline` describing the run and an `e2e` equal to the line's own value."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "0.05",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "Mpix/s" and d["value"] > 0 and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_gpu_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--scale", "0.05", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)


# ----------------------------------------------------------------------------- boundary: stub satpy / pyresample
def _install_stub_satpy(monkeypatch):
    """Minimal stand-ins for the satpy / pyresample surface the plug-in touches (neither can be installed here):
    ``pyresample.resampler.BaseResampler``, ``satpy.modifiers.ModifierBase`` with the ``id`` property the dependency
    tree reads (satpy/node.py:163), ``satpy.composites.core.IncompatibleAreas``, ``satpy.resample.base`` with
    ``RESAMPLER_MODULES`` / ``get_all_resampler_classes`` (base.py:77-105), ``satpy.config``."""
    import importlib
    import sys
    import types

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        monkeypatch.setitem(sys.modules, name, m)
        return m

    class BaseResampler:
        def __init__(self, source_geo_def, target_geo_def):
            self.source_geo_def, self.target_geo_def = source_geo_def, target_geo_def

        def _create_cache_filename(self, cache_dir, prefix="", fmt=".zarr", **kwargs):
            return os.path.join(cache_dir, prefix + "stubhash" + fmt)

    class IncompatibleAreas(Exception):
        pass

    class ModifierBase:
        def __init__(self, name, prerequisites=None, optional_prerequisites=None, **kwargs):
            kwargs.update(name=name, prerequisites=prerequisites or [], optional_prerequisites=optional_prerequisites or [])
            self.attrs = kwargs

        @property
        def id(self):  # noqa: A003
            return self.attrs["_satpy_id"]

        def apply_modifier_info(self, origin, destination):
            destination.attrs["modifiers"] = self.attrs.get("modifiers")

    class Config(dict):
        def set(self, **kw):  # noqa: A003
            self.update(kw)

    pr = mod("pyresample")
    pr.resampler = mod("pyresample.resampler", BaseResampler=BaseResampler)
    modules = ["satpy.resample.kdtree"]

    def get_all_resampler_classes():
        out = {}
        for name in modules:
            if name.startswith("satpy_b200"):
                out.update(importlib.import_module(name).get_resampler_classes())
        return out
    sp = mod("satpy", config=Config(config_path=[]))
    sp.resample = mod("satpy.resample")
    sp.resample.base = mod("satpy.resample.base", RESAMPLER_MODULES=modules,
                           get_all_resampler_classes=get_all_resampler_classes)
    sp.modifiers = mod("satpy.modifiers", ModifierBase=ModifierBase)
    sp.composites = mod("satpy.composites")
    sp.composites.core = mod("satpy.composites.core", IncompatibleAreas=IncompatibleAreas)
    return BaseResampler, ModifierBase, IncompatibleAreas, get_all_resampler_classes


def test_register_and_base_classes_against_stub_satpy(monkeypatch):
    """With satpy / pyresample importable (stubs): ``register()`` runs its whole body, the resampler classes ARE
    ``BaseResampler``s (the ``isinstance`` check of satpy/resample/base.py:151-156), the modifiers ARE ``ModifierBase``s
    with a working ``id``, and ``IncompatibleAreas`` is satpy's own exception type."""
    import importlib
    import sys
    BaseResampler, ModifierBase, IncompatibleAreas, get_all = _install_stub_satpy(monkeypatch)
    saved = {k: sys.modules[k] for k in list(sys.modules) if k == "satpy_b200" or k.startswith("satpy_b200.")}
    try:
        for k in saved:
            del sys.modules[k]
        pkg = importlib.import_module("satpy_b200")
        assert pkg.register() is True
        import satpy
        from satpy.resample import base
        assert "satpy_b200.resample" in base.RESAMPLER_MODULES
        assert satpy.config.get("config_path") == []            # modifiers are opt-in
        assert pkg.register(modifiers=True) is True
        assert pkg.ETC_DIR in satpy.config["config_path"]
        classes = get_all()
        assert {"b200_nearest", "b200_bilinear", "b200_ewa", "b200_native"} <= set(classes)
        src = pkg.AreaDefinition("s", "s", "s", {"proj": "geos", "lon_0": 0.0, "h": 35785831.0, "ellps": "WGS84"}, 8, 8,
                                 (-4e6, -4e6, 4e6, 4e6))
        dst = pkg.AreaDefinition("d", "d", "d", {"proj": "eqc", "lon_0": 0.0, "ellps": "WGS84"}, 4, 4, (-2e6, -2e6, 2e6, 2e6))
        for name in ("b200_nearest", "b200_bilinear", "b200_ewa", "b200_native"):
            r = classes[name](src, dst)
            assert isinstance(r, BaseResampler), name
        r = classes["b200_nearest"](src, dst)
        assert r._cache_filename("/tmp/c", 5000.0).endswith("nn_lut-stubhash.zarr")   # pyresample's own naming
        mods = importlib.import_module("satpy_b200.modifiers")
        m = mods.SunZenithCorrector(name="sunz_corrected", prerequisites=[], modifiers=("sunz_corrected",), _satpy_id="ID")
        assert isinstance(m, ModifierBase) and m.id == "ID" and m.max_sza == 95.0 and m.correction_limit == 88.0
        assert mods.IncompatibleAreas is IncompatibleAreas
        red = mods.SunZenithReducer(name="sunz_reduced", _satpy_id="ID2")
        assert isinstance(red, ModifierBase) and red.id == "ID2"
    finally:
        for k in [k for k in sys.modules if k == "satpy_b200" or k.startswith("satpy_b200.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_dataarraylite_rewrap_carries_coordinates():
    """What ``_update_resampled_coords`` does (satpy/resample/base.py:45-73): non-spatial coordinates and the name are
    carried over, x / y come from the target area; modifiers (same grid) keep everything."""
    from satpy_b200 import AreaDefinition, DataArrayLite
    from satpy_b200.dataset import rewrap
    dst = AreaDefinition("d", "d", "d", {"proj": "eqc", "lon_0": 0.0, "ellps": "WGS84"}, 5, 4, (-2.5e6, -2e6, 2.5e6, 2e6))
    src = DataArrayLite(np.zeros((3, 6, 7), np.float32), ("bands", "y", "x"), {"name": "rgb"},
                        coords={"bands": np.array(["R", "G", "B"]), "x": np.arange(7.0), "y": np.arange(6.0), "crs": "old"},
                        name="rgb")
    res = rewrap(src, np.zeros((3, 4, 5), np.float32), attrs={"area": dst}, new_area=dst)
    assert res.name == "rgb" and list(res.coords["bands"]) == ["R", "G", "B"] and "crs" not in res.coords
    x, y = dst.get_proj_vectors()
    np.testing.assert_array_equal(res.coords["x"], x)
    np.testing.assert_array_equal(res.coords["y"], y)
    same = rewrap(src, np.ones((3, 6, 7), np.float32))
    assert set(same.coords) == {"bands", "x", "y", "crs"} and same.name == "rgb"
    band = rewrap(src, np.zeros((3, 2, 5), np.float32), attrs={}, new_area=dst)   # a row window: no x / y invented
    assert "x" not in band.coords and "bands" in band.coords


def test_zarr2_store_round_trip(tmp_path):
    """The hand-written zarr-v2 layout (satpy/resample/kdtree.py:158-177 writes it through xarray/zarr): .zgroup,
    per-array .zarray / .zattrs with _ARRAY_DIMENSIONS, raw C-order chunks named i.j.k; zlib chunks are readable too."""
    import json
    import zlib
    from satpy_b200 import _zarr2
    rng = np.random.default_rng(0)
    idx = rng.integers(0, 5000, size=(301, 77, 1)).astype(np.int64)
    vin = rng.random((40, 50)) > 0.3
    path = str(tmp_path / "nn_lut-x.zarr")
    old = _zarr2._TARGET_CHUNK_BYTES
    _zarr2._TARGET_CHUNK_BYTES = 40000
    try:
        _zarr2.write_store(path, {"index_array": (("y2", "x2", "z2"), idx), "valid_input_index": (("y1", "x1"), vin)})
    finally:
        _zarr2._TARGET_CHUNK_BYTES = old
    meta = json.load(open(os.path.join(path, "index_array", ".zarray")))
    assert meta["zarr_format"] == 2 and meta["compressor"] is None and meta["order"] == "C" and meta["dtype"] == "<i8"
    assert meta["shape"] == [301, 77, 1] and meta["chunks"][1:] == [77, 1] and meta["chunks"][0] < 301
    assert os.path.exists(os.path.join(path, "index_array", "0.0.0")) and os.path.exists(os.path.join(path, ".zmetadata"))
    np.testing.assert_array_equal(_zarr2.read_array(path, "index_array"), idx)
    np.testing.assert_array_equal(_zarr2.read_array(path, "valid_input_index"), vin)
    assert _zarr2.array_dims(path, "index_array") == ("y2", "x2", "z2")
    # a zlib-compressed chunk, as another writer might leave it
    meta["compressor"] = {"id": "zlib", "level": 1}
    json.dump(meta, open(os.path.join(path, "index_array", ".zarray"), "w"))
    for f in os.listdir(os.path.join(path, "index_array")):
        if not f.startswith("."):
            fn = os.path.join(path, "index_array", f)
            raw = open(fn, "rb").read()
            open(fn, "wb").write(zlib.compress(raw, 1))
    np.testing.assert_array_equal(_zarr2.read_array(path, "index_array"), idx)
    with pytest.raises(IOError):
        _zarr2.read_array(path, "nope")


def test_block_column_span_is_conservative():
    """Outside ``block_column_span`` no target pixel has a neighbour (checked against the oracle's search on a
    down-scaled config 5), and the spans cover well under half of the global grid."""
    from oracle import nearest as onn
    from oracle.projections import Area as OArea
    from satpy_b200 import demo, parallel
    s_scale, d_scale = 300 / 10848, 1202 / 80150
    src, dst = demo.make_area("goes_east_abi_f_1km", s_scale), demo.make_area("global_eqc_500m", d_scale)
    o_src = OArea(*demo.area_params("goes_east_abi_f_1km", s_scale))
    o_dst = OArea(*demo.area_params("global_eqc_500m", d_scale))
    radius = 2.5 * 1002 / s_scale
    vin, _, idx = onn.get_neighbour_info(*o_src.get_lonlats(), *o_dst.get_lonlats(), radius)
    hit = idx[..., 0] < vin.sum()
    total = 0
    for row0 in range(0, dst.height, 37):
        n = min(37, dst.height - row0)
        c_lo, c_hi = parallel.block_column_span(src, dst, row0, n, radius)
        assert c_lo % 32 == 0 and (c_hi % 32 == 0 or c_hi == dst.width)
        assert not hit[row0:row0 + n, :c_lo].any() and not hit[row0:row0 + n, c_hi:].any()
        total += (c_hi - c_lo) * n
    assert hit.mean() < total / dst.size < 0.5
    # pairs the test does not know: the whole width
    laea = demo.make_area("euro_laea_1km", 0.05)
    assert parallel.block_column_span(src, laea, 0, 10, radius) == (0, laea.width)
    assert parallel.block_column_span(laea, dst, 0, 10, radius) == (0, dst.width)


# ------------------------------------------------------------------ host side of the transfers (no device needed)
def test_fill_async_writes_exactly_the_outside_of_the_spans():
    """``_xfer.fill_async`` (the library's host helper with non-temporal stores, a few batched tasks per thread): the
    columns outside every band's span get the fill value, the spans and the rows of no band are not touched."""
    import torch
    from satpy_b200 import _xfer
    rng = np.random.default_rng(3)
    nb, rows, width = 2, 97, 1013
    host = torch.from_numpy(rng.random((nb, rows, width), dtype=np.float32))
    before = host.numpy().copy()
    bands = [(0, 10, (100, 900)), (10, 30, (0, 0)), (40, 7, (0, width)), (47, 40, (3, 1013)), (87, 5, (512, 513))]
    _xfer.fill_async(host, [b for b in bands if b[2] != (0, width)], float("nan")).result()
    got = host.numpy()
    expect = before.copy()
    for r0, n, (lo, hi) in bands:
        if hi <= lo:
            expect[:, r0:r0 + n] = np.nan
        else:
            expect[:, r0:r0 + n, :lo] = np.nan
            expect[:, r0:r0 + n, hi:] = np.nan
    np.testing.assert_array_equal(got, expect)          # rows 92..96 belong to no band: untouched
    assert np.isfinite(got[:, 92:]).all()
    # integer images take the generic path
    u8 = torch.zeros((1, 16, 64), dtype=torch.uint8)
    _xfer.fill_async(u8, [(0, 16, (8, 24))], 255).result()
    assert (u8.numpy()[0, :, :8] == 255).all() and (u8.numpy()[0, :, 24:] == 255).all() and (u8.numpy()[0, :, 8:24] == 0).all()


def test_threaded_host_copy_is_a_memcpy():
    from satpy_b200 import _xfer
    a = np.random.default_rng(4).integers(0, 255, size=(9 << 20) + 13, dtype=np.uint8)
    b = np.zeros_like(a)
    _xfer._host_copy(b.ctypes.data, a.ctypes.data, a.nbytes)
    np.testing.assert_array_equal(a, b)
    c = np.zeros(1000, np.uint8)
    _xfer._host_copy(c.ctypes.data, a.ctypes.data, 1000)       # small: single memmove
    np.testing.assert_array_equal(c, a[:1000])


def test_row_bands_cover_any_width_on_multiples_of_four_rows():
    """Host results are produced in row bands whatever the width (config 5's 80150 columns are not a multiple of 4):
    bands tile the rows exactly and start on multiples of four rows (16-byte boundaries of the output rows)."""
    from satpy_b200 import demo
    from satpy_b200.resample import B200NearestResampler
    src, dst = demo.make_area("goes_east_abi_f_1km"), demo.make_area("global_eqc_500m")
    r = B200NearestResampler(src, dst)
    for nrows in (dst.height, 5010, 1253, 3, 0):
        bands = r._row_bands(nrows, True)
        assert sum(n for _, n in bands) == nrows
        assert all(r0 % 4 == 0 for r0, _ in bands)
        assert [r0 for r0, _ in bands] == list(np.cumsum([0] + [n for _, n in bands])[:-1])
    assert len(r._row_bands(dst.height, True)) == 24 and len(r._row_bands(dst.height, False)) == 1
