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
