"""ctypes binding of ``libsatpy_b200.so`` (the C ABI declared in ``include/satpy_b200.h``).

There is no CPU fallback: if the shared library is missing or a call fails the
caller gets an exception (``B200LibraryError`` / ``ValueError`` /
``RuntimeError`` / ``MemoryError``), never a silently different code path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsatpy_b200.so")

B200_OK, B200_EINVAL, B200_ECUDA, B200_ENOMEM, B200_EUNSUPPORTED = range(5)

PROJ_LONGLAT, PROJ_EQC, PROJ_MERC, PROJ_GEOS, PROJ_LAEA, PROJ_LCC, PROJ_STERE = range(7)


class B200LibraryError(ImportError):
    """The CUDA library could not be loaded (not built, or CUDA runtime missing)."""


class ProjStruct(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mode", C.c_int32),
                ("a", C.c_double), ("es", C.c_double), ("e", C.c_double), ("one_es", C.c_double),
                ("lam0", C.c_double), ("phi0", C.c_double), ("x0", C.c_double), ("y0", C.c_double),
                ("k0", C.c_double), ("p", C.c_double * 12)]


class AreaStruct(C.Structure):
    _fields_ = [("proj", ProjStruct), ("width", C.c_int64), ("height", C.c_int64),
                ("pixel_size_x", C.c_double), ("pixel_size_y", C.c_double),
                ("x_ul", C.c_double), ("y_ul", C.c_double)]


class GeoStruct(C.Structure):
    _fields_ = [("is_swath", C.c_int32), ("lonlat_is_f32", C.c_int32), ("area", AreaStruct),
                ("lons", C.c_void_p), ("lats", C.c_void_p),
                ("width", C.c_int64), ("height", C.c_int64), ("row0", C.c_int64), ("nrows", C.c_int64)]


class AbiCalStruct(C.Structure):
    _fields_ = [("mode", C.c_int32), ("is_unsigned", C.c_int32), ("has_fill", C.c_int32), ("fill", C.c_int32),
                ("clip_negative_radiances", C.c_int32),
                ("scale_factor", C.c_float), ("add_offset", C.c_float), ("refl_factor", C.c_float),
                ("min_rad", C.c_float), ("fk1", C.c_float), ("fk2", C.c_float), ("bc1", C.c_float),
                ("bc2", C.c_float)]


class SunzStruct(C.Structure):
    _fields_ = [("gmst", C.c_double), ("ra", C.c_double), ("dec", C.c_double),
                ("limit_deg", C.c_double), ("max_sza_deg", C.c_double),
                ("has_max_sza", C.c_int32), ("lonlat_f32", C.c_int32), ("method", C.c_int32),
                ("reserved", C.c_int32)]


class AhiCalStruct(C.Structure):
    _fields_ = [("mode", C.c_int32), ("has_invalid", C.c_int32), ("count_outside_scan", C.c_int32),
                ("count_error", C.c_int32), ("bt_in_f64", C.c_int32), ("reserved", C.c_int32),
                ("gain", C.c_float), ("offset", C.c_float), ("coeff", C.c_float), ("pad", C.c_float),
                ("a", C.c_double), ("num", C.c_double), ("cwl5", C.c_double),
                ("c0", C.c_double), ("c1", C.c_double), ("c2", C.c_double)]


class SeviriCalStruct(C.Structure):
    _fields_ = [("mode", C.c_int32), ("reserved", C.c_int32), ("gain", C.c_float), ("offset", C.c_float),
                ("pi_f", C.c_float), ("solar_irradiance", C.c_float), ("sun_earth_dist2", C.c_float),
                ("c1", C.c_float), ("nu3", C.c_float), ("c2nu", C.c_float), ("alpha", C.c_float), ("beta", C.c_float),
                ("fit_a", C.c_float), ("fit_b", C.c_float), ("fit_c", C.c_float)]


class EwaParamsStruct(C.Structure):
    _fields_ = [("weight_count", C.c_int32), ("maximum_weight_mode", C.c_int32), ("weight_min", C.c_float),
                ("weight_distance_max", C.c_float), ("weight_delta_max", C.c_float), ("weight_sum_min", C.c_float)]


_P = C.c_void_p
_GEO = C.POINTER(GeoStruct)

# name -> (restype, argtypes); must list EVERY symbol include/satpy_b200.h declares
SIGNATURES = {
    "b200_version": (C.c_char_p, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "b200_area_lonlats": (C.c_int, [_GEO, _P, _P, _P]),
    "b200_abi_calibrate": (C.c_int, [_P, _P, C.c_int64, C.POINTER(AbiCalStruct), _P]),
    "b200_abi_calibrate_variant": (C.c_int, [_P, _P, C.c_int64, C.POINTER(AbiCalStruct), C.c_int32, _P]),
    "b200_ahi_calibrate": (C.c_int, [_P, _P, C.c_int64, C.POINTER(AhiCalStruct), _P]),
    "b200_seviri_calibrate": (C.c_int, [_P, C.c_int32, _P, C.c_int64, C.POINTER(SeviriCalStruct), _P]),
    "b200_viirs_scale": (C.c_int, [_P, C.c_int32, _P, C.c_int64, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                   C.c_int32, C.c_int32, C.c_float, _P]),
    "b200_get_angles": (C.c_int, [_GEO, C.POINTER(SunzStruct), C.c_double, C.c_double, C.c_double, _P, _P, _P, _P, _P]),
    "b200_relative_azimuth": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, _P]),
    "b200_sunz_reduce": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float, _P]),
    "b200_sunz_reduce_area": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _GEO, C.POINTER(SunzStruct), C.c_float, C.c_float,
                                        C.c_float, _P]),
    "b200_cos_sza": (C.c_int, [_GEO, C.POINTER(SunzStruct), _P, _P]),
    "b200_sunz_correct": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _GEO, C.POINTER(SunzStruct), _P]),
    "b200_sunz_correct_f64": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _GEO, C.POINTER(SunzStruct), _P]),
    "b200_sunz_correct_sza_f64": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, C.c_int64, C.POINTER(SunzStruct), _P]),
    "b200_sunz_correct_sza": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, C.c_int64, C.POINTER(SunzStruct), _P]),
    "b200_abi_vis_sunz": (C.c_int, [C.POINTER(_P), C.POINTER(_P), C.POINTER(AbiCalStruct), C.c_int32, _GEO,
                                    C.POINTER(SunzStruct), _P]),
    "b200_nn_grid_create": (C.c_int, [_GEO, _P, C.c_double, _P, C.POINTER(C.c_int64), C.POINTER(_P), _P]),
    "b200_nn_grid_create_ex": (C.c_int, [_GEO, _P, C.c_double, _P, C.POINTER(C.c_int64), C.POINTER(_P), C.c_int32, _P]),
    "b200_nn_grid_method": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "b200_nn_grid_n_valid": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "b200_nn_grid_destroy": (C.c_int, [_P]),
    "b200_nn_grid_nbytes": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "b200_nn_query": (C.c_int, [_P, _GEO, C.c_double, _P, _P, _P, _P]),
    "b200_nn_gather": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _P, _P]),
    "b200_nn_gather_variant": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _P, C.c_int32, _P]),
    "b200_nn_query_gather_f32": (C.c_int, [_P, _GEO, C.c_double, _P, _P, C.c_int32, C.c_int64, C.c_int64,
                                           C.c_float, _P]),
    "b200_nn_compressed_to_raw": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int64, _P]),
    "b200_native_replicate": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P]),
    "b200_native_aggregate_f32": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P]),
    "b200_simulated_green": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, _P]),
    "b200_ratio_sharpen_f32": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, _P]),
    "b200_ratio_sharpen_f64": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, _P]),
    "b200_fill_f32_2d": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_float, _P]),
    "b200_fill_f32_rects": (C.c_int, [_P, C.c_int64, C.POINTER(C.c_int64), C.c_int32, C.c_float, _P]),
    "b200_host_fill_f32": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64, C.c_float]),
    "b200_copy_2d": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int64, _P]),
    "b200_bil_info": (C.c_int, [_P, _P, _GEO, C.c_double, C.c_int32, _P, _P, _P, _P, _P]),
    "b200_bil_sample_f32": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_float, _P]),
    "b200_ll2cr": (C.c_int, [_GEO, _GEO, _P, _P, C.POINTER(C.c_int64), _P]),
    "b200_fornav_accum": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int64, C.c_int32, _P, C.c_int32, C.c_int64,
                                    C.c_float, C.c_int64, C.c_int64, _P, _P, C.POINTER(EwaParamsStruct), _P]),
    "b200_fornav_finalize": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, C.POINTER(EwaParamsStruct), C.c_float,
                                       C.POINTER(C.c_int64), _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library once; raise ``B200LibraryError`` if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200LibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C satpy_b200/csrc`). satpy_b200 has no CPU fallback.")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as err:  # e.g. libcudart / libstdc++ not resolvable
        raise B200LibraryError(f"cannot load {LIB_PATH}: {err}") from err
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().b200_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a B200_E* status to the Python exception the reference would raise."""
    if rc == B200_OK:
        return
    msg = last_error()
    if rc == B200_EINVAL:
        raise ValueError(msg)
    if rc == B200_ENOMEM:
        raise MemoryError(msg)
    if rc == B200_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def version() -> str:
    return load().b200_version().decode()


def device_count() -> int:
    n = C.c_int(0)
    rc = load().b200_device_count(C.byref(n))
    return n.value if rc == B200_OK else 0


def require_cuda():
    """Import torch, make sure a CUDA device is visible, return ``torch``."""
    import torch
    load()
    if not torch.cuda.is_available():
        raise RuntimeError("satpy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def stream_ptr(torch) -> int:
    return int(torch.cuda.current_stream().cuda_stream)
