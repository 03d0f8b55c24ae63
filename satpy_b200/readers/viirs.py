"""VIIRS SDR scaling on the device ("next" row of SURVEY.md section 8f).

Host-side mirror of ``mask_fill_values`` + ``scale_swath_data`` (satpy/readers/core/viirs_atms_sdr.py:168-214,328-361):
integers ``>= fill_min_int`` (65528) and floats ``<= fill_max_float`` (-999) become NaN, then every granule of rows is
scaled with its own ``(factor, offset)`` pair (factors ``<= -999`` are NaN).  M-band granules have 48 scans of 16 rows
(``rows_per_scan``, viirs_atms_sdr.py:247-255), which is what the EWA resampler is told through ``attrs``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from ._common import new_f32_like, to_device


def scale_viirs(data, factors, rows_per_granule, fill_min_int=65528, fill_max_float=-999.0, out=None):
    """``data``: (rows, cols) uint16 or float32; ``factors``: flat or (n_granules, 2); returns float32."""
    t, was_numpy, name = to_device(data, ("uint16", "int16", "float32"))
    if t.dim() != 2:
        raise ValueError("VIIRS swath data must be 2-D (rows, cols)")
    torch = _lib.require_cuda()
    f = np.ascontiguousarray(np.asarray(factors, dtype=np.float32).reshape(-1, 2))
    rpg = np.ascontiguousarray(np.asarray(rows_per_granule, dtype=np.int32))
    if f.shape[0] != rpg.size:
        raise ValueError("one (factor, offset) pair per granule is required")
    if out is None:
        out = new_f32_like(t)
    _lib.check(_lib.load().b200_viirs_scale(
        t.data_ptr(), 1 if name == "float32" else 0, out.data_ptr(), t.shape[0], t.shape[1],
        f.ctypes.data_as(C.POINTER(C.c_float)), rpg.ctypes.data_as(C.POINTER(C.c_int32)), int(rpg.size),
        int(fill_min_int), float(fill_max_float), _lib.stream_ptr(torch)))
    return out.cpu().numpy() if was_numpy else out
