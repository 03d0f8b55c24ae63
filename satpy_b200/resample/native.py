"""``b200_native``: integer-factor replicate / aggregate resampler on the device ("next" row of SURVEY.md section 8f).

Host-side mirror of ``NativeResampler`` (satpy/resample/native.py:20-159): constructed as
``cls(source_geo_def, target_geo_def)``, ``resample(data, mask_area=False, **kw)`` -> ``compute(data, expand=True)``;
the target may be a list of areas (the largest is used when ``expand`` else the smallest, native.py:56-63); repeat
factors come from the shapes (``_get_repeats``); same errors for non-integer factors and for reducing arrays with more
than two dimensions (native.py:101-113,145-148).
"""
from __future__ import annotations

import numpy as np

from .. import _compat, _lib, _xfer
from ..dataset import is_dataarray, payload, rewrap
from ..geometry import as_geometry


class B200NativeResampler(_compat.resampler_base()):
    def __init__(self, source_geo_def, target_geo_def):
        if _compat.resampler_base() is not object:
            super().__init__(source_geo_def, target_geo_def)
        self.source_geo_def = source_geo_def
        self.target_geo_def = target_geo_def

    def resample(self, data, cache_dir=None, mask_area=False, **kwargs):
        del cache_dir, mask_area
        return self.compute(data, **kwargs)

    def compute(self, data, expand=True, **kwargs):
        del kwargs
        torch = _lib.require_cuda()
        lib = _lib.load()
        if isinstance(self.target_geo_def, (list, tuple)):
            target = (max if expand else min)(self.target_geo_def, key=lambda x: tuple(as_geometry(x).shape))
        else:
            target = self.target_geo_def
        out_shape = tuple(as_geometry(target).shape)
        arr = payload(data)
        was_numpy = not isinstance(arr, torch.Tensor)
        t = torch.from_numpy(np.ascontiguousarray(arr)) if was_numpy else arr
        if t.dim() not in (2, 3):
            raise ValueError("Can only handle 2D or 3D arrays without dimensions.")
        t = t.to("cuda", non_blocking=True).contiguous()
        in_shape = tuple(t.shape[-2:])
        y_rep, x_rep = out_shape[0] / float(in_shape[0]), out_shape[1] / float(in_shape[1])
        stream = _lib.stream_ptr(torch)
        if y_rep == 1 and x_rep == 1:
            res = t
        elif y_rep >= 1 and x_rep >= 1:
            if not (float(y_rep).is_integer() and float(x_rep).is_integer()):
                raise ValueError("Expand factor must be a whole number")
            nb = 1 if t.dim() == 2 else t.shape[0]
            res = torch.empty(tuple(t.shape[:-2]) + out_shape, dtype=t.dtype, device="cuda")
            _lib.check(lib.b200_native_replicate(t.data_ptr(), res.data_ptr(), nb, in_shape[0], in_shape[1], int(y_rep),
                                                 int(x_rep), t.element_size(), stream))
        elif y_rep <= 1 and x_rep <= 1:
            if t.dim() != 2:
                raise ValueError("Can't aggregrate (reduce) data arrays with more than 2 dimensions.")
            y_size, x_size = 1.0 / y_rep, 1.0 / x_rep
            if not (x_size.is_integer() and y_size.is_integer()):
                raise ValueError("Aggregation factors are not integers")
            if t.dtype != torch.float32:
                raise TypeError(f"the device aggregation works on float32 data, got {t.dtype}")
            res = torch.empty(out_shape, dtype=torch.float32, device="cuda")
            _lib.check(lib.b200_native_aggregate_f32(t.data_ptr(), res.data_ptr(), out_shape[0], out_shape[1],
                                                     int(y_size), int(x_size), stream))
        else:
            raise ValueError("Must either expand or reduce in both directions")
        res = _xfer.to_host(res) if was_numpy else res
        if not is_dataarray(data):
            return res
        attrs = dict(data.attrs)
        attrs["area"] = target
        return rewrap(data, res, dims=tuple(data.dims), attrs=attrs, new_area=target)
