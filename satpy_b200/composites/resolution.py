"""``RatioSharpenedRGB`` / ``SelfSharpenedRGB`` on the device (satpy/composites/resolution.py:32-233).

Same constructor keywords (``high_resolution_band="red"``, ``neutral_resolution_band=None``, ``common_channel_mask``)
and call convention ``compositor((red, green, blue), optional_datasets=(high_res,), **info)``; the result is a
``(bands, y, x)`` array with ``bands = ["R", "G", "B"]``, attrs combined like ``_combined_sharpened_info`` (:145-153)
and ``units`` removed (:97).  The arithmetic -- ``_get_sharpening_ratio`` (:157-165), ``_mean4`` (:168-192), the band
products (:143-154) and ``GenericCompositor``'s common channel mask (satpy/composites/core.py:489-490) -- is one kernel
launch (``b200_ratio_sharpen_f32`` / ``_f64``, csrc/native.cu).
"""
from __future__ import annotations

import logging

import numpy as np

from .. import _lib, _xfer
from ..dataset import is_dataarray, payload, rewrap
from ..modifiers.geometry import IncompatibleAreas

LOG = logging.getLogger(__name__)
_COLORS = ["red", "green", "blue"]


class RatioSharpenedRGB:
    """Sharpen RGB bands with the ratio of a high resolution band to a lower resolution version of it."""

    def __init__(self, name, high_resolution_band="red", neutral_resolution_band=None, common_channel_mask=True,
                 **kwargs):
        if high_resolution_band not in _COLORS + [None]:
            raise ValueError("RatioSharpenedRGB.high_resolution_band must be one of ['red', 'green', 'blue', None]. "
                             "Not '{}'".format(high_resolution_band))
        if neutral_resolution_band not in _COLORS + [None]:
            raise ValueError("RatioSharpenedRGB.neutral_resolution_band must be one of ['red', 'green', 'blue', None]. "
                             "Not '{}'".format(neutral_resolution_band))
        self.high_resolution_color = high_resolution_band
        self.neutral_resolution_color = neutral_resolution_band
        self.common_channel_mask = common_channel_mask
        self.attrs = dict(kwargs, name=name)

    _self_mean = False

    def _run(self, bands, high, info, template, offset=(0, 0)):
        torch = _lib.require_cuda()
        was_numpy = any(not isinstance(payload(b), torch.Tensor) for b in bands)
        tens = [_xfer.to_device(payload(b)).contiguous() for b in bands]
        dtype = tens[0].dtype
        if dtype not in (torch.float32, torch.float64) or any(t.dtype != dtype for t in tens):
            raise TypeError("ratio sharpening works on three float32 or three float64 bands")
        h, w = tens[0].shape
        th = None
        if high is not None:
            th = _xfer.to_device(payload(high)).to(dtype).contiguous()
        out = torch.empty((3, h, w), dtype=dtype, device="cuda")
        hi = _COLORS.index(self.high_resolution_color) if self.high_resolution_color else 0
        neutral = _COLORS.index(self.neutral_resolution_color) if self.neutral_resolution_color else -1
        fn = _lib.load().b200_ratio_sharpen_f32 if dtype == torch.float32 else _lib.load().b200_ratio_sharpen_f64
        _lib.check(fn(tens[0].data_ptr(), tens[1].data_ptr(), tens[2].data_ptr(), None if th is None else th.data_ptr(),
                      out.data_ptr(), h, w, hi, neutral, 1 if self._self_mean else 0, int(offset[0]) % 2,
                      int(offset[1]) % 2, 1 if self.common_channel_mask else 0, _lib.stream_ptr(torch)))
        res = _xfer.to_host(out) if was_numpy else out
        if not is_dataarray(template):
            return res
        new_attrs = {}
        if high is not None and is_dataarray(high):
            if "rows_per_scan" in high.attrs:
                new_attrs["rows_per_scan"] = high.attrs["rows_per_scan"]
            if "resolution" in high.attrs:
                new_attrs["resolution"] = high.attrs["resolution"]
        attrs = {k: v for k, v in template.attrs.items() if k not in ("units", "wavelength", "calibration")}
        attrs.update(info)
        attrs.update(new_attrs)
        attrs.update({k: v for k, v in self.attrs.items() if v is not None})
        attrs.setdefault("standard_name", "true_color")
        attrs["mode"] = "RGB"
        attrs.pop("units", None)
        res = rewrap(template, res, dims=("bands",) + tuple(template.dims[-2:]), attrs=attrs)
        if hasattr(res, "coords") and isinstance(res.coords, dict):
            res.coords["bands"] = np.array(["R", "G", "B"])
        return res

    def __call__(self, datasets, optional_datasets=None, **info):
        if len(datasets) != 3:
            raise ValueError("Expected 3 datasets, got %d" % (len(datasets), ))
        shapes = [tuple(payload(d).shape) for d in datasets]
        optional = tuple(optional_datasets or ())
        if any(s != shapes[0] for s in shapes[1:]) or (optional and tuple(payload(optional[0]).shape) != shapes[0]):
            raise IncompatibleAreas("RatioSharpening requires datasets of the same size. Must resample first.")
        high = optional[0] if (optional and self.high_resolution_color is not None) else None
        if high is None:
            LOG.debug("No sharpening band specified for ratio sharpening")
        return self._run(datasets, high, info, datasets[0])


class SelfSharpenedRGB(RatioSharpenedRGB):
    """Sharpen RGB with the ratio of a band to a 2 x 2 averaged version of itself (resolution.py:195-233)."""

    _self_mean = True

    def __call__(self, datasets, optional_datasets=None, **attrs):
        if self.high_resolution_color not in _COLORS:
            raise ValueError("SelfSharpenedRGB requires at least one high resolution band, not "
                             "'{}'".format(self.high_resolution_color))
        if len(datasets) != 3:
            raise ValueError("Expected 3 datasets, got %d" % (len(datasets), ))
        shapes = [tuple(payload(d).shape) for d in datasets]
        if any(s != shapes[0] for s in shapes[1:]):
            raise IncompatibleAreas("RatioSharpening requires datasets of the same size. Must resample first.")
        high = datasets[_COLORS.index(self.high_resolution_color)]
        try:
            offset = high.attrs["area"].crop_offset
        except (KeyError, AttributeError):
            offset = (0, 0)
        return self._run(datasets, high, attrs, high, offset)
