"""``SimulatedGreen`` on the device (satpy/composites/abi.py:27-61): ``d1 * f1 + d2 * f2 + d3 * f3`` in float32."""
from __future__ import annotations

import numpy as np

from .. import _lib
from ..dataset import is_dataarray, payload, rewrap
from ..modifiers.geometry import IncompatibleAreas


class SimulatedGreen:
    """A single band resembling a green (0.55 um) band from C01, C02, C03; attrs are taken from C03 (abi.py:60)."""

    def __init__(self, name, fractions=(0.465, 0.465, 0.07), **kwargs):
        self.fractions = fractions
        self.attrs = dict(kwargs, name=name)

    def __call__(self, projectables, optional_datasets=None, **attrs):
        torch = _lib.require_cuda()
        c01, c02, c03 = projectables
        tens = []
        was_numpy = False
        for d in (c01, c02, c03):
            a = payload(d)
            if not isinstance(a, torch.Tensor):
                was_numpy = True
                a = torch.from_numpy(np.ascontiguousarray(a))
            if a.dtype != torch.float32:
                raise TypeError(f"SimulatedGreen works on float32 data, got {a.dtype}")
            tens.append(a.to("cuda", non_blocking=True).contiguous())
        if len({tuple(t.shape) for t in tens}) != 1:
            raise IncompatibleAreas("All areas should be of the same size")
        out = torch.empty_like(tens[0])
        f = [float(np.float32(v)) for v in self.fractions]
        _lib.check(_lib.load().b200_simulated_green(tens[0].data_ptr(), tens[1].data_ptr(), tens[2].data_ptr(),
                                                    out.data_ptr(), out.numel(), f[0], f[1], f[2],
                                                    _lib.stream_ptr(torch)))
        res = out.cpu().numpy() if was_numpy else out
        if not is_dataarray(c03):
            return res
        new_attrs = dict(c03.attrs)
        new_attrs.update({k: v for k, v in self.attrs.items() if v is not None})
        new_attrs.update(attrs)
        return rewrap(c03, res, attrs=new_attrs)
