"""Shared host helpers of the calibration wrappers."""
from __future__ import annotations

import numpy as np

from .. import _lib


def to_device(arr, allowed):
    """numpy array or torch tensor -> contiguous CUDA tensor.  uint16 travels as int16 bits.
    Returns ``(tensor, was_numpy, numpy_dtype_name)``; ``allowed``: accepted numpy dtype names."""
    torch = _lib.require_cuda()
    was_numpy = not isinstance(arr, torch.Tensor)
    if was_numpy:
        a = np.ascontiguousarray(arr)
        name = a.dtype.name
        if name not in allowed:
            raise TypeError(f"expected one of {allowed}, got {a.dtype}")
        if name == "uint16":
            a = a.view(np.int16)
        return torch.from_numpy(a).to("cuda", non_blocking=True), True, name
    name = str(arr.dtype).replace("torch.", "")
    if name not in allowed:
        raise TypeError(f"expected one of {allowed}, got {arr.dtype}")
    t = arr if arr.is_cuda else arr.to("cuda")
    return t.contiguous(), False, name


def new_f32_like(t):
    torch = _lib.require_cuda()
    return torch.empty(tuple(t.shape), dtype=torch.float32, device="cuda")
