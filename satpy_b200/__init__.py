"""satpy_b200 -- the per-pixel hot path of pytroll/satpy on NVIDIA B200 (sm_100a).

Reader calibration (ABI counts -> radiance -> reflectance / brightness temperature),
``SunZenithCorrector`` and nearest-neighbour resampling, as hand-written CUDA kernels behind
satpy's own plug points.  Host code is Python over PyTorch tensors and a ctypes C ABI
(``include/satpy_b200.h``); there is no CPU fallback.
"""
from __future__ import annotations

import os

from ._lib import B200LibraryError, device_count, version  # noqa: F401
from .dataset import DataArrayLite  # noqa: F401
from .geometry import AreaDefinition, SwathDefinition, as_geometry  # noqa: F401

__version__ = "0.1.0"

ETC_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "etc")


def register(modifiers: bool = False) -> bool:
    """Hook the plug-ins into an installed satpy; returns False when satpy is not importable.

    * appends ``satpy_b200.resample`` to ``satpy.resample.base.RESAMPLER_MODULES`` so that
      ``Scene.resample(area, resampler="b200_nearest")`` resolves (satpy/resample/base.py:77-105); the resampler
      classes are ``pyresample.resampler.BaseResampler`` subclasses whenever pyresample is importable, which is what
      ``satpy.resample.base.resample`` checks before re-using an instance (base.py:151-156);
    * ``modifiers=True`` additionally puts ``satpy_b200/etc`` on satpy's ``config_path`` so that the YAML modifier
      names (``sunz_corrected``, ``effective_solar_pathlength_corrected``, ``sunz_reduced``) bind to the device
      classes (satpy/_config.py:158-173, satpy/composites/config_loader.py:176-183); those are
      ``satpy.modifiers.ModifierBase`` subclasses when satpy is importable (``.id`` / ``.attrs`` as the dependency tree
      reads them, satpy/node.py:163).  Off by default: satpy cannot be installed in the image this package was built
      in, so the re-binding has only been exercised against stub modules (tests/test_host_cpu.py), never against a
      real Scene.
    """
    try:
        import satpy
        from satpy.resample import base as _base
    except ImportError:
        return False
    mod = "satpy_b200.resample"
    if mod not in _base.RESAMPLER_MODULES:
        _base.RESAMPLER_MODULES.append(mod)
    if modifiers:
        paths = list(satpy.config.get("config_path", []))
        if ETC_DIR not in paths:
            satpy.config.set(config_path=paths + [ETC_DIR])
        try:
            from satpy.composites.config_loader import load_compositor_configs_for_sensor
            load_compositor_configs_for_sensor.cache_clear()
        except (ImportError, AttributeError):
            pass
    return True
