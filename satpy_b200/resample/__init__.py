"""Resampler plug-ins of the B200 hot path.

``get_resampler_classes()`` is the hook satpy's registry calls for every module
listed in ``satpy.resample.base.RESAMPLER_MODULES`` (satpy/resample/base.py:77-105);
``satpy_b200.register()`` appends this module there so that
``Scene.resample(area, resampler="b200_nearest")`` resolves.
"""
from .bilinear import B200BilinearResampler  # noqa: F401
from .ewa import B200EWAResampler  # noqa: F401
from .native import B200NativeResampler  # noqa: F401
from .nearest import NN_COORDINATES, B200NearestResampler  # noqa: F401


def get_resampler_classes():
    """Name -> class map merged into satpy's registry (same shape as satpy/resample/kdtree.py:314-320)."""
    return {
        "b200_nearest": B200NearestResampler,
        "b200_bilinear": B200BilinearResampler,
        "b200_ewa": B200EWAResampler,
        "b200_native": B200NativeResampler,
    }
