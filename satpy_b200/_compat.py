"""Base classes of the plug-ins: the reference's own when it is importable, small stand-ins otherwise.

satpy checks ``isinstance(resampler, pyresample.resampler.BaseResampler)`` before re-using a resampler instance
(satpy/resample/base.py:151-156) and reads ``compositor.id`` / ``.attrs`` of every modifier it puts in the dependency
tree (satpy/node.py:163, satpy/composites/core.py:65-97), so the device classes must BE those types when they run
inside satpy.  Neither satpy nor pyresample can be installed in this image: the stand-ins
below carry the same surface, and tests/test_host_cpu.py runs ``register()`` and the plug-in classes against stub
``satpy`` / ``pyresample`` modules to exercise the other branch.
"""
from __future__ import annotations


def resampler_base():
    """``pyresample.resampler.BaseResampler`` or ``object``."""
    try:
        from pyresample.resampler import BaseResampler
        return BaseResampler
    except Exception:  # noqa: BLE001 -- ImportError, or a half-installed pyresample
        return object


class LocalModifierBase:
    """What ``satpy.modifiers.ModifierBase`` (a ``CompositeBase``, satpy/composites/core.py:46-125) gives a modifier:
    keyword storage in ``attrs``, the ``id`` property the dependency tree reads, ``apply_modifier_info``."""

    def __init__(self, name=None, prerequisites=None, optional_prerequisites=None, **kwargs):
        kwargs["name"] = name
        kwargs["prerequisites"] = prerequisites or []
        kwargs["optional_prerequisites"] = optional_prerequisites or []
        self.attrs = kwargs

    @property
    def id(self):  # noqa: A003 -- the reference's name
        """The identifier of the modified dataset: the ``_satpy_id`` satpy passes in, else the name."""
        return self.attrs.get("_satpy_id", self.attrs.get("name"))

    def __str__(self):
        return f"{type(self).__name__}({self.attrs.get('name')!r})"

    def apply_modifier_info(self, origin, destination):
        """satpy/composites/core.py:99-125 for the default key set (``name``, ``modifiers``)."""
        o, d = origin.attrs, destination.attrs
        for k in ("name", "modifiers"):
            if k == "modifiers" and k in self.attrs:
                d[k] = self.attrs[k]
            elif d.get(k) is None:
                if self.attrs.get(k) is not None:
                    d[k] = self.attrs[k]
                elif o.get(k) is not None:
                    d[k] = o[k]


def modifier_base():
    """``satpy.modifiers.ModifierBase`` or the local stand-in."""
    try:
        from satpy.modifiers import ModifierBase
        return ModifierBase
    except Exception:  # noqa: BLE001
        return LocalModifierBase


def incompatible_areas():
    """``satpy.composites.core.IncompatibleAreas`` when available, so that satpy's own ``except`` clauses see it."""
    try:
        from satpy.composites.core import IncompatibleAreas
        return IncompatibleAreas
    except Exception:  # noqa: BLE001
        class IncompatibleAreas(Exception):
            """Same role as satpy.composites.core.IncompatibleAreas (composites/core.py:41)."""
        return IncompatibleAreas
