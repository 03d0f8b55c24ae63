"""Modifier plug-ins of the B200 hot path (same class names as ``satpy.modifiers``)."""
from .geometry import (EffectiveSolarPathLengthCorrector, IncompatibleAreas, SunZenithCorrector,  # noqa: F401
                       SunZenithReducer, cos_sza, sunz_correct, sunz_reduce)
