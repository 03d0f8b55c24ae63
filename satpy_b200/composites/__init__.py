"""Compositors of the ABI true-colour chain that sit next to the hot path ("next" row of SURVEY.md section 8f)."""
from .abi import SimulatedGreen  # noqa: F401
from .resolution import RatioSharpenedRGB, SelfSharpenedRGB  # noqa: F401
