"""Reader-side arithmetic (calibration) of the B200 hot path."""
from .abi import ABICalibration, calibrate_abi  # noqa: F401
