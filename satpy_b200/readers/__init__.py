"""Reader-side arithmetic (calibration) of the B200 hot path."""
from .abi import ABICalibration, calibrate_abi  # noqa: F401
from .ahi import AHICalibration, calibrate_ahi  # noqa: F401
from .seviri import SEVIRICalibration, calibrate_seviri  # noqa: F401
from .viirs import scale_viirs  # noqa: F401
