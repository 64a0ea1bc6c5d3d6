"""fbdsp -- B200-native batch demodulation engine for FileBeep's receive hot path.

Python host layer over libfbdsp.so (include/fbdsp.h).  `fbdsp.modem` keeps the reference's
modem.py signatures; `fbdsp.Engine` is the batch interface; there is no CPU fallback.
"""
from ._lib import FbdspError, LIB_PATH  # noqa: F401
from .engine import Engine, DemodResult, default_engine  # noqa: F401
from .design import psk_design, PskDesign  # noqa: F401
