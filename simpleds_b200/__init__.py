"""B200-native delay-spectrum hot path with the simpleDS ``DelaySpectrum`` API.

Python/PyTorch host code calling hand-written sm_100a CUDA kernels through the
C ABI of ``libsds.so`` (``include/sds.h``).  There is no CPU fallback."""
from . import units, utils  # noqa: F401
from .delay_spectrum import DelaySpectrum, UVDataLite  # noqa: F401

__version__ = "0.1.0"
