"""poulpy_b200 -- B200-native backend for the poulpy-hal hot path (NTT120 / FFT64 DFT, svp, vmp, big normalize and the
key-switch / external-product / blind-rotate compositions).  The product is the C-ABI library `libpoulpy_b200.so`
(include/poulpy_b200.h); `poulpy_b200.hal` is its ctypes binding."""
from . import hal  # noqa: F401
from .hal import FFT64, NTT120, DevBuf, Module, PoulpyError, lib, pinned_empty  # noqa: F401
