"""fvfi -- B200-native (sm_100a) frame-synthesis hot path of
stefan01/Fusion-Method-for-Video-Frame-Interpolation.

Host-side mirror of the reference's operator interface; every op crosses the
C-ABI of include/fvfi.h into hand-written CUDA kernels (libfvfi.so).  There is
no CPU fallback: importing works anywhere (so the interface can be inspected),
but calling an op without the CUDA library or without a GPU raises.
"""
from ._lib import lib, lib_path, FvfiError  # noqa: F401

__version__ = "0.1.0"
