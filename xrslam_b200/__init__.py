"""xrslam_b200 -- B200-native visual-inertial odometry hot path behind the xrslam interface.

The product is the C-ABI shared library ``xrslam_b200/lib/libxrvio.so`` (CUDA kernels for sm_100a +
host glue, built by ``xrslam_b200.build``); this package only carries the build script and a thin
ctypes binding used by tests and bench.py.  There is no CPU fallback: loading fails loudly when the
library is missing, and every call fails when no CUDA device is present.
"""
from .capi import lib, XrvioError  # noqa: F401
