"""pnfft_b200: B200-native window-convolution path of PNFFT behind PNFFT's own C API.

The product is the C-ABI library pnfft_b200/lib/libpnfft_b200.so (sources in pnfft_b200/csrc, public
headers in include/); this package is only the thin ctypes mirror of that interface used by the
tests and the benchmark.
"""
from . import api  # noqa: F401
