"""larvio_b200 — B200-native batched MSCKF-VIO hot path behind LARVIO's call surface.

The compute lives in ``lib/liblarvio_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/larvio_b200.h``).  This package is the thin Python mirror of the reference's
``ImageProcessor`` / ``LarVio`` interface used by the tests and the benchmark; it never falls
back to a CPU implementation — if the library is missing, importing ``larvio_b200.api`` raises.
"""
__version__ = "0.1.0"
