"""nimble_b200 — B200-native (sm_100a) implementation of nimble's read-assignment hot path.

The product is the CUDA shared library `libnimble_b200.so` (C ABI in include/nimble_b200.h);
this package is the thin Python host side that mirrors nimble's own operator interface
(`python -m nimble_b200 align|report|generate`, same flags as `python -m nimble`).
There is no CPU fallback: importing works anywhere, creating an Engine needs a CUDA device.
"""
__version__ = "0.1.0"

from ._lib import NimbleB200Error, RESULT_DTYPE  # noqa: F401


def Engine(*a, **kw):
    from .engine import Engine as _E
    return _E(*a, **kw)
