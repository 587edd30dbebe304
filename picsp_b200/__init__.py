"""picsp_b200 — B200-native implementation of PICSP's per-timestep particle loop.

The product is the shared library ``picsp_b200/libpicsp_b200.so`` (hand-written sm_100a
CUDA kernels behind the C ABI of ``include/picsp_b200.h``).  This Python package is only
a thin ctypes mirror of that ABI, named after the reference's functions, used by the
parity tests and the benchmark.  There is no CPU implementation here: if the library is
missing or no CUDA device is present, calls fail loudly.
"""
from .lib import PicspError, abi_symbols, load_library  # noqa: F401
from .sim import ELECTRON, ION, Params, Simulation  # noqa: F401

__all__ = ["Simulation", "Params", "ION", "ELECTRON", "PicspError", "load_library", "abi_symbols"]
