"""gp_compressor_b200 — B200 (sm_100a) implementation of the gp_compressor compress /
decompress hot path behind the C ABI of include/gpc.h.

The product is gp_compressor_b200/libgpc_b200.so (CUDA kernels in csrc/).  This package
only carries the ctypes binding used by tests and bench.py, the build script and the
synthetic workload generators.  Nothing here falls back to the CPU.
"""
from . import binding, synth  # noqa: F401
from .binding import Handle, GpcError, load  # noqa: F401
