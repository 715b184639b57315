"""idencomp_b200 -- B200 (sm_100a) implementation of idencomp's context-binned rANS hot path.

The product is `libidn_gpu.so` (CUDA kernels + C-ABI, include/idn_gpu.h) and the host-side mirror of the
reference's compressor/decompressor API built on top of it.  `capi` is the ctypes binding of the C-ABI.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
