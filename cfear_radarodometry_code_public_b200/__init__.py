"""B200-native CFEAR per-scan hot path (k-strongest filter -> oriented surface points -> scan-to-keyframes
registration).  The product is `libcfear_b200.so` (C ABI, include/cfear_b200.h, hand-written sm_100a CUDA) and the
C++ host mirror of the reference classes (include/cfear_b200.hpp); `capi` is the ctypes plumbing used by tests and
bench.py; `synth` generates seeded synthetic Navtech-shaped polar images.  No CPU fallback exists in this package.
"""
from . import capi, synth  # noqa: F401

__all__ = ["capi", "synth"]
