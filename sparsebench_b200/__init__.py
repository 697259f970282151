"""sparsebench_b200 -- B200-native SpMV + CG hot path of SparseBench behind the reference's C entry points.

The product is the CUDA library (csrc/, built in-tree by sparsebench_b200.build); this package only loads it
through ctypes and mirrors the reference's call shapes for tests and benchmarks. There is no CPU fallback:
loading fails loudly when the library is missing, every compute call exits when no CUDA device is present.
"""
from . import api  # noqa: F401
