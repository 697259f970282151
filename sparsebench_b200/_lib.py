"""ctypes loader of libsparsebench_b200.so (the C ABI declared in include/sparsebench_b200.h)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SB_LIB") or os.path.join(HERE, "libsparsebench_b200.so")   # SB_LIB: an experimental build
_lib = None


def load(variant=""):
    """variant: "" (double / unsigned int), "f32", "u64", "f32u64" -- one per process"""
    global _lib, LIB_PATH
    if _lib is None:
        if variant and not os.environ.get("SB_LIB"):
            LIB_PATH = os.path.join(HERE, "libsparsebench_b200_%s.so" % variant)
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "%s is missing: build it with `python -m sparsebench_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback." % LIB_PATH)
        # RTLD_LOCAL on purpose: the library exports the reference's own names (allocate, waxpby, ...); a global load
        # would interpose them on any other copy of the reference in the process (the test oracle). The drop-in shims
        # find it through their DT_NEEDED entry + $ORIGIN rpath.
        _lib = C.CDLL(LIB_PATH)
    return _lib


def load_dropin(fmt, variant=""):
    load(variant)
    return C.CDLL(os.path.join(HERE, "libsparsebench_b200_%s%s.so" % (fmt, "_" + variant if variant else "")))
