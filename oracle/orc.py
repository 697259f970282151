"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front-end of the CPU parity oracle (oracle/sb_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (sparsebench_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")

ENTRY_DTYPE = np.dtype([("col", np.uint32), ("pad", np.uint32), ("val", np.float64)])  # matrix.h:24-27


class OrcRank(C.Structure):
    _fields_ = [
        ("nr", C.c_uint32), ("startRow", C.c_uint32), ("stopRow", C.c_uint32),
        ("rowPtr", C.c_void_p), ("col", C.c_void_p),
        ("externalCount", C.c_int), ("totalSendCount", C.c_int), ("indegree", C.c_int), ("outdegree", C.c_int),
        ("sources", C.POINTER(C.c_int)), ("recvCounts", C.POINTER(C.c_int)), ("rdispls", C.POINTER(C.c_int)),
        ("destinations", C.POINTER(C.c_int)), ("sendCounts", C.POINTER(C.c_int)), ("sdispls", C.POINTER(C.c_int)),
        ("elementsToSend", C.POINTER(C.c_int)), ("externalsReordered", C.POINTER(C.c_int)),
    ]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_generate.restype = C.c_int64
        L.orc_generate.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_init_vectors.argtypes = [C.c_uint32, u32p, C.c_int, f64p, f64p, C.c_void_p]
        L.orc_spmv_crs.argtypes = [C.c_uint32, u32p, u32p, f64p, f64p, f64p]
        L.orc_spmv_ccrs.argtypes = [C.c_uint32, u32p, C.c_void_p, f64p, f64p]
        L.orc_scs_structure.restype = C.c_int64
        L.orc_scs_structure.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, u32p, u32p, u32p, u32p, u32p]
        L.orc_scs_fill.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, f64p, u32p, u32p, C.c_int64, u32p, f64p]
        L.orc_spmv_scs.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, u32p, f64p, f64p, f64p]
        L.orc_waxpby.argtypes = [C.c_uint32, C.c_double, f64p, C.c_double, f64p, f64p]
        L.orc_ddot.restype = C.c_double
        L.orc_ddot.argtypes = [C.c_uint32, f64p, f64p]
        L.orc_cg_crs.restype = C.c_int
        L.orc_cg_crs.argtypes = [C.c_uint32, C.c_uint32, u32p, u32p, f64p, f64p, f64p, C.c_int, C.c_double, f64p,
                                 C.POINTER(C.c_int)]
        L.orc_partition_all.argtypes = [C.c_int, C.POINTER(OrcRank)]
        L.orc_partition_free.argtypes = [C.c_int, C.POINTER(OrcRank)]
        L.orc_exchange_all.argtypes = [C.c_int, C.POINTER(OrcRank), C.POINTER(C.c_void_p)]
        L.orc_cg_multi.restype = C.c_int
        L.orc_cg_multi.argtypes = [C.c_int, C.POINTER(OrcRank), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_void_p), C.c_int, C.c_double, f64p, C.POINTER(C.c_int)]
        _LIB = L
    return _LIB


class Csr:
    """rowPtr/col/val of one rank's row block (GMatrix after matrixGenerate, matrix.h:29-35)."""

    def __init__(self, rowPtr, col, val, nc=None, startRow=0, totalNr=None):
        self.rowPtr = np.ascontiguousarray(rowPtr, np.uint32)
        self.col = np.ascontiguousarray(col, np.uint32)
        self.val = np.ascontiguousarray(val, np.float64)
        self.nr = len(self.rowPtr) - 1
        self.nc = self.nr if nc is None else nc
        self.startRow = startRow
        self.stopRow = startRow + self.nr - 1
        self.totalNr = self.nr if totalNr is None else totalNr

    @property
    def nnz(self):
        return int(self.rowPtr[-1])

    def entries(self):
        e = np.zeros(self.nnz, ENTRY_DTYPE)
        e["col"] = self.col
        e["val"] = self.val
        return e


def generate(nx, ny, nz, rank=0, size=1, use7pt=False):
    """matrix.c:30-121. Column ids are GLOBAL (commPartition renumbers them)."""
    L = lib()
    n = nx * ny * nz
    nnz = L.orc_generate(nx, ny, nz, rank, size, int(use7pt), None, None, None, 0)
    rowPtr = np.zeros(n + 1, np.uint32)
    col = np.zeros(nnz, np.uint32)
    val = np.zeros(nnz, np.float64)
    got = L.orc_generate(nx, ny, nz, rank, size, int(use7pt), rowPtr.ctypes.data, col.ctypes.data,
                         val.ctypes.data, nnz)
    assert got == nnz
    return Csr(rowPtr, col, val, startRow=n * rank, totalNr=n * size)


def init_vectors(m, generated=True):
    x = np.zeros(m.nr)
    b = np.zeros(m.nr)
    xe = np.zeros(m.nr) if generated else None
    lib().orc_init_vectors(m.nr, m.rowPtr, int(generated), x, b, xe.ctypes.data if generated else None)
    return x, b, xe


def spmv_crs(m, x):
    y = np.zeros(m.nr)
    lib().orc_spmv_crs(m.nr, m.rowPtr, m.col, m.val, np.ascontiguousarray(x, np.float64), y)
    return y


def spmv_ccrs(m, x):
    y = np.zeros(m.nr)
    e = m.entries()
    lib().orc_spmv_ccrs(m.nr, m.rowPtr, e.ctypes.data, np.ascontiguousarray(x, np.float64), y)
    return y


class Scs:
    pass


def scs_convert(m, Cc, sigma):
    """matrix-SCS.c:31-196 without the :42-43 overwrite."""
    L = lib()
    s = Scs()
    s.C, s.sigma, s.nr = Cc, sigma, m.nr
    s.nChunks = (m.nr + Cc - 1) // Cc
    s.nrPadded = s.nChunks * Cc
    s.oldToNewPerm = np.zeros(max(m.nr, 1), np.uint32)
    s.newToOldPerm = np.zeros(max(m.nr, 1), np.uint32)
    s.chunkLens = np.zeros(max(s.nChunks, 1), np.uint32)
    s.chunkPtr = np.zeros(s.nChunks + 1, np.uint32)
    s.nElems = L.orc_scs_structure(m.nr, Cc, sigma, m.rowPtr, s.oldToNewPerm, s.newToOldPerm, s.chunkLens, s.chunkPtr)
    s.colInd = np.zeros(max(s.nElems, 1), np.uint32)
    s.val = np.zeros(max(s.nElems, 1), np.float64)
    L.orc_scs_fill(m.nr, Cc, m.rowPtr, m.col, m.val, s.oldToNewPerm, s.chunkPtr, s.nElems, s.colInd, s.val)
    s.oldToNewPerm = s.oldToNewPerm[:m.nr]
    s.newToOldPerm = s.newToOldPerm[:m.nr]
    s.chunkLens = s.chunkLens[:s.nChunks]
    s.colInd = s.colInd[:s.nElems]
    s.val = s.val[:s.nElems]
    return s


def spmv_scs(s, x):
    y = np.zeros(max(s.nrPadded, 1))
    lib().orc_spmv_scs(s.nChunks, s.C, s.chunkPtr, np.ascontiguousarray(s.chunkLens), np.ascontiguousarray(s.colInd),
                       np.ascontiguousarray(s.val), np.ascontiguousarray(x, np.float64), y)
    return y[:s.nrPadded]


def waxpby(alpha, x, beta, y):
    w = np.zeros(len(x))
    lib().orc_waxpby(len(x), alpha, np.ascontiguousarray(x), beta, np.ascontiguousarray(y), w)
    return w


def ddot(x, y):
    return lib().orc_ddot(len(x), np.ascontiguousarray(x), np.ascontiguousarray(y))


class dot_threads:
    """`with orc.dot_threads(16): ...` -- dot products summed like the reference's OpenMP build with that many
    threads (solver.c:46-61, static schedule); 1 = sequential (default); 0 = long double yardstick."""

    def __init__(self, t):
        self.t = t

    def __enter__(self):
        self.old = lib().orc_get_dot_threads()
        lib().orc_set_dot_threads(self.t)

    def __exit__(self, *exc):
        lib().orc_set_dot_threads(self.old)
        return False


def cg_crs(m, b, x0, itermax, eps):
    """CGSolver.c:62-141. Returns (k, history, x): history[0] initial residual, history[k] = normr of iteration k."""
    x = np.array(x0, np.float64)
    hist = np.zeros(itermax + 2)
    nh = C.c_int(0)
    k = lib().orc_cg_crs(m.nr, m.nc, m.rowPtr, m.col, m.val, np.ascontiguousarray(b), x, itermax, eps, hist,
                         C.byref(nh))
    return k, hist[:nh.value].copy(), x


class Partition:
    """commPartition (comm.c:414-625) of all P ranks at once; ranks[r] is a dict of int32 arrays."""

    def __init__(self, mats):
        L = lib()
        self.P = len(mats)
        self.mats = mats
        self._arr = (OrcRank * self.P)()
        for r, m in enumerate(mats):
            a = self._arr[r]
            a.nr, a.startRow, a.stopRow = m.nr, m.startRow, m.stopRow
            a.rowPtr = m.rowPtr.ctypes.data
            a.col = m.col.ctypes.data
        L.orc_partition_all(self.P, self._arr)
        self.ranks = []
        for r, m in enumerate(mats):
            a = self._arr[r]

            def arr(p, n):
                return np.array([p[i] for i in range(n)], np.int32)
            d = dict(externalCount=a.externalCount, totalSendCount=a.totalSendCount, indegree=a.indegree,
                     outdegree=a.outdegree,
                     sources=arr(a.sources, a.indegree), recvCounts=arr(a.recvCounts, a.indegree),
                     rdispls=arr(a.rdispls, a.indegree), destinations=arr(a.destinations, a.outdegree),
                     sendCounts=arr(a.sendCounts, a.outdegree), sdispls=arr(a.sdispls, a.outdegree),
                     elementsToSend=np.ctypeslib.as_array(a.elementsToSend, (max(a.totalSendCount, 1),))[
                         :a.totalSendCount].astype(np.int32),
                     externalsReordered=np.ctypeslib.as_array(a.externalsReordered, (max(a.externalCount, 1),))[
                         :a.externalCount].astype(np.int32))
            m.nc = m.nr + a.externalCount
            self.ranks.append(d)

    def _ptrs(self, arrays):
        p = (C.c_void_p * self.P)()
        for r, a in enumerate(arrays):
            assert a.dtype == np.float64 and a.flags.c_contiguous
            p[r] = a.ctypes.data
        return p

    def exchange(self, xs):
        lib().orc_exchange_all(self.P, self._arr, self._ptrs(xs))

    def cg(self, bs, xs, itermax, eps):
        hist = np.zeros(itermax + 2)
        nh = C.c_int(0)
        vals = [m.val for m in self.mats]
        k = lib().orc_cg_multi(self.P, self._arr, self._ptrs(vals), self._ptrs(bs), self._ptrs(xs), itermax, eps,
                               hist, C.byref(nh))
        return k, hist[:nh.value].copy()

    def __del__(self):
        try:
            lib().orc_partition_free(self.P, self._arr)
        except Exception:
            pass
