"""Host-side mirror of the reference's operator interface over the C ABI (include/sparsebench_b200.h).

Same names, argument meaning and error behaviour as the reference's C functions (matrix.h, solver.h, comm.h,
allocate.h, timing.h); numpy arrays stand in for host arrays, DeviceBuffer for what allocate() returns.
"""
import ctypes as C
import os

import numpy as np

from ._lib import load

# The reference's compile-time type switches (util.h:35-53) select one of four builds of the library; this mirror
# follows the environment variable SB_VARIANT ("" = double / unsigned int, "f32", "u64", "f32u64"), read at import.
VARIANT = os.environ.get("SB_VARIANT", "")
assert VARIANT in ("", "f32", "u64", "f32u64"), VARIANT
F = C.c_float if "f32" in VARIANT else C.c_double          # CG_FLOAT
U = C.c_uint64 if "u64" in VARIANT else C.c_uint32         # CG_UINT
RDT = np.dtype(np.float32 if "f32" in VARIANT else np.float64)
IDT = np.dtype(np.uint64 if "u64" in VARIANT else np.uint32)
_HEAD = [("nr", U), ("nc", U), ("nnz", U), ("totalNr", U), ("totalNnz", U), ("startRow", U), ("stopRow", U)]
# matrix.h:24-27 `struct { CG_UINT col; CG_FLOAT val; }` with the C compiler's padding
ENTRY_DTYPE = np.dtype({"": [("col", np.uint32), ("pad", np.uint32), ("val", np.float64)],
                        "f32": [("col", np.uint32), ("val", np.float32)],
                        "u64": [("col", np.uint64), ("val", np.float64)],
                        "f32u64": [("col", np.uint64), ("val", np.float32), ("pad", np.uint32)]}[VARIANT])

FMT_CRS, FMT_SCS, FMT_CCRS = 0, 1, 2
FMT_NAMES = {FMT_CRS: "CRS", FMT_SCS: "SCS", FMT_CCRS: "CCRS"}
CG_FUSED, CG_PRINT, CG_HOST_VECTORS, CG_NO_OVERLAP, CG_PROFILE = 1, 2, 4, 8, 16
REGIONS = ("update_p", "exchange", "spmv", "allreduce", "update_xr", "halo_wait", "spmv_boundary")
OP_MAX, OP_SUM = 0, 1


class GMatrix(C.Structure):
    _fields_ = _HEAD + [("rowPtr", C.c_void_p), ("entries", C.c_void_p)]


class CRSMatrix(C.Structure):
    _fields_ = _HEAD + [("rowPtr", C.c_void_p), ("colInd", C.c_void_p), ("val", C.c_void_p)]


class SCSMatrix(C.Structure):
    _fields_ = _HEAD + [("colInd", C.c_void_p), ("val", C.c_void_p), ("C", U), ("sigma", U), ("nrPadded", U),
                        ("nChunks", U), ("nElems", U), ("chunkPtr", C.c_void_p), ("chunkLens", C.c_void_p),
                        ("oldToNewPerm", C.c_void_p), ("newToOldPerm", C.c_void_p)]


class CCRSMatrix(C.Structure):
    _fields_ = _HEAD + [("rowPtr", C.c_void_p), ("entries", C.c_void_p)]


MATRIX_TYPES = {FMT_CRS: CRSMatrix, FMT_SCS: SCSMatrix, FMT_CCRS: CCRSMatrix}


class MMMatrix(C.Structure):                     # matrix.h:43-49
    _fields_ = [("count", C.c_size_t), ("nr", C.c_int), ("nnz", C.c_int), ("totalNr", C.c_int), ("totalNnz", C.c_int),
                ("startRow", C.c_int), ("stopRow", C.c_int), ("entries", C.c_void_p)]


class Parameter(C.Structure):
    _fields_ = [("filename", C.c_char_p), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("itermax", C.c_int),
                ("eps", C.c_double)]


class Comm(C.Structure):
    _fields_ = [("rank", C.c_int), ("size", C.c_int), ("logFile", C.c_void_p), ("externalCount", C.c_int),
                ("totalSendCount", C.c_int), ("elementsToSend", C.POINTER(C.c_int)), ("indegree", C.c_int),
                ("outdegree", C.c_int), ("sources", C.POINTER(C.c_int)), ("recvCounts", C.POINTER(C.c_int)),
                ("rdispls", C.POINTER(C.c_int)), ("destinations", C.POINTER(C.c_int)),
                ("sendCounts", C.POINTER(C.c_int)), ("sdispls", C.POINTER(C.c_int)), ("sendBuffer", C.c_void_p),
                ("communicator", C.c_void_p)]

    def lists(self):
        def arr(p, n):
            return np.array([p[i] for i in range(n)], np.int32)
        return dict(externalCount=self.externalCount, totalSendCount=self.totalSendCount,
                    sources=arr(self.sources, self.indegree), recvCounts=arr(self.recvCounts, self.indegree),
                    rdispls=arr(self.rdispls, self.indegree), destinations=arr(self.destinations, self.outdegree),
                    sendCounts=arr(self.sendCounts, self.outdegree), sdispls=arr(self.sdispls, self.outdegree),
                    elementsToSend=np.ctypeslib.as_array(self.elementsToSend, (max(self.totalSendCount, 1),))[
                        :self.totalSendCount].astype(np.int32))


class CGInfo(C.Structure):
    _fields_ = [("flags", C.c_int), ("b", C.c_void_p), ("x", C.c_void_p), ("history", C.POINTER(C.c_double)),
                ("historyCap", C.c_int), ("nhist", C.c_int), ("solveMs", C.c_double), ("maxError", C.c_double),
                ("regionMs", C.c_double * 7), ("createMs", C.c_double), ("finishMs", C.c_double)]


_configured = False


def lib():
    global _configured
    L = load(VARIANT)
    if not _configured:
        L.allocate.restype = C.c_void_p
        L.allocate.argtypes = [C.c_size_t, C.c_size_t]
        L.sbAllocateDevice.restype = C.c_void_p
        L.sbAllocateDevice.argtypes = [C.c_size_t, C.c_size_t]
        L.sbPrefetchManaged.argtypes = [C.c_void_p]
        L.sbPrefetchManaged.restype = C.c_int
        L.sbFree.argtypes = [C.c_void_p]
        L.sbAllocateHost.restype = C.c_void_p
        L.sbAllocateHost.argtypes = [C.c_size_t]
        L.sbFreeHost.argtypes = [C.c_void_p]
        L.sbCopyToDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.sbCopyToHost.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.getTimeStamp.restype = C.c_double
        L.getTimeResolution.restype = C.c_double
        L.sbTimerCreate.restype = C.c_void_p
        L.sbTimerStart.argtypes = [C.c_void_p]
        L.sbTimerStopMs.argtypes = [C.c_void_p]
        L.sbTimerStopMs.restype = C.c_double
        L.sbTimerDestroy.argtypes = [C.c_void_p]
        L.sbKernelLaunchCount.restype = C.c_size_t
        L.sbMeasureReadBandwidth.restype = C.c_double
        L.sbMeasureReadBandwidth.argtypes = [C.c_size_t, C.c_int]
        L.sbSetDevice.argtypes = [C.c_int]
        L.matrixGenerate.argtypes = [C.POINTER(GMatrix), C.POINTER(Parameter), C.c_int, C.c_int, C.c_bool]
        L.sbGenerateDevice.argtypes = [C.POINTER(GMatrix), C.POINTER(Parameter), C.c_int, C.c_int, C.c_bool]
        L.sbFreeGMatrix.argtypes = [C.POINTER(GMatrix)]
        L.MMMatrixRead.argtypes = [C.POINTER(MMMatrix), C.c_char_p]
        L.matrixConvertfromMM.argtypes = [C.POINTER(MMMatrix), C.POINTER(GMatrix)]
        L.commDistributeMatrix.argtypes = [C.POINTER(Comm), C.POINTER(MMMatrix), C.POINTER(MMMatrix)]
        L.matrixBinWrite.argtypes = [C.POINTER(GMatrix), C.POINTER(Comm), C.c_char_p]
        L.matrixBinRead.argtypes = [C.POINTER(GMatrix), C.POINTER(Comm), C.c_char_p]
        for f in ("CRS", "SCS", "CCRS"):
            getattr(L, "sb%s_convertMatrix" % f).argtypes = [C.c_void_p, C.POINTER(GMatrix)]
            getattr(L, "sb%s_spMVM" % f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            getattr(L, "sb%s_destroyMatrix" % f).argtypes = [C.c_void_p]
            getattr(L, "sb%s_solveCG" % f).argtypes = [C.POINTER(Comm), C.POINTER(Parameter), C.c_void_p]
            getattr(L, "sb%s_solveCG" % f).restype = C.c_int
        L.sbSpmvOrdered.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, U, U]
        L.sbSpmvOrdered.restype = C.c_int
        L.sbSpmvDot.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sbSpmvKernelFamily.restype = C.c_int
        L.sbSpmvKernelFamily.argtypes = [C.c_void_p, C.c_int]
        L.sbTrimPool.argtypes = []
        L.waxpby.argtypes = [U, F, C.c_void_p, F, C.c_void_p, C.c_void_p]
        L.ddot.argtypes = [U, C.c_void_p, C.c_void_p, C.POINTER(F)]
        L.sbSolveCG.argtypes = [C.POINTER(Comm), C.POINTER(Parameter), C.c_void_p, C.c_int, C.POINTER(CGInfo)]
        L.sbSolveCG.restype = C.c_int
        L.sbCGCreate.argtypes = [C.POINTER(Comm), C.POINTER(Parameter), C.c_void_p, C.c_int, C.POINTER(CGInfo)]
        L.sbCGCreate.restype = C.c_void_p
        L.sbCGIterate.argtypes = [C.c_void_p, C.c_int]
        L.sbCGIterate.restype = C.c_int
        L.sbCGFinish.argtypes = [C.c_void_p, C.POINTER(CGInfo), C.c_double]
        L.sbCGFinish.restype = C.c_int
        L.sbSolveGMRES.argtypes = [C.POINTER(Comm), C.POINTER(Parameter), C.c_void_p, C.c_int, C.POINTER(CGInfo), C.c_int]
        L.sbSolveGMRES.restype = C.c_int
        L.sbChebyshevFilter.argtypes = [C.POINTER(Comm), C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
        L.commInit.argtypes = [C.POINTER(Comm), C.c_int, C.c_void_p]
        L.commFinalize.argtypes = [C.POINTER(Comm)]
        L.commPartition.argtypes = [C.POINTER(Comm), C.POINTER(GMatrix)]
        L.commExchange.argtypes = [C.POINTER(Comm), U, C.c_void_p]
        L.commReduction.argtypes = [C.POINTER(F), C.c_int]
        L.sbCommGetUniqueId.argtypes = [C.c_void_p]
        L.sbCommInitRank.argtypes = [C.POINTER(Comm), C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.sbPartitionLocal.restype = C.c_void_p
        L.sbPartitionLocal.argtypes = [C.POINTER(GMatrix), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.sbPartitionRequestSlice.restype = C.POINTER(C.c_int)
        L.sbPartitionRequestSlice.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.sbPartitionFinish.argtypes = [C.c_void_p, C.POINTER(Comm), C.c_void_p, C.c_void_p]
        _configured = True
    return L


# ------------------------------------------------------------------ device memory (allocate.c)
class DeviceBuffer:
    """Plain device memory (sbAllocateDevice): a device pointer plus its size."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = lib().sbAllocateDevice(64, self.nbytes)

    def free(self):
        if self.ptr:
            lib().sbFree(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class UnifiedBuffer(DeviceBuffer):
    """What the reference-ABI allocate() returns (allocate.h:9): unified memory the host may write with plain stores;
    `host(dtype)` is a numpy view of it. The entry points move it to the GPU before their first kernel."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = lib().allocate(64, self.nbytes)

    def host(self, dtype=None):
        dtype = RDT if dtype is None else dtype
        n = self.nbytes // np.dtype(dtype).itemsize
        return np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), (n,))


def allocate(alignment, bytesize):
    assert alignment <= 256
    return UnifiedBuffer(bytesize)


def to_device(a, slots=None):
    """numpy -> device; `slots` over-allocates (zero-filled) in elements, e.g. nrPadded or nc."""
    a = np.ascontiguousarray(a)
    n = len(a) if slots is None else max(slots, len(a))
    if n > len(a):
        a = np.concatenate([a, np.zeros(n - len(a), a.dtype)])
    buf = DeviceBuffer(max(a.nbytes, 8))
    if a.nbytes:
        lib().sbCopyToDevice(buf.ptr, a.ctypes.data, a.nbytes)
    return buf


def vec(a):
    """numpy vector in the library's value type (CG_FLOAT)"""
    return np.ascontiguousarray(a, RDT)


def to_host(ptr, dtype, count):
    out = np.zeros(max(count, 1), dtype)
    p = ptr.ptr if isinstance(ptr, DeviceBuffer) else ptr
    if count:
        lib().sbCopyToHost(out.ctypes.data, p, out.itemsize * count)
    return out[:count]


def getTimeStamp():
    return lib().getTimeStamp()


class EventTimer:
    """CUDA-event timing on the library's stream (replaces getTimeStamp() pairs, timing.c:8-13)."""

    def __init__(self):
        self.h = lib().sbTimerCreate()

    def start(self):
        lib().sbTimerStart(self.h)

    def stop_ms(self):
        return lib().sbTimerStopMs(self.h)


# ------------------------------------------------------------------ matrix sources (matrix.c)
def matrixGenerate(nx, ny, nz, rank=0, size=1, use7pt=False, device=False):
    """matrix.c:30-121. device=False: host arrays like the reference; device=True: arrays in HBM."""
    g = GMatrix()
    p = Parameter(b"generate7P" if use7pt else b"generate", nx, ny, nz, 0, 0.0)
    if device:
        lib().sbGenerateDevice(C.byref(g), C.byref(p), rank, size, use7pt)
    else:
        lib().matrixGenerate(C.byref(g), C.byref(p), rank, size, use7pt)
    g._device = device
    return g


def matrixRead(filename, comm=None):
    """main.c:64-71: MMMatrixRead on the master, commDistributeMatrix, matrixConvertfromMM -> host GMatrix of this rank."""
    L = lib()
    if comm is None:
        comm = Comm()
        comm.rank, comm.size = 0, 1
    mm, local = MMMatrix(), MMMatrix()
    if comm.rank == 0:
        L.MMMatrixRead(C.byref(mm), filename.encode())
    L.commDistributeMatrix(C.byref(comm), C.byref(mm), C.byref(local))
    g = GMatrix()
    L.matrixConvertfromMM(C.byref(local), C.byref(g))
    g._device = False
    return g


def gmatrix_from_csr(rowPtr, col, val, nc=None, startRow=0, totalNr=None):
    """Host GMatrix over numpy arrays (what MMMatrixRead + matrixConvertfromMM would hand over)."""
    rowPtr = np.ascontiguousarray(rowPtr, IDT)
    e = np.zeros(max(len(col), 1), ENTRY_DTYPE)
    e["col"][:len(col)] = col
    e["val"][:len(col)] = val
    g = GMatrix()
    nr = len(rowPtr) - 1
    g.nr, g.nc, g.nnz = nr, (nr if nc is None else nc), len(col)
    g.totalNr, g.totalNnz = (nr if totalNr is None else totalNr), len(col)
    g.startRow, g.stopRow = startRow, startRow + nr - 1
    g.rowPtr, g.entries = rowPtr.ctypes.data, e.ctypes.data
    g._keep = (rowPtr, e)
    g._device = False
    return g


def gmatrix_arrays(g):
    """(rowPtr, col, val) of a host or device GMatrix as numpy arrays."""
    if getattr(g, "_device", False):
        rp = to_host(g.rowPtr, IDT, g.nr + 1)
        e = to_host(g.entries, ENTRY_DTYPE, int(rp[-1]))
    else:
        rp = np.ctypeslib.as_array(C.cast(g.rowPtr, C.POINTER(U)), (g.nr + 1,)).copy()
        n = int(rp[-1])
        e = np.frombuffer((C.c_char * (ENTRY_DTYPE.itemsize * max(n, 1))).from_address(g.entries), ENTRY_DTYPE, count=n)
    return rp, e["col"].copy(), e["val"].copy()


# ------------------------------------------------------------------ format plugins (matrix-<FMT>.c)
def convertMatrix(fmt, g, C_=None, sigma=None):
    m = MATRIX_TYPES[fmt]()
    if fmt == FMT_SCS:
        m.C, m.sigma = C_, sigma          # inputs read from the struct (matrix-SCS.c:40)
    getattr(lib(), "sb%s_convertMatrix" % FMT_NAMES[fmt])(C.byref(m), C.byref(g))
    m._fmt = fmt
    return m


def destroyMatrix(m):
    getattr(lib(), "sb%s_destroyMatrix" % FMT_NAMES[m._fmt])(C.byref(m))


def spMVM(m, x, y):
    """x, y: DeviceBuffer. SCS: y needs nrPadded slots and comes back in permuted row order (matrix-SCS.c:198-228)."""
    getattr(lib(), "sb%s_spMVM" % FMT_NAMES[m._fmt])(C.byref(m), x.ptr, y.ptr)


def crs_arrays(m):
    rp = to_host(m.rowPtr, IDT, m.nr + 1)
    n = int(rp[-1])
    return rp, to_host(m.colInd, IDT, n), to_host(m.val, RDT, n)


def scs_arrays(m):
    return dict(C=m.C, sigma=m.sigma, nr=m.nr, nc=m.nc, nChunks=m.nChunks, nrPadded=m.nrPadded, nElems=m.nElems,
                oldToNewPerm=to_host(m.oldToNewPerm, IDT, m.nr), newToOldPerm=to_host(m.newToOldPerm, IDT, m.nr),
                chunkLens=to_host(m.chunkLens, IDT, m.nChunks), chunkPtr=to_host(m.chunkPtr, IDT, m.nChunks + 1),
                colInd=to_host(m.colInd, IDT, m.nElems), val=to_host(m.val, RDT, m.nElems))


# ------------------------------------------------------------------ Krylov kernels (solver.c)
def waxpby(n, alpha, x, beta, y, w):
    lib().waxpby(n, alpha, x.ptr, beta, y.ptr, w.ptr)


def ddot(n, x, y):
    r = F(0.0)
    lib().ddot(n, x.ptr, y.ptr, C.byref(r))
    return r.value


# ------------------------------------------------------------------ CG (CGSolver.c)
def solveCG(m, itermax, eps, comm=None, generated=True, b=None, x=None, flags=CG_FUSED, want_x=False):
    """Returns (k, history, x_or_None, info). b/x: numpy (host) arrays or DeviceBuffers; history[0] is the initial
    residual norm, history[i] the normr of iteration i (CGSolver.c:116)."""
    if comm is None:
        comm = Comm()
        comm.rank, comm.size = 0, 1
    p = Parameter(b"generate" if generated else b"matrix.mtx", 0, 0, 0, itermax, eps)
    info = CGInfo()
    info.flags = flags
    hist = np.zeros(itermax + 4)
    info.history = hist.ctypes.data_as(C.POINTER(C.c_double))
    info.historyCap = len(hist)
    keep = []
    if b is not None:
        if isinstance(b, DeviceBuffer):
            info.b = b.ptr
        else:
            b = np.ascontiguousarray(b, RDT); keep.append(b); info.b = b.ctypes.data
    xo = None
    if x is not None or want_x:
        if isinstance(x, DeviceBuffer):
            info.x = x.ptr
            xo = x
        else:
            xo = np.zeros(m.nr, RDT) if x is None else np.array(x, RDT)
            info.x = xo.ctypes.data
    k = lib().sbSolveCG(C.byref(comm), C.byref(p), C.byref(m), m._fmt, C.byref(info))
    return k, hist[:info.nhist].copy(), xo, info


# ------------------------------------------------------------------ the solver types main.c:22 only names
def solveGMRES(m, itermax, eps, restart=30, comm=None, generated=True, b=None, x=None, want_x=False):
    """Restarted GMRES(restart). Returns (k, history, x_or_None, info); history[j] = residual estimate after j products."""
    if comm is None:
        comm = Comm()
        comm.rank, comm.size = 0, 1
    p = Parameter(b"generate" if generated else b"matrix.mtx", 0, 0, 0, itermax, eps)
    info = CGInfo()
    info.flags = 0
    hist = np.zeros(itermax + 4)
    info.history = hist.ctypes.data_as(C.POINTER(C.c_double))
    info.historyCap = len(hist)
    keep = []
    if b is not None:
        b = vec(b); keep.append(b); info.b = b.ctypes.data
    xo = None
    if x is not None or want_x:
        xo = np.zeros(m.nr, RDT) if x is None else np.array(x, RDT)
        info.x = xo.ctypes.data
    k = lib().sbSolveGMRES(C.byref(comm), C.byref(p), C.byref(m), m._fmt, C.byref(info), restart)
    return k, hist[:info.nhist].copy(), xo, info


def chebyshevFilter(m, x, degree, lmin, lmax, coef=None, comm=None, want_y=True):
    """y = sum_k coef[k] T_k(A~) x (coef None: T_degree) and moments[k] = x . T_k(A~) x. Returns (y_or_None, moments)."""
    if comm is None:
        comm = Comm()
        comm.rank, comm.size = 0, 1
    xv = vec(x)
    y = np.zeros(m.nr, RDT) if want_y else None
    mu = np.zeros(degree + 1, RDT)
    cf = None if coef is None else vec(coef)
    lib().sbChebyshevFilter(C.byref(comm), C.byref(m), m._fmt, degree, lmin, lmax, None if cf is None else cf.ctypes.data,
                            xv.ctypes.data, None if y is None else y.ctypes.data, mu.ctypes.data)
    return y, mu
