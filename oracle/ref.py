"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/_ref/*.so, i.e. the reference's OWN sources
compiled by oracle/Makefile (strict IEEE build unless noted). Used to pin oracle/sb_oracle.c and,
through bench.py --impl reference / cpu_baseline, as the CPU baseline. Never imported by the product.

Struct layouts mirror /root/reference/src/{matrix.h:29-35, CRSMatrix.h:9-16, SCSMatrix.h:13-27,
CCRSMatrix.h:14-20, comm.h:27-46, parameter.h:8-13} for CG_UINT = unsigned int, CG_FLOAT = double.
"""
import ctypes as C
import os
import re
import tempfile

import numpy as np

from .orc import ENTRY_DTYPE, Csr

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_libs = {}
_libc = C.CDLL(None)

U = C.c_uint32
_HEAD = [("nr", U), ("nc", U), ("nnz", U), ("totalNr", U), ("totalNnz", U), ("startRow", U), ("stopRow", U)]


class GMatrix(C.Structure):
    _fields_ = _HEAD + [("rowPtr", C.c_void_p), ("entries", C.c_void_p)]


class CRSMatrix(C.Structure):
    _fields_ = _HEAD + [("rowPtr", C.c_void_p), ("colInd", C.c_void_p), ("val", C.c_void_p)]


class SCSMatrix(C.Structure):
    _fields_ = _HEAD + [("colInd", C.c_void_p), ("val", C.c_void_p), ("C", U), ("sigma", U), ("nrPadded", U),
                        ("nChunks", U), ("nElems", U), ("chunkPtr", C.c_void_p), ("chunkLens", C.c_void_p),
                        ("oldToNewPerm", C.c_void_p), ("newToOldPerm", C.c_void_p)]


CCRSMatrix = GMatrix


class Comm(C.Structure):  # non-MPI build
    _fields_ = [("rank", C.c_int), ("size", C.c_int), ("logFile", C.c_void_p)]


class Parameter(C.Structure):
    _fields_ = [("filename", C.c_char_p), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("itermax", C.c_int),
                ("eps", C.c_double)]


class MMMatrix(C.Structure):
    _fields_ = [("count", C.c_size_t), ("nr", C.c_int), ("nnz", C.c_int), ("totalNr", C.c_int),
                ("totalNnz", C.c_int), ("startRow", C.c_int), ("stopRow", C.c_int), ("entries", C.c_void_p)]


class RefRankOut(C.Structure):  # oracle/mpi_shim/ref_driver.c
    _fields_ = [("nr", C.c_int), ("nc", C.c_int), ("externalCount", C.c_int), ("totalSendCount", C.c_int),
                ("indegree", C.c_int), ("outdegree", C.c_int), ("startRow", C.c_int), ("stopRow", C.c_int),
                ("sources", C.POINTER(C.c_int)), ("recvCounts", C.POINTER(C.c_int)), ("rdispls", C.POINTER(C.c_int)),
                ("destinations", C.POINTER(C.c_int)), ("sendCounts", C.POINTER(C.c_int)),
                ("sdispls", C.POINTER(C.c_int)), ("elementsToSend", C.POINTER(C.c_int)),
                ("rowPtr", C.POINTER(C.c_uint32)), ("cols", C.POINTER(C.c_uint32)), ("vals", C.POINTER(C.c_double)),
                ("haloProbe", C.POINTER(C.c_double)),
                ("k_solveCG", C.c_int), ("k_redriven", C.c_int), ("nhist", C.c_int),
                ("hist", C.POINTER(C.c_double)), ("x", C.POINTER(C.c_double))]


def available(name="CRS"):
    return os.path.exists(os.path.join(_REF, "libref_%s.so" % name))


def load(name):
    if name not in _libs:
        path = os.path.join(_REF, "libref_%s.so" % name)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(path)
        L.allocate.restype = C.c_void_p
        L.allocate.argtypes = [C.c_size_t, C.c_size_t]
        L.solveCG.restype = C.c_int
        L.getTimeStamp.restype = C.c_double
        L.waxpby.argtypes = [U, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.ddot.argtypes = [U, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.spMVM.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.convertMatrix.argtypes = [C.c_void_p, C.c_void_p]
        _libs[name] = L
    return _libs[name]


class capture_stdout:
    """Capture what C code prints to fd 1 (the reference reports through printf)."""

    def __enter__(self):
        _libc.fflush(None)
        self._saved = os.dup(1)
        self._tmp = tempfile.TemporaryFile(mode="w+b")
        os.dup2(self._tmp.fileno(), 1)
        self.text = ""
        return self

    def __exit__(self, *exc):
        _libc.fflush(None)
        os.dup2(self._saved, 1)
        os.close(self._saved)
        self._tmp.seek(0)
        self.text = self._tmp.read().decode()
        self._tmp.close()
        return False


def gmatrix_from_csr(m):
    """Build the GMatrix the reference's convertMatrix/commPartition consume (keeps numpy owners alive)."""
    g = GMatrix()
    e = m.entries()
    rp = np.ascontiguousarray(m.rowPtr, np.uint32)
    g.nr, g.nc, g.nnz = m.nr, m.nc, m.nnz
    g.totalNr, g.totalNnz = m.totalNr, m.nnz
    g.startRow, g.stopRow = m.startRow, m.startRow + m.nr - 1
    g.rowPtr, g.entries = rp.ctypes.data, e.ctypes.data
    g._keep = (e, rp)
    return g


def csr_from_gmatrix(g):
    rp = np.ctypeslib.as_array(C.cast(g.rowPtr, C.POINTER(C.c_uint32)), (g.nr + 1,)).copy()
    nnz = int(rp[-1])
    raw = (C.c_char * (16 * max(nnz, 1))).from_address(g.entries)
    e = np.frombuffer(raw, ENTRY_DTYPE, count=nnz)
    return Csr(rp, e["col"].copy(), e["val"].copy(), nc=g.nc, startRow=g.startRow, totalNr=g.totalNr)


def generate(nx, ny, nz, use7pt=False, lib="CRS"):
    """matrixGenerate of the reference itself (single rank)."""
    L = load(lib)
    p = Parameter(b"generate7P" if use7pt else b"generate", nx, ny, nz, 10, 0.0)
    g = GMatrix()
    with capture_stdout():
        L.matrixGenerate(C.byref(g), C.byref(p), 0, 1, C.c_bool(use7pt))
    return g


def read_mm(path, lib="CRS"):
    """MMMatrixRead + commDistributeMatrix(single rank) + matrixConvertfromMM of the reference itself."""
    L = load(lib)
    mm, ml, g = MMMatrix(), MMMatrix(), GMatrix()
    comm = Comm(0, 1, None)
    with capture_stdout():
        L.MMMatrixRead(C.byref(mm), path.encode())
        L.commDistributeMatrix(C.byref(comm), C.byref(mm), C.byref(ml))
    ml.totalNr, ml.totalNnz = mm.nr, mm.nnz  # not set by the non-MPI branch (comm.c:404-410)
    L.matrixConvertfromMM(C.byref(ml), C.byref(g))
    return g


def _u32(ptr, n):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), (max(n, 1),))[:n].copy()


def _f64(ptr, n):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), (max(n, 1),))[:n].copy()


def convert_crs(g, lib="CRS"):
    sm = CRSMatrix()
    load(lib).convertMatrix(C.byref(sm), C.byref(g))
    return sm


def crs_arrays(sm):
    rp = _u32(sm.rowPtr, sm.nr + 1)
    return rp, _u32(sm.colInd, int(rp[-1])), _f64(sm.val, int(rp[-1]))


def convert_scs(g, Cc, sigma):
    sm = SCSMatrix()
    sm.C, sm.sigma = Cc, sigma          # inputs read from the struct (matrix-SCS.c:40)
    load("SCS").convertMatrix(C.byref(sm), C.byref(g))
    return sm


def scs_arrays(sm):
    return dict(C=sm.C, sigma=sm.sigma, nr=sm.nr, nChunks=sm.nChunks, nrPadded=sm.nrPadded, nElems=sm.nElems,
                oldToNewPerm=_u32(sm.oldToNewPerm, sm.nr), newToOldPerm=_u32(sm.newToOldPerm, sm.nr),
                chunkLens=_u32(sm.chunkLens, sm.nChunks), chunkPtr=_u32(sm.chunkPtr, sm.nChunks + 1),
                colInd=_u32(sm.colInd, sm.nElems), val=_f64(sm.val, sm.nElems))


def dump_scs(sm):
    """Text of commMatrixDump's SCS branch (comm.c:755-803) -- the format of the golden files."""
    comm = Comm(0, 1, None)
    with capture_stdout() as cap:
        load("SCS").commMatrixDump(C.byref(comm), C.byref(sm))
    return cap.text


def spmv(lib, sm, x, ny):
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros(max(ny, 1))
    load(lib).spMVM(C.byref(sm), x.ctypes.data, y.ctypes.data)
    return y[:ny]


_RES = re.compile(r"Residual = (\S+)")


def solve_cg(sm, generated, itermax, eps, lib="CRS"):
    """The reference's own solveCG (CGSolver.c:62-141). Returns (k, printed residuals as floats, stdout).
    The strict builds print %.17g (see oracle/Makefile), the _fast build the shipped %E."""
    L = load(lib)
    comm = Comm(0, 1, None)
    p = Parameter(b"generate" if generated else b"file.mtx", 0, 0, 0, itermax, eps)
    with capture_stdout() as cap:
        k = L.solveCG(C.byref(comm), C.byref(p), C.byref(sm))
    return k, [float(v) for v in _RES.findall(cap.text)], cap.text


def mpi_run(P, nx, ny, nz, use7pt=False, itermax=10, eps=0.0, do_cg=True):
    """P in-process 'MPI ranks' of the unmodified comm.c (oracle/mpi_shim). Returns a list of dicts."""
    L = load("mpi_CRS")
    out = (RefRankOut * P)()
    with capture_stdout() as cap:
        L.refdrv_run(P, nx, ny, nz, int(use7pt), itermax, C.c_double(eps), int(do_cg), out)
    res = []
    for r in range(P):
        o = out[r]

        def ia(p, n):
            return np.array([p[i] for i in range(n)], np.int32) if n < 64 else \
                np.ctypeslib.as_array(p, (n,)).astype(np.int32)
        nnz = int(o.rowPtr[o.nr])
        d = dict(nr=o.nr, nc=o.nc, externalCount=o.externalCount, totalSendCount=o.totalSendCount,
                 indegree=o.indegree, outdegree=o.outdegree, startRow=o.startRow, stopRow=o.stopRow,
                 sources=ia(o.sources, o.indegree), recvCounts=ia(o.recvCounts, o.indegree),
                 rdispls=ia(o.rdispls, o.indegree), destinations=ia(o.destinations, o.outdegree),
                 sendCounts=ia(o.sendCounts, o.outdegree), sdispls=ia(o.sdispls, o.outdegree),
                 elementsToSend=ia(o.elementsToSend, o.totalSendCount),
                 rowPtr=np.ctypeslib.as_array(o.rowPtr, (o.nr + 1,)).copy(),
                 cols=np.ctypeslib.as_array(o.cols, (max(nnz, 1),))[:nnz].copy(),
                 vals=np.ctypeslib.as_array(o.vals, (max(nnz, 1),))[:nnz].copy(),
                 haloProbe=np.ctypeslib.as_array(o.haloProbe, (max(o.externalCount, 1),))[:o.externalCount].copy(),
                 k_solveCG=o.k_solveCG, k_redriven=o.k_redriven)
        if do_cg:
            d["hist"] = np.ctypeslib.as_array(o.hist, (o.nhist,)).copy()
            d["x"] = np.ctypeslib.as_array(o.x, (o.nr,)).copy()
        res.append(d)
    L.refdrv_free(P, out)
    return res, cap.text


class RefGmOut(C.Structure):  # oracle/mpi_shim/ref_driver.c
    _fields_ = [("nr", C.c_int), ("nc", C.c_int), ("nnz", C.c_int), ("totalNr", C.c_int), ("totalNnz", C.c_int),
                ("startRow", C.c_int), ("stopRow", C.c_int), ("rowPtr", C.POINTER(C.c_uint32)),
                ("cols", C.POINTER(C.c_uint32)), ("vals", C.POINTER(C.c_double))]


def _gm_list(L, P, out):
    res = []
    for r in range(P):
        o = out[r]
        stored = int(o.rowPtr[o.nr])
        res.append(dict(nr=o.nr, nc=o.nc, nnz=o.nnz, totalNr=o.totalNr, totalNnz=o.totalNnz, startRow=o.startRow, stopRow=o.stopRow,
                        rowPtr=np.ctypeslib.as_array(o.rowPtr, (o.nr + 1,)).copy(),
                        cols=np.ctypeslib.as_array(o.cols, (max(stored, 1),))[:stored].copy(),
                        vals=np.ctypeslib.as_array(o.vals, (max(stored, 1),))[:stored].copy()))
    L.refdrv_gm_free(P, out)
    return res


def mpi_mm_read(P, mtx):
    """main.c:64-71 on P shim ranks of the unmodified sources: MMMatrixRead on the master, commDistributeMatrix
    (comm.c:311-402, MPI branch), matrixConvertfromMM. Returns every rank's GMatrix as a dict."""
    L = load("mpi_CRS")
    out = (RefGmOut * P)()
    with capture_stdout():
        L.refdrv_mm_read(P, mtx.encode(), out)
    return _gm_list(L, P, out)


def mpi_bmx_write(mtx, bmx):
    """main.c:36-47 (`-c file.mtx`): the reference's own matrixBinWrite (matrixBinfile.c:38-105) on one shim rank."""
    L = load("mpi_CRS")
    out = (RefGmOut * 1)()
    with capture_stdout():
        L.refdrv_bmx_write(mtx.encode(), bmx.encode(), out)
    return _gm_list(L, 1, out)[0]


def mpi_bmx_read(P, bmx):
    """the reference's own matrixBinRead (matrixBinfile.c:107-236) on P shim ranks."""
    L = load("mpi_CRS")
    out = (RefGmOut * P)()
    with capture_stdout():
        L.refdrv_bmx_read(P, bmx.encode(), out)
    return _gm_list(L, P, out)
