"""Parity of one type variant of the library (the reference's compile-time switches PRECISION / UINT_TYPE,
util.h:35-53) against the REFERENCE'S OWN SOURCES compiled with the same switches (oracle/_ref/libref_{CRS,SCS}_<v>.so,
strict IEEE build, built by `make -C oracle ref`):

    SB_VARIANT=f32 python tests/variant_check.py        (also: u64, f32u64; needs a GPU)

One variant per process: the C ABI's struct layouts depend on the two types, sparsebench_b200/api.py follows the
environment variable. Run by tests/test_gpu_variants.py for every variant.

Bars: integer structures (generator, CRS arrays, SELL permutations / chunk tables / column ids) bit-exact; SELL SpMV and
waxpby bit-exact (same operation order, every operation rounded separately in the variant's precision); CRS / CCRS
SpMV within |dy_i| <= tol * sum_j |a_ij||x_j| and ddot within tol * sum |x_i y_i|, tol = 1e-12 (double) or 4e-6 (float:
27-term rows, 2^-24 per operation); CG: identical iteration count, residual history within 1e-10 (double) or 2e-3
(float: the reference prints 7 digits and sums its dot products left to right in float).
"""
import ctypes as C
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparsebench_b200 import api  # noqa: E402

V = api.VARIANT
RDT, IDT = api.RDT, api.IDT
TOL = 1e-12 if RDT == np.float64 else 4e-6
CG_TOL = 1e-10 if RDT == np.float64 else 2e-3
failures = []


def check(cond, msg):
    if not cond:
        failures.append(msg)
        print("FAIL", msg, flush=True)


def ref_lib(fmt):
    path = os.path.join(ROOT, "oracle", "_ref", "libref_%s_%s.so" % (fmt, V))
    L = C.CDLL(path)
    L.solveCG.restype = C.c_int
    L.spMVM.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.convertMatrix.argtypes = [C.c_void_p, C.c_void_p]
    return L


class capture_stdout:
    def __enter__(self):
        import tempfile
        C.CDLL(None).fflush(None)
        self._saved = os.dup(1)
        self._tmp = tempfile.TemporaryFile(mode="w+b")
        os.dup2(self._tmp.fileno(), 1)
        return self

    def __exit__(self, *exc):
        C.CDLL(None).fflush(None)
        os.dup2(self._saved, 1)
        os.close(self._saved)
        self._tmp.seek(0)
        self.text = self._tmp.read().decode()
        self._tmp.close()
        return False


def ref_generate(L, nx, ny, nz):
    g = api.GMatrix()
    p = api.Parameter(b"generate", nx, ny, nz, 10, 0.0)
    with capture_stdout():
        L.matrixGenerate(C.byref(g), C.byref(p), 0, 1, C.c_bool(False))
    g._device = False
    return g


def host_arr(ptr, dtype, n):
    ct = np.ctypeslib.as_ctypes_type(dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), (max(n, 1),))[:n].copy()


def spmv_dev(A, x, slots):
    xd, yd = api.to_device(api.vec(x)), api.to_device(np.full(max(slots, 1), np.nan, RDT))
    api.spMVM(A, xd, yd)
    y = api.to_host(yd, RDT, slots)
    xd.free(); yd.free()
    return y


def main():
    L = api.lib()
    RC, RS = ref_lib("CRS"), ref_lib("SCS")
    print("variant %r: CG_FLOAT = %s, CG_UINT = %s" % (V, RDT, IDT), flush=True)
    for (nx, ny, nz) in [(12, 12, 12), (9, 7, 5), (33, 2, 3)]:
        tag = "%dx%dx%d" % (nx, ny, nz)
        gr = ref_generate(RC, nx, ny, nz)                       # the reference's generator, host arrays
        rp_r, col_r, val_r = api.gmatrix_arrays(gr)
        g = api.matrixGenerate(nx, ny, nz, device=True)
        rp, col, val = api.gmatrix_arrays(g)
        check(rp.dtype == IDT and val.dtype == RDT, "%s: array types" % tag)
        check(np.array_equal(rp, rp_r) and np.array_equal(col, col_r) and np.array_equal(val, val_r), "%s: generator differs from the reference's" % tag)
        n = nx * ny * nz
        x = (1.0 + 0.001 * np.arange(n)).astype(RDT)
        absrow = np.add.reduceat(np.abs(val_r.astype(np.float64) * x.astype(np.float64)[col_r.astype(np.int64)]), rp_r[:-1].astype(np.int64))
        # ---- CRS: the reference's convertMatrix + spMVM on the same GMatrix
        Ar = api.CRSMatrix()
        RC.convertMatrix(C.byref(Ar), C.byref(gr))
        yr = np.zeros(n, RDT)
        RC.spMVM(C.byref(Ar), x.ctypes.data, yr.ctypes.data)
        A = api.convertMatrix(api.FMT_CRS, g)
        a_rp, a_col, a_val = api.crs_arrays(A)
        check(np.array_equal(a_rp, rp_r) and np.array_equal(a_col, host_arr(Ar.colInd, IDT, int(rp_r[-1]))) and
              np.array_equal(a_val, host_arr(Ar.val, RDT, int(rp_r[-1]))), "%s: CRS arrays" % tag)
        y = spmv_dev(A, x, n)
        check(np.all(np.abs(y.astype(np.float64) - yr.astype(np.float64)) <= TOL * absrow), "%s: CRS SpMV, worst %.3e" % (
            tag, float(np.max(np.abs(y.astype(np.float64) - yr.astype(np.float64)) / absrow))))
        # ---- CCRS: equals CRS bit for bit
        B = api.convertMatrix(api.FMT_CCRS, g)
        check(np.array_equal(spmv_dev(B, x, n), y), "%s: CCRS SpMV differs from CRS" % tag)
        # ---- SELL-C-sigma against the reference's own convertMatrix / spMVM (matrix-SCS.c without :42-43)
        for (Cc, sg) in [(32, 256), (4, 8), (1, 1)]:
            grs = ref_generate(RS, nx, ny, nz)
            Sr = api.SCSMatrix()
            Sr.C, Sr.sigma = Cc, sg
            RS.convertMatrix(C.byref(Sr), C.byref(grs))
            S = api.convertMatrix(api.FMT_SCS, g, Cc, sg)
            a = api.scs_arrays(S)
            check((S.nChunks, S.nrPadded, S.nElems) == (Sr.nChunks, Sr.nrPadded, Sr.nElems), "%s SCS %d/%d: scalars" % (tag, Cc, sg))
            for f, dt, cnt in (("oldToNewPerm", IDT, n), ("newToOldPerm", IDT, n), ("chunkLens", IDT, Sr.nChunks),
                               ("chunkPtr", IDT, Sr.nChunks + 1), ("colInd", IDT, Sr.nElems), ("val", RDT, Sr.nElems)):
                check(np.array_equal(a[f], host_arr(getattr(Sr, f), dt, cnt)), "%s SCS %d/%d: %s" % (tag, Cc, sg, f))
            ysr = np.zeros(Sr.nrPadded, RDT)
            RS.spMVM(C.byref(Sr), x.ctypes.data, ysr.ctypes.data)
            check(np.array_equal(spmv_dev(S, x, Sr.nrPadded), ysr), "%s SCS %d/%d: SpMV not bit-exact" % (tag, Cc, sg))
            api.destroyMatrix(S)
        # ---- waxpby (bit-exact: separately rounded multiply and add in the variant's precision) and ddot
        yv = (np.cos(np.arange(n)) * 3).astype(RDT)
        xd, yd, wd = api.to_device(x), api.to_device(yv), api.to_device(np.zeros(n, RDT))
        for (al, be) in [(1.0, 0.5), (-0.25, 1.0), (1.5, -2.5)]:
            api.waxpby(n, al, xd, be, yd, wd)
            al_, be_ = RDT.type(al), RDT.type(be)
            want = (x + be_ * yv) if al == 1.0 else (al_ * x + yv) if be == 1.0 else (al_ * x + be_ * yv)
            check(np.array_equal(api.to_host(wd, RDT, n), want.astype(RDT)), "%s: waxpby(%g, %g)" % (tag, al, be))
        d = api.ddot(n, xd, yd)
        exact = float(np.dot(x.astype(np.float64), yv.astype(np.float64)))
        check(abs(d - exact) <= TOL * float(np.dot(np.abs(x.astype(np.float64)), np.abs(yv.astype(np.float64)))), "%s: ddot %r vs %r" % (tag, d, exact))
        for b_ in (xd, yd, wd):
            b_.free()
        api.destroyMatrix(A)
        api.destroyMatrix(B)
        L.sbFreeGMatrix(C.byref(g))

    # ---- uneven row lengths: the row-block kernel (spmvRowsStreamKernel) against the reference's CRS spMVM
    rng = np.random.default_rng(9)
    n = 2500
    lens = np.minimum(2 + (rng.pareto(1.1, n) * 5).astype(np.int64), 1200)
    lens[::89] = 0
    lens[1000] = 2049                                           # longer than one block: the whole-CTA path
    rp = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=rp[1:])
    rows = np.repeat(np.arange(n), lens)
    col = rng.integers(0, n, int(rp[-1]))
    col = col[np.lexsort((col, rows))]
    val = rng.uniform(-1.0, 1.0, int(rp[-1])).astype(RDT)
    g = api.gmatrix_from_csr(rp, col, val)
    x = (1.0 + np.cos(np.arange(n) * 0.37)).astype(RDT)
    Ar = api.CRSMatrix()
    RC.convertMatrix(C.byref(Ar), C.byref(g))
    yr = np.zeros(n, RDT)
    RC.spMVM(C.byref(Ar), x.ctypes.data, yr.ctypes.data)
    absrow = np.zeros(n)
    np.add.at(absrow, rows, np.abs(val.astype(np.float64) * x.astype(np.float64)[col]))
    A = api.convertMatrix(api.FMT_CRS, g)
    B = api.convertMatrix(api.FMT_CCRS, g)
    check(L.sbSpmvKernelFamily(C.byref(A), api.FMT_CRS) == 2 and L.sbSpmvKernelFamily(C.byref(B), api.FMT_CCRS) == 2, "uneven rows: kernel family")
    y = spmv_dev(A, x, n)
    tol = TOL * np.maximum(1.0, lens / 27.0)                    # TOL is stated for 27-term rows; rounding errors add up per term
    check(np.all(np.abs(y.astype(np.float64) - yr.astype(np.float64)) <= tol * absrow), "uneven rows: CRS SpMV, worst %.3e" % float(
        np.max(np.abs(y.astype(np.float64) - yr.astype(np.float64)) / np.maximum(tol * absrow, 1e-300))))
    check(np.array_equal(spmv_dev(B, x, n), y), "uneven rows: CCRS differs from CRS")
    api.destroyMatrix(A)
    api.destroyMatrix(B)

    # ---- CG against the reference's own solveCG (strict build: residuals printed with %.17g)
    for (n, itermax, eps) in [(8, 12, 0.0), (16, 20, 1.0), (16, 40, 0.0)]:
        gr = ref_generate(RC, n, n, n)
        Ar = api.CRSMatrix()
        RC.convertMatrix(C.byref(Ar), C.byref(gr))
        comm = (C.c_int * 8)(0, 1, 0, 0, 0, 0, 0, 0)            # non-MPI Comm {rank, size, logFile}
        p = api.Parameter(b"generate", n, n, n, itermax, eps)
        with capture_stdout() as cap:
            kref = RC.solveCG(comm, C.byref(p), C.byref(Ar))
        href = np.array([float(v) for v in re.findall(r"Residual = (\S+)", cap.text)])
        g = api.matrixGenerate(n, n, n, device=True)
        for fmt in (api.FMT_CRS, api.FMT_SCS, api.FMT_CCRS):
            A = api.convertMatrix(fmt, g, 32, 256)
            k, hist, x, info = api.solveCG(A, itermax, eps, want_x=True)
            tag = "CG %d^3 itermax %d eps %g fmt %s" % (n, itermax, eps, api.FMT_NAMES[fmt])
            check(k == kref, "%s: k %d != %d" % (tag, k, kref))
            # the reference prints the initial residual and every printFreq-th iteration (CGSolver.c:85-91,118-120)
            freq = max(1, min(50, itermax // 10))
            printed = [0] + [i for i in range(1, k) if i % freq == 0 or i + 1 == itermax]
            if len(href) == len(printed) and max(printed) < len(hist):
                # relative, plus the absolute round-off level of a recursively updated residual (64 ulp of the
                # initial residual): deep into the convergence (1e-12 of the initial residual and below) the history
                # is summation-order noise in the reference as well
                noise = 64 * float(np.finfo(RDT).eps) * href[0]
                err = float(np.max((np.abs(hist[printed] - href) - noise) / np.maximum(href, 1e-300)))
                check(err <= CG_TOL, "%s: history error %.3e" % (tag, err))
            else:
                check(False, "%s: %d printed residuals vs %d expected" % (tag, len(href), len(printed)))
            if eps == 0.0 and itermax >= 40:
                check(float(np.max(np.abs(x.astype(np.float64) - 1.0))) <= (1e-6 if RDT == np.float64 else 1e-2), "%s: solution" % tag)
            api.destroyMatrix(A)
        L.sbFreeGMatrix(C.byref(g))
    print("variant_check %r: %s (%d failures)" % (V, "PASS" if not failures else "FAIL", len(failures)), flush=True)
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
