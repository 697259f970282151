"""GPU parity tests: every hot-path entry point, called through the C ABI, against the CPU oracle on the same
inputs, against the committed reference outputs (tests/golden), and -- at BASELINE.json's full sizes --
through size-independent properties.

Bars (BASELINE.json north_star):
  integer / index work ............ bit-exact
  SpMV ............................ elementwise relative error <= 1e-12 in fp64; the SELL and vector kernels keep
                                    the reference's summation order and rounding, so they are compared BIT-EXACTLY;
                                    the sub-warp CRS/CCRS kernels reassociate the row sum and are gated on the
                                    componentwise bound |dy_i| <= 1e-12 * sum_j |a_ij||x_j|
  CG residual histories ........... <= 1e-10 relative, identical returned iteration count
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import mmio, orc
from sparsebench_b200 import _lib, api

pytestmark = pytest.mark.gpu

SPMV_TOL = 1e-12
CG_TOL = 1e-10


def dev_spmv(m, x, ny):
    xd = api.to_device(x)
    yd = api.to_device(np.full(max(ny, 1), np.nan))
    api.spMVM(m, xd, yd)
    return api.to_host(yd, np.float64, ny)


def assert_spmv_close(y, yref, m, x):
    """componentwise backward-error bound + plain relative error where no cancellation occurs"""
    absA = orc.Csr(m.rowPtr, m.col, np.abs(m.val))
    scale = orc.spmv_crs(absA, np.abs(x))
    err = np.abs(y - yref)
    assert np.all(err <= SPMV_TOL * scale + 0.0), float(np.max(err / np.maximum(scale, 1e-300)))
    big = np.abs(yref) > 1e-3 * scale
    if big.any():
        assert np.max(err[big] / np.abs(yref[big])) <= 1e-9


STENCILS = [(12, 12, 12, False), (9, 7, 5, False), (6, 6, 6, True), (32, 32, 16, False), (1, 1, 7, False), (33, 2, 3, False)]


# ------------------------------------------------------------------------------------------- generator
@pytest.mark.parametrize("nx,ny,nz,use7", STENCILS)
@pytest.mark.parametrize("rank,size", [(0, 1), (1, 3), (2, 3)])
def test_device_generator_bit_exact(nx, ny, nz, use7, rank, size):
    g = api.matrixGenerate(nx, ny, nz, rank, size, use7, device=True)
    rp, col, val = api.gmatrix_arrays(g)
    m = orc.generate(nx, ny, nz, rank, size, use7)
    assert np.array_equal(rp, m.rowPtr) and np.array_equal(col, m.col) and np.array_equal(val, m.val)
    api.lib().sbFreeGMatrix(C.byref(g))


# ------------------------------------------------------------------------------------------- CRS / CCRS
@pytest.mark.parametrize("nx,ny,nz,use7", STENCILS)
@pytest.mark.parametrize("device_input", [False, True])
def test_crs_convert_and_spmv(nx, ny, nz, use7, device_input):
    m = orc.generate(nx, ny, nz, use7pt=use7)
    g = api.matrixGenerate(nx, ny, nz, 0, 1, use7, device=device_input)
    A = api.convertMatrix(api.FMT_CRS, g)
    rp, col, val = api.crs_arrays(A)
    assert np.array_equal(rp, m.rowPtr) and np.array_equal(col, m.col) and np.array_equal(val, m.val)
    assert (A.nr, A.nc, A.nnz, A.totalNr, A.totalNnz, A.startRow, A.stopRow) == \
        (g.nr, g.nc, g.nnz, g.totalNr, g.totalNnz, g.startRow, g.stopRow)          # matrix-CRS.c:14-20
    rng = np.random.default_rng(1)
    for x in (np.ones(m.nr), 1.0 + 0.001 * np.arange(m.nr), rng.standard_normal(m.nr)):
        assert_spmv_close(dev_spmv(A, x, m.nr), orc.spmv_crs(m, x), m, x)
    Cc = api.convertMatrix(api.FMT_CCRS, g)
    x = rng.standard_normal(m.nr)
    assert_spmv_close(dev_spmv(Cc, x, m.nr), orc.spmv_ccrs(m, x), m, x)
    assert np.array_equal(dev_spmv(Cc, x, m.nr), dev_spmv(A, x, m.nr))              # CCRS == CRS
    api.destroyMatrix(A)
    api.destroyMatrix(Cc)


def random_csr(rng, nr, nc, maxlen, empty_frac=0.2):
    lens = rng.integers(0, maxlen + 1, nr)
    lens[rng.random(nr) < empty_frac] = 0
    rp = np.zeros(nr + 1, np.uint32)
    rp[1:] = np.cumsum(lens)
    col = rng.integers(0, nc, int(rp[-1])).astype(np.uint32)
    val = rng.standard_normal(int(rp[-1]))
    return orc.Csr(rp, col, val, nc=nc)


@pytest.mark.parametrize("nr,maxlen", [(1, 3), (257, 5), (1000, 11), (999, 40), (300, 130), (64, 700)])
def test_crs_ccrs_spmv_ragged_rows(nr, maxlen):
    """empty rows, rows longer than a warp, every sub-warp width"""
    rng = np.random.default_rng(nr + maxlen)
    m = random_csr(rng, nr, nr, maxlen)
    g = api.gmatrix_from_csr(m.rowPtr, m.col, m.val)
    x = rng.standard_normal(nr)
    for fmt, ref in ((api.FMT_CRS, orc.spmv_crs), (api.FMT_CCRS, orc.spmv_ccrs)):
        A = api.convertMatrix(fmt, g)
        assert_spmv_close(dev_spmv(A, x, nr), ref(m, x), m, x)
        api.destroyMatrix(A)


# ------------------------------------------------------------------------------------------- SELL-C-sigma
SCS_CASES = [(1, 1), (2, 1), (4, 1), (2, 4), (4, 8), (32, 256), (3, 5)]


@pytest.mark.parametrize("t", range(11))
@pytest.mark.parametrize("Cc,sigma", SCS_CASES)
def test_scs_convert_and_spmv_reference_matrices(golden, fixtures_dir, t, Cc, sigma):
    """the reference's hand-drawn matrices (tests/data/testMatrices) against outputs of the reference itself"""
    m = mmio.read_mm(os.path.join(fixtures_dir, "test%d.mtx" % t))
    g = api.gmatrix_from_csr(m.rowPtr, m.col, m.val)
    A = api.convertMatrix(api.FMT_SCS, g, Cc, sigma)
    a = api.scs_arrays(A)
    key = "scs_test%d_C%d_s%d_" % (t, Cc, sigma)
    assert list(golden[key + "scalars"]) == [a["nChunks"], a["nrPadded"], a["nElems"]]
    for f in ("oldToNewPerm", "newToOldPerm", "chunkLens", "chunkPtr", "colInd", "val"):
        assert np.array_equal(golden[key + f], a[f]), f
    assert (A.nr, A.nc, A.C, A.sigma) == (10, 10, Cc, sigma)                       # matrix-SCS.c:33-39 (C kept: :42-43 skipped)
    x = 1.0 + 0.25 * np.arange(10)
    assert np.array_equal(dev_spmv(A, x, a["nrPadded"]), golden[key + "spmv"])      # bit-exact
    api.destroyMatrix(A)


@pytest.mark.parametrize("Cc", [1, 2, 4])
def test_spmv_reference_golden_file(fixtures_dir, Cc):
    """tests/solver/spmvSCS.c with x = 1: golden test0_spmv_x_1.in"""
    text = open(os.path.join(fixtures_dir, "test0_spmv_x_1.in")).read()
    gold = [float(v) for v in text.split("=")[1].replace(",", " ").split()]
    m = mmio.read_mm(os.path.join(fixtures_dir, "test0.mtx"))
    g = api.gmatrix_from_csr(m.rowPtr, m.col, m.val)
    A = api.convertMatrix(api.FMT_SCS, g, Cc, 1)
    assert list(dev_spmv(A, np.ones(10), A.nrPadded)[:10]) == gold
    B = api.convertMatrix(api.FMT_CRS, g)
    assert list(dev_spmv(B, np.ones(10), 10)) == gold
    E = api.convertMatrix(api.FMT_CCRS, g)
    assert list(dev_spmv(E, np.ones(10), 10)) == gold


@pytest.mark.parametrize("nx,ny,nz,use7", STENCILS)
@pytest.mark.parametrize("Cc,sigma", [(32, 1), (32, 256), (32, 1000), (8, 64), (32, 7)])
def test_scs_stencils_bit_exact(nx, ny, nz, use7, Cc, sigma):
    m = orc.generate(nx, ny, nz, use7pt=use7)
    s = orc.scs_convert(m, Cc, sigma)
    g = api.matrixGenerate(nx, ny, nz, 0, 1, use7, device=True)
    A = api.convertMatrix(api.FMT_SCS, g, Cc, sigma)
    a = api.scs_arrays(A)
    assert (a["nChunks"], a["nrPadded"], a["nElems"]) == (s.nChunks, s.nrPadded, s.nElems)
    for f in ("oldToNewPerm", "newToOldPerm", "chunkLens", "chunkPtr", "colInd", "val"):
        assert np.array_equal(getattr(s, f), a[f]), f
    rng = np.random.default_rng(7)
    for x in (np.ones(m.nr), rng.standard_normal(m.nr)):
        assert np.array_equal(dev_spmv(A, x, s.nrPadded), orc.spmv_scs(s, x))       # bit-exact incl. padded rows
    api.destroyMatrix(A)


@pytest.mark.parametrize("nr,maxlen,sigma", [(1, 3, 1), (31, 9, 4), (33, 9, 64), (1000, 40, 128), (517, 70, 517)])
def test_scs_ragged_rows_bit_exact(nr, maxlen, sigma):
    rng = np.random.default_rng(nr)
    m = random_csr(rng, nr, nr, maxlen)
    s = orc.scs_convert(m, 32, sigma)
    A = api.convertMatrix(api.FMT_SCS, api.gmatrix_from_csr(m.rowPtr, m.col, m.val), 32, sigma)
    a = api.scs_arrays(A)
    for f in ("oldToNewPerm", "newToOldPerm", "chunkLens", "chunkPtr", "colInd", "val"):
        assert np.array_equal(getattr(s, f), a[f]), f
    x = rng.standard_normal(nr)
    assert np.array_equal(dev_spmv(A, x, s.nrPadded), orc.spmv_scs(s, x))
    api.destroyMatrix(A)


@pytest.mark.parametrize("fmt", [api.FMT_CRS, api.FMT_SCS, api.FMT_CCRS])
@pytest.mark.parametrize("nx,ny,nz", [(16, 16, 12), (33, 7, 9), (40, 40, 40)])
def test_ordered_single_launch_spmv_matches_plain_kernel(fmt, nx, ny, nz):
    """the interior-first / boundary-last kernel of the multi-GPU CG (open gate) gives the plain kernel's bits for
    any split, including an empty interior and an empty boundary"""
    m = orc.generate(nx, ny, nz)
    g = api.matrixGenerate(nx, ny, nz, device=True)
    A = make_matrix(fmt, g, 256)
    units = A.nChunks if fmt == api.FMT_SCS else A.nr
    slots = A.nrPadded if fmt == api.FMT_SCS else A.nr
    x = np.random.default_rng(3).standard_normal(m.nr)
    xd = api.to_device(x)
    yref = dev_spmv(A, x, slots)
    plane = max(1, (nx * ny) // (32 if fmt == api.FMT_SCS else 1))
    for lo, hi in [(0, units), (plane, units - plane), (0, units - plane), (plane, units), (units // 2, units // 2),
                   (units, units), (1, 2), (units - 1, units)]:
        yd = api.to_device(np.full(slots, np.nan))
        assert api.lib().sbSpmvOrdered(C.byref(A), fmt, xd.ptr, yd.ptr, lo, hi) == 1
        assert np.array_equal(api.to_host(yd, np.float64, slots), yref), (lo, hi)
    api.destroyMatrix(A)


# ------------------------------------------------------------------------------------------- vector kernels
@pytest.mark.parametrize("n", [1, 2, 3, 255, 1000, 4097, 1 << 20])
def test_waxpby_bit_exact_and_in_place(n):
    rng = np.random.default_rng(n)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    for (a, b) in [(1.0, 0.37), (-2.5, 1.0), (0.3, -0.7), (1.0, 0.0), (1.0, 1.0)]:       # solver.c:23-38
        xd, yd, wd = api.to_device(x), api.to_device(y), api.to_device(np.zeros(n))
        api.waxpby(n, a, xd, b, yd, wd)
        assert np.array_equal(api.to_host(wd, np.float64, n), orc.waxpby(a, x, b, y))
        api.waxpby(n, a, xd, b, yd, yd)                 # w aliases y (CGSolver.c:114)
        assert np.array_equal(api.to_host(yd, np.float64, n), orc.waxpby(a, x, b, y))
        api.waxpby(n, a, xd, b, api.to_device(y), xd)   # w aliases x (CGSolver.c:127)
        assert np.array_equal(api.to_host(xd, np.float64, n), orc.waxpby(a, x, b, y))


@pytest.mark.parametrize("n", [1, 2, 3, 255, 1000, 4097, 1 << 20, 3_000_001])
def test_ddot(n):
    rng = np.random.default_rng(n)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    xd, yd = api.to_device(x), api.to_device(y)
    for (a, b, ah, bh) in ((xd, yd, x, y), (xd, xd, x, x)):               # solver.c:48-58: x == y branch
        got, ref = api.ddot(n, a, b), orc.ddot(ah, bh)
        assert abs(got - ref) <= 1e-12 * float(np.sum(np.abs(ah * bh)))
        assert got == api.ddot(n, a, b)                                   # deterministic reduction order
    assert api.ddot(0, xd, yd) == 0.0


# ------------------------------------------------------------------------------------------- CG
def cg_reference(n, itermax, eps):
    m = orc.generate(n, n, n)
    x, b, _ = orc.init_vectors(m)
    return m, orc.cg_crs(m, b, x, itermax, eps)


def make_matrix(fmt, g, sigma=256):
    return api.convertMatrix(fmt, g, 32, sigma) if fmt == api.FMT_SCS else api.convertMatrix(fmt, g)


def assert_history(hist, href):
    """|hist_k - href_k| <= 1e-10 * max(href_k, 1e-10 * href_0): 1e-10 relative over ten orders of magnitude of
    residual reduction; below that the reference's own history is summation-order noise (its
    1-thread strict and 8-thread fast-math builds already differ by 1e-11 at iteration 15 and by 4x at 1e-20)."""
    assert len(hist) == len(href)
    scale = np.maximum(href, 1e-10 * href[0])
    assert np.max(np.abs(hist - href) / scale) <= CG_TOL


@pytest.mark.parametrize("fmt,sigma", [(api.FMT_CRS, 0), (api.FMT_SCS, 1), (api.FMT_SCS, 256), (api.FMT_CCRS, 0)])
@pytest.mark.parametrize("n,itermax,eps", [(8, 12, 0.0), (16, 20, 1.0), (16, 60, 1e-6), (10, 150, 1e-9), (24, 40, 0.0)])
@pytest.mark.parametrize("flags", [api.CG_FUSED, 0])
def test_cg_history_and_iteration_count(fmt, sigma, n, itermax, eps, flags):
    """CGSolver.c:62-141: identical k (lagging test), history <= 1e-10, same solution -- fused and call-by-call paths"""
    m, (kref, href, xref) = cg_reference(n, itermax, eps)
    g = api.matrixGenerate(n, n, n, device=True)
    A = make_matrix(fmt, g, sigma)
    k, hist, x, info = api.solveCG(A, itermax, eps, flags=flags, want_x=True)
    assert k == kref
    assert_history(hist, href)
    assert np.max(np.abs(x - xref)) <= 1e-9 * max(1.0, np.max(np.abs(xref)))
    assert abs(info.maxError - np.max(np.abs(xref - 1.0))) <= 1e-9
    api.destroyMatrix(A)


@pytest.mark.parametrize("fmt,sigma", [(api.FMT_CRS, 0), (api.FMT_SCS, 1), (api.FMT_SCS, 64), (api.FMT_SCS, 256), (api.FMT_CCRS, 0)])
@pytest.mark.parametrize("n,extra", [(1574, 20), (3001, 6), (700, 60)])
def test_cg_irregular_spd_matrix(fmt, sigma, n, extra):
    """file-input rules (b = 1, CGSolver.c:34-36) on a matrix whose rows really get permuted by the SELL sort and whose
    row lengths select different lanes-per-row variants of the CRS/CCRS kernel"""
    from matrices import irregular_spd
    m = irregular_spd(n, max_extra=extra)
    kref, href, xref = orc.cg_crs(m, np.ones(n), np.zeros(n), 80, 1e-9)
    g = api.gmatrix_from_csr(m.rowPtr, m.col, m.val)
    A = make_matrix(fmt, g, sigma)
    for flags in (api.CG_FUSED, 0):
        k, hist, x, _ = api.solveCG(A, 80, 1e-9, generated=False, flags=flags, want_x=True)
        assert k == kref
        assert_history(hist, href)
        assert np.max(np.abs(x - xref)) <= 1e-9
    api.destroyMatrix(A)


def test_cg_golden_transcripts(golden):
    """printed residuals of the reference's own solveCG (tests/golden/ref_vectors.npz)"""
    for (n, itermax, eps) in [(8, 12, 0.0), (16, 20, 1.0), (16, 60, 1e-6), (10, 150, 1e-9)]:
        g = api.matrixGenerate(n, n, n, device=True)
        A = api.convertMatrix(api.FMT_SCS, g, 32, 256)
        k, hist, _, _ = api.solveCG(A, itermax, eps)
        key = "cg_%d_%d_%g_" % (n, itermax, eps)
        assert k == int(golden[key + "k"][0])
        pf = min(50, max(1, itermax // 10))
        mine = np.array([hist[0]] + [hist[i] for i in range(1, k) if i % pf == 0 or i + 1 == itermax])
        assert_history(mine, golden[key + "printed"])


def test_cg_klein_anchor(fixtures_dir):
    """BASELINE.json configs[0]: history 10, 10, 0; k = 3; x = NaN after the lagging extra iteration"""
    m = mmio.read_mm(os.path.join(fixtures_dir, "matrix_band_klein.mtx"))
    g = api.gmatrix_from_csr(m.rowPtr, m.col, m.val)
    for fmt in (api.FMT_CRS, api.FMT_SCS, api.FMT_CCRS):
        for flags in (api.CG_FUSED, 0):
            A = make_matrix(fmt, g, 16)
            k, hist, x, info = api.solveCG(A, 10, 0.0, generated=False, flags=flags, want_x=True)
            assert k == 3 and list(hist) == [10.0, 10.0, 0.0]
            assert np.isnan(x).all() and info.maxError == -1.0
            api.destroyMatrix(A)


def test_cg_klein_end_to_end_through_the_mm_path(fixtures_dir):
    """main.c:64-71 + :170-198 through the C ABI only: MMMatrixRead -> commDistributeMatrix -> matrixConvertfromMM ->
    commPartition -> convertMatrix -> solveCG on data/matrix_band_klein.mtx"""
    comm = api.Comm()
    api.lib().commInit(C.byref(comm), 0, None)
    g = api.matrixRead(os.path.join(fixtures_dir, "matrix_band_klein.mtx"), comm)
    api.lib().commPartition(C.byref(comm), C.byref(g))
    for fmt in (api.FMT_CRS, api.FMT_SCS, api.FMT_CCRS):
        A = make_matrix(fmt, g, 16)
        k, hist, x, _ = api.solveCG(A, 10, 0.0, comm=comm, generated=False, want_x=True)
        assert k == 3 and list(hist) == [10.0, 10.0, 0.0] and np.isnan(x).all()
        api.destroyMatrix(A)
    api.lib().commFinalize(C.byref(comm))


def test_cg_host_vectors_and_custom_rhs():
    """explicit b / x0 in host memory (SB_CG_HOST_VECTORS path used by bench.py's e2e leg), SELL permutation undone"""
    n = 12
    m = orc.generate(n, n, n)
    rng = np.random.default_rng(2)
    b = rng.standard_normal(m.nr)
    x0 = 0.1 * rng.standard_normal(m.nr)
    kref, href, xref = orc.cg_crs(m, b, x0, 30, 1e-7)
    g = api.matrixGenerate(n, n, n, device=True)
    for fmt, sigma in ((api.FMT_SCS, 256), (api.FMT_CRS, 0)):
        A = make_matrix(fmt, g, sigma)
        k, hist, x, _ = api.solveCG(A, 30, 1e-7, b=b, x=x0, flags=api.CG_FUSED | api.CG_HOST_VECTORS)
        assert k == kref
        assert_history(hist, href)
        assert np.max(np.abs(x - xref)) <= 1e-10
        api.destroyMatrix(A)


def test_cg_itermax_edge_cases():
    g = api.matrixGenerate(6, 6, 6, device=True)
    A = api.convertMatrix(api.FMT_SCS, g, 32, 256)
    m = orc.generate(6, 6, 6)
    x, b, _ = orc.init_vectors(m)
    for itermax in (0, 1, 2, 3):
        kref, href, _ = orc.cg_crs(m, b, x, itermax, 0.0)
        k, hist, _, _ = api.solveCG(A, itermax, 0.0)
        assert k == kref and len(hist) == len(href)
        assert np.allclose(hist, href, rtol=1e-12)


# ------------------------------------------------------------------------------------------- allocate(): host-writable
@pytest.mark.parametrize("fmt", [api.FMT_CRS, api.FMT_SCS, api.FMT_CCRS])
def test_abi_allocate_is_host_writable_and_feeds_the_kernels(fmt):
    """allocate.h:9 through the C ABI: the reference's callers fill what it returns with plain host stores
    (main.c:208-211) and hand it to spMVM / waxpby / ddot. Unified memory: written on the host, multiplied on the GPU
    (moved there once by the entry point), read back on the host."""
    L = api.lib()
    m = orc.generate(11, 9, 7)
    g = api.matrixGenerate(11, 9, 7, device=True)
    A = make_matrix(fmt, g, 4)
    slots = A.nrPadded if fmt == api.FMT_SCS else m.nr
    x, y, w = api.allocate(64, 8 * m.nr), api.allocate(64, 8 * slots), api.allocate(64, 8 * m.nr)
    xv = 1.0 + 0.01 * np.arange(m.nr)
    x.host()[:m.nr] = xv                                    # plain host stores into what allocate() returned
    y.host()[:slots] = 1.0
    assert L.sbPrefetchManaged(x.ptr) == 1 and L.sbPrefetchManaged(api.to_device(xv).ptr) == 0
    api.spMVM(A, x, y)
    L.sbDeviceSynchronize()
    got = np.array(y.host()[:slots])
    if fmt == api.FMT_SCS:
        s = orc.scs_convert(m, 32, 4)
        assert np.array_equal(got, orc.spmv_scs(s, xv))
    else:
        assert_spmv_close(got, orc.spmv_crs(m, xv), m, xv)
    api.waxpby(m.nr, 2.0, x, -1.0, x, w)                    # w = 2x - x
    d = api.ddot(m.nr, x, w)
    L.sbDeviceSynchronize()
    assert np.array_equal(np.array(w.host()[:m.nr]), orc.waxpby(2.0, xv, -1.0, xv))
    assert abs(d - float(np.dot(xv, xv))) <= 1e-12 * float(np.dot(xv, xv))
    x.host()[:m.nr] = 2.0 * xv                              # the host writes again: pages come back, results stay right
    api.spMVM(A, x, y)
    L.sbDeviceSynchronize()
    got2 = np.array(y.host()[:slots])
    assert np.allclose(got2, 2.0 * got, rtol=1e-13, atol=0)
    for b in (x, y, w):
        b.free()
    api.destroyMatrix(A)


# ------------------------------------------------------------------------------------------- drop-in shims
def test_link_time_dropin_names():
    """libsparsebench_b200_CRS.so: the reference's bare symbols convertMatrix / spMVM / solveCG (Makefile:32-34)"""
    S = _lib.load_dropin("CRS")
    m = orc.generate(10, 10, 10)
    g = api.matrixGenerate(10, 10, 10)                       # host GMatrix, like the reference's main.c:168
    A = api.CRSMatrix()
    S.convertMatrix(C.byref(A), C.byref(g))
    A._fmt = api.FMT_CRS
    x = 1.0 + 0.001 * np.arange(m.nr)
    xd, yd = api.to_device(x), api.to_device(np.zeros(m.nr))
    S.spMVM(C.byref(A), C.c_void_p(xd.ptr), C.c_void_p(yd.ptr))
    assert_spmv_close(api.to_host(yd, np.float64, m.nr), orc.spmv_crs(m, x), m, x)
    comm = api.Comm()
    api.lib().commInit(C.byref(comm), 0, None)
    api.lib().commPartition(C.byref(comm), C.byref(g))
    p = api.Parameter(b"generate", 10, 10, 10, 25, 0.0)
    S.solveCG.restype = C.c_int
    xs, b, _ = orc.init_vectors(m)
    assert S.solveCG(C.byref(comm), C.byref(p), C.byref(A)) == orc.cg_crs(m, b, xs, 25, 0.0)[0]
    api.lib().commFinalize(C.byref(comm))


# ------------------------------------------------------------------------------------------- full-size properties
@pytest.mark.parametrize("n,fmt", [(128, api.FMT_CRS), (128, api.FMT_CCRS), (256, api.FMT_SCS)])
def test_full_size_properties(n, fmt):
    """BASELINE.json configs[1]/[2] sizes, no oracle needed: A*1 == 27-(len-1) exactly (integer-valued sums, any
    order), nnz == (3n-2)^3, CG converges to xexact = 1 with a monotone-enough residual history."""
    g = api.matrixGenerate(n, n, n, device=True)
    N = n ** 3
    rp = api.to_host(g.rowPtr, np.uint32, N + 1)
    assert int(rp[-1]) == (3 * n - 2) ** 3
    A = make_matrix(fmt, g, 256)
    if fmt != api.FMT_CCRS:
        api.lib().sbFreeGMatrix(C.byref(g))
    lens = np.diff(rp.astype(np.int64))
    b = 27.0 - (lens - 1.0)
    slots = A.nrPadded if fmt == api.FMT_SCS else N
    y = dev_spmv(A, np.ones(N), slots)
    if fmt == api.FMT_SCS:
        perm = api.to_host(A.oldToNewPerm, np.uint32, N)
        assert np.array_equal(np.sort(perm), np.arange(N, dtype=np.uint32))          # a permutation
        assert np.array_equal(y[perm], b)
        assert A.nElems == int(api.to_host(A.chunkPtr, np.uint32, A.nChunks + 1)[-1])
    else:
        assert np.array_equal(y, b)
    k, hist, _, info = api.solveCG(A, 60, 0.0)
    assert k == 60 and hist[0] == np.sqrt(np.sum(b * b)) and hist[-1] < 1e-6 * hist[0]
    assert info.maxError < 1e-6
    api.destroyMatrix(A)
