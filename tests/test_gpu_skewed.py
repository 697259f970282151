"""CRS / CCRS matrices with skewed row lengths (rows a1, a3 off the stencil): the nnz-balanced row-block kernel
(spmvRowsStreamKernel) against the CPU oracle's CRS product (matrix-CRS.c:54-64 restated) through the
C ABI -- empty rows, rows longer than a thread's share, than a warp's, than a whole block; the kernel-family
decision; the fused dot; run-to-run
identical bits; a CG solve on a heavy-tailed SPD matrix."""
import ctypes as C

import numpy as np
import pytest

from oracle import orc
from sparsebench_b200 import api
from matrices import irregular_spd

pytestmark = pytest.mark.gpu

FAMILY_TILED, FAMILY_BLOCKS, FAMILY_SELL_LONG = 1, 2, 4


def skewed(n, lens, seed, local=False):
    """rows with the given lengths, sorted random columns, values in (-1, 1)"""
    rng = np.random.default_rng(seed)
    lens = np.asarray(lens, np.int64)
    rp = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=rp[1:])
    rows = np.repeat(np.arange(n), lens)
    col = (rows + rng.integers(-40, 41, int(rp[-1]))) % n if local else rng.integers(0, n, int(rp[-1]))
    order = np.lexsort((col, rows))
    return orc.Csr(rp.astype(np.uint32), col[order].astype(np.uint32), rng.uniform(-1.0, 1.0, int(rp[-1])))


def cases():
    rng = np.random.default_rng(5)
    n = 3000
    tail = np.minimum(2 + (rng.pareto(1.1, n) * 5).astype(np.int64), 1500)
    tail[::97] = 0                                       # empty rows
    tail[7] = 64
    tail[8] = 65                                         # either side of the thread-per-row limit
    tail[1200] = 2048                                    # exactly one block
    tail[1201] = 2049                                    # one more: a block of its own, straight from global memory
    tail[2500] = 9000
    yield "heavy_tail", skewed(n, tail, 1)
    yield "uniform_5_45", skewed(4000, rng.integers(5, 46, 4000), 2, local=True)
    ends = np.zeros(700, np.int64)
    ends[0] = ends[-1] = 3000                            # long rows first and last, nothing in between but singletons
    ends[1:-1] = 1
    yield "long_ends", skewed(700, ends, 3)
    yield "all_empty_but_one", skewed(300, np.where(np.arange(300) == 150, 37, 0), 4)
    yield "many_tiny_rows", skewed(9000, np.where(np.arange(9000) % 50 == 0, 60, 1), 6)     # > 2048 rows per 2048 non-zeros


CASES = list(cases())


def spmv(A, x, n):
    xd, yd = api.to_device(x), api.to_device(np.full(n, np.nan))
    api.spMVM(A, xd, yd)
    y = api.to_host(yd, np.float64, n)
    xd.free(); yd.free()
    return y


@pytest.mark.parametrize("name,m", CASES, ids=[c[0] for c in CASES])
def test_skewed_spmv_against_the_oracle(name, m):
    n = m.nr
    x = 1.0 + np.cos(np.arange(n) * 0.37)
    yref = orc.spmv_crs(m, x)
    bound = orc.spmv_crs(orc.Csr(m.rowPtr, m.col, np.abs(m.val)), np.abs(x))
    g = api.gmatrix_from_csr(m.rowPtr, m.col, m.val)
    A = api.convertMatrix(api.FMT_CRS, g)
    B = api.convertMatrix(api.FMT_CCRS, g)
    lens = np.diff(m.rowPtr.astype(np.int64))
    assert lens.max() > 1.2 * lens.mean() + 4.0                                        # the documented rule
    family = FAMILY_BLOCKS
    assert api.lib().sbSpmvKernelFamily(C.byref(A), api.FMT_CRS) == family
    assert api.lib().sbSpmvKernelFamily(C.byref(B), api.FMT_CCRS) == family
    y = spmv(A, x, n)
    assert np.all(np.abs(y - yref) <= 1e-12 * bound), float(np.max(np.abs(y - yref) / np.maximum(bound, 1e-300)))
    assert np.array_equal(y[bound == 0.0], np.zeros(int((bound == 0.0).sum())))          # empty rows give exact zeros
    assert np.array_equal(spmv(B, x, n), y)                                              # CCRS equals CRS bit for bit
    assert np.array_equal(spmv(A, x, n), y)                                              # and every run itself
    # fused x . (A x): same y, dot within the summation bound
    xd, yd, dd = api.to_device(x), api.to_device(np.zeros(n)), api.to_device(np.zeros(8))
    for M, fmt in ((A, api.FMT_CRS), (B, api.FMT_CCRS)):
        api.lib().sbSpmvDot(C.byref(M), fmt, xd.ptr, yd.ptr, dd.ptr)
        assert np.array_equal(api.to_host(yd, np.float64, n), y)
        d = float(api.to_host(dd, np.float64, 1)[0])
        assert abs(d - float(np.dot(x, y))) <= 1e-12 * float(np.dot(np.abs(x), bound))
    for b in (xd, yd, dd):
        b.free()
    api.destroyMatrix(A)
    api.destroyMatrix(B)


@pytest.mark.parametrize("sigma", [1, 256, 4096])
@pytest.mark.parametrize("name,m", [c for c in CASES if c[0] in ("heavy_tail", "long_ends")], ids=["heavy_tail", "long_ends"])
def test_sell_long_chunks_against_the_oracle(name, m, sigma):
    """SELL-32-sigma with chunks longer than 256 columns: those rows are summed by eight warps (componentwise bound),
    every other row keeps the reference's order bit for bit"""
    n = m.nr
    x = 1.0 + np.cos(np.arange(n) * 0.37)
    s = orc.scs_convert(m, 32, sigma)
    yref = orc.spmv_scs(s, x)
    bound = np.zeros(s.nrPadded)
    bound[s.oldToNewPerm.astype(np.int64)] = orc.spmv_crs(orc.Csr(m.rowPtr, m.col, np.abs(m.val)), np.abs(x))
    long_rows = np.repeat(s.chunkLens.astype(np.int64) > 256, 32)
    assert long_rows.any() and not long_rows.all()
    A = api.convertMatrix(api.FMT_SCS, api.gmatrix_from_csr(m.rowPtr, m.col, m.val), 32, sigma)
    assert api.lib().sbSpmvKernelFamily(C.byref(A), api.FMT_SCS) == FAMILY_SELL_LONG
    y = spmv(A, x, s.nrPadded)
    assert np.array_equal(y[~long_rows], yref[~long_rows])
    assert np.all(np.abs(y - yref)[long_rows] <= 1e-12 * bound[long_rows])
    assert np.array_equal(spmv(A, x, s.nrPadded), y)
    # fused dot (vectors in the permuted order of the solver): same y, dot within the summation bound
    xp = np.zeros(s.nrPadded)
    xp[s.oldToNewPerm.astype(np.int64)] = x
    xd, yd, dd = api.to_device(xp), api.to_device(np.zeros(s.nrPadded)), api.to_device(np.zeros(8))
    api.lib().sbSpmvDot(C.byref(A), api.FMT_SCS, xd.ptr, yd.ptr, dd.ptr)
    yp = api.to_host(yd, np.float64, s.nrPadded)
    d = float(api.to_host(dd, np.float64, 1)[0])
    assert abs(d - float(np.dot(xp, yp))) <= 1e-12 * float(np.dot(np.abs(xp), np.abs(yp)) + 1.0) * 64
    for b in (xd, yd, dd):
        b.free()
    api.destroyMatrix(A)


def test_the_stencil_keeps_the_tiled_pipeline():
    g = api.matrixGenerate(24, 24, 24, device=True)
    for fmt in (api.FMT_CRS, api.FMT_CCRS):
        A = api.convertMatrix(fmt, g)
        assert api.lib().sbSpmvKernelFamily(C.byref(A), fmt) == FAMILY_TILED
        api.destroyMatrix(A)
    A = api.convertMatrix(api.FMT_SCS, g, 32, 256)
    assert api.lib().sbSpmvKernelFamily(C.byref(A), api.FMT_SCS) == 0           # the ring kernel alone
    api.destroyMatrix(A)
    api.lib().sbFreeGMatrix(C.byref(g))


def heavy_tailed_spd(n, seed=21):
    """symmetric, strictly diagonally dominant; five hub rows couple (weakly: the CG history must not hinge on the
    summation order -- checked with the oracle's three dot orders, which agree to 3e-14) to ~560 others"""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    i = rng.integers(0, n, 6 * n)
    j = rng.integers(0, n, 6 * n)
    w = -rng.uniform(0.1, 1.0, i.size)
    hi = np.repeat(rng.choice(n, 5, replace=False), 600)
    hj = rng.integers(0, n, hi.size)
    hw = -0.01 * rng.uniform(0.1, 1.0, hi.size)
    i, j, w = np.concatenate([i, hi]), np.concatenate([j, hj]), np.concatenate([w, hw])
    keep = i != j
    S = sp.coo_matrix((w[keep], (i[keep], j[keep])), shape=(n, n)).tocsr()
    S = S + S.T
    S = (S + sp.diags(np.asarray(abs(S).sum(axis=1)).ravel() + 1.0)).tocsr()
    S.sort_indices()
    return orc.Csr(S.indptr.astype(np.uint32), S.indices.astype(np.uint32), S.data.astype(np.float64))


@pytest.mark.parametrize("fmt", [api.FMT_CRS, api.FMT_CCRS, api.FMT_SCS])
def test_cg_on_a_heavy_tailed_spd_matrix(fmt):
    m = heavy_tailed_spd(2500)
    b = np.cos(np.arange(m.nr) * 0.11) + 1.5
    kref, href, xref = orc.cg_crs(m, b, np.zeros(m.nr), 60, 1e-9)
    g = api.gmatrix_from_csr(m.rowPtr, m.col, m.val)
    A = api.convertMatrix(fmt, g, 32, 256) if fmt == api.FMT_SCS else api.convertMatrix(fmt, g)
    assert api.lib().sbSpmvKernelFamily(C.byref(A), fmt) == (FAMILY_SELL_LONG if fmt == api.FMT_SCS else FAMILY_BLOCKS)
    for flags in (api.CG_FUSED, 0):
        k, hist, x, _ = api.solveCG(A, 60, 1e-9, generated=False, b=b, flags=flags, want_x=True)
        assert k == kref and len(hist) == len(href)
        assert np.max(np.abs(hist - href) / np.maximum(href, 1e-10 * href[0])) <= 1e-10
        assert np.max(np.abs(x - xref)) <= 1e-9
    api.destroyMatrix(A)
