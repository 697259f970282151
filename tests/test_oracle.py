"""Pins the CPU oracle (oracle/sb_oracle.c) before anything is compared against it:
  1. the reference's 7 golden files + klein anchor (tests/golden/reference_fixtures),
  2. committed outputs of the reference's own sources (tests/golden/ref_vectors.npz),
  3. the live reference libraries under oracle/_ref when they are present.
"""
import os
import re

import numpy as np
import pytest

from oracle import mmio, orc, ref

SCS_CASES = [(1, 1), (2, 1), (4, 1), (2, 4), (4, 8), (32, 256), (3, 5)]


def parse_dump(text):
    """Golden format = commMatrixDump SCS branch (comm.c:755-803)."""
    d = {}
    for line in text.splitlines():
        m = re.match(r"m->(\w+) = (\d+)", line)
        if m:
            d[m.group(1)] = int(m.group(2))
            continue
        m = re.match(r"(\w+): (.*)", line)
        if m:
            d[m.group(1)] = [float(t) for t in m.group(2).replace(",", " ").split()]
    return d


@pytest.mark.parametrize("name,Cc", [("test0", 1), ("test0", 2), ("test0", 4), ("test8", 1), ("test8", 2), ("test8", 4)])
def test_scs_convert_reference_goldens(fixtures_dir, name, Cc):
    """tests/matrix/convertSCS.c:11-90 with (C,sigma) in {(1,1),(2,1),(4,1)} (tests/matrix/matrixTests.c:44-46)."""
    gold = parse_dump(open(os.path.join(fixtures_dir, "%s_C_%d_sigma_1.in" % (name, Cc))).read())
    m = mmio.read_mm(os.path.join(fixtures_dir, name + ".mtx"))
    s = orc.scs_convert(m, Cc, 1)
    assert gold["nr"] == m.nr and gold["nnz"] == m.nnz and gold["C"] == Cc and gold["sigma"] == 1
    assert gold["nChunks"] == s.nChunks and gold["nrPadded"] == s.nrPadded and gold["nElems"] == s.nElems
    assert gold["stopRow"] == m.nr          # the old harness stored an exclusive stopRow (convertSCS.c:56-59)
    for f in ("oldToNewPerm", "newToOldPerm", "chunkLens", "chunkPtr", "colInd"):
        assert list(getattr(s, f)) == [int(v) for v in gold[f]], f
    assert list(s.val) == gold["val"]


@pytest.mark.parametrize("fmt", [(0, 0), (1, 1), (2, 1), (4, 1)])
def test_spmv_reference_golden(fixtures_dir, fmt):
    """tests/solver/spmvSCS.c:19-139 with x = 1; golden test0_spmv_x_1.in."""
    text = open(os.path.join(fixtures_dir, "test0_spmv_x_1.in")).read()
    gold = [float(t) for t in text.split("=")[1].replace(",", " ").split()]
    m = mmio.read_mm(os.path.join(fixtures_dir, "test0.mtx"))
    if fmt == (0, 0):
        y = orc.spmv_crs(m, np.ones(10))
    else:
        s = orc.scs_convert(m, fmt[0], fmt[1])
        y = orc.spmv_scs(s, np.ones(10))[s.oldToNewPerm]   # sigma=1: identity permutation
    assert list(y[:10]) == gold
    assert list(orc.spmv_ccrs(m, np.ones(10))) == gold


@pytest.mark.parametrize("t", range(11))
@pytest.mark.parametrize("Cc,sigma", SCS_CASES)
def test_scs_convert_vs_reference_outputs(golden, fixtures_dir, t, Cc, sigma):
    m = mmio.read_mm(os.path.join(fixtures_dir, "test%d.mtx" % t))
    s = orc.scs_convert(m, Cc, sigma)
    key = "scs_test%d_C%d_s%d_" % (t, Cc, sigma)
    assert list(golden[key + "scalars"]) == [s.nChunks, s.nrPadded, s.nElems]
    for f in ("oldToNewPerm", "newToOldPerm", "chunkLens", "chunkPtr", "colInd", "val"):
        assert np.array_equal(golden[key + f], getattr(s, f)), f
    x = 1.0 + 0.25 * np.arange(10)
    assert np.array_equal(golden[key + "spmv"], orc.spmv_scs(s, x))       # bit-exact, same summation order
    assert np.array_equal(golden["spmv_test%d_crs" % t], orc.spmv_crs(m, x))


def test_scs_sigma_gt1_known_answer(fixtures_dir):
    """test9, C=2, sigma=4 (behaviour of the fixed source, no upstream golden)."""
    m = mmio.read_mm(os.path.join(fixtures_dir, "test9.mtx"))
    s = orc.scs_convert(m, 2, 4)
    assert list(s.oldToNewPerm) == [0, 2, 1, 3, 6, 4, 7, 5, 9, 8]
    assert list(s.newToOldPerm) == [0, 2, 1, 3, 5, 7, 4, 6, 9, 8]
    assert list(s.chunkLens) == [10, 2, 3, 1, 2] and list(s.chunkPtr) == [0, 20, 24, 30, 32, 36]
    assert list(orc.spmv_scs(s, np.ones(10))) == [245, 96, 48, 44, 745, 261, 55, 77, 1111, 99]


@pytest.mark.parametrize("n,use7", [(12, False), (6, True)])
def test_stencil_spmv_and_scs(golden, n, use7):
    m = orc.generate(n, n, n, use7pt=use7)
    N = n ** 3
    x = 1.0 + 0.001 * np.arange(N)
    y = orc.spmv_crs(m, x)
    assert np.array_equal(golden["spmv_sten%d_%d_crs" % (n, use7)], y)
    assert np.array_equal(golden["spmv_sten%d_%d_ccrs" % (n, use7)], orc.spmv_ccrs(m, x))
    if not use7:   # appendix A3
        assert m.nnz == 34 ** 3 and y[0] == 19.372 and y[777] == 1.7770000000000072
    for (Cc, sg) in [(32, 1), (32, 256), (8, 64)]:
        s = orc.scs_convert(m, Cc, sg)
        key = "scs_sten%d_%d_C%d_s%d_" % (n, use7, Cc, sg)
        assert np.array_equal(golden[key + "oldToNewPerm"], s.oldToNewPerm)
        assert np.array_equal(golden[key + "chunkLens"], s.chunkLens)
        assert np.array_equal(golden[key + "chunkPtr"], s.chunkPtr)
        assert list(golden[key + "colInd_sum"]) == [int(s.colInd.astype(np.uint64).sum()), s.nElems]
        assert np.array_equal(golden[key + "spmv"], orc.spmv_scs(s, x))


@pytest.mark.parametrize("n,itermax,eps", [(8, 12, 0.0), (16, 20, 1.0), (16, 60, 1e-6), (10, 150, 1e-9)])
def test_cg_history_and_iteration_count(golden, n, itermax, eps):
    """CGSolver.c:62-141: same returned k (lagging convergence test) and bit-identical printed residuals."""
    m = orc.generate(n, n, n)
    x, b, _ = orc.init_vectors(m)
    k, hist, xs = orc.cg_crs(m, b, x, itermax, eps)
    key = "cg_%d_%d_%g_" % (n, itermax, eps)
    assert k == int(golden[key + "k"][0])
    printed = golden[key + "printed"]           # initial + every printFreq-th iteration (CGSolver.c:85-91,118)
    pf = min(50, max(1, itermax // 10))
    want = [hist[0]] + [hist[i] for i in range(1, k) if i % pf == 0 or i + 1 == itermax]
    assert list(printed) == want


def test_cg_lagging_test_transcript():
    """-x16 -y16 -z16 -i20 -e1.0 returns 13 although ||r|| < eps at k=12."""
    m = orc.generate(16, 16, 16)
    x, b, _ = orc.init_vectors(m)
    k, hist, _ = orc.cg_crs(m, b, x, 20, 1.0)
    assert k == 13 and hist[0] == 408.1078288883956 and hist[10] == 4.6756296922759963 and hist[12] == 0.958729111558553


def test_klein_anchor(golden, fixtures_dir):
    """BASELINE.json configs[0]: data/matrix_band_klein.mtx, CRS, history 10, 10, 0 and k = 3 (appendix A4)."""
    m = mmio.read_mm(os.path.join(fixtures_dir, "matrix_band_klein.mtx"))
    assert m.nr == 100 and m.nnz == 298
    x, b, _ = orc.init_vectors(m, generated=False)
    k, hist, xs = orc.cg_crs(m, b, x, 10, 0.0)
    assert k == int(golden["klein_k"][0]) == 3
    assert list(hist) == list(golden["klein_printed"]) == [10.0, 10.0, 0.0]
    assert np.array_equal(orc.spmv_crs(m, np.ones(100)), golden["klein_spmv_ones"])
    assert np.isnan(xs).all()        # alpha = 0/0 in the extra (lagging) iteration


MPI_CASES = [(3, 3, 3, 2, False, 5), (2, 4, 3, 2, False, 12), (4, 5, 4, 3, True, 15), (8, 16, 16, 4, False, 40),
             (1, 4, 4, 4, False, 8)]


@pytest.mark.parametrize("P,nx,ny,nz,use7,itermax", MPI_CASES)
def test_partition_lists_vs_unmodified_comm_c(golden, P, nx, ny, nz, use7, itermax):
    """comm.c:414-625 run as P pthread ranks (oracle/mpi_shim) vs the serial restatement: bit-exact lists."""
    mats = [orc.generate(nx, ny, nz, r, P, use7) for r in range(P)]
    part = orc.Partition(mats)
    cfg = "mpi_P%d_%dx%dx%d_%d_" % (P, nx, ny, nz, int(use7))
    for r in range(P):
        d = part.ranks[r]
        sc = golden[cfg + "r%d_scalars" % r]
        assert [mats[r].nr, mats[r].nc, d["externalCount"], d["totalSendCount"]] == list(sc[:4])
        for f in ("sources", "recvCounts", "rdispls", "destinations", "sendCounts", "sdispls", "elementsToSend"):
            assert np.array_equal(golden[cfg + "r%d_%s" % (r, f)], d[f]), (r, f)
        assert np.array_equal(golden[cfg + "r%d_cols" % r], mats[r].col)
        assert np.array_equal(golden[cfg + "r%d_rowPtr" % r], mats[r].rowPtr)
    # halo probe: exchange a vector of global row ids (comm.c:627-651)
    xs = [np.concatenate([m.startRow + np.arange(m.nr, dtype=np.float64), -np.ones(m.nc - m.nr)]) for m in mats]
    part.exchange(xs)
    for r in range(P):
        assert np.array_equal(xs[r][mats[r].nr:], golden[cfg + "r%d_haloProbe" % r])
        assert np.array_equal(xs[r][mats[r].nr:], part.ranks[r]["externalsReordered"].astype(np.float64))
    # P-rank CG: identical k and bit-identical history (dots combined in ascending rank order in both)
    bs, x0 = [], []
    for m in mats:
        x, b, _ = orc.init_vectors(m)
        bs.append(b); x0.append(x)
    k, hist = part.cg(bs, x0, itermax, 0.0)
    assert k == int(golden[cfg + "r0_scalars"][4]) == int(golden[cfg + "r0_scalars"][5])
    assert np.array_equal(hist, golden[cfg + "hist"])
    assert np.array_equal(np.concatenate(x0), golden[cfg + "x"])


def test_partition_worked_example():
    """3x3x2 per rank, 3 ranks."""
    mats = [orc.generate(3, 3, 2, r, 3) for r in range(3)]
    part = orc.Partition(mats)
    assert list(part.ranks[1]["externalsReordered"]) == [9, 10, 12, 13, 11, 14, 15, 16, 17, 36, 37, 39, 40, 38, 41, 42, 43, 44]
    assert list(part.ranks[0]["elementsToSend"]) == [9, 10, 12, 13, 11, 14, 15, 16, 17]
    assert list(part.ranks[1]["elementsToSend"][:9]) == [0, 1, 3, 4, 2, 5, 6, 7, 8]
    assert list(mats[1].col[:12]) == [18, 19, 20, 21, 0, 1, 3, 4, 9, 10, 12, 13]


def test_multi_rank_cg_matches_single_rank():
    """Row-block CG over 4 ranks == single-rank CG on the stacked domain up to dot-product rounding."""
    mats = [orc.generate(6, 5, 3, r, 4) for r in range(4)]
    part = orc.Partition(mats)
    bs, xs = zip(*[orc.init_vectors(m)[1::-1] for m in mats])
    k, hist = part.cg(list(bs), list(xs), 30, 1e-8)
    m1 = orc.generate(6, 5, 12)
    x, b, _ = orc.init_vectors(m1)
    k1, h1, x1 = orc.cg_crs(m1, b, x, 30, 1e-8)
    assert k == k1
    assert np.allclose(hist, h1, rtol=1e-10, atol=0)


# ---- live cross-checks against the compiled reference (skipped where oracle/_ref was not built)
needs_ref = pytest.mark.skipif(not ref.available("CRS"), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("nx,ny,nz,use7", [(7, 5, 3, False), (4, 4, 4, True), (20, 20, 20, False)])
def test_live_generator_convert_spmv(nx, ny, nz, use7):
    g = ref.generate(nx, ny, nz, use7)
    mr = ref.csr_from_gmatrix(g)
    m = orc.generate(nx, ny, nz, use7pt=use7)
    assert np.array_equal(m.rowPtr, mr.rowPtr) and np.array_equal(m.col, mr.col) and np.array_equal(m.val, mr.val)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(m.nr)
    assert np.array_equal(ref.spmv("CRS", ref.convert_crs(g), x, m.nr), orc.spmv_crs(m, x))
    gs = ref.generate(nx, ny, nz, use7, "SCS")
    for (Cc, sg) in [(32, 256), (16, 7), (5, 1000)]:
        a = ref.scs_arrays(ref.convert_scs(gs, Cc, sg))
        s = orc.scs_convert(m, Cc, sg)
        for f in ("oldToNewPerm", "newToOldPerm", "chunkLens", "chunkPtr", "colInd", "val"):
            assert np.array_equal(a[f], getattr(s, f)), f


@needs_ref
def test_live_mm_reader(fixtures_dir):
    for name in ["test%d.mtx" % t for t in range(11)] + ["matrix_band_klein.mtx"]:
        path = os.path.join(fixtures_dir, name)
        mr = ref.csr_from_gmatrix(ref.read_mm(path))
        m = mmio.read_mm(path)
        assert np.array_equal(m.rowPtr, mr.rowPtr) and np.array_equal(m.col, mr.col) and np.array_equal(m.val, mr.val)


@needs_ref
def test_live_waxpby_ddot():
    import ctypes as C
    L = ref.load("CRS")
    rng = np.random.default_rng(5)
    x, y = rng.standard_normal(1001), rng.standard_normal(1001)
    for (a, b) in [(1.0, 0.37), (-2.5, 1.0), (0.3, -0.7), (1.0, 0.0)]:
        w = np.zeros(1001)
        L.waxpby(1001, a, x.ctypes.data, b, y.ctypes.data, w.ctypes.data)
        assert np.array_equal(w, orc.waxpby(a, x, b, y))
    res = C.c_double(0)
    L.ddot(1001, x.ctypes.data, y.ctypes.data, C.byref(res))
    assert res.value == orc.ddot(x, y)


def test_krylov_restatement_against_dense_math():
    """oracle/krylov_ref.py (GMRES(m), Chebyshev filter: no reference behaviour exists) against dense linear algebra"""
    import scipy.sparse as sp

    from oracle import krylov_ref as kr
    m = orc.generate(6, 5, 4)
    x0, b, _ = orc.init_vectors(m)
    for restart in (4, 30):
        k, h, x = kr.gmres(m, b, x0, 80, 1e-10, restart)
        assert h[-1] <= 1e-10 and np.max(np.abs(x - 1.0)) <= 1e-9
        assert abs(np.linalg.norm(b - kr.spmv(m, x)) - h[-1]) <= 1e-9
        assert len(h) == k + 1 and np.all(np.diff(h) <= 1e-12 * h[0])          # GMRES residuals never grow
    D = sp.csr_matrix((m.val, m.col.astype(np.int64), m.rowPtr.astype(np.int64)), shape=(m.nr, m.nr)).toarray()
    xx = 1.0 + 0.01 * np.arange(m.nr)
    w, Q = np.linalg.eigh((D - 27.0 * np.eye(m.nr)) / 27.0)
    for degree in (0, 1, 2, 9):
        y, mu = kr.chebyshev(m, xx, degree, 0.0, 54.0)
        T = Q @ np.diag(np.cos(degree * np.arccos(np.clip(w, -1, 1)))) @ Q.T
        assert np.max(np.abs(y - T @ xx)) <= 1e-11 and abs(mu[degree] - xx @ (T @ xx)) <= 1e-10
