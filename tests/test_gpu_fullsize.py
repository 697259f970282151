"""Oracle parity at BASELINE.json's REAL sizes (configs[1]/[2]/[4]), not only through size-independent properties:

  128^3 (2.1 M rows, 55.7 M non-zeros)  CRS, CCRS, SELL-32-256: random-x SpMV against the oracle (SELL bit-exact,
                                        CRS/CCRS componentwise bound), SELL integer arrays bit-exact, 30-iteration CG
                                        history <= 1e-10 with the identical iteration count;
  256^3 (16.8 M rows, 449 M non-zeros)  SELL-32-256 integer arrays bit-exact against the oracle (matrix-SCS.c:31-196),
                                        random-x SpMV bit-exact on every row;
  512^3 (134 M rows, 3.6 G non-zeros)   SELL and CRS: chunkPtr[-1] == nElems == 3 614 447 616 (4 % below the u32 wrap),
                                        rowPtr[-1] == 3 609 741 304, random-x SpMV on sampled rows against rows recomputed
                                        on the host from the generator's definition (matrix.c:63-96) -- exercises the
                                        64-bit offset arithmetic of the kernels.
Pins: src/matrix-SCS.c:31-196, :198-228, src/matrix-CRS.c:46-65, src/CGSolver.c:62-141.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import orc
from sparsebench_b200 import api

pytestmark = pytest.mark.gpu

SPMV_TOL = 1e-12
CG_TOL = 1e-10


def dev_spmv(m, x, ny):
    xd = api.to_device(x)
    yd = api.to_device(np.full(max(ny, 1), np.nan))
    api.spMVM(m, xd, yd)
    y = api.to_host(yd, np.float64, ny)
    xd.free(); yd.free()
    return y


@pytest.fixture(scope="module")
def stencil128():
    m = orc.generate(128, 128, 128)
    x = np.random.default_rng(128).standard_normal(m.nr)
    yref = orc.spmv_crs(m, x)
    scale = orc.spmv_crs(orc.Csr(m.rowPtr, m.col, np.abs(m.val)), np.abs(x))
    x0, b, _ = orc.init_vectors(m)
    # The reference's ddot is an OpenMP reduction (solver.c:46-61): its summation order depends on the thread count.
    # Over 2.1 M mostly IDENTICAL terms (the interior of the stencil is translation invariant) a sequential sum rounds
    # in the same direction again and again and drifts by ~1e-10 relative -- measured below against a long double
    # accumulation -- so the history is pinned against three orders: 16 threads (the bench host), 1 thread, and the
    # rounding-free yardstick.
    hists = {}
    for t in (1, 16, 0):
        with orc.dot_threads(t):
            kref, hists[t], xref = orc.cg_crs(m, b, x0, 31, 0.0)
    return dict(m=m, x=x, yref=yref, scale=scale, kref=kref, hists=hists, xref=xref)


@pytest.mark.parametrize("fmt", [api.FMT_CRS, api.FMT_CCRS, api.FMT_SCS])
def test_128_cubed_against_the_oracle(stencil128, fmt):
    """BASELINE.json configs[1] size. Random x sees a wrong column inside the row's neighbour set and any offset
    bug that preserves row sums -- which the x = 1 property test cannot."""
    s = stencil128
    m = s["m"]
    g = api.matrixGenerate(128, 128, 128, device=True)
    A = api.convertMatrix(fmt, g, 32, 256)
    if fmt != api.FMT_CCRS:
        api.lib().sbFreeGMatrix(C.byref(g))
    if fmt == api.FMT_SCS:
        o = orc.scs_convert(m, 32, 256)
        a = api.scs_arrays(A)
        assert (a["nChunks"], a["nrPadded"], a["nElems"]) == (o.nChunks, o.nrPadded, o.nElems)
        for f in ("oldToNewPerm", "newToOldPerm", "chunkLens", "chunkPtr", "colInd", "val"):
            assert np.array_equal(a[f], getattr(o, f)), f
        y = dev_spmv(A, s["x"], o.nrPadded)
        assert np.array_equal(y, orc.spmv_scs(o, s["x"]))            # same summation order: bit-exact
        assert np.all(np.abs(y[o.oldToNewPerm] - s["yref"]) <= SPMV_TOL * s["scale"])
    else:
        rp, col, val = api.crs_arrays(A) if fmt == api.FMT_CRS else api.gmatrix_arrays(g)
        assert np.array_equal(rp, m.rowPtr) and np.array_equal(col, m.col) and np.array_equal(val, m.val)
        y = dev_spmv(A, s["x"], m.nr)
        err = np.abs(y - s["yref"])
        assert np.all(err <= SPMV_TOL * s["scale"]), float(np.max(err / s["scale"]))
    k, hist, x, info = api.solveCG(A, 31, 0.0, want_x=True)
    assert k == s["kref"]
    dev = {t: float(np.max(np.abs(hist - h) / h)) for t, h in s["hists"].items() if len(h) == len(hist)}
    assert len(dev) == 3
    seq_rounding = float(np.max(np.abs(s["hists"][1] - s["hists"][0]) / s["hists"][0]))   # what a sequential sum loses
    assert dev[0] <= CG_TOL and dev[16] <= CG_TOL, dev          # vs exactly-summed dots, vs the 16-thread reference order
    assert dev[1] <= CG_TOL + 2 * seq_rounding, (dev, seq_rounding)  # vs the 1-thread order: its own drift on top
    assert float(np.max(np.abs(x - s["xref"]))) <= 1e-9
    api.destroyMatrix(A)
    if fmt == api.FMT_CCRS:
        api.lib().sbFreeGMatrix(C.byref(g))


def test_256_cubed_sell_arrays_and_spmv_bit_exact():
    """BASELINE.json configs[2] size: every integer array of SELL-32-256 against the oracle's restatement of
    matrix-SCS.c:31-196 (about a minute of CPU), and a random-x SpMV bit-exact on all 16.8 M rows."""
    n = 256
    m = orc.generate(n, n, n)
    assert m.nnz == (3 * n - 2) ** 3
    o = orc.scs_convert(m, 32, 256)
    g = api.matrixGenerate(n, n, n, device=True)
    A = api.convertMatrix(api.FMT_SCS, g, 32, 256)
    api.lib().sbFreeGMatrix(C.byref(g))
    assert (A.nChunks, A.nrPadded, A.nElems) == (o.nChunks, o.nrPadded, o.nElems) == (524288, n ** 3, 450628608)
    for f, dt, cnt in (("oldToNewPerm", np.uint32, A.nr), ("newToOldPerm", np.uint32, A.nr), ("chunkLens", np.uint32, A.nChunks),
                       ("chunkPtr", np.uint32, A.nChunks + 1), ("colInd", np.uint32, A.nElems), ("val", np.float64, A.nElems)):
        got = api.to_host(getattr(A, f), dt, cnt)
        assert np.array_equal(got, getattr(o, f)), f
        del got
    x = np.random.default_rng(256).standard_normal(m.nr)
    y = dev_spmv(A, x, o.nrPadded)
    assert np.array_equal(y, orc.spmv_scs(o, x))
    api.destroyMatrix(A)


def _host_rows(n, rows, x):
    """rows of the 27-point stencil on an n^3 grid recomputed from the generator's definition (matrix.c:63-96):
    entries in (sz, sy, sx) ascending order, 27 on the diagonal, -1 elsewhere. Returns (sequential sum, sum |a||x|)."""
    out, scale = np.zeros(len(rows)), np.zeros(len(rows))
    for t, row in enumerate(rows):
        row = int(row)
        ix, iy, iz = row % n, (row // n) % n, row // (n * n)
        acc, sc = np.float64(0.0), 0.0
        for sz in (-1, 0, 1):
            for sy in (-1, 0, 1):
                for sx in (-1, 0, 1):
                    if 0 <= ix + sx < n and 0 <= iy + sy < n and 0 <= iz + sz < n:
                        col = row + sz * n * n + sy * n + sx
                        v = 27.0 if col == row else -1.0
                        acc = np.float64(acc + np.float64(v * x[col]))       # separate multiply and add, stored order
                        sc += abs(v * x[col])
        out[t], scale[t] = acc, sc
    return out, scale


@pytest.mark.parametrize("fmt", [api.FMT_SCS, api.FMT_CRS])
def test_512_cubed_offsets_and_sampled_rows(fmt):
    """BASELINE.json configs[4] on one GPU: 3.6 G non-zeros, element offsets beyond 2^31 and within 4 % of 2^32."""
    n = 512
    N = n ** 3
    g = api.matrixGenerate(n, n, n, device=True)
    assert int(api.to_host(g.rowPtr + 4 * N, np.uint32, 1)[0]) == (3 * n - 2) ** 3 == 3609741304
    A = api.convertMatrix(fmt, g, 32, 256)
    api.lib().sbFreeGMatrix(C.byref(g))
    rng = np.random.default_rng(512)
    x = rng.standard_normal(N)
    # sampled rows: the first and last planes (offsets near 0 and near the u32 limit), the 2^31 element boundary, random
    rows = np.unique(np.concatenate([np.arange(0, 3), np.arange(N - 3, N), rng.integers(0, N, 600),
                                     N // 2 + rng.integers(-2000, 2000, 200), (N * 16) // 27 + rng.integers(-5000, 5000, 200)]))
    want, scale = _host_rows(n, rows, x)
    if fmt == api.FMT_SCS:
        assert A.nElems == 3614447616 and A.nChunks == N // 32 and A.nrPadded == N
        cp = api.to_host(A.chunkPtr, np.uint32, A.nChunks + 1)
        assert int(cp[-1]) == A.nElems and np.all(np.diff(cp.astype(np.int64)) > 0)
        cl = api.to_host(A.chunkLens, np.uint32, A.nChunks)
        assert np.array_equal(np.diff(cp.astype(np.int64)), cl.astype(np.int64) * 32)
        perm = api.to_host(A.oldToNewPerm, np.uint32, N)
        y = dev_spmv(A, x, N)
        got = y[perm[rows]]
        assert np.array_equal(got, want)                                # sequential stored-order sum: bit-exact
    else:
        rp = api.to_host(A.rowPtr, np.uint32, N + 1)
        assert int(rp[-1]) == 3609741304 and np.all(np.diff(rp.astype(np.int64)) >= 8)
        y = dev_spmv(A, x, N)
        assert np.all(np.abs(y[rows] - want) <= SPMV_TOL * scale)
    # and the whole vector through the x = 1 identity: A 1 = 27 - (len - 1), exact in any summation order
    y1 = dev_spmv(A, np.ones(N), N)
    c = np.full(n, 3.0); c[0] = c[-1] = 2.0
    lens = (c[:, None, None] * c[None, :, None] * c[None, None, :]).reshape(-1)
    b = 27.0 - (lens - 1.0)
    assert np.array_equal(y1[perm] if fmt == api.FMT_SCS else y1, b)
    api.destroyMatrix(A)
