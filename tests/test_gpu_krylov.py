"""GMRES(m) and the Chebyshev filter (sparsebench_b200/csrc/krylov.cu): the two solver types main.c:22 names but the
reference never implemented. There is no reference behaviour; the pins are a numpy restatement of the same algorithms
(oracle/krylov_ref.py, itself checked against dense linear algebra in tests/test_oracle.py) and mathematical invariants:
the residual estimate equals the true residual b - A x, T_k(A~) equals the dense Chebyshev polynomial."""
import numpy as np
import pytest

from matrices import irregular_spd
from oracle import krylov_ref as kr
from oracle import orc
from sparsebench_b200 import api

pytestmark = pytest.mark.gpu
FMTS = [api.FMT_CRS, api.FMT_SCS, api.FMT_CCRS]


def to_matrix(m, fmt, sigma=256):
    g = api.gmatrix_from_csr(m.rowPtr, m.col, m.val)
    return api.convertMatrix(fmt, g, 32, sigma) if fmt == api.FMT_SCS else api.convertMatrix(fmt, g), g


def nonsymmetric(n, seed=3):
    """diagonally dominant, nonsymmetric: GMRES territory (CG does not apply)"""
    m = irregular_spd(n, seed=seed, max_extra=6)
    rng = np.random.default_rng(seed)
    val = m.val.copy()
    rows = np.repeat(np.arange(m.nr), np.diff(m.rowPtr.astype(np.int64)))
    off = rows != m.col
    val[off] *= 1.0 + 0.8 * rng.random(int(off.sum()))            # breaks the symmetry, keeps the dominance
    return orc.Csr(m.rowPtr, m.col, val)


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("case", ["stencil", "stencil_restarts", "irregular", "nonsymmetric"])
def test_gmres_against_the_numpy_restatement(fmt, case):
    if case.startswith("stencil"):
        m = orc.generate(12, 11, 7)
        _, b, _ = orc.init_vectors(m)
        restart, itermax, eps, generated = (30, 60, 1e-9, True) if case == "stencil" else (5, 80, 1e-8, True)
    else:
        m = irregular_spd(700) if case == "irregular" else nonsymmetric(700)
        b = np.ones(m.nr)
        restart, itermax, eps, generated = 12, 120, 1e-9, False
    kref, href, xref = kr.gmres(m, b, np.zeros(m.nr), itermax, eps, restart)
    A, g = to_matrix(m, fmt, 64)
    k, hist, x, info = api.solveGMRES(A, itermax, eps, restart=restart, generated=generated, b=None if generated else b, want_x=True)
    assert k == kref and len(hist) == len(href)
    scale = np.maximum(href, 1e-10 * href[0])
    # classical Gram-Schmidt in another summation order; every restart recomputes b - A x, whose round-off floor
    # (a few ulp of the initial residual) is an ABSOLUTE error, visible once the residual has dropped by 1e8
    noise = 64 * np.finfo(np.float64).eps * href[0]
    assert float(np.max((np.abs(hist - href) - noise) / scale)) <= 1e-7
    assert float(np.max(np.abs(x - xref))) <= 1e-8 * max(1.0, float(np.max(np.abs(xref))))
    true = float(np.linalg.norm(b - kr.spmv(m, x.astype(np.float64))))
    assert abs(true - hist[-1]) <= 1e-6 * hist[0] and true <= 10 * eps + 1e-9 * hist[0] or k == itermax - 1
    if generated:
        assert info.maxError < 1e-6
    api.destroyMatrix(A)


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("degree", [0, 1, 2, 3, 12, 40])
def test_chebyshev_filter_and_moments(fmt, degree):
    m = orc.generate(10, 9, 8)
    x = 1.0 + 0.01 * np.arange(m.nr) % 7
    lmin, lmax = 0.0, 54.0                                           # Gershgorin bounds of the 27-point stencil
    A, g = to_matrix(m, fmt)
    yref, muref = kr.chebyshev(m, x, degree, lmin, lmax)
    y, mu = api.chebyshevFilter(A, x, degree, lmin, lmax)
    assert np.max(np.abs(y - yref)) <= 1e-11 * max(1.0, float(np.max(np.abs(yref))))
    assert np.max(np.abs(mu - muref)) <= 1e-11 * float(np.max(np.abs(muref)))
    # a filter with coefficients, and moments only
    coef = 1.0 / (1.0 + np.arange(degree + 1))
    yref2, _ = kr.chebyshev(m, x, degree, lmin, lmax, coef)
    y2, mu2 = api.chebyshevFilter(A, x, degree, lmin, lmax, coef=coef)
    assert np.max(np.abs(y2 - yref2)) <= 1e-11 * max(1.0, float(np.max(np.abs(yref2))))
    _, mu3 = api.chebyshevFilter(A, x, degree, lmin, lmax, want_y=False)
    assert np.array_equal(mu2, mu) and np.array_equal(mu3, mu)       # deterministic reductions
    api.destroyMatrix(A)


def test_chebyshev_against_dense_polynomial():
    """the recurrence really is T_k: against cos(k arccos(lambda)) on the eigen-decomposition of a small dense matrix"""
    import scipy.sparse as sp
    m = orc.generate(5, 4, 3)
    x = np.cos(np.arange(m.nr))
    D = sp.csr_matrix((m.val, m.col.astype(np.int64), m.rowPtr.astype(np.int64)), shape=(m.nr, m.nr)).toarray()
    lmin, lmax = 0.0, 54.0
    w, Q = np.linalg.eigh((D - 27.0 * np.eye(m.nr)) / 27.0)
    A, g = to_matrix(m, api.FMT_SCS)
    for degree in (1, 5, 16):
        y, mu = api.chebyshevFilter(A, x, degree, lmin, lmax)
        T = Q @ np.diag(np.cos(degree * np.arccos(np.clip(w, -1, 1)))) @ Q.T
        assert np.max(np.abs(y - T @ x)) <= 1e-11
        assert abs(mu[degree] - x @ (T @ x)) <= 1e-11 * abs(x @ x)
    api.destroyMatrix(A)
