"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the two solver types the reference names but never implemented
(main.c:22 GMRES, CHEBFD; `parity unpinned`: there is no reference behaviour). Same algorithms as
sparsebench_b200/csrc/krylov.cu: restarted GMRES(m) with classical Gram-Schmidt and Givens rotations; Chebyshev filter
y = sum_k c_k T_k(A~) x with the moments mu_k = x . T_k(A~) x."""
import numpy as np


def spmv(m, x):
    """m: oracle Csr (rowPtr, col, val); plain numpy"""
    rows = np.repeat(np.arange(m.nr), np.diff(m.rowPtr.astype(np.int64)))
    y = np.zeros(m.nr)
    np.add.at(y, rows, m.val * x[m.col.astype(np.int64)])
    return y


def gmres(m, b, x0, itermax, eps, restart):
    """returns (k = number of matrix-vector products in Arnoldi steps, history, x)"""
    n = m.nr
    x = np.array(x0, np.float64)
    hist = []
    k = 0
    done = False
    while not done:
        r = b - spmv(m, x)
        beta = float(np.sqrt(r @ r))
        if not hist:
            hist.append(beta)
        if not (beta > eps) or k >= itermax - 1:
            break
        V = np.zeros((restart + 1, n))
        V[0] = r / beta
        H = np.zeros((restart + 1, restart))
        cs, sn, g = np.zeros(restart), np.zeros(restart), np.zeros(restart + 1)
        g[0] = beta
        j = 0
        while j < restart and k < itermax - 1:
            w = spmv(m, V[j])
            h = V[:j + 1] @ w                         # classical Gram-Schmidt: all projections of the same w
            for i in range(j + 1):
                w = w - h[i] * V[i]
            hn = float(np.sqrt(w @ w))
            V[j + 1] = w / hn if hn > 0 else 0.0
            k += 1
            col = np.concatenate([h, [hn]])
            for i in range(j):
                a = cs[i] * col[i] + sn[i] * col[i + 1]
                col[i + 1] = -sn[i] * col[i] + cs[i] * col[i + 1]
                col[i] = a
            d = float(np.hypot(col[j], col[j + 1]))
            cs[j], sn[j] = (col[j] / d, col[j + 1] / d) if d > 0 else (1.0, 0.0)
            col[j], col[j + 1] = d, 0.0
            H[:j + 2, j] = col
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            hist.append(abs(float(g[j + 1])))
            j += 1
            if not (hist[-1] > eps):
                done = True
                break
        if k >= itermax - 1:
            done = True
        y = np.zeros(j)
        for i in range(j - 1, -1, -1):
            y[i] = (g[i] - H[i, i + 1:j] @ y[i + 1:]) / H[i, i] if H[i, i] != 0 else 0.0
        x = x + V[:j].T @ y
    return k, np.array(hist), x


def chebyshev(m, x, degree, lmin, lmax, coef=None):
    """returns (y, moments) for A~ = (A - c I) / e"""
    c, e = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    cf = np.zeros(degree + 1) if coef is None else np.asarray(coef, np.float64)
    if coef is None:
        cf[degree] = 1.0
    t_prev, t = None, np.array(x, np.float64)
    y = cf[0] * t
    mu = [float(x @ t)]
    for k in range(1, degree + 1):
        q = spmv(m, t)
        nxt = (q - c * t) / e if k == 1 else 2.0 / e * (q - c * t) - t_prev
        y = y + cf[k] * nxt
        mu.append(float(x @ nxt))
        t_prev, t = t, nxt
    return y, np.array(mu)
