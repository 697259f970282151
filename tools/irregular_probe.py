"""SpMV off the stencil: synthetic irregular matrices at a size that streams from HBM, every format through the C ABI,
each result checked against scipy's CSR product on the host (componentwise bound; SELL against the same bound, its
summation order differs from scipy's), CUDA-event times of back-to-back launches.

    python tools/irregular_probe.py [--rows 2000000] [--reps 20]

Patterns (rows, seeded):
  banded     row lengths uniform 5..45, columns within +-2000 of the diagonal (a reordered FEM / finite-volume matrix)
  fem81      81 entries in every row, in 3 clusters of 27 around the diagonal (3 unknowns per node of a stencil)
  powerlaw   lengths 3 + Pareto tail up to 20 000 (a few rows longer than one pipeline stage), columns anywhere
  scattered  lengths uniform 8..24, columns anywhere (the worst case for the x gathers: every one misses L1)

Bytes counted: the format's own stream (values + column ids + row pointers / chunk tables, padding included) + 8 N
for y + 8 N for x (x counted once: what an ideal cache would move). For `scattered` and `powerlaw` the gathers cost
one 32-byte sector each, so the real bound is far below this one -- the table says how far.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsebench_b200 import api  # noqa: E402


def make(pattern, n, seed=11):
    rng = np.random.default_rng(seed)
    if pattern == "banded":
        lens = rng.integers(5, 46, n)
    elif pattern == "fem81":
        lens = np.full(n, 81)
    elif pattern == "powerlaw":
        lens = np.minimum(3 + (rng.pareto(1.3, n) * 6).astype(np.int64), 20000)
    else:
        lens = rng.integers(8, 25, n)
    rp = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=rp[1:])
    nnz = int(rp[-1])
    rows = np.repeat(np.arange(n, dtype=np.int64), lens)
    if pattern == "banded":
        col = rows + rng.integers(-2000, 2001, nnz)
    elif pattern == "fem81":
        k = np.arange(nnz, dtype=np.int64) - rp[rows]
        col = rows + (k // 27 - 1) * 40000 + (k % 27 - 13)
    else:
        col = rng.integers(0, n, nnz)
    col = np.clip(col, 0, n - 1)
    # one diagonal entry per row (position 0), rows sorted by column like the MatrixMarket path delivers them
    col[rp[:-1]] = np.arange(n)
    order = np.lexsort((col, rows))
    col = col[order]
    val = rng.uniform(-1.0, 1.0, nnz)
    return rp, col, val


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2000000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--patterns", default="banded,fem81,powerlaw,scattered")
    ap.add_argument("--formats", default="CRS,CCRS,SELL-32-1,SELL-32-256,SELL-32-4096")
    a = ap.parse_args()
    L = api.lib()
    n = a.rows
    t = api.EventTimer()
    peak = L.sbMeasureReadBandwidth(C.c_size_t(4 << 30), 5) if hasattr(L, "sbMeasureReadBandwidth") else 0.0
    print("read-only stream of this GPU: %.0f GB/s" % peak)
    print("%-10s %-12s %9s %9s %9s %8s  %s" % ("pattern", "format", "ms", "GB/s", "GFLOP/s", "of peak", "check vs scipy (max |dy| / sum|a||x|)"))
    bad = 0
    for pattern in a.patterns.split(","):
        rp, col, val = make(pattern, n)
        nnz = int(rp[-1])
        xh = 1.0 + 1e-3 * (np.arange(n) % 1000)
        # copies: scipy canonicalises (sums duplicates) IN PLACE in arrays it was handed
        yref = sp.csr_matrix((val.copy(), col.copy(), rp.copy()), shape=(n, n)) @ xh
        bound = sp.csr_matrix((np.abs(val), col.copy(), rp.copy()), shape=(n, n)) @ np.abs(xh)
        g = api.gmatrix_from_csr(rp.astype(api.IDT), col.astype(api.IDT), val)
        x = api.to_device(xh)
        for name, fmt, sigma in (("CRS", api.FMT_CRS, 0), ("CCRS", api.FMT_CCRS, 0), ("SELL-32-1", api.FMT_SCS, 1),
                                 ("SELL-32-256", api.FMT_SCS, 256), ("SELL-32-4096", api.FMT_SCS, 4096)):
            if name not in a.formats.split(","):
                continue
            A = api.convertMatrix(fmt, g, 32, sigma) if fmt == api.FMT_SCS else api.convertMatrix(fmt, g)
            slots = int(A.nrPadded) if fmt == api.FMT_SCS else n
            y = api.to_device(np.zeros(slots + 64))
            api.spMVM(A, x, y)
            yh = api.to_host(y, np.float64, slots)
            if fmt == api.FMT_SCS:
                yh = yh[api.to_host(A.oldToNewPerm, api.IDT, n).astype(np.int64)]      # SELL rows are permuted
            err = float(np.max(np.abs(yh[:n] - yref) / np.maximum(bound, 1e-300)))
            ok = err <= 1e-12
            bad += not ok
            stream = {"CRS": 12 * nnz + 4 * n, "CCRS": 16 * nnz + 4 * n}.get(name, 12 * int(getattr(A, "nElems", 0)) + 8 * int(getattr(A, "nChunks", 0)))
            B = stream + 16 * n
            for i in range(3):
                api.spMVM(A, x, y)
            t.start()
            for i in range(a.reps):
                api.spMVM(A, x, y)
            ms = t.stop_ms() / a.reps
            print("%-10s %-12s %9.4f %9.1f %9.1f %8.3f  %.2e %s  (nnz %d, stored %d)" % (
                pattern, name, ms, B / ms / 1e6, 2 * nnz / ms / 1e6, B / ms / 1e6 / peak if peak else 0.0, err,
                "ok" if ok else "MISMATCH", nnz, int(getattr(A, "nElems", nnz))), flush=True)
            y.free()
            api.destroyMatrix(A)
        x.free()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
