"""Test matrices shared by the GPU parity tests."""
import numpy as np

from oracle import orc


def irregular_spd(n, seed=7, max_extra=20):
    """Symmetric, strictly diagonally dominant, random long-range couplings, 2..2*max_extra entries per row."""
    rng = np.random.default_rng(seed)
    rowsets = [set([i]) for i in range(n)]
    for i in range(n):
        for j in rng.integers(0, n, int(rng.integers(0, max_extra))):
            rowsets[i].add(int(j)); rowsets[int(j)].add(i)
        if i + 1 < n:
            rowsets[i].add(i + 1); rowsets[i + 1].add(i)
    rp = np.zeros(n + 1, np.uint32)
    cols, vals = [], []
    for i in range(n):
        for c in sorted(rowsets[i]):
            cols.append(c)
            vals.append(float(len(rowsets[i]) + len(rowsets[c])) if c == i else -1.0 / (1 + ((i + c) % 3)))
        rp[i + 1] = len(cols)
    return orc.Csr(rp, np.array(cols, np.uint32), np.array(vals))
