"""TEST INFRASTRUCTURE ONLY -- MatrixMarket -> CSR the way the reference does it.

Restates MMMatrixRead (/root/reference/src/matrix.c:123-229) and matrixConvertfromMM (:231-269):
coordinate real/integer/pattern, general or symmetric (off-diagonals mirrored right after the entry
they come from), then a sort by column followed by a STABLE sort by row (:221-228), i.e. rows in
ascending order, columns ascending inside a row, duplicates in file order.
"""
import numpy as np

from .orc import Csr


def read_mm(path):
    with open(path) as f:
        banner = f.readline().lower().split()
        if banner[:3] != ["%%matrixmarket", "matrix", "coordinate"]:
            raise ValueError("unsupported MatrixMarket banner: %s" % banner)
        field, symm = banner[3], banner[4]
        if field not in ("real", "integer", "pattern") or symm not in ("general", "symmetric"):
            raise ValueError("unsupported MatrixMarket type")       # matrix.c:139-173
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        M, N, nz = (int(t) for t in line.split())
        rows, cols, vals = [], [], []
        for _ in range(nz):
            t = f.readline().split()
            r, c = int(t[0]) - 1, int(t[1]) - 1                      # :201-202
            v = 1.0 if field == "pattern" else float(t[2])
            rows.append(r); cols.append(c); vals.append(v)
            if symm == "symmetric" and r != c:                       # :208-212
                rows.append(c); cols.append(r); vals.append(v)
    rows = np.array(rows, np.int64)
    cols = np.array(cols, np.int64)
    vals = np.array(vals, np.float64)
    o = np.argsort(cols, kind="stable")                              # :221
    rows, cols, vals = rows[o], cols[o], vals[o]
    o = np.argsort(rows, kind="stable")                              # :224-228
    rows, cols, vals = rows[o], cols[o], vals[o]
    rowPtr = np.zeros(M + 1, np.uint32)                              # :246-257
    np.add.at(rowPtr, rows + 1, 1)
    rowPtr = np.cumsum(rowPtr, dtype=np.uint64).astype(np.uint32)
    return Csr(rowPtr, cols.astype(np.uint32), vals, totalNr=M)
