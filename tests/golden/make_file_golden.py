"""Generates the file-input goldens from the reference's OWN sources run as shim ranks (oracle/_ref/libref_mpi_CRS.so:
comm.c, matrix.c, mmio.c and matrixBinfile.c compiled with -D_MPI against oracle/mpi_shim, whose MPI-IO subset works on
POSIX files):

  tests/golden/reference_fixtures/klein_ref.bmx   written by the reference's matrixBinWrite (matrixBinfile.c:38-105)
                                                  from data/matrix_band_klein.mtx, i.e. `sparseBench -c` (main.c:36-47)
  tests/golden/ref_files.npz
      bmx_klein_P<P>_r<r>_{scalars,rowPtr,cols,vals}   matrixBinRead (matrixBinfile.c:107-236) of that file on P ranks
      mm_<mtx>_P<P>_r<r>_{scalars,rowPtr,cols,vals}    MMMatrixRead + commDistributeMatrix (comm.c:311-402, MPI branch)
                                                       + matrixConvertfromMM on P ranks (main.c:64-71)
  scalars = nr, nc, nnz, totalNr, totalNnz, startRow, stopRow.   Re-run with:  python tests/golden/make_file_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

FIX = os.path.join(ROOT, "tests", "golden", "reference_fixtures")
out = {}


def keep(prefix, ranks):
    for r, d in enumerate(ranks):
        out["%s_r%d_scalars" % (prefix, r)] = np.array([d[k] for k in ("nr", "nc", "nnz", "totalNr", "totalNnz", "startRow", "stopRow")], np.int64)
        for f in ("rowPtr", "cols", "vals"):
            out["%s_r%d_%s" % (prefix, r, f)] = d[f]


klein = os.path.join(FIX, "matrix_band_klein.mtx")
bmx = os.path.join(FIX, "klein_ref.bmx")
if os.path.exists(bmx):
    os.unlink(bmx)
ref.mpi_bmx_write(klein, bmx)
for P in (1, 2, 3, 7):
    keep("bmx_klein_P%d" % P, ref.mpi_bmx_read(P, bmx))
for name, Ps in (("matrix_band_klein", (1, 2, 3, 4, 8)), ("test9", (2, 3))):
    for P in Ps:
        keep("mm_%s_P%d" % (name, P), ref.mpi_mm_read(P, os.path.join(FIX, name + ".mtx")))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_files.npz"), **out)
print("wrote %s (%d bytes) and %d arrays" % (bmx, os.path.getsize(bmx), len(out)))
