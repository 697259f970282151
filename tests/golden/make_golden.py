"""Generates tests/golden/ref_vectors.npz from the reference's OWN sources (oracle/_ref/*.so, built by
`make -C oracle ref` where /root/reference exists). The .npz is committed so that the GPU box and
any machine without /root/reference can still pin the oracle and the CUDA path against outputs of
the real reference. Re-run with:  python tests/golden/make_golden.py

Contents (all produced by reference code, strict IEEE build, 1 thread):
  scs_<mtx>_C<c>_s<s>_*      convertMatrix (matrix-SCS.c without :42-43) on the 11 test matrices
  spmv_*                     spMVM results (CRS / SCS / CCRS)
  cg_<n>_*                   solveCG return value + printed %.17g residuals + re-driven full history
  mpi_<cfg>_r<rank>_*        commPartition lists / renumbered columns / halo probe from the unmodified
                             comm.c run as P pthread "ranks" (oracle/mpi_shim), plus P-rank CG histories
  klein_*                    data/matrix_band_klein.mtx anchor
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

FIX = os.path.join(ROOT, "tests", "golden", "reference_fixtures")
out = {}

# ---- SELL-C-sigma structures on the hand-drawn 10x10 matrices (C, sigma) incl. sigma > 1 (no golden upstream)
for t in range(11):
    path = os.path.join(FIX, "test%d.mtx" % t)
    for (Cc, sg) in [(1, 1), (2, 1), (4, 1), (2, 4), (4, 8), (32, 256), (3, 5)]:
        g = ref.read_mm(path, "SCS")
        sm = ref.convert_scs(g, Cc, sg)
        a = ref.scs_arrays(sm)
        key = "scs_test%d_C%d_s%d_" % (t, Cc, sg)
        for f in ("oldToNewPerm", "newToOldPerm", "chunkLens", "chunkPtr", "colInd", "val"):
            out[key + f] = a[f]
        out[key + "scalars"] = np.array([a["nChunks"], a["nrPadded"], a["nElems"]], np.int64)
        x = 1.0 + 0.25 * np.arange(10)
        out[key + "spmv"] = ref.spmv("SCS", sm, x, a["nrPadded"])
    g = ref.read_mm(path, "CRS")
    sm = ref.convert_crs(g, "CRS")
    out["spmv_test%d_crs" % t] = ref.spmv("CRS", sm, 1.0 + 0.25 * np.arange(10), 10)

# ---- stencil SpMV and SCS on stencils
for (n, use7) in [(12, False), (6, True)]:
    g = ref.generate(n, n, n, use7, "CRS")
    N = n ** 3
    x = 1.0 + 0.001 * np.arange(N)
    out["spmv_sten%d_%d_crs" % (n, use7)] = ref.spmv("CRS", ref.convert_crs(g, "CRS"), x, N)
    gc = ref.generate(n, n, n, use7, "CCRS")
    out["spmv_sten%d_%d_ccrs" % (n, use7)] = ref.spmv("CCRS", gc, x, N)   # GMatrix aliased as CCRS Matrix
    gs = ref.generate(n, n, n, use7, "SCS")
    for (Cc, sg) in [(32, 1), (32, 256), (8, 64)]:
        sm = ref.convert_scs(gs, Cc, sg)
        a = ref.scs_arrays(sm)
        key = "scs_sten%d_%d_C%d_s%d_" % (n, use7, Cc, sg)
        out[key + "oldToNewPerm"] = a["oldToNewPerm"]
        out[key + "chunkLens"] = a["chunkLens"]
        out[key + "chunkPtr"] = a["chunkPtr"]
        out[key + "colInd_sum"] = np.array([a["colInd"].astype(np.uint64).sum(), a["nElems"]], np.uint64)
        out[key + "spmv"] = ref.spmv("SCS", sm, x, a["nrPadded"])

# ---- CG (CRS): return value, printed residuals, e.g. lagging test with eps > 0 (appendix A1, A2)
for (n, itermax, eps) in [(8, 12, 0.0), (16, 20, 1.0), (16, 60, 1e-6), (10, 150, 1e-9)]:
    g = ref.generate(n, n, n, False, "CRS")
    sm = ref.convert_crs(g, "CRS")
    k, res, _ = ref.solve_cg(sm, True, itermax, eps, "CRS")
    key = "cg_%d_%d_%g_" % (n, itermax, eps)
    out[key + "k"] = np.array([k])
    out[key + "printed"] = np.array(res)

# ---- klein anchor (appendix A4)
g = ref.read_mm(os.path.join(FIX, "matrix_band_klein.mtx"), "CRS")
sm = ref.convert_crs(g, "CRS")
k, res, _ = ref.solve_cg(sm, False, 10, 0.0, "CRS")
out["klein_k"] = np.array([k])
out["klein_printed"] = np.array(res)
out["klein_spmv_ones"] = ref.spmv("CRS", sm, np.ones(100), 100)

# ---- multi-rank: unmodified comm.c under the MPI shim
for (P, nx, ny, nz, use7, itermax) in [(3, 3, 3, 2, False, 5), (2, 4, 3, 2, False, 12), (4, 5, 4, 3, True, 15),
                                        (8, 16, 16, 4, False, 40), (1, 4, 4, 4, False, 8)]:
    ranks, _ = ref.mpi_run(P, nx, ny, nz, use7, itermax, 0.0, True)
    cfg = "mpi_P%d_%dx%dx%d_%d_" % (P, nx, ny, nz, use7)
    for r, d in enumerate(ranks):
        for f in ("sources", "recvCounts", "rdispls", "destinations", "sendCounts", "sdispls", "elementsToSend",
                  "rowPtr", "cols", "haloProbe"):
            out[cfg + "r%d_%s" % (r, f)] = d[f]
        out[cfg + "r%d_scalars" % r] = np.array([d["nr"], d["nc"], d["externalCount"], d["totalSendCount"],
                                                  d["k_solveCG"], d["k_redriven"]], np.int64)
    out[cfg + "hist"] = ranks[0]["hist"]
    out[cfg + "x"] = np.concatenate([d["x"] for d in ranks])

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"), **out)
print("wrote %d arrays" % len(out))
