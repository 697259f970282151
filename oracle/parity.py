"""TEST INFRASTRUCTURE ONLY -- reduced multi-GPU parity check against the CPU oracle, callable from a running
multi-rank job (bench.py runs it, untimed, before it measures anything at N > 1, and prints the outcome in its JSON
line; tests/mgpu_check.py is the full version). Every rank compares what the CUDA path produced for ITS row block
with the oracle's restatement of the reference:

  * halo index lists and renumbered columns of commPartition, bit-exact (comm.c:414-625),
  * halo values delivered by commExchange, exact (comm.c:627-651),
  * CG residual history <= 1e-10 relative and an identical iteration count against the single-rank oracle CG on the
    global problem (CGSolver.c:62-141), solution block <= 1e-9, for SELL-32-256 and CRS on the product path
    (fused kernels, peer-window halo delivery, gated SpMV), two solves per matrix on one partition.
"""
import ctypes as C

import numpy as np

from . import orc

CG_TOL = 1e-10
CASES = [(16, 16, 6, 40, 1e-6), (8, 8, 4, 10, 0.0), (32, 32, 8, 30, 0.0)]


def multi_gpu_parity(api, L, comm, rank, world):
    failures = []
    max_hist_err = 0.0
    ncases = 0

    def check(cond, msg):
        if not cond:
            failures.append(msg)

    for (nx, ny, nz, itermax, eps) in CASES:
        mg = orc.generate(nx, ny, nz * world)
        x0, b, _ = orc.init_vectors(mg)
        kref, href, xref = orc.cg_crs(mg, b, x0, itermax, eps)
        omats = [orc.generate(nx, ny, nz, r, world) for r in range(world)]
        part = orc.Partition(omats)
        n = nx * ny * nz
        g = api.matrixGenerate(nx, ny, nz, rank, world, device=True)
        L.commPartition(C.byref(comm), C.byref(g))
        tag = "%dx%dx%d" % (nx, ny, nz)
        d, o = comm.lists(), part.ranks[rank]
        for f in ("externalCount", "totalSendCount"):
            check(d[f] == o[f], "%s: %s %r != %r" % (tag, f, d[f], o[f]))
        for f in ("sources", "recvCounts", "rdispls", "destinations", "sendCounts", "sdispls", "elementsToSend"):
            check(np.array_equal(d[f], o[f]), "%s: list %s differs" % (tag, f))
        check(np.array_equal(api.gmatrix_arrays(g)[1], omats[rank].col), "%s: renumbered columns differ" % tag)
        nc = n + comm.externalCount
        xs = np.zeros(nc)
        xs[:n] = rank * n + np.arange(n)
        xd = api.to_device(xs)
        L.commExchange(C.byref(comm), n, xd.ptr)
        got = api.to_host(xd, np.float64, nc)
        check(np.array_equal(got[n:], np.asarray(o["externalsReordered"], np.float64)), "%s: halo values differ" % tag)
        xd.free()
        mats = [(api.convertMatrix(api.FMT_SCS, g, 32, 256), "SELL-32-256"), (api.convertMatrix(api.FMT_CRS, g), "CRS")]
        L.sbFreeGMatrix(C.byref(g))
        for _rep in range(2):
            for A, name in mats:
                k, hist, x, _ = api.solveCG(A, itermax, eps, comm=comm, flags=api.CG_FUSED, want_x=True)
                ncases += 1
                check(k == kref, "%s %s: k %d != %d" % (tag, name, k, kref))
                if len(hist) == len(href):
                    err = float(np.max(np.abs(hist - href) / np.maximum(href, 1e-10 * href[0])))
                    max_hist_err = max(max_hist_err, err)
                    check(err <= CG_TOL, "%s %s: history error %.3e" % (tag, name, err))
                else:
                    check(False, "%s %s: history length %d != %d" % (tag, name, len(hist), len(href)))
                xe = float(np.max(np.abs(x - xref[rank * n:(rank + 1) * n])))
                check(xe <= 1e-9 * max(1.0, float(np.max(np.abs(xref)))), "%s %s: solution error %.3e" % (tag, name, xe))
        for A, _ in mats:
            api.destroyMatrix(A)
    return {"checked": True, "ok": not failures, "cases": ncases, "max_hist_err": max_hist_err, "tolerance": CG_TOL,
            "what": "per rank vs CPU oracle: partition lists + renumbered columns bit-exact, halo values exact, CG history "
                    "and k vs single-rank oracle CG on the global problem (SELL-32-256 and CRS, fused peer-window path, "
                    "2 solves per matrix), %d stencil shapes" % len(CASES),
            "failures": failures[:5]}
