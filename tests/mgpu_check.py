"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py

Every rank builds its z-slab on its GPU through the C ABI (matrixGenerate -> commPartition -> convertMatrix ->
solveCG) and compares with the CPU oracle: halo index lists bit-exact against the oracle's restatement of
comm.c:414-625, halo exchange values exact, CG residual history within 1e-10 of the single-rank oracle run on the
global problem with an identical iteration count, solution slab within 1e-9. Exits non-zero on any mismatch.
Also imported by tests/test_gpu_multi.py, which launches it when at least two GPUs are visible.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import orc  # noqa: E402
from sparsebench_b200 import api  # noqa: E402

CG_TOL = 1e-10


def check(cond, msg, failures):
    if not cond:
        failures.append(msg)


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    L = api.lib()
    comm = api.Comm()
    L.commInit(C.byref(comm), 0, None)           # reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (comm.h:48)
    assert (comm.rank, comm.size) == (rank, world)
    failures = []
    cases = [(6, 5, 1, 8, 0.0), (8, 8, 4, 10, 0.0), (16, 16, 6, 60, 1e-6), (12, 10, 3, 30, 1e-3), (32, 32, 8, 40, 0.0)]
    fmts = [(api.FMT_CRS, 0), (api.FMT_SCS, 1), (api.FMT_SCS, 256), (api.FMT_CCRS, 0)]
    for (nx, ny, nz, itermax, eps) in cases:
        # oracle: global problem on one rank + the reference's partition lists for every rank
        mg = orc.generate(nx, ny, nz * world)
        x0, b, _ = orc.init_vectors(mg)
        kref, href, xref = orc.cg_crs(mg, b, x0, itermax, eps)
        omats = [orc.generate(nx, ny, nz, r, world) for r in range(world)]
        part = orc.Partition(omats)
        n = nx * ny * nz
        # ONE partition per case, shared by every format and every solve below: the Comm's peer-mapped halo vector
        # is registered by the first gated solve and only borrowed by the later ones (different matrices, different
        # row permutations, the non-overlapped and call-by-call paths in between)
        g = api.matrixGenerate(nx, ny, nz, rank, world, device=True)
        L.commPartition(C.byref(comm), C.byref(g))
        tag = "%dx%dx%d" % (nx, ny, nz)
        d, o = comm.lists(), part.ranks[rank]
        for f in ("externalCount", "totalSendCount"):
            check(d[f] == o[f], "%s: %s %r != %r" % (tag, f, d[f], o[f]), failures)
        for f in ("sources", "recvCounts", "rdispls", "destinations", "sendCounts", "sdispls", "elementsToSend"):
            check(np.array_equal(d[f], o[f]), "%s: list %s differs" % (tag, f), failures)
        check(np.array_equal(api.gmatrix_arrays(g)[1], omats[rank].col), "%s: renumbered columns differ" % tag, failures)
        # halo exchange through the drop-in entry point (comm.c:627-651): x = global row id
        nc = n + comm.externalCount
        xs = np.zeros(nc)
        xs[:n] = rank * n + np.arange(n)
        xd = api.to_device(xs)
        for _ in range(3):                                   # slot parity + acknowledge path
            L.commExchange(C.byref(comm), n, xd.ptr)
        got = api.to_host(xd, np.float64, nc)
        check(np.array_equal(got[n:], np.array(part.ranks[rank]["externalsReordered"], np.float64)), "%s: halo values differ" % tag, failures)
        mats = []
        for fmt, sigma in fmts:
            tag = "%dx%dx%d fmt=%s sigma=%d" % (nx, ny, nz, api.FMT_NAMES[fmt], sigma)
            A = api.convertMatrix(fmt, g, 32, sigma)
            mats.append(A)
            for flags in (api.CG_FUSED, api.CG_FUSED, api.CG_FUSED | api.CG_NO_OVERLAP, 0, api.CG_FUSED):
                k, hist, x, info = api.solveCG(A, itermax, eps, comm=comm, flags=flags, want_x=True)
                check(k == kref, "%s flags=%d: k %d != %d" % (tag, flags, k, kref), failures)
                if len(hist) == len(href):
                    scale = np.maximum(href, 1e-10 * href[0])
                    err = float(np.max(np.abs(hist - href) / scale))
                    check(err <= CG_TOL, "%s flags=%d: history error %.3e" % (tag, flags, err), failures)
                else:
                    check(False, "%s flags=%d: history length %d != %d" % (tag, flags, len(hist), len(href)), failures)
                xe = float(np.max(np.abs(x - xref[rank * n:(rank + 1) * n])))
                check(xe <= 1e-9 * max(1.0, float(np.max(np.abs(xref)))), "%s flags=%d: solution error %.3e" % (tag, flags, xe), failures)
        # interleaved: one fused solve per matrix again, now that every permutation has been seen once
        for A, (fmt, sigma) in zip(mats, fmts):
            k, hist, _, _ = api.solveCG(A, itermax, eps, comm=comm, flags=api.CG_FUSED)
            ok = k == kref and len(hist) == len(href) and float(np.max(np.abs(hist - href) / np.maximum(href, 1e-10 * href[0]))) <= CG_TOL
            check(ok, "%dx%dx%d fmt=%s sigma=%d: interleaved re-solve k=%d/%d" % (nx, ny, nz, api.FMT_NAMES[fmt], sigma, k, kref), failures)
        if (nx, ny, nz) == (16, 16, 6):
            # GMRES(m) and the Chebyshev filter on the same partition (no reference behaviour: vs the numpy restatement
            # of the same algorithms on the global problem, oracle/krylov_ref.py)
            from oracle import krylov_ref as kr
            kg, hg, xg = kr.gmres(mg, b, x0, 50, 1e-8, 8)
            xin = 1.0 + 0.01 * (np.arange(mg.nr) % 13)
            yc, muc = kr.chebyshev(mg, xin, 9, 0.0, 54.0)
            for A, (fmt, sigma) in zip(mats, fmts):
                tag = "%dx%dx%d fmt=%s sigma=%d" % (nx, ny, nz, api.FMT_NAMES[fmt], sigma)
                k, hist, x, _ = api.solveGMRES(A, 50, 1e-8, restart=8, comm=comm, want_x=True)
                # 64 ulp of the initial residual: the absolute round-off floor of the b - A x every restart recomputes
                gerr = float(np.max((np.abs(hist - hg) - 64 * np.finfo(np.float64).eps * hg[0]) / np.maximum(hg, 1e-10 * hg[0]))) if len(hist) == len(hg) else np.inf
                check(k == kg and gerr <= 1e-7, "%s: GMRES k=%d/%d, history error %.3e" % (tag, k, kg, gerr), failures)
                check(float(np.max(np.abs(x - xg[rank * n:(rank + 1) * n]))) <= 1e-8, "%s: GMRES solution" % tag, failures)
                y, mu = api.chebyshevFilter(A, xin[rank * n:(rank + 1) * n], 9, 0.0, 54.0, comm=comm)
                check(float(np.max(np.abs(y - yc[rank * n:(rank + 1) * n]))) <= 1e-11 * float(np.max(np.abs(yc))), "%s: Chebyshev filter" % tag, failures)
                check(float(np.max(np.abs(mu - muc))) <= 1e-11 * float(np.max(np.abs(muc))), "%s: Chebyshev moments" % tag, failures)
        for A in mats:
            api.destroyMatrix(A)
        L.sbFreeGMatrix(C.byref(g))
    # MatrixMarket path over the ranks (main.c:64-71, comm.c:311-402): rank 0 reads, row blocks are scattered
    nxm, nym, nzm = 7, 6, 3 * world + 1                      # row count not divisible by the rank count
    mg = orc.generate(nxm, nym, nzm)
    path = "/tmp/sb_mgpu_check_%d.mtx" % os.getppid()
    if rank == 0:
        with open(path, "w") as f:
            f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (mg.nr, mg.nr, mg.nnz))
            rows = np.repeat(np.arange(mg.nr), np.diff(mg.rowPtr.astype(np.int64)))
            order = np.random.default_rng(1).permutation(mg.nnz)    # unsorted file
            for j in order:
                f.write("%d %d %r\n" % (rows[j] + 1, mg.col[j] + 1, float(mg.val[j])))
    g = api.matrixRead(path, comm)
    per = [mg.nr // world + (1 if mg.nr % world > r else 0) for r in range(world)]      # sizeOfRank, comm.c:35-38
    lo = sum(per[:rank])
    hi = lo + per[rank]
    check((g.nr, g.startRow, g.stopRow, g.totalNr) == (hi - lo, lo, hi - 1, mg.nr), "mm: row block %r" % ((g.nr, g.startRow, g.stopRow),), failures)
    rp, col, val = api.gmatrix_arrays(g)
    check(np.array_equal(rp, mg.rowPtr[lo:hi + 1] - mg.rowPtr[lo]), "mm: rowPtr slice differs", failures)
    check(np.array_equal(col, mg.col[mg.rowPtr[lo]:mg.rowPtr[hi]]), "mm: columns differ", failures)
    check(np.array_equal(val, mg.val[mg.rowPtr[lo]:mg.rowPtr[hi]]), "mm: values differ", failures)
    L.commPartition(C.byref(comm), C.byref(g))
    kref, href, xref = orc.cg_crs(mg, np.ones(mg.nr), np.zeros(mg.nr), 40, 1e-8)      # file input: b = 1 (CGSolver.c:34-36)
    for fmt, sigma in fmts:
        A = api.convertMatrix(fmt, g, 32, sigma)
        k, hist, x, _ = api.solveCG(A, 40, 1e-8, comm=comm, generated=False, want_x=True)
        ok = k == kref and len(hist) == len(href) and float(np.max(np.abs(hist - href) / np.maximum(href, 1e-10 * href[0]))) <= CG_TOL
        check(ok, "mm: CG fmt=%s sigma=%d k=%d/%d" % (api.FMT_NAMES[fmt], sigma, k, kref), failures)
        check(float(np.max(np.abs(x - xref[lo:hi]))) <= 1e-9, "mm: solution fmt=%s" % api.FMT_NAMES[fmt], failures)
        api.destroyMatrix(A)
    if rank == 0:
        os.unlink(path)

    # the reference's OWN multi-rank distribution of the klein file (unmodified comm.c:311-402 on shim ranks,
    # tests/golden/ref_files.npz): this rank's block must be identical
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_files.npz"))
    key = "mm_matrix_band_klein_P%d_r%d_" % (world, rank)
    if key + "scalars" in gold:
        gk = api.matrixRead(os.path.join(ROOT, "tests", "golden", "reference_fixtures", "matrix_band_klein.mtx"), comm)
        check([gk.nr, gk.nc, gk.nnz, gk.totalNr, gk.totalNnz, gk.startRow, gk.stopRow] == list(gold[key + "scalars"]),
              "klein distribution: header differs from the reference's", failures)
        rpk, colk, valk = api.gmatrix_arrays(gk)
        check(np.array_equal(rpk, gold[key + "rowPtr"]) and np.array_equal(colk, gold[key + "cols"]) and np.array_equal(valk, gold[key + "vals"]),
              "klein distribution: arrays differ from the reference's", failures)

    # an irregular SPD matrix (random long-range couplings, 1..40 entries per row): every rank neighbours every other
    # one, halo lists are not contiguous planes, the SELL sort really permutes rows, row lengths vary
    rng = np.random.default_rng(7)
    nI = 1500 + 37 * world
    rowsets = [set([i]) for i in range(nI)]
    for i in range(nI):
        for j in rng.integers(0, nI, int(rng.integers(0, 20))):
            rowsets[i].add(int(j)); rowsets[int(j)].add(i)
        if i + 1 < nI:
            rowsets[i].add(i + 1); rowsets[i + 1].add(i)
    rp = np.zeros(nI + 1, np.uint32)
    cols, vals = [], []
    for i in range(nI):
        cs = sorted(rowsets[i])
        for c2 in cs:
            cols.append(c2)
            vals.append(float(len(rowsets[i]) + len(rowsets[c2])) if c2 == i else -1.0 / (1 + ((i + c2) % 3)))
        rp[i + 1] = len(cols)
    mi = orc.Csr(rp, np.array(cols, np.uint32), np.array(vals))
    pathI = "/tmp/sb_mgpu_irregular_%d.mtx" % os.getppid()
    if rank == 0:
        with open(pathI, "w") as f:
            f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (nI, nI, mi.nnz))
            rowsI = np.repeat(np.arange(nI), np.diff(mi.rowPtr.astype(np.int64)))
            for j in range(mi.nnz):
                f.write("%d %d %r\n" % (rowsI[j] + 1, mi.col[j] + 1, float(mi.val[j])))
    gI = api.matrixRead(pathI, comm)
    loI = int(gI.startRow)
    L.commPartition(C.byref(comm), C.byref(gI))
    krefI, hrefI, xrefI = orc.cg_crs(mi, np.ones(nI), np.zeros(nI), 60, 1e-9)
    for fmt, sigma in fmts + [(api.FMT_SCS, 64)]:
        A = api.convertMatrix(fmt, gI, 32, sigma)
        for flags in (api.CG_FUSED, api.CG_FUSED | api.CG_NO_OVERLAP):
            k, hist, x, _ = api.solveCG(A, 60, 1e-9, comm=comm, generated=False, flags=flags, want_x=True)
            ok = k == krefI and len(hist) == len(hrefI) and float(np.max(np.abs(hist - hrefI) / np.maximum(hrefI, 1e-10 * hrefI[0]))) <= CG_TOL
            check(ok, "irregular: CG fmt=%s sigma=%d flags=%d k=%d/%d" % (api.FMT_NAMES[fmt], sigma, flags, k, krefI), failures)
            check(float(np.max(np.abs(x - xrefI[loI:loI + gI.nr]))) <= 1e-9, "irregular: solution fmt=%s sigma=%d flags=%d" % (api.FMT_NAMES[fmt], sigma, flags), failures)
        api.destroyMatrix(A)
    if rank == 0:
        os.unlink(pathI)

    # global reductions (comm.c:653-662)
    v = C.c_double(float(rank + 1))
    L.commReduction(C.byref(v), api.OP_SUM)
    check(v.value == world * (world + 1) / 2.0, "commReduction SUM %r" % v.value, failures)
    v = C.c_double(float(rank + 1))
    L.commReduction(C.byref(v), api.OP_MAX)
    check(v.value == float(world), "commReduction MAX %r" % v.value, failures)
    L.commFinalize(C.byref(comm))
    for f in failures[:20]:
        print("[rank %d] FAIL %s" % (rank, f), flush=True)
    print("[rank %d] mgpu_check: %s (%d failures, mode %s)" % (rank, "PASS" if not failures else "FAIL", len(failures),
                                                              os.environ.get("SB_COMM", "default")), flush=True)
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
