"""Small driver for profiling: builds one stencil matrix on the device and launches the SpMV kernel of one format a
few times through the C ABI (sb<FMT>_spMVM), printing CUDA-event times. Used under ncu (profiles/README.md).

    python tools/spmv_probe.py --n 256 --fmt SCS --reps 5 [--cg 3]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsebench_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--nz", type=int, default=0)
    ap.add_argument("--fmt", default="SCS", choices=["CRS", "SCS", "CCRS"])
    ap.add_argument("--sigma", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--dirty", action="store_true")
    ap.add_argument("--ordered", action="store_true")
    ap.add_argument("--dot", action="store_true", help="also time the fused SpMV + dot kernel (sbSpmvDot)")
    ap.add_argument("--cg", type=int, default=0, help="also run this many fused CG iterations")
    ap.add_argument("--gmres", type=int, default=0, help="also run this many GMRES(30) steps")
    ap.add_argument("--cheb", type=int, default=0, help="also apply a Chebyshev filter of this degree")
    a = ap.parse_args()
    L = api.lib()
    n, nz = a.n, (a.nz or a.n)
    g = api.matrixGenerate(n, n, nz, device=True)
    fmt = {"CRS": api.FMT_CRS, "SCS": api.FMT_SCS, "CCRS": api.FMT_CCRS}[a.fmt]
    A = api.convertMatrix(fmt, g, 32, a.sigma)
    if a.fmt != "CCRS":
        L.sbFreeGMatrix(C.byref(g))
    N = n * n * nz
    nnz = (3 * n - 2) ** 2 * (3 * nz - 2)
    x = api.to_device(1.0 + 1e-3 * (np.arange(N) % 1000))
    y = api.DeviceBuffer(8 * (N + 64))
    t = api.EventTimer()
    B = 12 * nnz + 16 * N
    times = []
    for i in range(a.reps):
        t.start()
        api.spMVM(A, x, y)
        times.append(t.stop_ms())
    # back-to-back block (what a solver loop sees): one event pair around `reps` launches
    t.start()
    for i in range(a.reps):
        api.spMVM(A, x, y)
    blk = t.stop_ms() / a.reps
    ms, med = min(times), sorted(times)[len(times) // 2]
    print("spmv %s %d^2x%d: min %.4f ms (%.1f GB/s, %.1f GFLOP/s)  median %.4f  back-to-back %.4f ms (%.1f GB/s)  first %.4f"
          % (a.fmt, n, nz, ms, B / ms / 1e6, 2 * nnz / ms / 1e6, med, blk, B / blk / 1e6, times[0]))
    if a.dot:
        # the fused kernel of the CG loop: y = A x with x . y in the epilogue
        d = api.DeviceBuffer(64)
        times = []
        for i in range(a.reps):
            t.start()
            L.sbSpmvDot(C.byref(A), fmt, x.ptr, y.ptr, d.ptr)
            times.append(t.stop_ms())
        # interleaved blocks (plain, fused, plain, fused, ...) so that clock / temperature drift hits both alike
        plain_b, fused_b = [], []
        for rnd in range(6):
            t.start()
            for i in range(a.reps):
                api.spMVM(A, x, y)
            plain_b.append(t.stop_ms() / a.reps)
            t.start()
            for i in range(a.reps):
                L.sbSpmvDot(C.byref(A), fmt, x.ptr, y.ptr, d.ptr)
            fused_b.append(t.stop_ms() / a.reps)
        pb, fb = sorted(plain_b)[len(plain_b) // 2], sorted(fused_b)[len(fused_b) // 2]
        print("spmv+dot %s: isolated min %.4f median %.4f | interleaved back-to-back blocks: plain %.4f ms, fused %.4f ms (%.1f GB/s)"
              "  -> fused-dot cost %+.2f %%" % (a.fmt, min(times), sorted(times)[len(times) // 2], pb, fb, B / fb / 1e6, 100.0 * (fb / pb - 1.0)))
    if a.ordered:
        units = A.nChunks if a.fmt == "SCS" else A.nr
        plane = (n * n) // (32 if a.fmt == "SCS" else 1)
        for (lo, hi) in ((0, units), (plane, units - plane)):
            times = []
            for i in range(a.reps):
                t.start()
                L.sbSpmvOrdered(C.byref(A), fmt, x.ptr, y.ptr, lo, hi)
                times.append(t.stop_ms())
            print("ordered single-launch kernel, interior [%d,%d): min %.4f median %.4f ms" % (lo, hi, min(times), sorted(times)[len(times) // 2]))
        # interleaved back-to-back blocks: what the gated kernel variant itself costs (no gate to wait for, ordinary memory)
        pb, ob = [], []
        for rnd in range(6):
            t.start()
            for i in range(a.reps):
                api.spMVM(A, x, y)
            pb.append(t.stop_ms() / a.reps)
            t.start()
            for i in range(a.reps):
                L.sbSpmvOrdered(C.byref(A), fmt, x.ptr, y.ptr, plane, units - plane)
            ob.append(t.stop_ms() / a.reps)
        pm, om = sorted(pb)[len(pb) // 2], sorted(ob)[len(ob) // 2]
        print("interleaved back-to-back blocks: plain %.4f ms, ordered/gated variant %.4f ms (%+.1f us)" % (pm, om, (om - pm) * 1e3))
    if a.dirty:
        # the CG context: a vector update that leaves x dirty in L2 right before every SpMV
        r = api.to_device(np.zeros(N))
        times = []
        for i in range(a.reps):
            api.waxpby(N, 1.0, r, 0.5, x, x)
            t.start()
            api.spMVM(A, x, y)
            times.append(t.stop_ms())
        print("spmv after waxpby(x): min %.4f median %.4f" % (min(times), sorted(times)[len(times) // 2]))
        times = []
        for i in range(a.reps):
            api.waxpby(N, 1.0, r, 0.5, r, r)
            t.start()
            api.spMVM(A, x, y)
            times.append(t.stop_ms())
        print("spmv after waxpby(r): min %.4f median %.4f" % (min(times), sorted(times)[len(times) // 2]))
    if a.cg:
        k, hist, _, info = api.solveCG(A, a.cg + 1, 0.0)
        k, hist, _, info = api.solveCG(A, a.cg + 1, 0.0)
        print("cg k=%d residual %.6e  loop %.4f ms/it" % (k, hist[-1], info.solveMs / a.cg))
        k, hist, _, info = api.solveCG(A, a.cg + 1, 0.0, flags=api.CG_FUSED | api.CG_PROFILE)
        print("cg regions ms/it:", {r: round(info.regionMs[i] / a.cg, 4) for i, r in enumerate(api.REGIONS)})
    if a.gmres:
        api.solveGMRES(A, 8, 0.0, restart=30)
        k, hist, _, info = api.solveGMRES(A, a.gmres + 1, 0.0, restart=30)
        print("gmres(30) k=%d residual %.6e  %.4f ms per step (SpMV + 2 Gram-Schmidt passes over on average %d basis vectors + host Givens)"
              % (k, hist[-1], info.solveMs / max(k, 1), min(k, 30) // 2))
    if a.cheb:
        import time
        xh = np.ones(N)
        api.chebyshevFilter(A, xh, 4, 0.0, 54.0)

        def timed(deg):
            L.sbDeviceSynchronize()
            t0 = time.perf_counter()
            _, mu = api.chebyshevFilter(A, xh, deg, 0.0, 54.0, want_y=False)
            return time.perf_counter() - t0, mu
        lo = max(2, a.cheb // 5)
        (t_lo, _), (t_hi, mu) = timed(lo), timed(a.cheb)
        print("chebyshev moments: %.4f ms per degree (degree %d minus degree %d: the upload of x cancels; SpMV + one fused vector pass), mu[-1] = %.6e"
              % ((t_hi - t_lo) * 1e3 / (a.cheb - lo), a.cheb, lo, mu[-1]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
