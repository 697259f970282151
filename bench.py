#!/usr/bin/env python
"""bench.py -- CG throughput (GFLOP/s over all GPUs, with CG iterations/s, SpMV GFLOP/s and HBM GB/s against the
roofline beside it) of the SparseBench hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sell256|crs128|ccrs128|...] [--impl reference]

A step is ONE CG iteration (CGSolver.c:107-129: p-update, halo exchange, SpMV, two dot products, x/r update)
over the synthetic HPCG 27-point stencil matrix named by the workload; the default workload is BASELINE.json
configs[2]/[3]: 256^3 rows per GPU, SELL-C-sigma (C=32, sigma=256), fp64, z-stacked row blocks over N GPUs (weak
scaling). `value` is the whole-job aggregate: floating-point operations of one CG iteration summed over all ranks
(2 nnz + 10 N per rank; the reference's own accounting, profiler.c:19-22) times iterations per second,
so it grows with N under weak scaling while `cg.iterations_per_sec` stays flat. One JSON line is printed by rank 0.
See DESIGN.md section "Measurement" for every field.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nx, ny, nz per GPU, format, C, sigma, description)
    "sell256": (256, 256, 256, "SCS", 32, 256, "HPCG 27-pt stencil 256^3 per GPU, SELL-C-sigma (C=32, sigma=256), fp64 CG"),
    "crs256": (256, 256, 256, "CRS", 0, 0, "HPCG 27-pt stencil 256^3 per GPU, CRS, fp64 CG"),
    "ccrs256": (256, 256, 256, "CCRS", 0, 0, "HPCG 27-pt stencil 256^3 per GPU, CCRS, fp64 CG"),
    "crs128": (128, 128, 128, "CRS", 0, 0, "HPCG 27-pt stencil 128^3 per GPU, CRS, fp64 CG"),
    "sell128": (128, 128, 128, "SCS", 32, 256, "HPCG 27-pt stencil 128^3 per GPU, SELL-C-sigma (C=32, sigma=256), fp64 CG"),
    "ccrs128": (128, 128, 128, "CCRS", 0, 0, "HPCG 27-pt stencil 128^3 per GPU, CCRS, fp64 CG"),
    "sell64": (64, 64, 64, "SCS", 32, 256, "HPCG 27-pt stencil 64^3 per GPU, SELL-C-sigma (C=32, sigma=256), fp64 CG (smoke size)"),
    # BASELINE.json configs[4]: 512^3 GLOBAL problem split over the ranks (strong scaling), three formats
    "strong512sell": (512, 512, -512, "SCS", 32, 256, "HPCG 27-pt stencil 512^3 global (strong scaling), SELL-C-sigma (C=32, sigma=256), fp64 CG"),
    "strong512crs": (512, 512, -512, "CRS", 0, 0, "HPCG 27-pt stencil 512^3 global (strong scaling), CRS, fp64 CG"),
    "strong512ccrs": (512, 512, -512, "CCRS", 0, 0, "HPCG 27-pt stencil 512^3 global (strong scaling), CCRS, fp64 CG"),
    "strong256sell": (256, 256, -256, "SCS", 32, 256, "HPCG 27-pt stencil 256^3 global (strong scaling), SELL-C-sigma (C=32, sigma=256), fp64 CG"),
}


def stencil_nnz(nx, ny, nz_total):
    return (3 * nx - 2) * (3 * ny - 2) * (3 * nz_total - 2)


def local_nnz(nx, ny, nz, rank, size):
    """stored non-zeros of rank's z-slab (matrix.c:63-96)"""
    per_line = (3 * nx - 2) * (3 * ny - 2)
    total = 0
    for z in range(nz):
        gz = rank * nz + z
        total += per_line * (1 + (gz > 0) + (gz < nz * size - 1))
    return total


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML DURING the timed region (same counters as the
    nvidia-smi line of B200_PROFILING.md, without a subprocess whose output would be block-buffered)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, device):
        self.device = device
        self.samples, self.bits, self.power = [], 0, []
        self.max_mhz = None
        self._stop = False
        self._thread = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            phys = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.device
            if phys:
                try:
                    idx = int(phys.split(",")[self.device])
                except Exception:
                    idx = self.device
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self._stop:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                try:
                    self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                time.sleep(0.02)
        except Exception as e:
            self.error = repr(e)

    def start(self):
        import threading
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop = True
        if self._thread:
            self._thread.join(timeout=5)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        if self.samples:
            out["sm_mhz"] = float(np.median(self.samples))
            out["reasons"] = sorted(k for k, b in self.REASONS.items() if self.bits & b)
            out["samples"] = len(self.samples)
            if self.power:
                out["power_w_max"] = max(self.power)
        if getattr(self, "error", None):
            out["error"] = self.error
        return out


# ----------------------------------------------------------------------------------------------- reference arm
def reference_arm(args, nx, ny, nz, desc, world):
    """The reference's own CPU implementation (oracle/_ref/libref_CRS_fast.so = its sources with its shipped flags
    -O3 -ffast-math -fopenmp, CRS format: the only format whose reference build works), all host threads, driven
    in the CGSolver.c:94-128 order so that exactly K iterations are timed after W warm-up iterations."""
    from oracle import ref
    kind = "reference"
    # all host cores; torchrun exports OMP_NUM_THREADS=1 to its workers, which is not what this arm measures
    threads = int(os.environ.get("SB_REF_THREADS", "0")) or (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
    if not ref.available("CRS_fast"):
        return None
    L = ref.load("CRS_fast")
    # bounded sample: one rank's block of the workload (nx*ny*nz rows), at most 128 z-planes
    snz = min(nz, args.ref_planes)
    frac = (nx * ny * snz) / float(nx * ny * nz * world)
    g = ref.generate(nx, ny, snz, False, "CRS_fast")
    A = ref.convert_crs(g, "CRS_fast")
    n = A.nr
    rp = np.ctypeslib.as_array(C.cast(A.rowPtr, C.POINTER(C.c_uint32)), (n + 1,))
    lens = np.diff(rp.astype(np.int64))
    b = 27.0 - (lens - 1.0)
    x = np.zeros(n); r = np.zeros(n); p = np.zeros(n); Ap = np.zeros(n)
    P = lambda a: a.ctypes.data
    res = C.c_double(0.0)
    L.waxpby(n, 1.0, P(x), 0.0, P(x), P(p))
    L.spMVM(C.byref(A), P(p), P(Ap))
    L.waxpby(n, 1.0, P(b), -1.0, P(Ap), P(r))
    L.ddot(n, P(r), P(r), C.byref(res))
    rtrans = res.value
    t0 = None
    K, W = args.steps, args.warmup
    for k in range(1, W + K + 1):
        if k == W + 1:
            t0 = time.perf_counter()
        if k == 1:
            L.waxpby(n, 1.0, P(r), 0.0, P(r), P(p))
        else:
            old = rtrans
            L.ddot(n, P(r), P(r), C.byref(res)); rtrans = res.value
            L.waxpby(n, 1.0, P(r), rtrans / old, P(p), P(p))
        L.spMVM(C.byref(A), P(p), P(Ap))
        L.ddot(n, P(p), P(Ap), C.byref(res))
        alpha = rtrans / res.value
        L.waxpby(n, 1.0, P(x), alpha, P(p), P(x))
        L.waxpby(n, 1.0, P(r), -alpha, P(Ap), P(r))
    dt = time.perf_counter() - t0
    sample_its = K / dt
    value = sample_its * frac          # a full step covers 1/frac times the sampled rows
    sample = ("%dx%dx%d rows of the %dx%dx%d global problem (%.4g of one step's rows), CRS, %d OpenMP threads, "
              "%d iterations after %d warm-up; full-step rate = sample rate x %.4g"
              % (nx, ny, snz, nx, ny, nz * world, frac, threads, K, W, frac))
    return dict(value=value, sample_its=sample_its, ms_per_step=1e3 / value, cores=threads, kind=kind, sample=sample,
                final_residual=float(np.sqrt(rtrans)))


# ----------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sell256", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-planes", type=int, default=128, help="z-planes of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE line, the JSON result: native libraries (NCCL's version banner, the reference's
    # printf) write to file descriptor 1 directly, so fd 1 is pointed at stderr for the run and the line goes to a
    # private duplicate of the original stdout
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(result_fd, (json.dumps(obj) + "\n").encode())
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nx, ny, nz, fmt, Cc, sigma, desc = WORKLOADS[args.workload]
    scaling = "weak"
    if nz < 0:                       # negative: global z extent, divided over the ranks
        scaling = "strong"
        if (-nz) % world:
            raise SystemExit("workload %s needs a rank count that divides %d" % (args.workload, -nz))
        nz = (-nz) // world
    K, W = args.steps, args.warmup
    metric, unit = "cg_gflops", "GFLOP/s"
    # flops of one CG iteration over the whole job (all ranks): 2 per stored non-zero + 10 per row
    flops_it_job = sum(2 * local_nnz(nx, ny, nz, r, world) + 10 * nx * ny * nz for r in range(world))

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = reference_arm(args, nx, ny, nz, desc, world)
        if r is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref/libref_CRS_fast.so was not built"})
            return 0
        gf = r["value"] * flops_it_job / 1e9
        line = {"impl": "reference", "metric": metric, "value": gf, "unit": unit, "n_gpus": args.gpus, "steps": K,
                "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "nx": nx, "ny": ny, "nz_per_gpu": nz, "format": "CRS (reference CPU build)"},
                "cpu_baseline": {"value": gf, "unit": unit, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                 "iterations_per_sec": r["value"]},
                "e2e": {"value": gf, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "cg": {"iterations_per_sec": r["value"]},
                "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    from sparsebench_b200 import api
    L = api.lib()
    if L.sbDeviceCount() < 1:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    ndev = L.sbDeviceCount()
    dev = local % ndev
    torch.cuda.set_device(dev)
    L.sbSetDevice(dev)
    comm = api.Comm()
    comm.rank, comm.size = 0, 1
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        nbytes = L.sbCommUniqueIdBytes()
        idbuf = (C.c_char * nbytes)()
        if rank == 0:
            L.sbCommGetUniqueId(idbuf)
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().numpy().tobytes())
        L.sbCommInitRank(C.byref(comm), rank, world, dev, raw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        L.sbDeviceSynchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def note(msg):
        if os.environ.get("SB_BENCH_VERBOSE"):
            sys.stderr.write("[bench r%d] %s\n" % (rank, msg))
            sys.stderr.flush()

    # ---- setup (untimed): generate on the device, partition, convert
    t_setup = time.perf_counter()
    g = api.matrixGenerate(nx, ny, nz, rank, world, device=True)
    note('generated')
    L.commPartition(C.byref(comm), C.byref(g))
    fmt_id = {"CRS": api.FMT_CRS, "SCS": api.FMT_SCS, "CCRS": api.FMT_CCRS}[fmt]
    A = api.convertMatrix(fmt_id, g, Cc, sigma)
    note('converted')
    if fmt != "CCRS":
        L.sbFreeGMatrix(C.byref(g))
    barrier()
    t_setup = time.perf_counter() - t_setup
    N = nx * ny * nz
    nnz = local_nnz(nx, ny, nz, rank, world)
    nc = N + comm.externalCount
    val_bytes = 16 if fmt == "CCRS" else 12
    B_spmv = 12 * nnz + 8 * nc + 8 * N            # algorithmic bytes, true nnz
    B_spmv_fmt = val_bytes * nnz + 8 * nc + 8 * N
    F_spmv = 2 * nnz
    B_it = B_spmv + 72 * N
    F_it = 2 * nnz + 10 * N

    def new_solver(flags, itermax, b=None, x=None):
        p = api.Parameter(b"generate", nx, ny, nz, itermax, 0.0)
        info = api.CGInfo()
        info.flags = flags
        hist = np.zeros(itermax + 4)
        info.history = hist.ctypes.data_as(C.POINTER(C.c_double))
        info.historyCap = len(hist)
        if b is not None:
            info.b = b
        if x is not None:
            info.x = x
        S = L.sbCGCreate(C.byref(comm), C.byref(p), C.byref(A), fmt_id, C.byref(info))
        return S, info, hist, p

    timer = api.EventTimer()
    sampler = ClockSampler(dev)

    # ---- timed region: W warm-up iterations, then exactly K iterations, inputs resident in HBM
    S, info, hist, _p = new_solver(api.CG_FUSED, W + K + 1)
    note('solver created')
    L.sbCGIterate(S, W + 1)
    barrier()
    sampler.start()
    launches0 = L.sbKernelLaunchCount()
    timer.start()
    kdone = L.sbCGIterate(S, W + K + 1)
    ms = timer.stop_ms()
    launches = L.sbKernelLaunchCount() - launches0
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(ms)
    L.sbCGFinish(S, C.byref(info), ms)
    assert kdone == W + K + 1, "CG stopped early: k=%d" % kdone
    resid0, resid = float(hist[0]), float(hist[info.nhist - 1])
    its = K / (ms * 1e-3)
    value = its * flops_it_job / 1e9
    note('timed region done: %.3f ms/it' % (ms / K))

    # ---- per-kernel device times inside the same loop (CUDA events on the launching stream)
    S2, info2, _h2, _p2 = new_solver(api.CG_FUSED | api.CG_PROFILE, W + K + 1)
    L.sbCGIterate(S2, W + K + 1)
    L.sbCGFinish(S2, C.byref(info2), 0.0)
    region = {api.REGIONS[i]: info2.regionMs[i] / (W + K) for i in range(len(api.REGIONS))}
    spmv_ms = region["spmv"] + region["spmv_boundary"]     # interior + boundary launches (the halo wait is not SpMV work)
    note('profile pass done')
    peak, peak_src = peaks()
    achieved = B_spmv / (spmv_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except Exception:
            traffic = None

    # ---- the reference's `-t spmv` mode (main.c:200-216): x = 1, back-to-back SpMVs, no fused dot
    xs = api.to_device(np.ones(8), slots=nc + 64)
    ones = np.ones(N)
    L.sbCopyToDevice(xs.ptr, ones.ctypes.data, 8 * N)
    del ones
    ys = api.DeviceBuffer(8 * (N + 64 + 32))
    for _ in range(3):
        api.spMVM(A, xs, ys)
    barrier()
    timer.start()
    for _ in range(K):
        api.spMVM(A, xs, ys)
    spmv_only_ms = max_over_ranks(timer.stop_ms()) / K
    xs.free(); ys.free()
    note('spmv mode done')

    # ---- e2e: the same iterations through sbSolveCG with HOST (pinned) b and x: upload, solve, download
    e2e = None
    if not args.no_e2e:
        hb = L.sbAllocateHost(8 * N)
        hx = L.sbAllocateHost(8 * N)
        b_np = np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_double)), (N,))
        x_np = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_double)), (N,))
        lens = np.full(N, 27.0)
        b_np[:] = 1.0   # any right-hand side; values do not change the work per iteration
        x_np[:] = 0.0
        del lens
        p = api.Parameter(b"generate", nx, ny, nz, K + 1, 0.0)
        einfo = api.CGInfo()
        einfo.flags = api.CG_FUSED | api.CG_HOST_VECTORS
        ehist = np.zeros(K + 8)
        einfo.history = ehist.ctypes.data_as(C.POINTER(C.c_double))
        einfo.historyCap = len(ehist)
        einfo.b, einfo.x = hb, hx
        barrier()
        t0 = time.perf_counter()
        ke = L.sbSolveCG(C.byref(comm), C.byref(p), C.byref(A), fmt_id, C.byref(einfo))
        L.sbDeviceSynchronize()
        dt = time.perf_counter() - t0
        dt = max_over_ranks(dt)
        e2e = {"value": (ke - 1) / dt * flops_it_job / 1e9, "unit": unit, "iterations_per_sec": (ke - 1) / dt,
               "h2d_bytes_per_step": 2 * 8 * N / (ke - 1),
               "d2h_bytes_per_step": (8 * N + 8 * ke) / (ke - 1),
               "what": "sbSolveCG(host b, host x0 -> host x): H2D of b and x0, %d iterations with a D2H residual scalar "
                       "each, D2H of x; matrix resident (convertMatrix is setup, as in the reference)" % (ke - 1),
               "ms_per_step": dt * 1e3 / (ke - 1)}
        L.sbFreeHost(hb); L.sbFreeHost(hx)
        note('e2e done')

    # ---- BASELINE.json configs[1] beside the headline (single GPU, default workload only): 128^3 CRS CG, same method
    also = None
    if world == 1 and args.workload == "sell256" and not args.no_e2e:
        try:
            wl = WORKLOADS["crs128"]
            ax, ay, az = wl[0], wl[1], wl[2]
            g1 = api.matrixGenerate(ax, ay, az, 0, 1, device=True)
            A1 = api.convertMatrix(api.FMT_CRS, g1)
            L.sbFreeGMatrix(C.byref(g1))
            p1 = api.Parameter(b"generate", ax, ay, az, W + K + 1, 0.0)
            i1 = api.CGInfo()
            i1.flags = api.CG_FUSED
            h1 = np.zeros(W + K + 5)
            i1.history = h1.ctypes.data_as(C.POINTER(C.c_double))
            i1.historyCap = len(h1)
            one = api.Comm()
            one.rank, one.size = 0, 1
            S1 = L.sbCGCreate(C.byref(one), C.byref(p1), C.byref(A1), api.FMT_CRS, C.byref(i1))
            L.sbCGIterate(S1, W + 1)
            L.sbDeviceSynchronize()
            timer.start()
            k1 = L.sbCGIterate(S1, W + K + 1)
            ms1 = timer.stop_ms()
            L.sbCGFinish(S1, C.byref(i1), ms1)
            api.destroyMatrix(A1)
            N1, nnz1 = ax * ay * az, local_nnz(ax, ay, az, 0, 1)
            done = k1 - (W + 1)
            b_it1 = 12 * nnz1 + 16 * N1 + 72 * N1
            also = {"workload": wl[6], "steps": done, "ms_per_step": ms1 / done, "iterations_per_sec": done / (ms1 * 1e-3),
                    "gflops": (2 * nnz1 + 10 * N1) * done / (ms1 * 1e-3) / 1e9, "cg_gbs": b_it1 * done / (ms1 * 1e-3) / 1e9,
                    "frac_of_peak": b_it1 * done / (ms1 * 1e-3) / 1e9 / peak}
        except Exception as e:
            also = {"workload": "crs128", "failed": repr(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sub = argparse.Namespace(**vars(args))
            sub.steps, sub.warmup = min(K, 20), 3
            r = reference_arm(sub, nx, ny, nz, desc, 1)
            if r:
                cpu = {"value": r["value"] * flops_it_job / 1e9, "unit": unit, "cores": r["cores"], "kind": r["kind"],
                       "sample": r["sample"], "iterations_per_sec": r["value"]}
        except Exception as e:  # the CPU leg must never take the GPU numbers down with it
            cpu = {"value": None, "unit": unit, "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}

    if rank == 0:
        line = {
            "metric": metric, "value": value * 1.0, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "nx": nx, "ny": ny, "nz_per_gpu": nz, "format": fmt, "C": Cc, "sigma": sigma,
                       "rows_per_gpu": N, "nnz_rank0": nnz, "parallelism": "row-block x%d" % world,
                       "l2": "inputs larger than L2: %.2f GB of matrix + vectors streamed per iteration vs 126 MB L2" % (B_it / 1e9),
                       "setup_s": round(t_setup, 2)},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "spmv (%s) fused with p.Ap" % fmt, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": B_spmv, "avg_launch_ms": spmv_ms,
                         "frac_of_nominal_8000": achieved / 8000.0},
            "cpu_baseline": cpu,
            "configs1": also,
            "spmv": {"gflops": F_spmv / (spmv_only_ms * 1e-3) / 1e9, "gbs": B_spmv / (spmv_only_ms * 1e-3) / 1e9,
                     "gbs_format_bytes": B_spmv_fmt / (spmv_only_ms * 1e-3) / 1e9, "ms": spmv_only_ms,
                     "frac_of_peak": B_spmv / (spmv_only_ms * 1e-3) / 1e9 / peak, "mode": "x=1, back-to-back (main.c:200-216)"},
            "cg": {"iterations_per_sec": its, "gbs_per_gpu": B_it / (ms / K * 1e-3) / 1e9, "gflops_per_gpu": F_it / (ms / K * 1e-3) / 1e9,
                   "frac_of_peak": B_it / (ms / K * 1e-3) / 1e9 / peak, "kernel_ms_per_iteration": region,
                   "residual_initial": resid0, "residual_final": resid, "max_error_vs_xexact": info.maxError},
        }
        emit(line)
    if world > 1:
        L.commFinalize(C.byref(comm))
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
