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
import re
import sys
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


def local_nnz(nx, ny, nz, rank, size):
    """stored non-zeros of rank's z-slab (matrix.c:63-96)"""
    per_line = (3 * nx - 2) * (3 * ny - 2)
    gz = rank * nz + np.arange(nz)
    return int(per_line * np.sum(1 + (gz > 0) + (gz < nz * size - 1)))


def stencil_rhs(nx, ny, nz, rank, size):
    """b = 27 - (rowLen - 1) of the generated matrix (CGSolver.c:26-33, matrix.c:63-96) without reading the matrix:
    a row has cx*cy*cz entries, c = 3 minus one per domain face the point lies on (z faces of the GLOBAL domain)."""
    def cnt(n, first=True, last=True):
        c = np.full(n, 3.0)
        if first:
            c[0] -= 1
        if last:
            c[-1] -= 1
        return c
    cz = cnt(nz, rank == 0, rank == size - 1)
    lens = cz[:, None, None] * cnt(ny)[None, :, None] * cnt(nx)[None, None, :]
    return (27.0 - (lens - 1.0)).reshape(-1)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def workload_config(desc, nx, ny, nz, world, B_it_rank0):
    """`config` of the JSON line: names the WORKLOAD and nothing arm-specific, so that the product arm and the
    reference arm print the same dict (what each arm ran it with is under `plugin`)."""
    return {"workload": desc, "nx": nx, "ny": ny, "nz_per_gpu": nz, "rows_per_gpu": nx * ny * nz,
            "global_rows": nx * ny * nz * world, "parallelism": "row-block x%d" % world,
            "l2": "inputs larger than L2: %.2f GB of matrix + vectors streamed per iteration and GPU vs 126 MB L2" % (B_it_rank0 / 1e9)}


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML DURING the timed region (same counters as the
    nvidia-smi line of B200_PROFILING.md, without a subprocess whose output would be block-buffered)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, device):
        import threading
        self.device = device
        self.samples, self.bits, self.power = [], 0, []
        self.max_mhz = None
        self._stop = False
        self._record = False
        self._ready = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

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
                mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    bits = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    bits = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self._ready.set()
                if self._record:
                    self.samples.append(mhz)
                    self.bits |= bits
                    if len(self.samples) % 8 == 1:      # every NVML query costs milliseconds: power only now and then
                        try:
                            self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                        except Exception:
                            pass
                time.sleep(0.001 if self._record else 0.02)
        except Exception as e:
            self.error = repr(e)
            self._ready.set()

    def start(self):
        """starts the NVML thread (before the warm-up) and waits until it has produced its first sample"""
        self._thread.start()
        self._ready.wait(timeout=10)

    def begin(self):
        self._record = True

    def end(self):
        self._record = False

    def stop(self):
        self._stop = True
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
def reference_arm(K, W, nx, ny, nz, world, planes=0):
    """The reference's own CPU implementation through its own entry point: oracle/_ref/libref_CRS_fast.so is the
    reference's sources compiled with its shipped flags (-O3 -ffast-math -fopenmp, mk/include_GCC.mk:15; CRS, the
    only format whose reference build works; OpenMP on all host cores, no MPI in the image). matrixGenerate ->
    convertMatrix -> solveCG (CGSolver.c:62-141) is called twice, with itermax = W+1 and W+K+1; the time of exactly
    K iterations after W warm-up iterations is the difference of what the reference's own profiler accumulated in
    _t[] (profiler.h:18-24) over the two calls. N=1: the full problem. N>1: one rank's block on the one host, rate
    scaled by 1/N (labelled in `sample`). `planes` > 0 (tests, small hosts) samples that many z-planes instead."""
    from oracle import ref
    threads = int(os.environ.get("SB_REF_THREADS", "0")) or (os.cpu_count() or 1)
    # all host cores; torchrun exports OMP_NUM_THREADS=1 to its workers, which is not what this arm measures
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
    if not ref.available("CRS_fast"):
        return None
    L = ref.load("CRS_fast")
    snz = min(nz, planes) if planes > 0 else nz
    frac = snz / float(nz * world)
    nz = snz
    t_setup = time.perf_counter()
    g = ref.generate(nx, ny, nz, False, "CRS_fast")
    A = ref.convert_crs(g, "CRS_fast")
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(g.rowPtr); libc.free(g.entries)            # 16 B x 27 N of staging the reference never releases
    t_setup = time.perf_counter() - t_setup
    regions = (C.c_double * 4).in_dll(L, "_t")
    comm = ref.Comm(0, 1, None)
    took = re.compile(r"Solution performed (\d+) iterations and took (\S+)s")
    resid = re.compile(r"Residual = (\S+)")

    def solve(itermax):
        p = ref.Parameter(b"generate", nx, ny, nz, itermax, 0.0)
        before = sum(regions)
        with ref.capture_stdout() as cap:
            k = L.solveCG(C.byref(comm), C.byref(p), C.byref(A))
        m = took.search(cap.text)
        res = resid.findall(cap.text)
        return sum(regions) - before, k, float(m.group(2)) if m else None, float(res[-1]) if res else None

    t_w, k_w, took_w, _ = solve(W + 1)
    t_all, k_all, took_all, final = solve(W + K + 1)
    assert k_w == W + 1 and k_all == W + K + 1, (k_w, k_all)
    dt = t_all - t_w
    if not dt > 0:
        raise RuntimeError("reference arm: the sample is too small to time (%.3g s for %d iterations)" % (dt, K))
    sample_its = K / dt
    full = "full %dx%dx%d problem" % (nx, ny, nz) if frac == 1.0 else \
        "%dx%dx%d rows (%.4g of one step's rows; one rank's block when N>1), full-step rate = sample rate x %.4g" % (
            nx, ny, nz, frac, frac)
    sample = ("%s through the reference's own solveCG (CRS, %d OpenMP threads): solveCG(itermax=%d) minus "
              "solveCG(itermax=%d) = %d iterations after %d warm-up, timed by the reference's profiler regions _t[]; "
              "its own report: %.2f s - %.2f s" % (full, threads, W + K + 1, W + 1, K, W, -1 if took_all is None else took_all, -1 if took_w is None else took_w))
    return dict(value=sample_its * frac, sample_its=sample_its, ms_per_step=1e3 / (sample_its * frac), cores=threads,
                kind="reference", sample=sample, same_config=(frac == 1.0), sampled_fraction=frac, setup_s=t_setup,
                final_residual=final)


# ----------------------------------------------------------------------------------------------- helpers of the main arm
class Problem:
    """one matrix on this rank: generated on the device, partitioned, converted"""

    def __init__(self, api, L, comm, nx, ny, nz, fmt, Cc, sigma, rank, world):
        self.api, self.L, self.comm = api, L, comm
        self.nx, self.ny, self.nz, self.fmt = nx, ny, nz, fmt
        t0 = time.perf_counter()
        g = api.matrixGenerate(nx, ny, nz, rank, world, device=True)
        L.commPartition(C.byref(comm), C.byref(g))
        self.fmt_id = {"CRS": api.FMT_CRS, "SCS": api.FMT_SCS, "CCRS": api.FMT_CCRS}[fmt]
        self.A = api.convertMatrix(self.fmt_id, g, Cc, sigma)
        self.g = g
        if fmt != "CCRS":                       # the CCRS Matrix aliases the GMatrix (matrix-CCRS.c:12)
            L.sbFreeGMatrix(C.byref(g))
        L.sbDeviceSynchronize()
        self.setup_s = time.perf_counter() - t0
        self.N = nx * ny * nz
        self.nnz = local_nnz(nx, ny, nz, rank, world)
        self.nc = self.N + comm.externalCount
        self.B_spmv = 12 * self.nnz + 8 * self.nc + 8 * self.N            # algorithmic bytes, true nnz
        self.B_spmv_fmt = (16 if fmt == "CCRS" else 12) * self.nnz + 8 * self.nc + 8 * self.N
        self.F_spmv = 2 * self.nnz
        self.B_it = self.B_spmv + 72 * self.N
        self.F_it = 2 * self.nnz + 10 * self.N

    def solver(self, flags, itermax, b=None, x=None):
        api = self.api
        p = api.Parameter(b"generate", self.nx, self.ny, self.nz, itermax, 0.0)
        info = api.CGInfo()
        info.flags = flags
        hist = np.zeros(itermax + 4)
        info.history = hist.ctypes.data_as(C.POINTER(C.c_double))
        info.historyCap = len(hist)
        if b is not None:
            info.b = b
        if x is not None:
            info.x = x
        S = self.L.sbCGCreate(C.byref(self.comm), C.byref(p), C.byref(self.A), self.fmt_id, C.byref(info))
        return S, info, hist, p

    def destroy(self):
        self.api.destroyMatrix(self.A)
        if self.fmt == "CCRS":
            self.L.sbFreeGMatrix(C.byref(self.g))


def timed_cg(P, timer, W, K, barrier, max_over_ranks, sampler=None):
    """W warm-up iterations, then exactly K iterations bracketed by barrier + synchronize; CUDA-event time, max over
    ranks. Returns (ms for K iterations, launches, info, history)."""
    api, L = P.api, P.L
    S, info, hist, _p = P.solver(api.CG_FUSED, W + K + 1)
    L.sbCGIterate(S, W + 1)
    barrier()
    if sampler:
        sampler.begin()
    launches0 = L.sbKernelLaunchCount()
    timer.start()
    kdone = L.sbCGIterate(S, W + K + 1)
    ms = timer.stop_ms()
    launches = L.sbKernelLaunchCount() - launches0
    barrier()
    if sampler:
        sampler.end()
    ms = max_over_ranks(ms)
    L.sbCGFinish(S, C.byref(info), ms)
    assert kdone == W + K + 1, "CG stopped early: k=%d" % kdone
    return ms, launches, info, hist


def profiled_cg(P, W, K):
    """the same loop once more with CUDA events around every kernel (on the launching stream) -> ms per iteration"""
    api, L = P.api, P.L
    S, info, _h, _p = P.solver(api.CG_FUSED | api.CG_PROFILE, W + K + 1)
    L.sbCGIterate(S, W + K + 1)
    L.sbCGFinish(S, C.byref(info), 0.0)
    region = {api.REGIONS[i]: info.regionMs[i] / (W + K) for i in range(len(api.REGIONS))}
    return region


# ----------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sell256", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-planes", type=int, default=0, help="z-planes of the CPU sample (0 = the full block; tests use small values)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[1] / configs[4] legs and the multi-GPU parity block")
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
    N0 = nx * ny * nz
    B_it0 = 12 * local_nnz(nx, ny, nz, 0, world) + 16 * N0 + 72 * N0
    config = workload_config(desc, nx, ny, nz, world, B_it0)

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = reference_arm(K, W, nx, ny, nz, world, args.ref_planes)
        if r is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref/libref_CRS_fast.so was not built"})
            return 0
        gf = r["value"] * flops_it_job / 1e9
        line = {"impl": "reference", "metric": metric, "value": gf, "unit": unit, "n_gpus": args.gpus, "steps": K,
                "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "plugin": {"format": "CRS", "why": "the reference's only working format build (SURVEY 0); same matrix, same CG"},
                "cpu_baseline": {"value": gf, "unit": unit, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                 "same_config": r["same_config"], "sampled_fraction": r["sampled_fraction"],
                                 "iterations_per_sec": r["value"], "setup_s": round(r["setup_s"], 2)},
                "e2e": {"value": gf, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "cg": {"iterations_per_sec": r["value"], "residual_final": r["final_residual"]},
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

    peak, peak_src = peaks()
    # second denominator, measured here: a read-only stream (the SpMV is ~98 % reads; the driver's peak is a copy)
    read_peak = L.sbMeasureReadBandwidth(4 << 30, 5)
    timer = api.EventTimer()
    sampler = ClockSampler(dev)
    sampler.start()

    # ---- multi-GPU parity against the CPU oracle (untimed, before anything is measured): a wrong answer must not
    # produce a bench line. The oracle is the checker here, never the thing measured.
    parity = None
    if world > 1 and not args.no_extra:
        from oracle import parity as oracle_parity
        parity = oracle_parity.multi_gpu_parity(api, L, comm, rank, world)
        flag = torch.tensor([0.0 if parity["ok"] else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.SUM)
        parity["ranks_failed"] = int(flag.item())
        parity["ok"] = parity["ranks_failed"] == 0
        note("parity %r" % (parity,))
        if not parity["ok"]:
            sys.stderr.write("[bench r%d] multi-GPU parity FAILED: %r\n" % (rank, parity))
            if rank == 0:
                emit({"metric": metric, "value": None, "unit": unit, "n_gpus": world, "parity": parity,
                      "error": "multi-GPU parity against the oracle failed; nothing was measured"})
            L.commFinalize(C.byref(comm))
            dist.destroy_process_group()
            return 3

    # ---- setup (untimed): generate on the device, partition, convert
    P = Problem(api, L, comm, nx, ny, nz, fmt, Cc, sigma, rank, world)
    barrier()
    N, nnz, nc = P.N, P.nnz, P.nc
    note("setup %.2f s" % P.setup_s)

    # ---- timed region: W warm-up iterations, then exactly K iterations, inputs resident in HBM
    ms, launches, info, hist = timed_cg(P, timer, W, K, barrier, max_over_ranks, sampler)
    clocks = sampler.stop()
    resid0, resid = float(hist[0]), float(hist[info.nhist - 1])
    its = K / (ms * 1e-3)
    value = its * flops_it_job / 1e9
    note('timed region done: %.3f ms/it' % (ms / K))

    # ---- per-kernel device times inside the same loop (CUDA events on the launching stream)
    region = profiled_cg(P, W, K)
    spmv_ms = region["spmv"] + region["spmv_boundary"]     # interior + boundary launches (the halo wait is not SpMV work)
    achieved = P.B_spmv / (spmv_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
            if traffic is not None:
                traffic_src = ("profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed "
                               "`ncu --set full` capture of this kernel (profiles/r2_sell256_spmv_and_vector_kernels.ncu.txt); a constant "
                               "read from file, NOT measured in this run")
        except Exception:
            traffic = None

    # ---- the reference's `-t spmv` mode (main.c:200-216): x = 1, back-to-back SpMVs, no fused dot
    xs = api.to_device(np.ones(8), slots=nc + 64)
    ones = np.ones(N)
    L.sbCopyToDevice(xs.ptr, ones.ctypes.data, 8 * N)
    del ones
    ys = api.DeviceBuffer(8 * (N + 64 + 32))
    for _ in range(3):
        api.spMVM(P.A, xs, ys)
    barrier()
    timer.start()
    for _ in range(K):
        api.spMVM(P.A, xs, ys)
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
        b_np[:] = stencil_rhs(nx, ny, nz, rank, world)     # the generated right-hand side (A 1 = b)
        p = api.Parameter(b"generate", nx, ny, nz, K + 1, 0.0)
        ehist = np.zeros(K + 8)
        best = None
        for rep in range(2):                                # first solve on this partition registers the halo window once
            x_np[:] = 0.0
            einfo = api.CGInfo()
            einfo.flags = api.CG_FUSED | api.CG_HOST_VECTORS
            einfo.history = ehist.ctypes.data_as(C.POINTER(C.c_double))
            einfo.historyCap = len(ehist)
            einfo.b, einfo.x = hb, hx
            barrier()
            t0 = time.perf_counter()
            ke = L.sbSolveCG(C.byref(comm), C.byref(p), C.byref(P.A), P.fmt_id, C.byref(einfo))
            L.sbDeviceSynchronize()
            dt = time.perf_counter() - t0
            dt = max_over_ranks(dt)
            cur = (dt, ke, einfo.createMs, einfo.solveMs, einfo.finishMs)
            if rep == 0:
                first = cur
            best = cur
        dt, ke, create_ms, loop_ms, finish_ms = best
        # the answer of the timed call: same residual history as the device-resident run, x consistent with maxError
        hd = float(np.max(np.abs(ehist[:ke] - hist[:ke]) / np.maximum(hist[:ke], 1e-300)))
        xerr = float(np.max(np.abs(x_np - 1.0)))
        assert hd <= 1e-12, "e2e residual history differs from the device-resident run: %.3e" % hd
        # maxError is the maximum over all ranks (CGSolver.c:55): equal to this rank's on one GPU, an upper bound otherwise
        assert xerr <= einfo.maxError * (1 + 1e-12) + 1e-15 and (world > 1 or abs(xerr - einfo.maxError) <= 1e-15 + 1e-12 * xerr), \
            (xerr, einfo.maxError)
        e2e = {"value": (ke - 1) / dt * flops_it_job / 1e9, "unit": unit, "iterations_per_sec": (ke - 1) / dt,
               "h2d_bytes_per_step": 2 * 8 * N / (ke - 1),
               "d2h_bytes_per_step": (8 * N + 8 * ke) / (ke - 1),
               "what": "sbSolveCG(host b, host x0 -> host x): H2D of b and x0, %d iterations with a D2H residual scalar "
                       "each, D2H of x; matrix resident (convertMatrix is setup, as in the reference); second of two "
                       "solves on the Comm (the first registers the peer window once per partition)" % (ke - 1),
               "ms_per_step": dt * 1e3 / (ke - 1),
               "breakdown_ms": {"create": max_over_ranks(create_ms), "loop": max_over_ranks(loop_ms), "finish": max_over_ranks(finish_ms),
                                "total": dt * 1e3},
               "first_solve_ms": {"create": max_over_ranks(first[2]), "total": first[0] * 1e3},
               "answer_check": {"history_max_rel_diff_vs_device_run": hd, "max_abs_x_minus_1": xerr, "maxError_reported": einfo.maxError}}
        L.sbFreeHost(hb); L.sbFreeHost(hx)
        note('e2e done')
    cg_line = {"iterations_per_sec": its, "gbs_per_gpu": P.B_it / (ms / K * 1e-3) / 1e9, "gflops_per_gpu": P.F_it / (ms / K * 1e-3) / 1e9,
               "frac_of_peak": P.B_it / (ms / K * 1e-3) / 1e9 / peak, "kernel_ms_per_iteration": region,
               "residual_initial": resid0, "residual_final": resid, "max_error_vs_xexact": info.maxError}
    plugin = {"format": fmt, "C": Cc, "sigma": sigma, "nnz_rank0": nnz, "setup_s": round(P.setup_s, 2)}
    B_spmv, B_spmv_fmt, F_spmv = P.B_spmv, P.B_spmv_fmt, P.F_spmv
    P.destroy()

    def side_leg(ax, ay, az, afmt, aC, asig, adesc, steps):
        """another BASELINE config measured by the same method in the same run (device-resident CG + per-kernel pass)"""
        Q = Problem(api, L, comm, ax, ay, az, afmt, aC, asig, rank, world)
        barrier()
        ms1, _l, i1, _h = timed_cg(Q, timer, W, steps, barrier, max_over_ranks)
        reg = profiled_cg(Q, W, steps)
        sp = reg["spmv"] + reg["spmv_boundary"]
        out = {"workload": adesc, "format": afmt, "steps": steps, "ms_per_step": ms1 / steps, "iterations_per_sec": steps / (ms1 * 1e-3),
               "gflops": sum(2 * local_nnz(ax, ay, az, r, world) + 10 * ax * ay * az for r in range(world)) * steps / (ms1 * 1e-3) / 1e9,
               "cg_gbs_per_gpu": Q.B_it * steps / (ms1 * 1e-3) / 1e9, "frac_of_peak": Q.B_it * steps / (ms1 * 1e-3) / 1e9 / peak,
               "spmv_ms": sp, "spmv_gbs": Q.B_spmv / (sp * 1e-3) / 1e9, "spmv_frac_of_peak": Q.B_spmv / (sp * 1e-3) / 1e9 / peak,
               "spmv_gbs_format_bytes": Q.B_spmv_fmt / (sp * 1e-3) / 1e9, "kernel_ms_per_iteration": reg,
               "max_error_vs_xexact": i1.maxError, "setup_s": round(Q.setup_s, 2)}
        Q.destroy()
        return out

    # ---- BASELINE.json configs[1] beside the headline (single GPU, default workload only): 128^3 CRS CG
    also = None
    if world == 1 and args.workload == "sell256" and not args.no_extra:
        try:
            wl = WORKLOADS["crs128"]
            also = side_leg(wl[0], wl[1], wl[2], wl[3], wl[4], wl[5], wl[6], K)
        except Exception as e:
            also = {"workload": "crs128", "failed": repr(e)}
        note("configs1 done")

    # ---- BASELINE.json configs[4] at this N: 512^3 GLOBAL problem (strong scaling), SELL vs CRS vs CCRS
    configs4 = None
    if args.workload == "sell256" and not args.no_extra and 512 % world == 0:
        configs4 = {"global": "512x512x512", "nz_per_gpu": 512 // world, "scaling": "strong", "formats": {}}
        for name in ("strong512sell", "strong512crs", "strong512ccrs"):
            wl = WORKLOADS[name]
            try:
                configs4["formats"][wl[3]] = side_leg(wl[0], wl[1], 512 // world, wl[3], wl[4], wl[5], wl[6], min(K, 20))
            except Exception as e:
                configs4["formats"][wl[3]] = {"failed": repr(e)}
            note("configs4 %s done" % name)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = reference_arm(min(K, 20), 3, nx, ny, nz, 1, args.ref_planes)
            if r:
                cpu = {"value": r["value"] * flops_it_job / 1e9, "unit": unit, "cores": r["cores"], "kind": r["kind"],
                       "sample": r["sample"], "same_config": r["same_config"], "sampled_fraction": r["sampled_fraction"],
                       "iterations_per_sec": r["value"], "setup_s": round(r["setup_s"], 2)}
        except Exception as e:  # the CPU leg must never take the GPU numbers down with it
            cpu = {"value": None, "unit": unit, "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}

    if rank == 0:
        line = {
            "metric": metric, "value": value * 1.0, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config,
            "plugin": plugin,
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "spmv (%s) fused with p.Ap" % fmt, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": B_spmv, "avg_launch_ms": spmv_ms,
                         "frac_of_nominal_8000": achieved / 8000.0,
                         "read_only_stream_gbs": read_peak, "frac_of_read_only_stream": achieved / read_peak,
                         "read_only_stream_source": "sbMeasureReadBandwidth: own ld.global.nc kernel over a fresh 4 GiB buffer, best of 5, this run"},
            "cpu_baseline": cpu,
            "parity": parity,
            "configs1": also,
            "configs4": configs4,
            "spmv": {"gflops": F_spmv / (spmv_only_ms * 1e-3) / 1e9, "gbs": B_spmv / (spmv_only_ms * 1e-3) / 1e9,
                     "gbs_format_bytes": B_spmv_fmt / (spmv_only_ms * 1e-3) / 1e9, "ms": spmv_only_ms,
                     "frac_of_peak": B_spmv / (spmv_only_ms * 1e-3) / 1e9 / peak, "mode": "x=1, back-to-back (main.c:200-216)"},
            "cg": cg_line,
        }
        emit(line)
    if world > 1:
        L.commFinalize(C.byref(comm))
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
