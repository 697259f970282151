// Conjugate-gradient driver (replaces solveCG / initVectors / solverCheckResidual, CGSolver.c:19-141).
//
// The operation order, the lagging `normr > eps` test and the returned loop counter are the reference's.
// What is B200-native is how an iteration runs: three kernels on one stream --
//   p = r + beta p                                   (beta from device scalars rho[k-1]/rho[k-2])
//   Ap = A p  fused with  pAp[k] = p.Ap
//   x += alpha p ; r -= alpha Ap  fused with  rho[k] = r.r      (alpha = rho[k-1]/pAp[k] on the device)
// -- with every scalar resident in HBM. The host only needs rho[k-2] to decide whether iteration k runs
// (that IS the reference's lagging test), so it always has one full iteration queued while it waits: no
// per-iteration pipeline drain.
#include <math.h>
#include <string.h>
#include <time.h>

#include <vector>

#include "sb_internal.h"

namespace sb {

namespace {

// hRho[j] holds this bit pattern (a NaN no arithmetic produces) until the device has stored the global rho[j] there
constexpr unsigned long long kRhoPending = 0xffffffffffffffffull;
enum Region { R_UPDATE_P = 0, R_EXCHANGE, R_SPMV, R_ALLREDUCE, R_UPDATE_XR, R_HALO_WAIT, R_SPMV_BOUNDARY, R_COUNT };

bool commActive(const Comm* c) { return c && c->size > 1; }

struct CgSolver {
  Comm* comm = nullptr;
  Operator A;
  cudaStream_t s = nullptr;
  double eps = 0.0;
  int itermax = 0, flags = 0, printFreq = 1;
  bool fused = true, print = false, generated = false, profile = false, overlap = true, gated = false;
  bool fusedReduce = false;                // multi-GPU: all-reduces split into push (producer) / collect (consumer)
  bool fusedPut = false;                   // multi-GPU: the p update stores boundary values straight into the neighbours' p
  PeerReduce pendingRho;                   // epoch of the newest rho that has been pushed but not yet collected
  const int* elems = nullptr;              // multi-GPU send list in solver numbering (owned by the Comm)
  bool pBorrowed = false;                  // p is the Comm's persistent, peer-mapped halo vector
  unsigned long long* syncTrace = nullptr; // SB_SYNC_TRACE=1: device counters of the waits inside the multi-GPU kernels
  idx_t intLo = 0, intHi = 0;           // SpMV units [intLo, intHi) reference no halo column
  idx_t n = 0;
  size_t rowSlots = 0, colSlots = 0;
  // vectors (CGSolver.c:69-79), solver order (SELL: permuted)
  real_t *r = nullptr, *p = nullptr, *Ap = nullptr, *x = nullptr, *b = nullptr, *tmp = nullptr;
  real_t *rho = nullptr, *pAp = nullptr;   // device scalars indexed by iteration: rho[j] = r_j.r_j, pAp[k] = p_k.Ap_k
  real_t* hRho = nullptr;                  // mapped pinned mirror of rho: written by the kernels themselves, polled by the host
  std::vector<double> hist;
  double normr = 0.0;                      // printed / compared in double like the reference's sqrt() result
  real_t rtrans = 0.0, oldrtrans = 0.0;
  int k = 1;
  bool stopped = false;
  // optional per-kernel event timing: a fixed pool of events created in setup(), drained into regionMs whenever
  // it is full, so the timed loop never creates an event and a long run never holds more than kProfEvents
  static constexpr int kProfEvents = 512;
  std::vector<cudaEvent_t> evPool;
  int evRegion[kProfEvents];
  int evUsed = 0;
  double regionMs[R_COUNT] = { 0, 0, 0, 0, 0, 0, 0 };
  double createMs = 0.0;

  void drainMarks()
  {
    if (evUsed < 2) return;
    SB_CUDA(cudaEventSynchronize(evPool[(size_t)evUsed - 1]));
    for (int i = 1; i < evUsed; i++) {
      float ms = 0.f;
      SB_CUDA(cudaEventElapsedTime(&ms, evPool[(size_t)i - 1], evPool[(size_t)i]));
      if (evRegion[i] >= 0) regionMs[evRegion[i]] += ms;
    }
    std::swap(evPool[0], evPool[(size_t)evUsed - 1]);         // the last stamp opens the next batch
    evRegion[0] = -1;
    evUsed = 1;
  }

  void mark(int region)
  {
    if (!profile) return;
    if (evUsed == kProfEvents) drainMarks();
    SB_CUDA(cudaEventRecord(evPool[(size_t)evUsed], s));
    evRegion[evUsed++] = region;
  }

  // Blocks until the device has published the global rho[j]. No copy and no event sits in the stream for this: the
  // kernel that completes the sum stores it into mapped host memory (gridSum's mirror / the p update's collect).
  static bool rhoPending(real_t v)
  {
    return memcmp(&v, &kRhoPending, sizeof(real_t)) == 0;     // all-ones: a NaN pattern no arithmetic produces
  }
  real_t waitRho(int j)
  {
    volatile real_t* slot = hRho + j;
    for (unsigned long spins = 1;; spins++) {
      const real_t v = *slot;
      if (!rhoPending(v)) return v;
      if ((spins & 0x3fff) == 0) {             // a failed or finished stream must not leave the host spinning
        const cudaError_t e = cudaStreamQuery(s);
        if (e == cudaSuccess) {
          if (rhoPending(*slot)) SB_FATAL("CG: the stream drained but rho[%d] never arrived", j);
        } else if (e != cudaErrorNotReady) {
          SB_CUDA(e);
        }
      }
    }
  }

  void allreduce(real_t* d, int op)
  {
    if (!commActive(comm)) return;
    commAllreduceDevice(comm, d, 1, op, s);
    mark(R_ALLREDUCE);
  }

  // commExchange + spMVM (CGSolver.c:95-96,122-123). With the peer-window transport my boundary values are stored
  // straight behind the neighbours' copy of p, and ONE SpMV launch multiplies the rows that reference no halo
  // column while those stores are in flight, waits on the arrival counters, and finishes with the boundary rows.
  void spmvWithHalo(const DotArgs* dot, const HaloGate* already = nullptr)
  {
    const idx_t units = spmvUnits(A);
    if (commActive(comm)) {
      if (gated) {
        // `already`: the p update itself delivered the halo (FusedPut); otherwise a put kernel does
        const HaloGate gate = already ? *already : commHaloPutDirect(comm, p, elems, s);
        mark(R_EXCHANGE);
        launchSpmvGated(A, p, Ap, intLo, intHi, gate, dot, s);
        mark(R_SPMV);
        return;
      }
      commExchangeOnStream(comm, A.nr, p, elems, s);
      mark(R_EXCHANGE);
    }
    launchSpmv(A, p, Ap, 0, units, dot, s);
    mark(R_SPMV);
  }

  // caller vector (host or device, original row order) -> device vector in solver order, on stream `st`;
  // `stage` receives a host vector that still has to be permuted
  void importVector(const real_t* src, real_t* dst, real_t* stage, cudaStream_t st)
  {
    const size_t bytes = sizeof(real_t) * n;
    const bool dev = isDevicePointer(src);
    if (dev) ensureOnDevice(src);
    if (!A.oldToNew) {
      SB_CUDA(cudaMemcpyAsync(dst, src, bytes, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    } else {
      const real_t* staged = src;
      if (!dev) {
        SB_CUDA(cudaMemcpyAsync(stage, src, bytes, cudaMemcpyHostToDevice, st));
        staged = stage;
      }
      launchScatter(n, A.oldToNew, staged, dst, st);          // dst[oldToNew[i]] = src[i]
    }
  }

  void exportVector(const real_t* src, real_t* dst)
  {
    const size_t bytes = sizeof(real_t) * n;
    const bool dev = isDevicePointer(dst);
    if (!A.oldToNew) {
      SB_CUDA(cudaMemcpyAsync(dst, src, bytes, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
    } else if (dev) {
      launchGather(n, A.oldToNew, src, dst, s);               // dst[i] = src[oldToNew[i]]
    } else {
      launchGather(n, A.oldToNew, src, tmp, s);
      SB_CUDA(cudaMemcpyAsync(dst, tmp, bytes, cudaMemcpyDeviceToHost, s));
    }
  }

  void printIteration(int iter, double value)
  {
    if (print && (iter % printFreq == 0 || iter + 1 == itermax)) printf("Iteration = %d Residual = %E\n", iter, value);   // :118-120
  }

  void setup(Comm* comm_, Parameter* param, const Operator& A_, const SbCGInfo* info)
  {
    Context& c = ctx();
    comm = comm_;
    A = A_;
    s = c.stream;
    eps = (double)(real_t)param->eps;                         // CG_FLOAT eps = (CG_FLOAT)param->eps, CGSolver.c:64
    itermax = param->itermax;
    flags = info ? info->flags : (SB_CG_FUSED | SB_CG_PRINT);
    fused = (flags & SB_CG_FUSED) != 0;
    profile = (flags & SB_CG_PROFILE) != 0;
    print = (flags & SB_CG_PRINT) != 0 && (!comm || comm->rank == 0);
    overlap = (flags & SB_CG_NO_OVERLAP) == 0 && getenv("SB_CG_NO_OVERLAP") == nullptr;
    generated = param->filename && (strcmp(param->filename, "generate") == 0 || strcmp(param->filename, "generate7P") == 0);
    n = A.nr;
    rowSlots = (size_t)(A.nrPadded > n ? A.nrPadded : n) + 2;
    colSlots = (A.nc > rowSlots ? (size_t)A.nc : rowSlots) + 2;
    const int nScal = (itermax > 0 ? itermax : 0) + 4;
    // direct halo delivery into p + gated single-launch SpMV: p is then the Comm's persistent vector, which every
    // peer mapped when the partition was first solved on (no per-solve registration)
    if (commActive(comm)) {
      if (commPeerMode(comm)) {
        const bool can = overlap && spmvGatedAvailable(A);
        if (can) spmvInteriorUnits(A, &intLo, &intHi, s);
        p = commAcquireHaloVector(comm, A.nr, colSlots, can);
        pBorrowed = gated = p != nullptr;
      }
      elems = commSolverElements(comm, A.permKey, A.oldToNew, s);   // vectors are row-permuted: send p[oldToNew[element]]
      if (gated && getenv("SB_SYNC_TRACE")) {
        syncTrace = (unsigned long long*)sbAllocateDevice(64, sizeof(unsigned long long) * 16);
        SB_CUDA(cudaMemsetAsync(syncTrace, 0, sizeof(unsigned long long) * 16, s));
      }
      if (gated) {
        fusedReduce = fused && getenv("SB_NO_FUSED_REDUCE") == nullptr;
        fusedPut = fusedReduce && getenv("SB_NO_FUSED_PUT") == nullptr && commPrepareFusedPut(comm, A.permKey, elems);
      }
    }
    r = (real_t*)sbAllocateDevice(64, sizeof(real_t) * rowSlots);
    if (!p) {
      p = (real_t*)sbAllocateDevice(64, sizeof(real_t) * colSlots);
      SB_CUDA(cudaMemsetAsync(p, 0, sizeof(real_t) * colSlots, s));
    }
    Ap = (real_t*)sbAllocateDevice(64, sizeof(real_t) * rowSlots);
    x = (real_t*)sbAllocateDevice(64, sizeof(real_t) * rowSlots);
    b = (real_t*)sbAllocateDevice(64, sizeof(real_t) * rowSlots);
    tmp = (real_t*)sbAllocateDevice(64, sizeof(real_t) * rowSlots);
    rho = (real_t*)sbAllocateDevice(64, sizeof(real_t) * nScal);
    pAp = (real_t*)sbAllocateDevice(64, sizeof(real_t) * nScal);
    hRho = (real_t*)sbAllocateHost(sizeof(real_t) * nScal);
    SB_CUDA(cudaMemsetAsync(x, 0, sizeof(real_t) * rowSlots, s));
    SB_CUDA(cudaMemsetAsync(rho, 0, sizeof(real_t) * nScal, s));
    SB_CUDA(cudaMemsetAsync(pAp, 0, sizeof(real_t) * nScal, s));
    if (profile) {
      evPool.resize(kProfEvents);
      for (cudaEvent_t& e : evPool) SB_CUDA(cudaEventCreate(&e));
    }
    hist.reserve((size_t)nScal);

    // initVectors (CGSolver.c:19-38), or caller-supplied b / x0. x0 is needed first (p = x0, A p); a host b travels
    // on the side stream behind it while the main stream already multiplies, and joins before r = b - A p.
    launchInitVectors(n, A.rowPtr, A.rowLen, generated, x, b, s);
    cudaEvent_t bReady = nullptr;
    if (info && info->b) {
      if (isDevicePointer(info->b)) {
        importVector(info->b, b, tmp, s);
      } else {
        cudaEvent_t bFree;
        SB_CUDA(cudaEventCreateWithFlags(&bFree, cudaEventDisableTiming));
        SB_CUDA(cudaEventCreateWithFlags(&bReady, cudaEventDisableTiming));
        if (info->x) importVector(info->x, x, Ap, s);          // queue x0 ahead of b on the link
        SB_CUDA(cudaEventRecord(bFree, s));
        SB_CUDA(cudaStreamWaitEvent(c.commStream, bFree, 0));
        importVector(info->b, b, tmp, c.commStream);
        SB_CUDA(cudaEventRecord(bReady, c.commStream));
        SB_CUDA(cudaEventDestroy(bFree));
      }
    }
    if (info && info->x && !bReady) importVector(info->x, x, Ap, s);

    // pre-loop (CGSolver.c:94-100)
    launchWaxpby(n, 1.0, x, 0.0, x, p, s);
    spmvWithHalo(nullptr);
    if (bReady) {
      SB_CUDA(cudaStreamWaitEvent(s, bReady, 0));
      SB_CUDA(cudaEventDestroy(bReady));
    }
    launchWaxpby(n, 1.0, b, -1.0, Ap, r, s);
    launchDot(n, r, r, rho, 0, s);
    allreduce(rho, SB_SUM);
    SB_CUDA(cudaMemcpyAsync(hRho, rho, sizeof(real_t), cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    rtrans = hRho[0];
    for (int j = 1; j < nScal; j++) memcpy(hRho + j, &kRhoPending, sizeof(real_t));
    normr = (double)(real_t)sqrt((double)rtrans);             // CG_FLOAT normr = sqrt(rtrans), CGSolver.c:100
    hist.push_back(normr);
    if (print) printf("Initial Residual = %E\n", normr);      // :102
    printFreq = itermax / 10;                                 // :85-91
    if (printFreq > 50) printFreq = 50;
    if (printFreq < 1) printFreq = 1;
    k = 1;
    if (profile) {   // drop the pre-loop marks, start the clock here
      evUsed = 0;
      mark(-1);
    }
  }

  // Runs iterations while k < min(untilK, itermax) and the lagging test passes. Asynchronous: returns with up to
  // two iterations still in flight.
  int iterate(int untilK)
  {
    const int stopK = untilK < itermax ? untilK : itermax;
    if (stopped) return k;
    if (fused) {
      // the normr tested before iteration k is sqrt(rho[max(k-2,0)]) -- the reference's lagging test (:107,:116)
      for (; k < stopK; k++) {
        if ((int)hist.size() < k) {                            // hist[k-1] = normr of iteration k-1 = sqrt(rho[k-2])
          if (k >= 3) normr = (double)(real_t)sqrt((double)waitRho(k - 2));
          hist.push_back(normr);
          printIteration(k - 1, normr);
        }
        if (!(normr > eps)) {
          stopped = true;
          break;
        }
        if (fusedReduce) {
          // rho[k-1] is summed over the ranks inside the p update, p.Ap inside the x/r update: no all-reduce launches
          FusedPut fp;
          HaloGate gate;
          if (fusedPut) gate = commFusedPutBegin(comm, &fp);    // the halo exchange (:122) rides on the p update
          if (syncTrace) {
            gate.trace = syncTrace;
            pendingRho.trace = pendingRho.size ? syncTrace + 8 : nullptr;
          }
          // the p update completes the global rho[k-1] and stores it to the host mirror itself
          launchCgUpdateP(n, k, rho, r, p, pendingRho.size ? &pendingRho : nullptr, fusedPut ? &fp : nullptr, hRho, s);   // :109 / :111-114
          mark(R_UPDATE_P);
          PeerReduce prPAp = commBeginReduce(comm);
          if (syncTrace) prPAp.trace = syncTrace + 4;
          DotArgs d { pAp + k, false, 1, &prPAp };
          spmvWithHalo(&d, fusedPut ? &gate : nullptr);         // :122-125
          pendingRho = commBeginReduce(comm);
          launchCgUpdateXR(n, k, rho, pAp, x, r, p, Ap, 2, &prPAp, &pendingRho, nullptr, s);   // :126-128 (+ :112 of iteration k+1)
          mark(R_UPDATE_XR);
          continue;
        }
        launchCgUpdateP(n, k, rho, r, p, nullptr, nullptr, nullptr, s); // :109 / :111-114
        mark(R_UPDATE_P);
        DotArgs d { pAp + k, false, 1 };
        spmvWithHalo(&d);                                      // :122-125
        allreduce(pAp + k, SB_SUM);
        // one GPU: the x/r update's grid reduction stores rho[k] to the host mirror; several GPUs with stand-alone
        // all-reduce kernels: the global value only exists after the all-reduce, copy it from there
        launchCgUpdateXR(n, k, rho, pAp, x, r, p, Ap, 2, nullptr, nullptr, commActive(comm) ? nullptr : hRho, s);   // :126-128 (+ :112 of iteration k+1)
        mark(R_UPDATE_XR);
        if (commActive(comm)) {
          allreduce(rho + k, SB_SUM);
          SB_CUDA(cudaMemcpyAsync(hRho + k, rho + k, sizeof(real_t), cudaMemcpyDeviceToHost, s));
        }
      }
    } else {
      // the reference's call sequence through the drop-in entry points (host scalars, 5 kernels + 2 syncs / iteration)
      for (; k < stopK; k++) {
        if (!(normr > eps)) {
          stopped = true;
          break;
        }
        if (k == 1) {
          launchWaxpby(n, 1.0, r, 0.0, r, p, s);
        } else {
          oldrtrans = rtrans;
          ddot(n, r, r, &rtrans);
          const real_t beta = rtrans / oldrtrans;
          launchWaxpby(n, 1.0, r, beta, p, p, s);
        }
        normr = (double)(real_t)sqrt((double)rtrans);
        hist.push_back(normr);
        printIteration(k, normr);
        spmvWithHalo(nullptr);
        real_t alpha = 0.0;
        ddot(n, p, Ap, &alpha);
        alpha = rtrans / alpha;
        launchWaxpby(n, 1.0, x, alpha, p, x, s);
        launchWaxpby(n, 1.0, r, -alpha, Ap, r, s);
      }
    }
    return k;
  }

  int finish(SbCGInfo* info, float loopMs)
  {
    Context& c = ctx();
    SB_CUDA(cudaStreamSynchronize(s));
    if (pBorrowed) commReleaseHaloVector(comm);
    if (syncTrace) {
      unsigned long long h[16];
      sbCopyToHost(h, syncTrace, sizeof(h));
      sbFree(syncTrace);
      fprintf(stderr, "[sync trace r%d] %d iterations | halo gate: %.2f us mean, %.2f us max per waiting CTA/warp (%llu waits) | "
                      "collect p.Ap (x/r update, block 0): %.2f us mean over %llu | collect r.r (p update, block 0): %.2f us mean over %llu\n",
          comm ? comm->rank : 0, k - 1, h[2] ? 1e-3 * (double)h[0] / (double)h[2] : 0.0, 1e-3 * (double)h[1], h[2],
          h[5] ? 1e-3 * (double)h[4] / (double)h[5] : 0.0, h[5], h[9] ? 1e-3 * (double)h[8] / (double)h[9] : 0.0, h[9]);
    }
    if (fused && (int)hist.size() < k) {
      // iteration k-1 was the last one executed; record its normr = sqrt(rho[k-2])
      const double last = k >= 3 ? (double)(real_t)sqrt((double)hRho[k - 2]) : normr;
      hist.push_back(last);
      printIteration(k - 1, last);
    }
    if (print) printf("Solution performed %d iterations and took %.2fs\n", k, loopMs * 1e-3);   // :133
    double maxErr = -1.0;
    if (generated) {                                           // solverCheckResidual, :40-60
      launchMaxErr(n, x, c.dWide, s);
      SB_CUDA(cudaMemcpyAsync(c.hWide, c.dWide, sizeof(double), cudaMemcpyDeviceToHost, s));
      SB_CUDA(cudaStreamSynchronize(s));
      CG_FLOAT worst = (CG_FLOAT)c.hWide[0];                     // CG_FLOAT residual, CGSolver.c:46-55
      commReduction(&worst, SB_MAX);
      maxErr = (double)worst;
      if (print) printf("Difference between computed and exact  = %f\n", maxErr);
    }
    if (info) {
      if (info->x) exportVector(x, info->x);
      SB_CUDA(cudaStreamSynchronize(s));
      info->nhist = (int)hist.size();
      if (info->history)
        for (int i = 0; i < info->nhist && i < info->historyCap; i++) info->history[i] = hist[(size_t)i];
      info->solveMs = loopMs;
      info->maxError = maxErr;
      for (int i = 0; i < R_COUNT; i++) info->regionMs[i] = 0.0;
      if (profile) {
        drainMarks();
        for (int i = 0; i < R_COUNT; i++) info->regionMs[i] = regionMs[i];
      }
    }
    for (cudaEvent_t e : evPool) cudaEventDestroy(e);
    sbFree(r); sbFree(Ap); sbFree(x); sbFree(b); sbFree(tmp); sbFree(rho); sbFree(pAp);
    if (!pBorrowed) sbFree(p);
    sbFreeHost(hRho);
    return k;
  }
};

} // namespace

} // namespace sb

using namespace sb;

extern "C" {

void* sbCGCreate(Comm* comm, Parameter* param, void* matrix, int fmt, SbCGInfo* info)
{
  CgSolver* S = new CgSolver();
  S->setup(comm, param, makeOperator(matrix, fmt), info);
  return S;
}

int sbCGIterate(void* solver, int untilK) { return ((CgSolver*)solver)->iterate(untilK); }

int sbCGFinish(void* solver, SbCGInfo* info, double loopMs)
{
  CgSolver* S = (CgSolver*)solver;
  const int k = S->finish(info, (float)loopMs);
  delete S;
  return k;
}

static double wallMs()
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec * 1e3 + (double)ts.tv_nsec * 1e-6;
}

int sbSolveCG(Comm* comm, Parameter* param, void* matrix, int fmt, SbCGInfo* info)
{
  const double t0 = wallMs();
  CgSolver* S = (CgSolver*)sbCGCreate(comm, param, matrix, fmt, info);   // ends with a drained stream (initial residual)
  const double t1 = wallMs();
  cudaEvent_t a, b;
  SB_CUDA(cudaEventCreate(&a));
  SB_CUDA(cudaEventCreate(&b));
  SB_CUDA(cudaEventRecord(a, S->s));                           // timeStart, CGSolver.c:106
  S->iterate(param->itermax);
  SB_CUDA(cudaEventRecord(b, S->s));                           // timeStop, :130
  SB_CUDA(cudaEventSynchronize(b));
  float ms = 0.f;
  SB_CUDA(cudaEventElapsedTime(&ms, a, b));
  SB_CUDA(cudaEventDestroy(a));
  SB_CUDA(cudaEventDestroy(b));
  const double t2 = wallMs();
  const int k = sbCGFinish(S, info, ms);
  const double t3 = wallMs();
  if (info) {
    info->createMs = t1 - t0;
    info->finishMs = t3 - t2;
  }
  if (getenv("SB_CG_TRACE"))
    fprintf(stderr, "[sbSolveCG] create %.2f ms, loop %.2f ms (device %.2f), finish %.2f ms\n", t1 - t0, t2 - t1, ms, t3 - t2);
  return k;
}

// The reference accumulates wall time per region in the global _t[] of its profiler.c (profiler.c:17,
// profiler.h:18-24: WAXPBY, SPMVM, DDOT, COMM) from PROFILE() macros inside CGSolver.c, and its profilerPrint turns
// them into MB/s and MFlop/s. CGSolver.c is replaced by this file, so when the program contains that array (weak
// reference: only then) the drop-in solveCG fills it from CUDA-event times of its own kernels. The dot products are
// fused into the vector and SpMV kernels here and have no time of their own: the vector-kernel time is split
// 24:16 between WAXPBY and DDOT like the reference's own per-iteration byte model (profiler.c:19-22).
extern double _t[4] __attribute__((weak));

static int solveDropIn(Comm* comm, Parameter* param, void* m, int fmt)
{
  if (!&_t[0] || getenv("SB_NO_PROFILER_T")) return sbSolveCG(comm, param, m, fmt, nullptr);
  SbCGInfo info;
  memset(&info, 0, sizeof(info));
  info.flags = SB_CG_FUSED | SB_CG_PRINT | SB_CG_PROFILE;
  const int k = sbSolveCG(comm, param, m, fmt, &info);
  const double vec = (info.regionMs[SB_REGION_UPDATE_P] + info.regionMs[SB_REGION_UPDATE_XR]) * 1e-3;
  _t[0] += vec * 0.6;
  _t[2] += vec * 0.4;
  _t[1] += (info.regionMs[SB_REGION_SPMV] + info.regionMs[SB_REGION_SPMV_BOUNDARY]) * 1e-3;
  _t[3] += (info.regionMs[SB_REGION_EXCHANGE] + info.regionMs[SB_REGION_HALO_WAIT] + info.regionMs[SB_REGION_ALLREDUCE]) * 1e-3;
  return k;
}

int sbCRS_solveCG(Comm* comm, Parameter* param, SbCRSMatrix* m) { return solveDropIn(comm, param, m, SB_FMT_CRS); }
int sbSCS_solveCG(Comm* comm, Parameter* param, SbSCSMatrix* m) { return solveDropIn(comm, param, m, SB_FMT_SCS); }
int sbCCRS_solveCG(Comm* comm, Parameter* param, SbCCRSMatrix* m) { return solveDropIn(comm, param, m, SB_FMT_CCRS); }

} // extern "C"
