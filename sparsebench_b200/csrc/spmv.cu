// Sparse matrix-vector kernels of the three format plugins (replace spMVM of matrix-CRS.c:46-65,
// matrix-SCS.c:198-228, matrix-CCRS.c:14-31). HBM-bound, no tensor cores: 12 B (16 B for CCRS) of matrix
// stream per non-zero against 2 flops; x is gathered through L1/L2, y written once.
//
// Common shape: one persistent CTA per SM walks the row/chunk range round-robin, so that at any instant the
// whole chip works on one contiguous window of the matrix (and of x: the window's x stays in L2 and is
// read from HBM once). The matrix stream -- values and column ids, read exactly once -- is moved
// global -> shared memory by the bulk-copy engine (cp.async.bulk, SASS UBLKCP) with mbarrier completion
// instead of through registers: ~200 KB of matrix are in flight per SM (Little's law needs ~45 KB per SM at
// 6.5 TB/s) independent of register allocation, warps only spend instructions on LDS + the x gather + the
// fp64 accumulate, and L1 holds nothing but the gathered x vector because the stream never passes through
// it. An optional fused epilogue accumulates sum_i x[i]*y[i] (the CG's p.Ap, CGSolver.c:125) with the
// deterministic one-kernel grid reduction from device_utils.cuh.
#include "device_utils.cuh"
#include "sb_internal.h"

namespace sb {

template <typename K>
static void allowLargeSmem(K kernel, size_t bytes)
{
  SB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

static int envInt(const char* name, int dflt)
{
  const char* e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}

// SB_SPMV_LEGACY=1 selects the register-staged kernels (kept as the measured baseline of profiles/)
static bool useLegacyKernels()
{
  static const int v = envInt("SB_SPMV_LEGACY", 0);
  return v != 0;
}

// ------------------------------------------------------------------------------------------- SELL-32-sigma
// One warp per chunk, lane = row of the chunk: the j-th column of a chunk is 32 consecutive values (256 B) and
// 32 consecutive column ids (128 B). Every warp owns a private ring of S stages of J chunk columns
// (J x 32 x 12 bytes) with one mbarrier each; producer and consumer of a ring are the same warp: a stage is
// refilled right after the warp has consumed it (__syncwarp orders the lanes' shared-memory reads before lane 0
// issues the overwrite), so no "empty" barriers are needed. A stage is consumed in batches of U columns: U
// independent x gathers in flight per lane, then the U accumulations in stored order, exactly like
// tmp[k] += val*x of matrix-SCS.c:216-222 (bit-identical row sums).
// Chunk lengths / offsets of a warp's next 32 chunks live one per lane and are broadcast by shuffle, so the
// metadata loads are off the critical path.
// x gather. Plain kernels use the read-only path (ld.global.nc). Gated kernels read halo values that peers store
// WHILE the kernel runs: those loads must be coherent ones -- then the acquire on the arrival counter plus the
// block / warp barrier behind it orders them after the peers' stores (ld.global.nc gives no such guarantee).
__device__ __forceinline__ double ldCoherent(const double* p)
{
  double v;
  asm("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldCoherent(const float* p)
{
  float v;
  asm("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
template <bool COHERENT>
__device__ __forceinline__ real_t gatherX(const real_t* p)
{
#ifndef SB_EXPERIMENT_NC       // -DSB_EXPERIMENT_NC: timing experiment only (read-only path even behind the gate)
  if (COHERENT) return ldCoherent(p);
#endif
  return __ldg(p);
}

// Device-side wait on a HaloGate (sb_internal.h): arrival counters written by the peers' put kernels.
// The gate's counter addresses are parked in shared memory at kernel start (statically indexed copy of the kernel
// parameter), so that the wait itself is a small dynamic loop that costs the hot path no registers.
struct GateSmem {
  const unsigned long long* flag[kMaxGateSources];
  unsigned long long target[kMaxGateSources];
  unsigned long long* trace;
};
__device__ __forceinline__ void gateStore(GateSmem& g, const HaloGate& gate)
{
#pragma unroll
  for (int i = 0; i < kMaxGateSources; i++) {
    g.flag[i] = gate.flag[i];
    g.target[i] = gate.target[i];
  }
  g.trace = gate.trace;
}
// Called by ONE thread.
__device__ __noinline__ void gateWait(const GateSmem* g, int nsrc)
{
  unsigned long long* trace = g->trace;
  const unsigned long long t0 = trace ? globalTimerNs() : 0ull;
  for (int i = 0; i < nsrc; i++) {
    const unsigned long long* p = g->flag[i];
    const unsigned long long target = g->target[i];
    unsigned long long v;
    const unsigned long long start = globalTimerNs();
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
      if (v >= target) break;
      __nanosleep(40);
      if (globalTimerNs() - start > kPeerTimeoutNs) __trap();   // dead peer: fail instead of hanging the GPU
    }
  }
  if (trace) {
    const unsigned long long dt = globalTimerNs() - t0;
    atomicAdd(trace, dt);
    atomicMax(trace + 1, dt);
    atomicAdd(trace + 2, 1ull);
  }
}

// GATED launches cover all chunks in one go, rotated so that the `nInterior` chunks that reference no halo column
// come first; before the CTA's first step outside that range it waits for the halo (gateWait). All x gathers of a
// gated kernel are coherent loads (gatherX), because the peers store halo values while it is already running.
template <bool DOT, int WARPS, int J, int S, int U, bool LOCKSTEP, bool GATED>
__global__ void __launch_bounds__(WARPS * 32, 1)
spmvSell32TmaKernel(SellView A, const real_t* __restrict__ x, real_t* __restrict__ y, idx_t lo, idx_t hi,
    real_t* partials, unsigned int* ticket, real_t* dotOut, bool accumulate, idx_t rot, idx_t nInterior, HaloGate gate,
    PeerReduce push)
{
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ real_t scratch[32];
  __shared__ uint64_t bars[WARPS * S];
  __shared__ GateSmem gateSmem;
  constexpr uint32_t kValBytes = 32 * sizeof(real_t), kColBytes = 32 * sizeof(idx_t);   // one chunk column: 32 values, 32 ids
  constexpr uint32_t kStageBytes = J * (kValBytes + kColBytes);
  constexpr uint32_t kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* mine = ring + (size_t)warp * S * kStageBytes;
  uint64_t* bar = bars + warp * S;
  griddepLaunchDependents();
  if (GATED && threadIdx.x == 0) gateStore(gateSmem, gate);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; s++) mbarInit(bar + s, 1);
    mbarFenceInit();
  }
  __syncwarp();

  // logical chunk l in [0, n) is physical chunk lo + (l + rot) mod n
  const uint64_t n = (uint64_t)hi - lo;
  auto phys = [&](uint64_t l) {
    uint64_t p = l + rot;
    if (p >= n) p -= n;
    return p + lo;
  };
  const uint64_t stride = (uint64_t)gridDim.x * WARPS;
  const uint64_t ctaFirst = (uint64_t)blockIdx.x * WARPS;
  const uint64_t first = ctaFirst + warp;
  // this warp's t-th chunk is first + t*stride; lane l of a metadata block B holds chunk t = 32*B + l
  auto loadMeta = [&](idx_t block, idx_t& lenL, idx_t& ptrL) {
    const uint64_t l = first + ((uint64_t)block * 32 + lane) * stride;
    const uint64_t ch = l < n ? phys(l) : 0;
    lenL = l < n ? __ldg(A.chunkLens + ch) : 0u;
    ptrL = l < n ? __ldg(A.chunkPtr + ch) : 0u;
  };

  // producer cursor: next slab to request = columns [pj, pj+J) of logical chunk pc (empty chunks have no slab)
  idx_t pLenL, pPtrL, pT = 0;
  loadMeta(0, pLenL, pPtrL);
  idx_t cLenL = pLenL, cT = 0;
  uint64_t pc = first;
  idx_t pj = 0, plen = __shfl_sync(kFull, pLenL, 0), pptr = __shfl_sync(kFull, pPtrL, 0);
  auto produce = [&](int s) {
    while (pc < n && pj >= plen) {
      pc += stride;
      pj = 0;
      pT++;
      if ((pT & 31u) == 0) loadMeta(pT >> 5, pLenL, pPtrL);
      plen = __shfl_sync(kFull, pLenL, pT & 31u);
      pptr = __shfl_sync(kFull, pPtrL, pT & 31u);
    }
    if (pc >= n) return;
    const idx_t cols = min((idx_t)J, plen - pj);
    if (lane == 0) {
      const uint64_t off = (uint64_t)pptr + (uint64_t)pj * 32;
      unsigned char* dst = mine + (size_t)s * kStageBytes;
      mbarExpectTx(bar + s, (uint32_t)cols * (kValBytes + kColBytes));
      bulkLoad(dst, A.val + off, (uint32_t)cols * kValBytes, bar + s);
      bulkLoad(dst + J * kValBytes, A.col + off, (uint32_t)cols * kColBytes, bar + s);
    }
    pj += cols;
  };
#pragma unroll
  for (int s = 0; s < S; s++) produce(s);
  // up to here only the matrix has been touched: the rings are filling while the previous kernel drains
  griddepWait();

  int cs = 0;
  idx_t phases = 0;
  real_t dotAcc = 0.0;
  bool gatePassed = false;
  // LOCKSTEP: the CTA's warps advance one chunk each per step and meet at a barrier, so that they keep working on
  // 32*WARPS consecutive rows (one shared window of x in L1) instead of drifting apart
  const uint64_t nSteps = ctaFirst < n ? (n - ctaFirst + stride - 1) / stride : 0;
  for (uint64_t step = 0; step < nSteps; step++, cT++) {
    const uint64_t lchunk = first + step * stride;
    if (GATED) {
      // CTA-uniform: is some warp of this step past the interior chunks?
      if (!gatePassed && ctaFirst + step * stride + WARPS > (uint64_t)nInterior) {
        if (threadIdx.x == 0) gateWait(&gateSmem, gate.nsrc);
        __syncthreads();
        gatePassed = true;
      }
    }
    if (lchunk >= n) {                               // only in the CTA's last step
      if (LOCKSTEP) __syncthreads();
      continue;
    }
    if (cT != 0 && (cT & 31u) == 0) {
      idx_t unused;
      loadMeta(cT >> 5, cLenL, unused);
    }
    const idx_t len = __shfl_sync(kFull, cLenL, cT & 31u);
    const uint64_t row = phys(lchunk) * 32 + lane;
    real_t xr = 0.0;
    if (DOT && row < A.nr) xr = __ldg(x + row);
    real_t sum = 0.0;
    for (idx_t j0 = 0; j0 < len; j0 += J) {
      const idx_t cols = min((idx_t)J, len - j0);
      mbarWait(bar + cs, (phases >> cs) & 1u);
      phases ^= 1u << cs;
      const real_t* v = reinterpret_cast<const real_t*>(mine + (size_t)cs * kStageBytes) + lane;
      const idx_t* c = reinterpret_cast<const idx_t*>(mine + (size_t)cs * kStageBytes + J * kValBytes) + lane;
#pragma unroll
      for (idx_t j = 0; j < (idx_t)J; j += U) {
        if (j < cols) {                              // warp-uniform
          real_t xx[U], vv[U];
#pragma unroll
          for (int u = 0; u < U; u++)
            if (j + u < cols) xx[u] = gatherX<GATED>(x + c[(j + u) * 32]);
#pragma unroll
          for (int u = 0; u < U; u++)
            if (j + u < cols) vv[u] = v[(j + u) * 32];
#pragma unroll
          for (int u = 0; u < U; u++)
            if (j + u < cols) sum = mulAdd(sum, vv[u], xx[u]);
        }
      }
      __syncwarp();
      produce(cs);
      cs = (cs + 1 == S) ? 0 : cs + 1;
    }
    y[row] = sum;                                  // padded rows are stored too (matrix-SCS.c:224-226)
    if (DOT && row < A.nr) dotAcc = fma(sum, xr, dotAcc);
    if (LOCKSTEP) __syncthreads();
  }
  if (DOT) {
    const real_t b = blockSum(dotAcc, scratch);
    gridSum(b, partials, ticket, dotOut, accumulate, scratch, push.size ? &push : nullptr);
  }
}

// SELL-32 chunks longer than 256 columns (LongChunks, sb_internal.h): one CTA per chunk, its warps split the COLUMNS
// -- a lane still owns one row -- and the per-warp partial sums are added in warp order. Only these rows leave the
// reference's strict left-to-right order (matrix-SCS.c:213-222); what they gain: a 20 000-column chunk is streamed by
// eight warps with eight loads each in flight instead of by one warp through a ring sized for 27 columns.
constexpr int kLongWarps = 8;
template <bool DOT>
__global__ void __launch_bounds__(kLongWarps * 32, 4)
spmvSellLongChunksKernel(SellView A, const idx_t* __restrict__ list, uint32_t count, const real_t* __restrict__ x,
    real_t* __restrict__ y, real_t* partials, unsigned int* ticket, real_t* dotOut, bool accumulate)
{
  __shared__ real_t part[kLongWarps][32];
  __shared__ real_t scratch[32];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  real_t dotAcc = 0.0;
  for (uint32_t q = blockIdx.x; q < count; q += gridDim.x) {
    const idx_t ch = __ldg(list + q);
    const uint64_t len = __ldg(A.chunkLens + ch);
    const uint64_t base = (uint64_t)__ldg(A.chunkPtr + ch) + lane;
    const uint64_t j1 = len * (warp + 1) / kLongWarps;
    uint64_t j = len * warp / kLongWarps;
    real_t sum = 0.0;
    for (; j + 8 <= j1; j += 8) {
      idx_t cc[8];
      real_t vv[8], xx[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        cc[u] = ldStream(A.col + base + (j + u) * 32);
        vv[u] = ldStream(A.val + base + (j + u) * 32);
      }
#pragma unroll
      for (int u = 0; u < 8; u++) xx[u] = __ldg(x + cc[u]);
#pragma unroll
      for (int u = 0; u < 8; u++) sum = mulAdd(sum, vv[u], xx[u]);
    }
    for (; j < j1; j++) sum = mulAdd(sum, ldStream(A.val + base + j * 32), __ldg(x + ldStream(A.col + base + j * 32)));
    part[warp][lane] = sum;
    __syncthreads();
    if (warp == 0) {
      real_t t = part[0][lane];
#pragma unroll
      for (int w = 1; w < kLongWarps; w++) t = addRn(t, part[w][lane]);
      const uint64_t row = (uint64_t)ch * 32 + lane;
      y[row] = t;
      if (DOT && row < A.nr) dotAcc = fma(t, __ldg(x + row), dotAcc);
    }
    __syncthreads();
  }
  if (DOT) {
    const real_t b = blockSum(dotAcc, scratch);
    gridSum(b, partials, ticket, dotOut, accumulate, scratch);
  }
}

struct SellGate {                                     // nullptr-able extra arguments of a gated launch
  idx_t rot, nInterior;
  HaloGate gate;
};

template <int WARPS, int J, int S, int U, bool LOCKSTEP, bool GATED>
static void launchSell32TmaCfg(const SellView& A, const real_t* x, real_t* y, idx_t lo, idx_t hi, const DotArgs* dot,
    const SellGate* g, cudaStream_t s)
{
  Context& c = ctx();
  const size_t smem = (size_t)WARPS * S * J * 32 * (sizeof(real_t) + sizeof(idx_t));
  static bool configured = false;
  if (!configured) {
    allowLargeSmem(spmvSell32TmaKernel<true, WARPS, J, S, U, LOCKSTEP, GATED>, smem);
    allowLargeSmem(spmvSell32TmaKernel<false, WARPS, J, S, U, LOCKSTEP, GATED>, smem);
    configured = true;
  }
  uint64_t blocks = ((uint64_t)(hi - lo) + WARPS - 1) / WARPS;
  if (blocks > (uint64_t)c.numSMs) blocks = c.numSMs;
  const idx_t rot = g ? g->rot : 0u, nInt = g ? g->nInterior : 0u;
  const HaloGate gate = g ? g->gate : HaloGate();
  if (dot)
    launchPdl(spmvSell32TmaKernel<true, WARPS, J, S, U, LOCKSTEP, GATED>, dim3((unsigned)blocks), dim3(WARPS * 32), smem, s, A, x, y, lo,
        hi, c.partials + (size_t)dot->slot * kMaxPartials, c.tickets + dot->slot, dot->out, dot->accumulate, rot, nInt, gate,
        dot->push ? *dot->push : PeerReduce());
  else
    launchPdl(spmvSell32TmaKernel<false, WARPS, J, S, U, LOCKSTEP, GATED>, dim3((unsigned)blocks), dim3(WARPS * 32), smem, s, A, x, y, lo,
        hi, (real_t*)nullptr, (unsigned int*)nullptr, (real_t*)nullptr, false, rot, nInt, gate, PeerReduce());
  countLaunch();
}

static void launchSell32Tma(const SellView& A, const real_t* x, real_t* y, idx_t lo, idx_t hi, const DotArgs* dot,
    const SellGate* g, cudaStream_t s)
{
  if (g) {
    launchSell32TmaCfg<32, 4, 3, 4, true, true>(A, x, y, lo, hi, dot, g, s);
    return;
  }
  static const int cfg = envInt("SB_SELL_CFG", 0);   // tuning knob, measured in profiles/
  switch (cfg) {
  case 1: launchSell32TmaCfg<32, 4, 3, 4, false, false>(A, x, y, lo, hi, dot, nullptr, s); break;   // free-running warps
  case 2: launchSell32TmaCfg<32, 4, 4, 4, true, false>(A, x, y, lo, hi, dot, nullptr, s); break;
  case 3: launchSell32TmaCfg<24, 8, 3, 8, true, false>(A, x, y, lo, hi, dot, nullptr, s); break;
  case 4: launchSell32TmaCfg<16, 16, 2, 16, true, false>(A, x, y, lo, hi, dot, nullptr, s); break;
  case 5: launchSell32TmaCfg<8, 32, 2, 8, true, false>(A, x, y, lo, hi, dot, nullptr, s); break;
  default: launchSell32TmaCfg<32, 4, 3, 4, true, false>(A, x, y, lo, hi, dot, nullptr, s); break;
  }
}

// Register-staged SELL-32 kernel (baseline): loads through LDG with an 8-deep unroll.
constexpr int kSellUnroll = 8;

template <bool DOT>
__global__ void __launch_bounds__(256, 4)
spmvSell32Kernel(SellView A, const real_t* __restrict__ x, real_t* __restrict__ y, idx_t lo, idx_t hi,
    real_t* partials, unsigned int* ticket, real_t* dotOut, bool accumulate)
{
  __shared__ real_t scratch[32];
  const int lane = threadIdx.x & 31;
  const idx_t warpsPerBlock = blockDim.x >> 5;
  real_t dotAcc = 0.0;
  for (uint64_t chunk = (uint64_t)lo + blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); chunk < hi;
       chunk += (uint64_t)gridDim.x * warpsPerBlock) {
    const uint64_t base = (uint64_t)A.chunkPtr[chunk] + lane;
    const idx_t len = A.chunkLens[chunk];
    const real_t* __restrict__ v = A.val + base;
    const idx_t* __restrict__ c = A.col + base;
    real_t sum = 0.0;
    idx_t j = 0;
    for (; j + kSellUnroll <= len; j += kSellUnroll) {
      idx_t cc[kSellUnroll];
      real_t vv[kSellUnroll], xx[kSellUnroll];
#pragma unroll
      for (int u = 0; u < kSellUnroll; u++) {
        cc[u] = ldStream(c + (uint64_t)(j + u) * 32);
        vv[u] = ldStream(v + (uint64_t)(j + u) * 32);
      }
#pragma unroll
      for (int u = 0; u < kSellUnroll; u++) xx[u] = __ldg(x + cc[u]);
#pragma unroll
      for (int u = 0; u < kSellUnroll; u++) sum = mulAdd(sum, vv[u], xx[u]);
    }
    if (j < len) {   // tail of the chunk: same batch, predicated (len is warp-uniform)
      idx_t cc[kSellUnroll];
      real_t vv[kSellUnroll], xx[kSellUnroll];
#pragma unroll
      for (int u = 0; u < kSellUnroll; u++)
        if (j + u < len) {
          cc[u] = ldStream(c + (uint64_t)(j + u) * 32);
          vv[u] = ldStream(v + (uint64_t)(j + u) * 32);
        }
#pragma unroll
      for (int u = 0; u < kSellUnroll; u++)
        if (j + u < len) xx[u] = __ldg(x + cc[u]);
#pragma unroll
      for (int u = 0; u < kSellUnroll; u++)
        if (j + u < len) sum = mulAdd(sum, vv[u], xx[u]);
    }
    const uint64_t row = chunk * 32 + lane;
    y[row] = sum;
    if (DOT && row < A.nr) dotAcc = fma(sum, __ldg(x + row), dotAcc);
  }
  if (DOT) {
    const real_t b = blockSum(dotAcc, scratch);
    gridSum(b, partials, ticket, dotOut, accumulate, scratch);
  }
}

// Any other chunk height (the reference's tests use C = 1, 2, 4): one thread per padded row, same
// per-row summation order. Correctness path, not tuned.
template <bool DOT>
__global__ void __launch_bounds__(256)
spmvSellAnyCKernel(SellView A, const real_t* __restrict__ x, real_t* __restrict__ y, idx_t lo, idx_t hi,
    real_t* partials, unsigned int* ticket, real_t* dotOut, bool accumulate)
{
  __shared__ real_t scratch[32];
  real_t dotAcc = 0.0;
  const uint64_t rowLo = (uint64_t)lo * A.C, rowHi = (uint64_t)hi * A.C;
  for (uint64_t row = rowLo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; row < rowHi;
       row += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t chunk = row / A.C;
    const uint64_t base = (uint64_t)A.chunkPtr[chunk] + row % A.C;
    const idx_t len = A.chunkLens[chunk];
    real_t sum = 0.0;
    for (idx_t j = 0; j < len; j++) {
      const uint64_t e = base + (uint64_t)j * A.C;
      sum = mulAdd(sum, A.val[e], __ldg(x + A.col[e]));
    }
    y[row] = sum;
    if (DOT && row < A.nr) dotAcc = fma(sum, __ldg(x + row), dotAcc);
  }
  if (DOT) {
    const real_t b = blockSum(dotAcc, scratch);
    gridSum(b, partials, ticket, dotOut, accumulate, scratch);
  }
}

// ------------------------------------------------------------------------------------------- CRS / CCRS
// Warp-specialised CTA pipeline ("CSR-stream"): the non-zeros of R consecutive rows are one contiguous range
// of the arrays, so one producer thread bulk-copies that range -- plus the R+1 row pointers -- into a ring of S
// shared-memory stages (full/empty mbarriers), and W consumer warps work the tile off: LPR consecutive lanes
// share a row (4 lanes x 7 strided elements for the 27-point rows), gather x, and combine their partial sums
// with xor-shuffles. The global stream is perfectly coalesced whatever the row lengths are, consumers never
// touch rowPtr/col/val in global memory, and the number of consumer warps is independent of the bytes in
// flight. Tiles whose non-zeros do not fit a stage are read straight from global memory by the same lanes.
constexpr uint32_t kPipeMaxRows = 768;               // row pointers per stage
// fused dot: the tile's own x entries are staged behind the ring, one slot of kPipeXBytes per stage
constexpr uint32_t kPipeXBytes = (kPipeMaxRows + 8) * sizeof(real_t);
constexpr uint64_t kXAlign = 16 / sizeof(real_t);     // bulk copies move multiples of 16 bytes from 16-byte aligned addresses

template <int WARPS, uint32_t CAP, uint32_t STAGES>
struct CrsPipe {                                        // stage: val[CAP+8] | col[CAP+8] | rowPtr[kPipeMaxRows+8]
  static constexpr int kWarps = WARPS;                  // consumer warps (+1 producer warp)
  static constexpr uint32_t kCap = CAP, kStages = STAGES;
  static constexpr uint32_t kColOff = (CAP + 8) * sizeof(real_t), kRpOff = kColOff + (CAP + 8) * sizeof(idx_t),
                            kBytes = kRpOff + (kPipeMaxRows + 8) * sizeof(idx_t);
  const idx_t* col;
  const real_t* val;
  // elements [s, e) -> shared; bulk copies need 16-byte granularity, so the range is widened to multiples of 4
  __device__ __forceinline__ static uint64_t origin(uint64_t s) { return s & ~3ull; }
  __device__ __forceinline__ static uint32_t bytes(uint64_t s, uint64_t e)
  {
    return (uint32_t)(((e + 3) & ~3ull) - (s & ~3ull)) * (uint32_t)(sizeof(real_t) + sizeof(idx_t));
  }
  __device__ __forceinline__ void request(unsigned char* dst, uint64_t s, uint64_t e, uint64_t* bar) const
  {
    const uint64_t a = s & ~3ull;
    const uint32_t n = (uint32_t)(((e + 3) & ~3ull) - a);
    bulkLoad(dst, val + a, n * (uint32_t)sizeof(real_t), bar);
    bulkLoad(dst + kColOff, col + a, n * (uint32_t)sizeof(idx_t), bar);
  }
  static constexpr bool kSplitFetch = true;             // column ids first: they die once their gather is issued
  __device__ __forceinline__ static void fetch(const unsigned char* st, uint32_t i, idx_t& c, real_t& v)
  {
    v = reinterpret_cast<const real_t*>(st)[i];
    c = reinterpret_cast<const idx_t*>(st + kColOff)[i];
  }
  __device__ __forceinline__ static idx_t fetchCol(const unsigned char* st, uint32_t i)
  {
    return reinterpret_cast<const idx_t*>(st + kColOff)[i];
  }
  __device__ __forceinline__ static real_t fetchVal(const unsigned char* st, uint32_t i)
  {
    return reinterpret_cast<const real_t*>(st)[i];
  }
  __device__ __forceinline__ void fetchGlobal(uint64_t j, idx_t& c, real_t& v) const
  {
    c = ldStream(col + j);
    v = ldStream(val + j);
  }
};

template <int WARPS, uint32_t CAP, uint32_t STAGES>
struct CcrsPipe {                                       // stage: {col, (pad,) val}[CAP] | rowPtr[kPipeMaxRows+8]
  static constexpr int kWarps = WARPS;
  static constexpr uint32_t kCap = CAP, kStages = STAGES;
  static constexpr uint32_t kRec = sizeof(Entry);       // 16 bytes (8 for float values with 32-bit ids)
  static constexpr uint64_t kAlign = 16 / kRec;         // records per 16 bytes: granularity of the bulk copies
  static constexpr uint32_t kRpOff = CAP * kRec, kBytes = kRpOff + (kPipeMaxRows + 8) * sizeof(idx_t);
  const Entry* entries;
  __device__ __forceinline__ static uint64_t origin(uint64_t s) { return s & ~(kAlign - 1); }
  __device__ __forceinline__ static uint32_t bytes(uint64_t s, uint64_t e)
  {
    return (uint32_t)(((e + kAlign - 1) & ~(kAlign - 1)) - origin(s)) * kRec;
  }
  __device__ __forceinline__ void request(unsigned char* dst, uint64_t s, uint64_t e, uint64_t* bar) const
  {
    bulkLoad(dst, entries + origin(s), bytes(s, e), bar);
  }
  static constexpr bool kSplitFetch = false;            // one LDS.128 per record beats two narrower reads
  __device__ __forceinline__ static Entry record(const unsigned char* st, uint32_t i)
  {
    Entry e;
    if constexpr (kRec == 16) {
      const uint4 r = reinterpret_cast<const uint4*>(st)[i];   // one 16-byte {col, pad, val} record
      memcpy(&e, &r, 16);
    } else {
      const uint2 r = reinterpret_cast<const uint2*>(st)[i];
      memcpy(&e, &r, 8);
    }
    return e;
  }
  __device__ __forceinline__ static void fetch(const unsigned char* st, uint32_t i, idx_t& c, real_t& v)
  {
    if constexpr (sizeof(real_t) == 8 && sizeof(idx_t) == 4) {
      // the default build reads the record as two doubles: the same LDS.128, but nvcc then interleaves the record reads
      // with the gathers they feed; through the byte-copy of record() it groups them and the kernel is 8 % slower
      const double2 e = reinterpret_cast<const double2*>(st)[i];
      c = (idx_t)__double_as_longlong(e.x);
      v = (real_t)e.y;
    } else {
      const Entry e = record(st, i);
      c = e.col;
      v = e.val;
    }
  }
  __device__ __forceinline__ static idx_t fetchCol(const unsigned char* st, uint32_t i) { return reinterpret_cast<const Entry*>(st)[i].col; }
  __device__ __forceinline__ static real_t fetchVal(const unsigned char* st, uint32_t i) { return reinterpret_cast<const Entry*>(st)[i].val; }
  __device__ __forceinline__ void fetchGlobal(uint64_t j, idx_t& c, real_t& v) const
  {
    const Entry e = ldStreamEntry(entries + j);
    c = e.col;
    v = e.val;
  }
};

// GATED launches cover all rows in one go as three runs of tiles: the interior rows [intLo, intHi) first, then the
// rows above and below, which reference halo columns; a consumer warp waits on the gate before its first tile of
// those (x gathers of a gated kernel are coherent loads: the peers store the halo while this kernel is running).
// VAR (bit mask; measured on the 27-point stencil, profiles/README.md):
//   1  the row's own x for the fused dot is requested BEFORE the gathers of the pass (no gain: not latency)
//   2  a row's lanes start at the row start rounded up to LPR elements -- their shared-memory reads then fall into one
//      aligned LPR*8-byte window per step, rows of a half-warp into different bank groups -- and the 0..LPR-1 leading
//      elements take one extra predicated slot (removes the bank conflicts, but no faster: not the limiter)
//   4  7 instead of 8 elements per lane and batch: lanes per row = avg/7, so one batch is a whole row (-1..2 %)
//   8  fused dot: the tile's own x entries arrive with the tile through the bulk-copy engine (one more 16-byte
//      granular copy per tile into a slot behind the ring) instead of through one more global load per row:
//      fused-dot cost +2.0 % -> +1.0 % at 256^3, +3.3 % -> +1.7 % at 128^3
// Defaults (Access::kVar): CRS 12; CCRS 4 (its 16-byte record fetch leaves no registers for the x slots: fused dot +15 % with 8).
template <bool DOT, int LPR, typename L, bool GATED, int VAR>
__global__ void __launch_bounds__((L::kWarps + 1) * 32, 1)
spmvRowsPipeKernel(L acc, const idx_t* __restrict__ rowPtr, const real_t* __restrict__ x, real_t* __restrict__ y,
    idx_t lo, idx_t hi, idx_t tileRows, real_t* partials, unsigned int* ticket, real_t* dotOut, bool accumulate,
    idx_t intLo, idx_t intHi, HaloGate gate, PeerReduce push)
{
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ real_t scratch[32];
  __shared__ uint64_t fullBar[L::kStages], emptyBar[L::kStages];
  __shared__ GateSmem gateSmem;
  constexpr idx_t S = L::kStages;
  constexpr int kPipeWarps = L::kWarps;
  constexpr bool EARLYX = (VAR & 1) != 0, ALIGN = (VAR & 2) != 0, XSMEM = (VAR & 8) != 0;
  constexpr int UN = (VAR & 4) ? 7 : 8;                  // slots per lane per batch
  constexpr int HEAD = ALIGN ? 1 : 0;                    // slot 0 of a row's first batch: its unaligned leading elements
  constexpr int GPW = 32 / LPR;                          // rows per warp per pass
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  griddepLaunchDependents();
  if (threadIdx.x == 0) {
    if (GATED) gateStore(gateSmem, gate);
    for (idx_t s = 0; s < S; s++) {
      mbarInit(fullBar + s, 1);
      mbarInit(emptyBar + s, kPipeWarps);
    }
    mbarFenceInit();
  }
  __syncthreads();

  // tile t -> rows [r0, r1); returns whether the tile lies behind the gate
  const uint64_t R = tileRows;
  const uint64_t s0 = GATED ? intLo : lo, e0 = GATED ? intHi : hi;
  const uint64_t nt0 = (e0 - s0 + R - 1) / R;
  const uint64_t nt1 = GATED ? ((uint64_t)hi - intHi + R - 1) / R : 0;
  const uint64_t nt2 = GATED ? ((uint64_t)intLo - lo + R - 1) / R : 0;
  const uint64_t nTiles = nt0 + nt1 + nt2;
  auto tileRange = [&](uint64_t t, uint64_t& r0, uint64_t& r1) {
    uint64_t base = s0, end = e0;
    bool gated = false;
    if (GATED && t >= nt0) {
      gated = true;
      t -= nt0;
      if (t < nt1) { base = intHi; end = hi; }
      else { t -= nt1; base = lo; end = intLo; }
    }
    r0 = base + t * R;
    r1 = r0 + R < end ? r0 + R : end;
    return gated;
  };

  real_t dotAcc = 0.0;
  if (warp == kPipeWarps) {
    // ---- producer: one thread keeps the ring full. The first S tiles are requested before griddepWait(): only the
    // matrix is touched, so the ring fills while the previous kernel drains; their x entries (fused dot) follow it.
    if (lane == 0) {
      idx_t i = 0;
      bool waited = false;
      for (uint64_t t = blockIdx.x; t < nTiles; t += gridDim.x, i++) {
        const idx_t s = i % S, k = i / S;
        if (k > 0 && !waited) {
          griddepWait();
          waited = true;
          if (DOT && XSMEM) {                                  // x entries of the tiles requested so far
            idx_t j = 0;
            for (uint64_t t2 = blockIdx.x; j < S && t2 < nTiles; t2 += gridDim.x, j++) {
              uint64_t q0, q1;
              tileRange(t2, q0, q1);
              const uint64_t ax = q0 & ~(kXAlign - 1);
              bulkLoadKeep(ring + (size_t)S * L::kBytes + (size_t)j * kPipeXBytes, x + ax,
                  (uint32_t)((q1 - ax + kXAlign - 1) & ~(kXAlign - 1)) * (uint32_t)sizeof(real_t), fullBar + j);
            }
          }
        }
        uint64_t r0, r1;
        tileRange(t, r0, r1);
        const uint64_t bs = __ldg(rowPtr + r0), be = __ldg(rowPtr + r1);
        const uint64_t a = r0 & ~3ull;
        const uint32_t nrp = (uint32_t)((r1 + 1 - a + 3) & ~3ull);
        const bool fits = be > bs && be - L::origin(bs) <= L::kCap;
        if (k > 0) mbarWait(emptyBar + s, (k - 1) & 1u);
        unsigned char* dst = ring + (size_t)s * L::kBytes;
        // fused dot: the tile's own x entries ride along (16-byte granularity: from the aligned row at or below r0)
        const uint64_t ax = r0 & ~(kXAlign - 1);
        const uint32_t nxr = DOT && XSMEM ? (uint32_t)((r1 - ax + kXAlign - 1) & ~(kXAlign - 1)) : 0u;
        mbarExpectTx(fullBar + s, nrp * (uint32_t)sizeof(idx_t) + nxr * (uint32_t)sizeof(real_t) + (fits ? L::bytes(bs, be) : 0u));
        bulkLoad(dst + L::kRpOff, rowPtr + a, nrp * (uint32_t)sizeof(idx_t), fullBar + s);
        if (DOT && XSMEM && waited)
          bulkLoadKeep(ring + (size_t)S * L::kBytes + (size_t)s * kPipeXBytes, x + ax, nxr * (uint32_t)sizeof(real_t), fullBar + s);
        if (fits) acc.request(dst, bs, be, fullBar + s);
      }
      if (!waited) {                                           // fewer than S + 1 tiles: nothing was waited for yet
        griddepWait();
        if (DOT && XSMEM) {
          idx_t j = 0;
          for (uint64_t t2 = blockIdx.x; j < S && t2 < nTiles; t2 += gridDim.x, j++) {
            uint64_t q0, q1;
            tileRange(t2, q0, q1);
            const uint64_t ax = q0 & ~(kXAlign - 1);
            bulkLoadKeep(ring + (size_t)S * L::kBytes + (size_t)j * kPipeXBytes, x + ax,
                (uint32_t)((q1 - ax + kXAlign - 1) & ~(kXAlign - 1)) * (uint32_t)sizeof(real_t), fullBar + j);
          }
        }
      }
    }
  } else {
    // ---- consumers
    griddepWait();                                             // x (and y) belong to the previous kernels
    const int sub = lane % LPR, grp = lane / LPR;
    bool gatePassed = false;
    idx_t i = 0;
    for (uint64_t t = blockIdx.x; t < nTiles; t += gridDim.x, i++) {
      const idx_t s = i % S, k = i / S;
      uint64_t r0, r1;
      const bool behindGate = tileRange(t, r0, r1);
      const idx_t nrows = (idx_t)(r1 - r0);
      if (GATED && behindGate && !gatePassed) {
        if (lane == 0) gateWait(&gateSmem, gate.nsrc);
        __syncwarp();
        gatePassed = true;
      }
      mbarWait(fullBar + s, k & 1u);
      const unsigned char* st = ring + (size_t)s * L::kBytes;
      const idx_t* rp = reinterpret_cast<const idx_t*>(st + L::kRpOff) + (uint32_t)(r0 & 3ull);
      const real_t* xrow = reinterpret_cast<const real_t*>(ring + (size_t)S * L::kBytes + (size_t)s * kPipeXBytes) + (uint32_t)(r0 & (kXAlign - 1));   // x[r0 + i] (DOT && XSMEM only)
      const uint64_t bs = rp[0], be = rp[nrows];
      const uint64_t org = L::origin(bs);
      const bool fits = be > bs && be - org <= L::kCap;
      for (idx_t g0 = warp * GPW; g0 < nrows; g0 += kPipeWarps * GPW) {
        const idx_t g = g0 + grp;
        const bool live = g < nrows;
        const idx_t rs = live ? rp[g] : 0u, re = live ? rp[g + 1] : 0u;
        real_t sum = 0.0, xr = 0.0;
        if (DOT && EARLYX && live && sub == 0) xr = gatherX<GATED>(x + r0 + g);   // in flight together with the gathers below
        if (fits) {
          // lanes without a row get an empty range (0 - org would wrap and alias real entries)
          const idx_t first = live ? rs - (idx_t)org : 0u;
          const idx_t end = live ? re - (idx_t)org : 0u;
          const idx_t start = ALIGN ? min((first + (idx_t)LPR - 1u) & ~((idx_t)LPR - 1u), end) : first;
          idx_t idx = start + sub;
          bool head = ALIGN && first + sub < start;     // this lane owns one of the leading elements
          do {
            real_t vv[UN], xx[UN];
            if constexpr (L::kSplitFetch) {
              if (HEAD && head) xx[0] = gatherX<GATED>(x + L::fetchCol(st, first + sub));
#pragma unroll
              for (int u = HEAD; u < UN; u++)
                if (idx + (u - HEAD) * LPR < end) xx[u] = gatherX<GATED>(x + L::fetchCol(st, idx + (u - HEAD) * LPR));
              if (HEAD && head) vv[0] = L::fetchVal(st, first + sub);
#pragma unroll
              for (int u = HEAD; u < UN; u++)
                if (idx + (u - HEAD) * LPR < end) vv[u] = L::fetchVal(st, idx + (u - HEAD) * LPR);
            } else {
              idx_t cc[UN];
              if (HEAD && head) L::fetch(st, first + sub, cc[0], vv[0]);
#pragma unroll
              for (int u = HEAD; u < UN; u++)
                if (idx + (u - HEAD) * LPR < end) L::fetch(st, idx + (u - HEAD) * LPR, cc[u], vv[u]);
              if (HEAD && head) xx[0] = gatherX<GATED>(x + cc[0]);
#pragma unroll
              for (int u = HEAD; u < UN; u++)
                if (idx + (u - HEAD) * LPR < end) xx[u] = gatherX<GATED>(x + cc[u]);
            }
            if (HEAD && head) sum = mulAdd(sum, vv[0], xx[0]);
#pragma unroll
            for (int u = HEAD; u < UN; u++)
              if (idx + (u - HEAD) * LPR < end) sum = mulAdd(sum, vv[u], xx[u]);
            idx += (UN - HEAD) * LPR;
            head = false;
          } while (__any_sync(0xffffffffu, idx < end));
        } else {
          for (uint64_t j = (uint64_t)rs + sub; j < re; j += LPR) {   // oversized tile: straight from global memory
            idx_t c;
            real_t v;
            acc.fetchGlobal(j, c, v);
            sum = mulAdd(sum, v, gatherX<GATED>(x + c));
          }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (live && sub == 0) {
          y[r0 + g] = sum;
          if (DOT) dotAcc = fma(sum, XSMEM ? xrow[g] : EARLYX ? xr : gatherX<GATED>(x + r0 + g), dotAcc);
        }
      }
      __syncwarp();
      if (lane == 0) mbarArrive(emptyBar + s);
    }
  }
  if (DOT) {
    const real_t b = blockSum(dotAcc, scratch);
    gridSum(b, partials, ticket, dotOut, accumulate, scratch, push.size ? &push : nullptr);
  }
}

struct RowsGate {                                     // extra arguments of a gated launch
  idx_t intLo, intHi;
  HaloGate gate;
};

template <int LPR, typename L, bool GATED, int VAR>
static void launchRowsPipe(L acc, const idx_t* rowPtr, const real_t* x, real_t* y, idx_t lo, idx_t hi,
    idx_t tileRows, const DotArgs* dot, const RowsGate* g, cudaStream_t s)
{
  Context& c = ctx();
  const size_t smemPlain = (size_t)L::kStages * L::kBytes;
  const size_t smemDot = smemPlain + ((VAR & 8) ? (size_t)L::kStages * kPipeXBytes : 0);
  const size_t smem = dot ? smemDot : smemPlain;
  static bool configured = false;
  if (!configured) {
    allowLargeSmem(spmvRowsPipeKernel<true, LPR, L, GATED, VAR>, smemDot);
    allowLargeSmem(spmvRowsPipeKernel<false, LPR, L, GATED, VAR>, smemPlain);
    configured = true;
  }
  uint64_t blocks = ((uint64_t)(hi - lo) + tileRows - 1) / tileRows + (GATED ? 2 : 0);
  if (blocks > (uint64_t)c.numSMs) blocks = c.numSMs;
  const int threads = (L::kWarps + 1) * 32;
  const idx_t intLo = g ? g->intLo : 0u, intHi = g ? g->intHi : 0u;
  const HaloGate gate = g ? g->gate : HaloGate();
  if (dot)
    launchPdl(spmvRowsPipeKernel<true, LPR, L, GATED, VAR>, dim3((unsigned)blocks), dim3(threads), smem, s, acc, rowPtr, x, y, lo, hi,
        tileRows, c.partials + (size_t)dot->slot * kMaxPartials, c.tickets + dot->slot, dot->out, dot->accumulate, intLo, intHi, gate,
        dot->push ? *dot->push : PeerReduce());
  else
    launchPdl(spmvRowsPipeKernel<false, LPR, L, GATED, VAR>, dim3((unsigned)blocks), dim3(threads), smem, s, acc, rowPtr, x, y, lo, hi,
        tileRows, (real_t*)nullptr, (unsigned int*)nullptr, (real_t*)nullptr, false, intLo, intHi, gate, PeerReduce());
  countLaunch();
}

// Register-staged sub-warp kernel (long rows, and the measured baseline): LANES consecutive lanes share a row,
// partial sums are combined with xor-shuffles.
struct CrsAccess {
  const idx_t* col;
  const real_t* val;
  template <int W, uint32_t CAP, uint32_t S> using Pipe = CrsPipe<W, CAP, S>;
  template <typename P> P pipe() const { return P { col, val }; }
  // stages that fit 227 KB next to the x slots of the fused dot: 12 bytes per non-zero (double + u32) -> 4 x 46 KB / 3 x 66 KB
  static constexpr bool kWide = sizeof(real_t) + sizeof(idx_t) > 12;     // 64-bit indices with double values: 16 bytes per non-zero
  static constexpr uint32_t kStagesFor3584 = kWide ? 3 : 4, kStagesFor5376 = 3;
  static constexpr int kDefaultCfg = kWide ? 1 : 0;     // 0: 23 consumer warps (+1 producer = 6 warps per scheduler, 80 registers), 3 stages of 66 KB
  static constexpr int kVar = 12;                       // spmvRowsPipeKernel VAR
  __device__ __forceinline__ void load(uint64_t j, idx_t& c, real_t& v) const
  {
    c = ldStream(col + j);
    v = ldStream(val + j);
  }
};
struct CcrsAccess {
  const Entry* entries;
  template <int W, uint32_t CAP, uint32_t S> using Pipe = CcrsPipe<W, CAP, S>;
  template <typename P> P pipe() const { return P { entries }; }
  static constexpr bool kWide = false;
  static constexpr uint32_t kStagesFor3584 = 3, kStagesFor5376 = 2;
  static constexpr int kDefaultCfg = 1;                 // 16 consumer warps, 3 stages of 59 KB (16-byte records)
  static constexpr int kVar = 4;
  __device__ __forceinline__ void load(uint64_t j, idx_t& c, real_t& v) const
  {
    const Entry e = ldStreamEntry(entries + j);       // one record
    c = e.col;
    v = e.val;
  }
};

template <int LANES, bool DOT, typename Access>
__global__ void __launch_bounds__(256, 4)
spmvRowsKernel(Access acc, const idx_t* __restrict__ rowPtr, idx_t nrTotal, const real_t* __restrict__ x,
    real_t* __restrict__ y, idx_t lo, idx_t hi, real_t* partials, unsigned int* ticket, real_t* dotOut,
    bool accumulate)
{
  __shared__ real_t scratch[32];
  constexpr int kUnroll = 4;
  const int sub = threadIdx.x % LANES;
  const idx_t groupsPerBlock = blockDim.x / LANES;
  real_t dotAcc = 0.0;
  for (uint64_t first = (uint64_t)lo + (uint64_t)blockIdx.x * groupsPerBlock; first < hi;
       first += (uint64_t)gridDim.x * groupsPerBlock) {
    const uint64_t row = first + threadIdx.x / LANES;
    const bool live = row < hi;
    uint64_t j = 0, end = 0;
    if (live) {
      j = (uint64_t)__ldg(rowPtr + row) + sub;
      end = __ldg(rowPtr + row + 1);
    }
    real_t sum = 0.0;
    while (j < end) {
      idx_t cc[kUnroll];
      real_t vv[kUnroll], xx[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; u++)
        if (j + (uint64_t)u * LANES < end) acc.load(j + (uint64_t)u * LANES, cc[u], vv[u]);
#pragma unroll
      for (int u = 0; u < kUnroll; u++)
        if (j + (uint64_t)u * LANES < end) xx[u] = __ldg(x + cc[u]);
#pragma unroll
      for (int u = 0; u < kUnroll; u++)
        if (j + (uint64_t)u * LANES < end) sum = mulAdd(sum, vv[u], xx[u]);
      j += (uint64_t)kUnroll * LANES;
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (live && sub == 0) {
      y[row] = sum;
      if (DOT) dotAcc = fma(sum, __ldg(x + row), dotAcc);
    }
  }
  if (DOT) {
    const real_t b = blockSum(dotAcc, scratch);
    gridSum(b, partials, ticket, dotOut, accumulate, scratch);
  }
}

template <int LANES, typename Access>
static void launchRows(Access acc, const idx_t* rowPtr, idx_t nr, const real_t* x, real_t* y, idx_t lo,
    idx_t hi, const DotArgs* dot, cudaStream_t s)
{
  Context& c = ctx();
  const idx_t groupsPerBlock = 256 / LANES;
  uint64_t blocks = ((uint64_t)(hi - lo) + groupsPerBlock - 1) / groupsPerBlock;
  const uint64_t cap = (uint64_t)c.numSMs * 8;
  if (blocks > cap) blocks = cap;
  if (blocks > (uint64_t)kMaxPartials) blocks = kMaxPartials;
  if (dot)
    spmvRowsKernel<LANES, true, Access><<<(int)blocks, 256, 0, s>>>(acc, rowPtr, nr, x, y, lo, hi,
        c.partials + (size_t)dot->slot * kMaxPartials, c.tickets + dot->slot, dot->out, dot->accumulate);
  else
    spmvRowsKernel<LANES, false, Access><<<(int)blocks, 256, 0, s>>>(acc, rowPtr, nr, x, y, lo, hi, nullptr, nullptr,
        nullptr, false);
  SB_CUDA(cudaGetLastError());
  countLaunch();
}

// ---- CRS / CCRS with uneven row lengths ----------------------------------------------------------------------------
// The pipelined kernel above gives every row the same number of lanes and walks them in lockstep: right for the
// stencil (27 +- 0), wasteful once the lengths vary (5..45: 0.24 ms where this kernel needs 0.19) and pathological
// with a heavy tail (one 20 000-entry row on 4 lanes: 3.3 ms for a 0.5 GB matrix, 0.29 here). Matrices whose longest
// row exceeds 1.2 avg + 4 take this kernel instead (one rule for CRS and CCRS, so the two stay bit-identical): the
// rows are cut into blocks of at most kStreamNnz non-zeros and kStreamMaxRows rows (RowBlocks, built once per matrix);
// a CTA
//   A  forms ALL products val * x[col] of its block with one thread per non-zero -- coalesced matrix reads, perfectly
//      balanced whatever the row lengths -- into shared memory,
//   B  adds every row's products left to right with one thread per row (the reference's own order,
//      matrix-CRS.c:58-61),
//   C  gives rows longer than kStreamSeqLen a warp; a row longer than a whole block gets the whole CTA straight
//      from global memory.
// Deterministic: every sum has one fixed order. (A variant that staged the blocks through the bulk-copy engine was
// slower, 0.27 ms on the 5..45 case: its 4 x 50 KB of stages leave the x gathers 28 KB of L1.)
constexpr int kStreamThreads = 256, kStreamPerThread = 8;
constexpr uint32_t kStreamNnz = kStreamThreads * kStreamPerThread;
constexpr uint32_t kStreamSeqLen = 64;
constexpr uint32_t kStreamMaxRows = 2048;

template <bool DOT, typename Access, int ROUND>
__global__ void __launch_bounds__(kStreamThreads, ROUND == 8 ? 4 : 5)
spmvRowsStreamKernel(Access acc, const idx_t* __restrict__ rowPtr, const idx_t* __restrict__ blockStart, uint32_t nBlocks,
    const real_t* __restrict__ x, real_t* __restrict__ y, real_t* partials, unsigned int* ticket, real_t* dotOut,
    bool accumulate, PeerReduce push)
{
  __shared__ real_t prod[kStreamNnz];
  __shared__ real_t scratch[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  real_t dotAcc = 0.0;
  // geometry of the next block is loaded one block ahead: blockStart -> rowPtr is a dependent pair of misses
  idx_t nr0 = 0, nr1 = 0, nbs = 0, nbe = 0;
  if (blockIdx.x < nBlocks) {
    nr0 = __ldg(blockStart + blockIdx.x);
    nr1 = __ldg(blockStart + blockIdx.x + 1);
    nbs = __ldg(rowPtr + nr0);
    nbe = __ldg(rowPtr + nr1);
  }
  for (uint32_t b = blockIdx.x; b < nBlocks; b += gridDim.x) {
    const idx_t r0 = nr0, r1 = nr1, bs = nbs, be = nbe;
    if ((uint64_t)b + gridDim.x < nBlocks) {
      nr0 = __ldg(blockStart + b + gridDim.x);
      nr1 = __ldg(blockStart + b + gridDim.x + 1);
      nbs = __ldg(rowPtr + nr0);
      nbe = __ldg(rowPtr + nr1);
    }
    const uint64_t nnz = (uint64_t)be - bs;
    if (nnz <= kStreamNnz) {
      const uint32_t n32 = (uint32_t)nnz;
      // A: products (slots past the end re-read the block's first element: defined values, nothing stored)
#pragma unroll 1
      for (uint32_t i0 = tid; i0 < n32; i0 += ROUND * kStreamThreads) {
        idx_t cc[ROUND];
        real_t vv[ROUND], xx[ROUND];
#pragma unroll
        for (int u = 0; u < ROUND; u++) acc.load((uint64_t)bs + (i0 + u * kStreamThreads < n32 ? i0 + u * kStreamThreads : 0u), cc[u], vv[u]);
#pragma unroll
        for (int u = 0; u < ROUND; u++) xx[u] = __ldg(x + cc[u]);
#pragma unroll
        for (int u = 0; u < ROUND; u++)
          if (i0 + u * kStreamThreads < n32) prod[i0 + u * kStreamThreads] = mulRn(vv[u], xx[u]);
      }
      __syncthreads();
      // B: one thread per row, left to right
      const uint32_t nrows = (uint32_t)(r1 - r0);
      int longSeen = 0;
      for (uint32_t r = tid; r < nrows; r += kStreamThreads) {
        const uint32_t a = (uint32_t)(__ldg(rowPtr + r0 + r) - bs), e = (uint32_t)(__ldg(rowPtr + r0 + r + 1) - bs);
        if (e - a <= kStreamSeqLen) {
          real_t sum = 0.0;
          for (uint32_t k = a; k < e; k++) sum = addRn(sum, prod[k]);
          y[r0 + r] = sum;
          if (DOT) dotAcc = fma(sum, __ldg(x + r0 + r), dotAcc);
        } else {
          longSeen = 1;
        }
      }
      if (__syncthreads_or(longSeen)) {                       // (also the barrier before prod is overwritten)
        // C: one warp per long row
        for (uint32_t r = warp; r < nrows; r += kStreamThreads / 32) {
          const uint32_t a = (uint32_t)(__ldg(rowPtr + r0 + r) - bs), e = (uint32_t)(__ldg(rowPtr + r0 + r + 1) - bs);
          if (e - a > kStreamSeqLen) {
            real_t sum = 0.0;
            for (uint32_t k = a + lane; k < e; k += 32) sum = addRn(sum, prod[k]);
            sum = warpSum(sum);
            if (lane == 0) {
              y[r0 + r] = sum;
              if (DOT) dotAcc = fma(sum, __ldg(x + r0 + r), dotAcc);
            }
          }
        }
        __syncthreads();
      }
    } else {
      // a single row longer than a block: the whole CTA, four independent accumulators per thread
      real_t part[4] = { 0.0, 0.0, 0.0, 0.0 };
      for (uint64_t j = (uint64_t)bs + tid; j < be; j += 4 * kStreamThreads) {
        idx_t cc[4];
        real_t vv[4], xx[4];
#pragma unroll
        for (int u = 0; u < 4; u++) acc.load(j + u * kStreamThreads < be ? j + u * kStreamThreads : (uint64_t)bs, cc[u], vv[u]);
#pragma unroll
        for (int u = 0; u < 4; u++) xx[u] = __ldg(x + cc[u]);
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (j + u * kStreamThreads < be) part[u] = mulAdd(part[u], vv[u], xx[u]);
      }
      const real_t tot = blockSum((part[0] + part[1]) + (part[2] + part[3]), scratch);
      if (tid == 0) {
        y[r0] = tot;
        if (DOT) dotAcc = fma(tot, __ldg(x + r0), dotAcc);
      }
    }
  }
  if (DOT) {
    const real_t bsum = blockSum(dotAcc, scratch);
    gridSum(bsum, partials, ticket, dotOut, accumulate, scratch, push.size ? &push : nullptr);
  }
}

__global__ void maxRowLenKernel(idx_t nr, const idx_t* __restrict__ rowPtr, unsigned long long* out)
{
  unsigned long long m = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nr; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long len = (unsigned long long)(rowPtr[i + 1] - rowPtr[i]);
    m = len > m ? len : m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
    m = other > m ? other : m;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// Decides once per matrix which kernel family takes it, and builds the block table for the skewed ones (host greedy
// pass over the row pointers: setup, like the SELL conversion). Blocks the stream.
static bool rowsAreSkewed(RowBlocks* rb, const idx_t* rowPtr, idx_t nr, uint64_t nnz, cudaStream_t s)
{
  if (!rb || useLegacyKernels()) return false;
  if (rb->state == 0) {
    static const int knob = envInt("SB_ROWS_BALANCED", -1);      // -1: decide by the row lengths, 0: never, 1: always
    Context& c = ctx();
    bool skewed = knob == 1;
    if (knob < 0 && nr > 0) {
      unsigned long long* dMax = (unsigned long long*)sbAllocateDevice(64, sizeof(unsigned long long));
      unsigned long long hMax = 0;
      SB_CUDA(cudaMemsetAsync(dMax, 0, sizeof(unsigned long long), s));
      uint64_t blocks = ((uint64_t)nr + 255) / 256;
      if (blocks > (uint64_t)c.numSMs * 8) blocks = (uint64_t)c.numSMs * 8;
      maxRowLenKernel<<<(int)blocks, 256, 0, s>>>(nr, rowPtr, dMax);
      SB_CUDA(cudaMemcpyAsync(&hMax, dMax, sizeof(hMax), cudaMemcpyDeviceToHost, s));
      SB_CUDA(cudaStreamSynchronize(s));
      sbFree(dMax);
      skewed = (double)hMax > 1.2 * ((double)nnz / (double)nr) + 4.0;
    }
    if (skewed && nr > 0) {
      std::vector<idx_t> rp((size_t)nr + 1), start;
      SB_CUDA(cudaMemcpyAsync(rp.data(), rowPtr, sizeof(idx_t) * ((size_t)nr + 1), cudaMemcpyDeviceToHost, s));
      SB_CUDA(cudaStreamSynchronize(s));
      start.reserve((size_t)(nnz / (kStreamNnz / 2)) + 16);
      idx_t r = 0;
      while (r < nr) {
        start.push_back(r);
        const idx_t base = rp[(size_t)r];
        idx_t e = r + 1;                                        // a row longer than a block stands alone
        while (e < nr && e - r < kStreamMaxRows && rp[(size_t)e + 1] - base <= kStreamNnz) e++;
        r = e;
      }
      start.push_back(nr);
      if (start.size() - 1 > 0xffffffffull) SB_FATAL("row block table: more than 2^32 blocks");
      rb->count = (uint32_t)(start.size() - 1);
      rb->start = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * start.size());
      SB_CUDA(cudaMemcpyAsync(rb->start, start.data(), sizeof(idx_t) * start.size(), cudaMemcpyHostToDevice, s));
      SB_CUDA(cudaStreamSynchronize(s));
    }
    rb->state = skewed && nr > 0 ? 2 : 1;
  }
  return rb->state == 2;
}

template <typename Access, int ROUND>
static void launchRowsStreamCfg(Access acc, const idx_t* rowPtr, const RowBlocks& rb, const real_t* x, real_t* y, const DotArgs* dot,
    cudaStream_t s)
{
  Context& c = ctx();
  uint64_t blocks = rb.count;
  const uint64_t cap = (uint64_t)c.numSMs * (ROUND == 8 ? 4 : 5);
  if (blocks > cap) blocks = cap;
  if (blocks > (uint64_t)kMaxPartials) blocks = kMaxPartials;
  if (dot)
    spmvRowsStreamKernel<true, Access, ROUND><<<(int)blocks, kStreamThreads, 0, s>>>(acc, rowPtr, rb.start, rb.count, x, y,
        c.partials + (size_t)dot->slot * kMaxPartials, c.tickets + dot->slot, dot->out, dot->accumulate, dot->push ? *dot->push : PeerReduce());
  else
    spmvRowsStreamKernel<false, Access, ROUND><<<(int)blocks, kStreamThreads, 0, s>>>(acc, rowPtr, rb.start, rb.count, x, y, nullptr, nullptr,
        nullptr, false, PeerReduce());
  SB_CUDA(cudaGetLastError());
  countLaunch();
}

template <typename Access>
static void launchRowsStream(Access acc, const idx_t* rowPtr, const RowBlocks& rb, const real_t* x, real_t* y, const DotArgs* dot,
    cudaStream_t s)
{
  static const int round = envInt("SB_STREAM_ROUND", 4);
  if (round == 8) launchRowsStreamCfg<Access, 8>(acc, rowPtr, rb, x, y, dot, s);
  else launchRowsStreamCfg<Access, 4>(acc, rowPtr, rb, x, y, dot, s);
}

// lanes per row ~ avg/7 (one 8-deep batch per row); tile = as many passes of the consumer warps as fit a stage.
// CRS and CCRS take the same decisions (same CAP), so their row sums are bit-identical to each other.
template <typename Access, typename P>
static bool tryRowsPipe(Access acc, const idx_t* rowPtr, real_t avg, const real_t* x, real_t* y, idx_t lo, idx_t hi,
    const DotArgs* dot, const RowsGate* g, bool probeOnly, cudaStream_t s)
{
  // every pipe layout has CAP / WARPS = 224 non-zeros per consumer warp and pass: lanes per row = avg / 7
  const int lpr = avg <= 7.0 ? 1 : avg <= 14.0 ? 2 : avg <= 28.0 ? 4 : avg <= 56.0 ? 8 : avg <= 112.0 ? 16 : 32;
  const idx_t rowsPerPass = (idx_t)P::kWarps * (32u / (idx_t)lpr);
  const idx_t fitRows = (idx_t)((real_t)P::kCap / (avg > 1.0 ? avg : 1.0));
  idx_t tileRows = (fitRows / rowsPerPass) * rowsPerPass;
  if (tileRows > kPipeMaxRows) tileRows = (kPipeMaxRows / rowsPerPass) * rowsPerPass;
  if (tileRows < rowsPerPass) return false;
  if (probeOnly) return true;
#define SB_PIPE_V(LPRV, V)                                                                                          \
  do {                                                                                                              \
    if (g) launchRowsPipe<LPRV, P, true, V>(acc.template pipe<P>(), rowPtr, x, y, lo, hi, tileRows, dot, g, s);      \
    else launchRowsPipe<LPRV, P, false, V>(acc.template pipe<P>(), rowPtr, x, y, lo, hi, tileRows, dot, nullptr, s); \
  } while (0)
#define SB_PIPE(LPRV) SB_PIPE_V(LPRV, Access::kVar)
  switch (lpr) {
  case 1: SB_PIPE(1); break;
  case 2: SB_PIPE(2); break;
  case 4: {
#ifdef SB_TUNING_SWEEPS                                              // builds for tools/gpu_rows_var.sh: every VAR of the stencil case
    static const int var = envInt("SB_ROWS_VAR", Access::kVar);
    switch (var) {
    case 0: SB_PIPE_V(4, 0); break;
    case 1: SB_PIPE_V(4, 1); break;
    case 2: SB_PIPE_V(4, 2); break;
    case 3: SB_PIPE_V(4, 3); break;
    case 4: SB_PIPE_V(4, 4); break;
    case 5: SB_PIPE_V(4, 5); break;
    case 8: SB_PIPE_V(4, 8); break;
    case 12: SB_PIPE_V(4, 12); break;
    default: SB_PIPE(4); break;
    }
#else
    SB_PIPE(4);
#endif
    break;
  }
  case 8: SB_PIPE(8); break;
  case 16: SB_PIPE(16); break;
  default: SB_PIPE(32); break;
  }
#undef SB_PIPE_V
#undef SB_PIPE
  return true;
}

// returns false if the pipelined kernel cannot take this matrix (rows too long for a stage)
template <typename Access>
static bool launchRowsPipeAuto(Access acc, const idx_t* rowPtr, idx_t nr, uint64_t nnz, const real_t* x, real_t* y,
    idx_t lo, idx_t hi, const DotArgs* dot, const RowsGate* g, bool probeOnly, cudaStream_t s)
{
  const real_t avg = nr ? (real_t)nnz / (real_t)nr : 0.0;
  if (useLegacyKernels()) return false;
  // the measured best per format (profiles/README.md); the other stage layout is compiled only into sweep builds
#ifdef SB_TUNING_SWEEPS
  static const int cfg = envInt("SB_ROWS_CFG", Access::kDefaultCfg);
  if (cfg == 1 || Access::kWide)                          // the 5376-element stages do not fit with 16 bytes per non-zero
    return tryRowsPipe<Access, typename Access::template Pipe<16, 3584, Access::kStagesFor3584>>(acc, rowPtr, avg, x, y, lo, hi, dot, g, probeOnly, s);
  return tryRowsPipe<Access, typename Access::template Pipe<23, 5376, Access::kStagesFor5376>>(acc, rowPtr, avg, x, y, lo, hi, dot, g, probeOnly, s);
#else
  if constexpr (Access::kDefaultCfg == 1 || Access::kWide)
    return tryRowsPipe<Access, typename Access::template Pipe<16, 3584, Access::kStagesFor3584>>(acc, rowPtr, avg, x, y, lo, hi, dot, g, probeOnly, s);
  else
    return tryRowsPipe<Access, typename Access::template Pipe<23, 5376, Access::kStagesFor5376>>(acc, rowPtr, avg, x, y, lo, hi, dot, g, probeOnly, s);
#endif
}

template <typename Access>
static void launchRowsAuto(Access acc, const idx_t* rowPtr, idx_t nr, uint64_t nnz, const real_t* x, real_t* y,
    idx_t lo, idx_t hi, const DotArgs* dot, RowBlocks* rb, cudaStream_t s)
{
  const real_t avg = nr ? (real_t)nnz / (real_t)nr : 0.0;
  if (lo == 0 && hi == nr && rowsAreSkewed(rb, rowPtr, nr, nnz, s)) {      // (a sub-range keeps the tiled kernels)
    launchRowsStream(acc, rowPtr, *rb, x, y, dot, s);
    return;
  }
  if (launchRowsPipeAuto(acc, rowPtr, nr, nnz, x, y, lo, hi, dot, nullptr, false, s)) return;
  if (avg <= 6.0) launchRows<2, Access>(acc, rowPtr, nr, x, y, lo, hi, dot, s);
  else if (avg <= 12.0) launchRows<4, Access>(acc, rowPtr, nr, x, y, lo, hi, dot, s);
  else if (avg <= 40.0) launchRows<8, Access>(acc, rowPtr, nr, x, y, lo, hi, dot, s);
  else if (avg <= 96.0) launchRows<16, Access>(acc, rowPtr, nr, x, y, lo, hi, dot, s);
  else launchRows<32, Access>(acc, rowPtr, nr, x, y, lo, hi, dot, s);
}

idx_t spmvUnits(const Operator& A) { return A.fmt == SB_FMT_SCS ? A.sell.nChunks : A.nr; }

// ---- interior / boundary split for the overlapped halo exchange (setup: once per converted matrix)
// one warp per unit; bounds[0] = max(u+1) over halo-touching units u below the middle, bounds[1] = min(u) over
// those at or above it
__global__ void haloTouchKernel(Operator A, idx_t units, idx_t* bounds)
{
  const idx_t mid = units / 2;
  const int lane = threadIdx.x & 31;
  const idx_t warpsTotal = (gridDim.x * blockDim.x) >> 5;
  for (idx_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < units; u += warpsTotal) {
    bool touch = false;
    if (A.fmt == SB_FMT_SCS) {
      const uint64_t b = A.sell.chunkPtr[u], e = b + (uint64_t)A.sell.chunkLens[u] * A.sell.C;
      for (uint64_t j = b + lane; j < e; j += 32) touch |= A.sell.col[j] >= A.nr;
    } else if (A.fmt == SB_FMT_CRS) {
      for (idx_t j = A.crs.rowPtr[u] + lane; j < A.crs.rowPtr[u + 1]; j += 32) touch |= A.crs.col[j] >= A.nr;
    } else {
      for (idx_t j = A.ccrs.rowPtr[u] + lane; j < A.ccrs.rowPtr[u + 1]; j += 32) touch |= A.ccrs.entries[j].col >= A.nr;
    }
    if (__any_sync(0xffffffffu, touch) && lane == 0) {
      if (u < mid) atomicMax(bounds, u + 1);
      else atomicMin(bounds + 1, u);
    }
  }
}

void spmvInteriorUnits(const Operator& A, idx_t* lo, idx_t* hi, cudaStream_t s)
{
  if (A.split && A.split->valid) {
    *lo = A.split->lo;
    *hi = A.split->hi;
    return;
  }
  const idx_t units = spmvUnits(A);
  idx_t h[2] = { 0u, units };
  if (units > 0) {
    idx_t* d = (idx_t*)sbAllocateDevice(64, 2 * sizeof(idx_t));
    SB_CUDA(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, s));
    haloTouchKernel<<<ctx().numSMs * 8, 256, 0, s>>>(A, units, d);
    SB_CUDA(cudaGetLastError());
    countLaunch();
    SB_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    sbFree(d);
  }
  *lo = h[0];
  *hi = h[1] > h[0] ? h[1] : h[0];
  if (A.split) {
    A.split->valid = true;
    A.split->lo = *lo;
    A.split->hi = *hi;
  }
}

void launchSpmv(const Operator& A, const real_t* x, real_t* y, idx_t lo, idx_t hi, const DotArgs* dot,
    cudaStream_t s)
{
  if (hi <= lo) {
    if (dot && !dot->accumulate) SB_CUDA(cudaMemsetAsync(dot->out, 0, sizeof(real_t), s));
    return;
  }
  Context& c = ctx();
  if (A.fmt == SB_FMT_SCS && A.sell.C != 32) {
    uint64_t blocks = ((uint64_t)(hi - lo) * A.sell.C + 255) / 256;
    const uint64_t cap = (uint64_t)c.numSMs * 8;
    if (blocks > cap) blocks = cap;
    if (dot)
      spmvSellAnyCKernel<true><<<(int)blocks, 256, 0, s>>>(A.sell, x, y, lo, hi,
          c.partials + (size_t)dot->slot * kMaxPartials, c.tickets + dot->slot, dot->out, dot->accumulate);
    else
      spmvSellAnyCKernel<false><<<(int)blocks, 256, 0, s>>>(A.sell, x, y, lo, hi, nullptr, nullptr, nullptr, false);
    SB_CUDA(cudaGetLastError());
    countLaunch();
  } else if (A.fmt == SB_FMT_SCS && !useLegacyKernels() && A.longc && lo == 0 && hi == A.sell.nChunks && !(dot && dot->push)) {
    // heavy-tailed chunk lengths: free-running warps for the ordinary chunks (nothing to keep in step), then the long
    // chunks by a CTA each; the second kernel adds its share of the fused dot to the first one's
    SellView shortView = A.sell;
    shortView.chunkLens = A.longc->shortLens;
    launchSell32TmaCfg<32, 4, 3, 4, false, false>(shortView, x, y, lo, hi, dot, nullptr, s);
    uint64_t blocks = A.longc->count;
    if (blocks > (uint64_t)c.numSMs * 4) blocks = (uint64_t)c.numSMs * 4;
    if (dot)
      spmvSellLongChunksKernel<true><<<(int)blocks, kLongWarps * 32, 0, s>>>(A.sell, A.longc->list, A.longc->count, x, y,
          c.partials + (size_t)dot->slot * kMaxPartials, c.tickets + dot->slot, dot->out, true);
    else
      spmvSellLongChunksKernel<false><<<(int)blocks, kLongWarps * 32, 0, s>>>(A.sell, A.longc->list, A.longc->count, x, y, nullptr,
          nullptr, nullptr, false);
    SB_CUDA(cudaGetLastError());
    countLaunch();
  } else if (A.fmt == SB_FMT_SCS && !useLegacyKernels()) {
    launchSell32Tma(A.sell, x, y, lo, hi, dot, nullptr, s);
  } else if (A.fmt == SB_FMT_SCS) {
    uint64_t blocks = ((uint64_t)(hi - lo) + 7) / 8;
    const uint64_t cap = (uint64_t)c.numSMs * 4;
    if (blocks > cap) blocks = cap;
    if (dot)
      spmvSell32Kernel<true><<<(int)blocks, 256, 0, s>>>(A.sell, x, y, lo, hi,
          c.partials + (size_t)dot->slot * kMaxPartials, c.tickets + dot->slot, dot->out, dot->accumulate);
    else
      spmvSell32Kernel<false><<<(int)blocks, 256, 0, s>>>(A.sell, x, y, lo, hi, nullptr, nullptr, nullptr, false);
    SB_CUDA(cudaGetLastError());
    countLaunch();
  } else if (A.fmt == SB_FMT_CRS) {
    launchRowsAuto(CrsAccess { A.crs.col, A.crs.val }, A.crs.rowPtr, A.nr, A.nnzTrue, x, y, lo, hi, dot, A.blocks, s);
  } else {
    launchRowsAuto(CcrsAccess { A.ccrs.entries }, A.ccrs.rowPtr, A.nr, A.nnzTrue, x, y, lo, hi, dot, A.blocks, s);
  }
}

bool spmvGatedAvailable(const Operator& A)
{
  if (useLegacyKernels()) return false;
  if (A.fmt == SB_FMT_SCS) return A.sell.C == 32 && !A.longc;       // long chunks: two launches, no gated form
  // skewed row lengths: the nnz-balanced kernel has no gated form; the halo exchange completes before its launch
  if (rowsAreSkewed(A.blocks, A.rowPtr, A.nr, A.nnzTrue, ctx().stream)) return false;
  if (A.fmt == SB_FMT_CRS)
    return launchRowsPipeAuto(CrsAccess { A.crs.col, A.crs.val }, A.crs.rowPtr, A.nr, A.nnzTrue, nullptr, nullptr, 0, A.nr, nullptr, nullptr, true, nullptr);
  return launchRowsPipeAuto(CcrsAccess { A.ccrs.entries }, A.ccrs.rowPtr, A.nr, A.nnzTrue, nullptr, nullptr, 0, A.nr, nullptr, nullptr, true, nullptr);
}

void launchSpmvGated(const Operator& A, const real_t* x, real_t* y, idx_t intLo, idx_t intHi, const HaloGate& gate,
    const DotArgs* dot, cudaStream_t s)
{
  const idx_t units = spmvUnits(A);
  if (units == 0) {
    if (dot && !dot->accumulate) SB_CUDA(cudaMemsetAsync(dot->out, 0, sizeof(real_t), s));
    return;
  }
  if (intHi < intLo) intHi = intLo;
  if (A.fmt == SB_FMT_SCS) {
    SellGate g { intLo, intHi - intLo, gate };       // rotate so that the interior chunks come first
    launchSell32Tma(A.sell, x, y, 0, units, dot, &g, s);
  } else {
    RowsGate g { intLo, intHi, gate };
    bool ok;
    if (A.fmt == SB_FMT_CRS)
      ok = launchRowsPipeAuto(CrsAccess { A.crs.col, A.crs.val }, A.crs.rowPtr, A.nr, A.nnzTrue, x, y, 0, A.nr, dot, &g, false, s);
    else
      ok = launchRowsPipeAuto(CcrsAccess { A.ccrs.entries }, A.ccrs.rowPtr, A.nr, A.nnzTrue, x, y, 0, A.nr, dot, &g, false, s);
    if (!ok) SB_FATAL("launchSpmvGated: no pipelined kernel for this matrix (check spmvGatedAvailable first)");
  }
}

} // namespace sb

using namespace sb;

extern "C" {

void sbCRS_spMVM(SbCRSMatrix* m, const CG_FLOAT* x, CG_FLOAT* y)
{
  Operator A = makeOperator(m, SB_FMT_CRS);
  ensureOnDevice(x);
  ensureOnDevice(y);
  launchSpmv(A, x, y, 0, A.nr, nullptr, ctx().stream);
}

void sbCCRS_spMVM(SbCCRSMatrix* m, const CG_FLOAT* x, CG_FLOAT* y)
{
  Operator A = makeOperator(m, SB_FMT_CCRS);
  ensureOnDevice(x);
  ensureOnDevice(y);
  launchSpmv(A, x, y, 0, A.nr, nullptr, ctx().stream);
}

int sbSpmvOrdered(void* matrix, int fmt, const CG_FLOAT* x, CG_FLOAT* y, CG_UINT intLo, CG_UINT intHi)
{
  // The single-launch kernel of the multi-GPU CG (interior units first, then the boundary units behind a halo
  // gate), launched with an open gate: lets a single-GPU test check its unit ordering against the plain kernel.
  Operator A = makeOperator(matrix, fmt);
  if (fmt == SB_FMT_SCS) A.sell.col = ((SbSCSMatrix*)matrix)->colInd;   // reference semantics, like sbSCS_spMVM
  if (!spmvGatedAvailable(A)) return 0;
  ensureOnDevice(x);
  ensureOnDevice(y);
  launchSpmvGated(A, x, y, intLo, intHi, HaloGate(), nullptr, ctx().stream);
  return 1;
}

int sbSpmvKernelFamily(void* matrix, int fmt)
{
  Operator A = makeOperator(matrix, fmt);
  if (useLegacyKernels()) return 3;
  if (fmt == SB_FMT_SCS) return A.sell.C != 32 ? 3 : A.longc ? 4 : 0;
  if (rowsAreSkewed(A.blocks, A.rowPtr, A.nr, A.nnzTrue, ctx().stream)) return 2;
  return spmvGatedAvailable(A) ? 1 : 3;
}

void sbSpmvDot(void* matrix, int fmt, const CG_FLOAT* x, CG_FLOAT* y, CG_FLOAT* dDot)
{
  // y = A x fused with *dDot = sum_i x[i] y[i] (device scalar; CGSolver.c:123-125 in one pass), the kernel the CG
  // loop runs. Vectors in solver order: for SCS with sigma > 1 that is the permuted row order for x AND y.
  Operator A = makeOperator(matrix, fmt);
  DotArgs d { dDot, false, 1 };
  ensureOnDevice(x);
  ensureOnDevice(y);
  ensureOnDevice(dDot);
  launchSpmv(A, x, y, 0, spmvUnits(A), &d, ctx().stream);
}

void sbSCS_spMVM(SbSCSMatrix* m, const CG_FLOAT* x, CG_FLOAT* y)
{
  // Reference semantics (matrix-SCS.c:198-228): x is indexed by the stored (un-permuted) column ids, y is
  // written in permuted row order and needs nrPadded slots.
  Operator A = makeOperator(m, SB_FMT_SCS);
  A.sell.col = m->colInd;
  ensureOnDevice(x);
  ensureOnDevice(y);
  launchSpmv(A, x, y, 0, A.sell.nChunks, nullptr, ctx().stream);
}

} // extern "C"
