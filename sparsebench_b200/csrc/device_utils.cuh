// Device-side helpers shared by the hot-path kernels: streaming loads, reference-order arithmetic and the
// deterministic one-kernel grid reduction used by every dot product.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "peer_window.h"
#include "sb_types.h"

namespace sb {

// Matrix values / indices are read exactly once per SpMV: keep them out of L1 so the cache stays with the
// gathered x vector (B200: 256 KB L1+smem per SM, 126 MB L2). One overload per width the two type switches produce.
__device__ __forceinline__ double ldStream(const double* p)
{
  double v;
  asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldStream(const float* p)
{
  float v;
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned int ldStream(const unsigned int* p)
{
  unsigned int v;
  asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long ldStream(const unsigned long long* p)
{
  unsigned long long v;
  asm("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
// one {col, val} record of the CCRS format (8 or 16 bytes, naturally aligned)
__device__ __forceinline__ Entry ldStreamEntry(const Entry* p)
{
  Entry e;
  if (sizeof(Entry) == 16) {
    unsigned long long a, b;
    asm("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    memcpy(&e, &a, 8);
    memcpy(reinterpret_cast<char*>(&e) + 8, &b, 8);
  } else {
    unsigned long long a;
    asm("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(a) : "l"(p));
    memcpy(&e, &a, sizeof(Entry) < 8 ? sizeof(Entry) : 8);
  }
  return e;
}

// wall-clock nanoseconds (the same on every SM, independent of the SM clock)
__device__ __forceinline__ unsigned long long globalTimerNs()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// A kernel that waits for a peer gives up after this long and traps (the host sees a launch failure and exits): a dead
// peer must not hang the GPU forever.
constexpr unsigned long long kPeerTimeoutNs = 20000000000ull;   // 20 s

// The reference accumulates `sum += val * x` with a separately rounded multiply and add (strict C, no
// contraction on baseline x86-64). Keeping that order and rounding makes row sums bit-identical to the
// reference's; fp64 pipes are <5 % utilised by an HBM-bound SpMV, so the extra instruction is free.
__device__ __forceinline__ double mulAdd(double acc, double a, double b) { return __dadd_rn(acc, __dmul_rn(a, b)); }
__device__ __forceinline__ float mulAdd(float acc, float a, float b) { return __fadd_rn(acc, __fmul_rn(a, b)); }
// x + a*y and a*x + b*y with every operation rounded separately (waxpby of solver.c:16-39, strict C semantics)
__device__ __forceinline__ double addRn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float addRn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double mulRn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mulRn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double divRn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float divRn(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ real_t warpSum(real_t v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the block in a fixed order; result valid in thread 0. `scratch` holds >= 32 values.
__device__ __forceinline__ real_t blockSum(real_t v, real_t* scratch)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  v = warpSum(v);
  __syncthreads();              // scratch may still be in use by a previous call
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  real_t t = 0.0;
  if (warp == 0) {
    t = lane < nwarps ? scratch[lane] : 0.0;
    t = warpSum(t);
  }
  return t;
}

// One-kernel grid reduction: every block deposits its partial, the last block to arrive (atomic ticket)
// adds all partials in a fixed order, so the result does not depend on block scheduling.
// out = (accumulate ? *out : 0) + sum(partials). The ticket resets itself for the next launch.
// With `push` (multi-GPU) the last block additionally stores the sum into slot [rank] of every peer's control
// window -- the first half of an all-reduce whose second half (peerCollect) runs in the prologue of the kernel
// that consumes the scalar, so no separate all-reduce launch sits between the two.
// `mirror` (optional): a second place the final value is stored to -- the solver passes mapped pinned host memory, so
// the host sees rho[k] without a copy or an event in the stream.
__device__ __forceinline__ void gridSum(real_t blockPartial, real_t* partials, unsigned int* ticket, real_t* out,
    bool accumulate, real_t* scratch, const PeerReduce* push = nullptr, real_t* mirror = nullptr)
{
  __shared__ bool amLast;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = blockPartial;
    __threadfence();
    const unsigned int t = atomicInc(ticket, gridDim.x - 1);
    amLast = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (amLast) {
    __threadfence();
    real_t v = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += __ldcg(partials + i);
    v = blockSum(v, scratch);
    if (threadIdx.x == 0) {
      v = accumulate ? (*out + v) : v;
      *out = v;
      if (mirror) *(volatile real_t*)mirror = v;
      scratch[0] = v;
    }
    if (push && push->size > 0) {
      __syncthreads();
      if ((int)threadIdx.x < push->size) {
        const real_t mine = scratch[0];
        CtrlWindow* w = push->peers[threadIdx.x];
        const int slot = (int)(push->epoch % kRedDepth);
        *(volatile double*)&w->redVal[slot][push->rank] = (double)mine;      // the window's slots are doubles in every build
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&w->redFlag[slot][push->rank]), "l"(push->epoch) : "memory");
      }
    }
  }
}

// Second half of the all-reduce: every thread of the calling block gets the global sum. `vals` is shared memory
// for kMaxRanks doubles. Contains block barriers: call from all threads.
__device__ __forceinline__ real_t peerCollect(const PeerReduce& pr, real_t* vals)
{
  const int slot = (int)(pr.epoch % kRedDepth);
  const bool traced = pr.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  const unsigned long long t0 = traced ? globalTimerNs() : 0ull;
  if ((int)threadIdx.x < pr.size) {
    const unsigned long long* flag = &pr.mine->redFlag[slot][threadIdx.x];
    unsigned long long seen;
    const unsigned long long start = globalTimerNs();
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
      if (seen >= pr.epoch) break;
      __nanosleep(40);
      if (globalTimerNs() - start > kPeerTimeoutNs) __trap();
    }
    vals[threadIdx.x] = (real_t) * (volatile double*)&pr.mine->redVal[slot][threadIdx.x];
  }
  __syncthreads();
  if (traced) {                                           // how long the slowest peer's partial kept this rank waiting
    atomicAdd(pr.trace, globalTimerNs() - t0);
    atomicAdd(pr.trace + 1, 1ull);
  }
  real_t acc = vals[0];
  for (int r = 1; r < pr.size; r++) acc += vals[r];      // rank order: same bits on every rank
  __syncthreads();
  return acc;
}

// ---- programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still draining. Everything it does before griddepWait() must touch
// only data no kernel ever writes (the matrix arrays, its own shared memory); griddepWait() returns when the
// predecessor has completed and its writes are visible. Every kernel of a PDL chain executes both calls, so the order
// is transitive. Without the launch attribute both are no-ops.
__device__ __forceinline__ void griddepLaunchDependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddepWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier + 1-D bulk copy (TMA, SASS UBLKCP): asynchronous global -> shared streaming of the matrix
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t arrivals)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbarFenceInit() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SB_DONE;\n"
      "bra SB_WAIT;\n"
      "SB_DONE:\n"
      "}\n" ::"r"(smemAddr(bar)),
      "r"(parity)
      : "memory");
}
// The matrix stream is read exactly once per SpMV: mark its lines evict-first in L2, so that they are replaced
// before the vectors (x gathers, and at small sizes all CG vectors, which then stay L2-resident between kernels).
// Measured: 128^3 CG 0.156 -> 0.147 ms per iteration, 256^3 unchanged (profiles/README.md).
__device__ __forceinline__ uint64_t l2EvictFirstPolicy()
{
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}
// size and both addresses must be multiples of 16 bytes
__device__ __forceinline__ void bulkLoad(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smemAddr(smemDst)),
               "l"(gmemSrc), "r"(bytes), "r"(smemAddr(bar)), "l"(l2EvictFirstPolicy())
               : "memory");
}

// the same copy without the evict-first hint: for data that is used again (vector entries)
__device__ __forceinline__ void bulkLoadKeep(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(smemDst)),
               "l"(gmemSrc), "r"(bytes), "r"(smemAddr(bar))
               : "memory");
}

} // namespace sb
