// Krylov vector kernels (replace waxpby / ddot of solver.c:16-62) and their fused single-pass forms used by
// the CG driver. All are pure HBM streams: grid = numSMs * 8 CTAs, 16-byte vector accesses, arithmetic in the
// reference's rounding (separate multiply and add), dot products through the deterministic grid reduction.
#include "device_utils.cuh"
#include "sb_internal.h"

namespace sb {

constexpr int kVecThreads = 256;

static inline int vecGrid(uint64_t n, int perThread)
{
  Context& c = ctx();
  uint64_t blocks = (n + (uint64_t)kVecThreads * perThread - 1) / ((uint64_t)kVecThreads * perThread);
  uint64_t cap = (uint64_t)c.numSMs * 8;
  if (blocks > cap) blocks = cap;
  if (blocks > (uint64_t)kMaxPartials) blocks = kMaxPartials;
  return (int)(blocks ? blocks : 1);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// 16 bytes of vector: 2 doubles or 4 floats, moved with one 128-bit access
constexpr int kVL = 16 / sizeof(real_t);
struct alignas(16) RVec {
  real_t v[kVL];
};

// MODE 0: w = x + beta*y   (alpha == 1, solver.c:24-27)
// MODE 1: w = alpha*x + y  (beta == 1,  solver.c:29-32)
// MODE 2: w = alpha*x + beta*y          (solver.c:34-37)
template <int MODE>
__device__ __forceinline__ real_t waxpbyOne(real_t alpha, real_t x, real_t beta, real_t y)
{
  if (MODE == 0) return addRn(x, mulRn(beta, y));
  if (MODE == 1) return addRn(mulRn(alpha, x), y);
  return addRn(mulRn(alpha, x), mulRn(beta, y));
}

// w may alias x or y (the CG calls it in place, CGSolver.c:114,127,128): every element is read before it is
// written by the same thread, so no __restrict__ here.
template <int MODE, bool VEC>
__global__ void __launch_bounds__(kVecThreads)
waxpbyKernel(idx_t n, real_t alpha, const real_t* x, real_t beta, const real_t* y, real_t* w)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (VEC) {
    const uint64_t n2 = n / kVL;
    const RVec* x2 = reinterpret_cast<const RVec*>(x);
    const RVec* y2 = reinterpret_cast<const RVec*>(y);
    RVec* w2 = reinterpret_cast<RVec*>(w);
    for (uint64_t i = tid; i < n2; i += stride) {
      const RVec a = x2[i], b = y2[i];
      RVec r;
#pragma unroll
      for (int c = 0; c < kVL; c++) r.v[c] = waxpbyOne<MODE>(alpha, a.v[c], beta, b.v[c]);
      w2[i] = r;
    }
    if (tid == 0)
      for (uint64_t i = n2 * kVL; i < n; i++) w[i] = waxpbyOne<MODE>(alpha, x[i], beta, y[i]);
  } else {
    for (uint64_t i = tid; i < n; i += stride) w[i] = waxpbyOne<MODE>(alpha, x[i], beta, y[i]);
  }
}

void launchWaxpby(idx_t n, real_t alpha, const real_t* x, real_t beta, const real_t* y, real_t* w, cudaStream_t s)
{
  if (n == 0) return;
  const bool vec = aligned16(x) && aligned16(y) && aligned16(w);
  const int grid = vecGrid(n, vec ? 4 : 2);
  const int mode = (alpha == (real_t)1.0) ? 0 : (beta == (real_t)1.0) ? 1 : 2;
#define SB_LAUNCH(M)                                                                        \
  do {                                                                                      \
    if (vec) waxpbyKernel<M, true><<<grid, kVecThreads, 0, s>>>(n, alpha, x, beta, y, w);   \
    else waxpbyKernel<M, false><<<grid, kVecThreads, 0, s>>>(n, alpha, x, beta, y, w);      \
  } while (0)
  if (mode == 0) SB_LAUNCH(0);
  else if (mode == 1) SB_LAUNCH(1);
  else SB_LAUNCH(2);
#undef SB_LAUNCH
  SB_CUDA(cudaGetLastError());
  countLaunch();
}

// per-lane partial sums of a thread -> one value, fixed order ((0+1)+(2+3))
__device__ __forceinline__ real_t sumLanes(const real_t (&part)[kVL])
{
  if (kVL == 2) return part[0] + part[1];
  return (part[0] + part[1]) + (part[kVL - 2] + part[kVL - 1]);
}

template <bool VEC>
__global__ void __launch_bounds__(kVecThreads)
dotKernel(idx_t n, const real_t* __restrict__ x, const real_t* __restrict__ y, real_t* partials, unsigned int* ticket,
    real_t* out)
{
  __shared__ real_t scratch[32];
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  real_t acc = 0.0;
  if (VEC) {
    const uint64_t n2 = n / kVL;
    const RVec* x2 = reinterpret_cast<const RVec*>(x);
    const RVec* y2 = reinterpret_cast<const RVec*>(y);
    real_t part[kVL] = {};
    for (uint64_t i = tid; i < n2; i += stride) {
      const RVec a = x2[i], b = y2[i];
#pragma unroll
      for (int c = 0; c < kVL; c++) part[c] = fma(a.v[c], b.v[c], part[c]);
    }
    acc = sumLanes(part);
    if (tid == 0)
      for (uint64_t i = n2 * kVL; i < n; i++) acc = fma(x[i], y[i], acc);
  } else {
    for (uint64_t i = tid; i < n; i += stride) acc = fma(x[i], y[i], acc);
  }
  const real_t b = blockSum(acc, scratch);
  gridSum(b, partials, ticket, out, false, scratch);
}

void launchDot(idx_t n, const real_t* x, const real_t* y, real_t* dResult, int slot, cudaStream_t s)
{
  Context& c = ctx();
  if (n == 0) {
    SB_CUDA(cudaMemsetAsync(dResult, 0, sizeof(real_t), s));
    return;
  }
  const bool vec = aligned16(x) && aligned16(y);
  const int grid = vecGrid(n, vec ? 4 : 2);
  if (vec)
    dotKernel<true><<<grid, kVecThreads, 0, s>>>(n, x, y, c.partials + (size_t)slot * kMaxPartials, c.tickets + slot, dResult);
  else
    dotKernel<false><<<grid, kVecThreads, 0, s>>>(n, x, y, c.partials + (size_t)slot * kMaxPartials, c.tickets + slot, dResult);
  SB_CUDA(cudaGetLastError());
  countLaunch();
}

// ------------------------------------------------------------------------------------------- fused CG passes
// rho[j] = r_j . r_j (rho[0] from the initial residual), pAp[k] = p_k . A p_k. Iteration k (1-based,
// CGSolver.c:107-129) uses rtrans = rho[k-1], oldrtrans = rho[k-2].

// Both kernels keep 4 independent 16-byte loads per array in flight per thread (kVecUnroll); at 128^3 the whole
// vector is one pass of the grid, so the kernels are a single DRAM/L2 round trip instead of a dependent chain.
constexpr int kVecUnroll = 4;

// KEEP: when all CG vectors together fit the 126 MB L2 (small problems), their lines are tagged evict_last, so that
// the matrix stream of the SpMV in between (tagged evict_first by the bulk copies) cannot push them out and the
// vector kernels run out of L2 instead of HBM.
__device__ __forceinline__ uint64_t l2EvictLastPolicy()
{
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}
template <bool KEEP>
__device__ __forceinline__ RVec ldVec(const RVec* p, uint64_t policy)
{
  if (!KEEP) return *p;
  uint4 raw;
  asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
               : "l"(p), "l"(policy)
               : "memory");
  RVec v;
  memcpy(&v, &raw, 16);
  return v;
}
template <bool KEEP>
__device__ __forceinline__ void stVec(RVec* p, const RVec& v, uint64_t policy)
{
  if (!KEEP) {
    *p = v;
    return;
  }
  uint4 raw;
  memcpy(&raw, &v, 16);
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(raw.x), "r"(raw.y), "r"(raw.z), "r"(raw.w),
               "l"(policy)
               : "memory");
}
// all CG vectors (r, p, Ap, x, b) of this size fit L2 with room for the stream
static inline bool vectorsFitL2(idx_t n)
{
  static const int knob = getenv("SB_VEC_KEEP") ? atoi(getenv("SB_VEC_KEEP")) : 1;
  return knob != 0 && (uint64_t)n * sizeof(real_t) * 5 <= ((uint64_t)96 << 20);
}

// p = r + beta*p with beta = rho[k-1]/rho[k-2]  (CGSolver.c:111-114); k == 1: p = r + 0*r (:109).
// PUT: this launch also delivers the halo elements of p to the neighbours (FusedPut); without it the kernel carries
// none of that code.
template <bool KEEP, bool PUT>
__global__ void __launch_bounds__(kVecThreads)
cgUpdatePKernel(idx_t n, int k, real_t* rho, const real_t* __restrict__ r, real_t* __restrict__ p, PeerReduce collectRho,
    FusedPut put, real_t* hostRho)
{
  __shared__ real_t peerVals[kMaxRanks];
  __shared__ unsigned int delivered[kMaxFusedDests];     // halo elements this block stored at each neighbour
  griddepLaunchDependents();
  griddepWait();
  if (PUT && threadIdx.x < kMaxFusedDests) delivered[threadIdx.x] = 0;
  // multi-GPU: a freshly computed p[e] that a neighbour needs goes straight behind that neighbour's local rows
  // (a rolled loop over the ACTUAL destinations: unrolled to kMaxFusedDests predicated range tests it cost the
  // 2-rank p update 6 us of issue slots; the parameter arrays are indexed in the constant bank)
  auto deliver = [&](idx_t e, real_t v) {
#pragma unroll 1
    for (int d = 0; d < put.ndest; d++)
      if (e >= put.lo[d] && e <= put.hi[d]) {
        const int pos = __ldg(put.inv[d] + (e - put.lo[d]));
        if (pos >= 0) {
          put.remote[d][pos] = v;
          atomicAdd(&delivered[d], 1u);
        }
      }
  };
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t n2 = n / kVL;
  const RVec* r2 = reinterpret_cast<const RVec*>(r);
  RVec* p2 = reinterpret_cast<RVec*>(p);
  const uint64_t keep = KEEP ? l2EvictLastPolicy() : 0;
  // multi-GPU: rho[k-1] was pushed to the peer windows by the previous x/r update; sum it here (every block gets
  // the same bits) and let one thread store the global value for the host's convergence test
  real_t rtrans = 0.0;
  if (k > 1) {
    if (collectRho.size > 0) {
      rtrans = peerCollect(collectRho, peerVals);
      if (tid == 0) {
        rho[k - 1] = rtrans;
        if (hostRho) *(volatile real_t*)(hostRho + k - 1) = rtrans;   // the global value, for the host's convergence test
      }
    } else {
      rtrans = rho[k - 1];
    }
  }
  // k == 1: p = r + 0*r, i.e. beta = 0 applied to r itself (waxpby(1, r, 0, r, p))
  const real_t beta = k == 1 ? (real_t)0.0 : divRn(rtrans, rho[k - 2]);
  for (uint64_t i0 = tid; i0 < n2; i0 += kVecUnroll * stride) {
    RVec a[kVecUnroll], b[kVecUnroll];
#pragma unroll
    for (int u = 0; u < kVecUnroll; u++)
      if (i0 + u * stride < n2) {
        a[u] = ldVec<KEEP>(r2 + i0 + u * stride, keep);
        b[u] = k == 1 ? a[u] : ldVec<KEEP>(p2 + i0 + u * stride, keep);
      }
#pragma unroll
    for (int u = 0; u < kVecUnroll; u++)
      if (i0 + u * stride < n2) {
        RVec o;
#pragma unroll
        for (int c = 0; c < kVL; c++) o.v[c] = addRn(a[u].v[c], mulRn(beta, b[u].v[c]));
        stVec<KEEP>(p2 + i0 + u * stride, o, keep);
        if (PUT) {
          const idx_t e = (idx_t)(kVL * (i0 + u * stride));
#pragma unroll
          for (int c = 0; c < kVL; c++) deliver(e + c, o.v[c]);
        }
      }
  }
  if (tid == 0)
    for (uint64_t i = n2 * kVL; i < n; i++) {
      const real_t v = addRn(r[i], mulRn(beta, k == 1 ? r[i] : p[i]));
      p[i] = v;
      if (PUT) deliver((idx_t)i, v);
    }
  if (PUT) {
    // the barrier orders every thread's peer stores before the signalling threads' release (system scope)
    __syncthreads();
    if ((int)threadIdx.x < put.ndest && delivered[threadIdx.x] > 0)
      asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(put.remoteFlag[threadIdx.x]),
                   "l"((unsigned long long)delivered[threadIdx.x])
                   : "memory");
  }
}

// alpha = rho[k-1]/pAp[k]; x += alpha*p; r += (-alpha)*Ap; rho[k] = r.r   (CGSolver.c:126-128 + :112 of the
// next iteration, which reads the same r)
template <bool KEEP>
__global__ void __launch_bounds__(kVecThreads)
cgUpdateXRKernel(idx_t n, int k, real_t* rho, real_t* pAp, real_t* __restrict__ x,
    real_t* __restrict__ r, const real_t* __restrict__ p, const real_t* __restrict__ Ap, real_t* partials,
    unsigned int* ticket, PeerReduce collectPAp, PeerReduce pushRho, real_t* hostRho)
{
  __shared__ real_t scratch[32];
  __shared__ real_t peerVals[kMaxRanks];
  griddepLaunchDependents();
  griddepWait();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  real_t pApK;
  if (collectPAp.size > 0) {                         // multi-GPU: p.Ap was pushed to the peer windows by the SpMV
    pApK = peerCollect(collectPAp, peerVals);
    if (tid == 0) pAp[k] = pApK;
  } else {
    pApK = pAp[k];
  }
  const real_t alpha = divRn(rho[k - 1], pApK);
  const real_t nalpha = -alpha;
  const uint64_t n2 = n / kVL;
  RVec* x2 = reinterpret_cast<RVec*>(x);
  RVec* r2 = reinterpret_cast<RVec*>(r);
  const RVec* p2 = reinterpret_cast<const RVec*>(p);
  const RVec* q2 = reinterpret_cast<const RVec*>(Ap);
  const uint64_t keep = KEEP ? l2EvictLastPolicy() : 0;
  real_t part[kVL] = {};
  for (uint64_t i0 = tid; i0 < n2; i0 += 2 * stride) {
    RVec xv[2], pv[2], rv[2], qv[2];
#pragma unroll
    for (int u = 0; u < 2; u++)
      if (i0 + u * stride < n2) {
        xv[u] = ldVec<KEEP>(x2 + i0 + u * stride, keep);
        pv[u] = ldVec<KEEP>(p2 + i0 + u * stride, keep);
        rv[u] = ldVec<KEEP>(r2 + i0 + u * stride, keep);
        qv[u] = ldVec<KEEP>(q2 + i0 + u * stride, keep);
      }
#pragma unroll
    for (int u = 0; u < 2; u++)
      if (i0 + u * stride < n2) {
        RVec xo, ro;
#pragma unroll
        for (int c = 0; c < kVL; c++) {
          xo.v[c] = addRn(xv[u].v[c], mulRn(alpha, pv[u].v[c]));
          ro.v[c] = addRn(rv[u].v[c], mulRn(nalpha, qv[u].v[c]));
          part[c] = fma(ro.v[c], ro.v[c], part[c]);
        }
        stVec<KEEP>(x2 + i0 + u * stride, xo, keep);
        stVec<KEEP>(r2 + i0 + u * stride, ro, keep);
      }
  }
  real_t acc = sumLanes(part);
  if (tid == 0)
    for (uint64_t i = n2 * kVL; i < n; i++) {
      x[i] = addRn(x[i], mulRn(alpha, p[i]));
      const real_t ro = addRn(r[i], mulRn(nalpha, Ap[i]));
      r[i] = ro;
      acc = fma(ro, ro, acc);
    }
  const real_t b = blockSum(acc, scratch);
  gridSum(b, partials, ticket, rho + k, false, scratch, pushRho.size ? &pushRho : nullptr, hostRho ? hostRho + k : nullptr);
}

void launchCgUpdateP(idx_t n, int k, real_t* rho, const real_t* r, real_t* p, const PeerReduce* collectRho,
    const FusedPut* put, real_t* hostRho, cudaStream_t s)
{
  if (n == 0 && !collectRho) return;
  const bool keep = vectorsFitL2(n), fused = put && put->ndest > 0;
  launchPdl(fused ? (keep ? cgUpdatePKernel<true, true> : cgUpdatePKernel<false, true>)
                  : (keep ? cgUpdatePKernel<true, false> : cgUpdatePKernel<false, false>), dim3((unsigned)vecGrid(n, 2 * kVecUnroll)),
      dim3(kVecThreads), 0, s, n, k, rho, r, p, collectRho ? *collectRho : PeerReduce(), put ? *put : FusedPut(),
      collectRho ? hostRho : (real_t*)nullptr);
  countLaunch();
}

void launchCgUpdateXR(idx_t n, int k, real_t* rho, real_t* pAp, real_t* x, real_t* r, const real_t* p,
    const real_t* Ap, int slot, const PeerReduce* collectPAp, const PeerReduce* pushRho, real_t* hostRho, cudaStream_t s)
{
  Context& c = ctx();
  if (n == 0 && !collectPAp && !pushRho) {
    SB_CUDA(cudaMemsetAsync(rho + k, 0, sizeof(real_t), s));
    if (hostRho) SB_CUDA(cudaMemcpyAsync(hostRho + k, rho + k, sizeof(real_t), cudaMemcpyDeviceToHost, s));
    return;
  }
  // with pushRho the local sum is not the global one yet: the next p update publishes that
  launchPdl(vectorsFitL2(n) ? cgUpdateXRKernel<true> : cgUpdateXRKernel<false>, dim3((unsigned)vecGrid(n, 4)), dim3(kVecThreads), 0,
      s, n, k, rho, pAp, x, r, p, Ap,
      c.partials + (size_t)slot * kMaxPartials, c.tickets + slot, collectPAp ? *collectPAp : PeerReduce(),
      pushRho ? *pushRho : PeerReduce(), pushRho ? (real_t*)nullptr : hostRho);
  countLaunch();
}

// ------------------------------------------------------------------------------------------- setup helpers
// initVectors (CGSolver.c:19-38): x = 0, b = 27 - (rowLen - 1) for generated matrices, else b = 1.
__global__ void initVectorsKernel(idx_t n, const idx_t* __restrict__ rowPtr, const idx_t* __restrict__ rowLen,
    int generated, real_t* __restrict__ x, real_t* __restrict__ b)
{
  for (idx_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int len = rowLen ? (int)rowLen[i] : (int)(rowPtr[i + 1] - rowPtr[i]);
    x[i] = 0.0;
    b[i] = generated ? 27.0 - (real_t)(len - 1) : 1.0;
  }
}

void launchInitVectors(idx_t n, const idx_t* rowPtr, const idx_t* rowLen, bool generated, real_t* x, real_t* b,
    cudaStream_t s)
{
  if (n == 0) return;
  initVectorsKernel<<<vecGrid(n, 1), kVecThreads, 0, s>>>(n, rowPtr, rowLen, generated ? 1 : 0, x, b);
  SB_CUDA(cudaGetLastError());
  countLaunch();
}

// out[map[i]] = in[i]  /  out[i] = in[map[i]]  (SELL row permutation of the CG vectors)
__global__ void scatterKernel(idx_t n, const idx_t* __restrict__ map, const real_t* __restrict__ in, real_t* __restrict__ out)
{
  for (idx_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[map[i]] = in[i];
}
__global__ void gatherKernel(idx_t n, const idx_t* __restrict__ map, const real_t* __restrict__ in, real_t* __restrict__ out)
{
  for (idx_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[map[i]];
}
void launchScatter(idx_t n, const idx_t* map, const real_t* in, real_t* out, cudaStream_t s)
{
  if (n == 0) return;
  scatterKernel<<<vecGrid(n, 1), kVecThreads, 0, s>>>(n, map, in, out);
  SB_CUDA(cudaGetLastError());
  countLaunch();
}
void launchGather(idx_t n, const idx_t* map, const real_t* in, real_t* out, cudaStream_t s)
{
  if (n == 0) return;
  gatherKernel<<<vecGrid(n, 1), kVecThreads, 0, s>>>(n, map, in, out);
  SB_CUDA(cudaGetLastError());
  countLaunch();
}

__global__ void permuteIndicesKernel(idx_t n, const idx_t* __restrict__ map, const int* __restrict__ in, int* __restrict__ out)
{
  for (idx_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = (int)map[in[i]];
}
void launchPermuteIndices(idx_t n, const idx_t* map, const int* in, int* out, cudaStream_t s)
{
  if (n == 0) return;
  permuteIndicesKernel<<<vecGrid(n, 1), kVecThreads, 0, s>>>(n, map, in, out);
  SB_CUDA(cudaGetLastError());
  countLaunch();
}

// max_i |x[i] - 1| (solverCheckResidual, CGSolver.c:40-60 with xexact == 1). NaN never compares greater, as in the
// reference loop (CGSolver.c:50-53). Non-negative doubles order like their bit patterns, so the grid maximum is an
// integer atomicMax.
__global__ void __launch_bounds__(kVecThreads)
maxErrKernel(idx_t n, const real_t* __restrict__ x, unsigned long long* out)
{
  __shared__ double sm[kVecThreads / 32];
  double m = 0.0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double d = fabs((double)(real_t)(x[i] - (real_t)1.0));      // CG_FLOAT diff = fabs(x[i] - xexact[i]), CGSolver.c:50
    if (d > m) m = d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, m, o);
    if (t > m) m = t;
  }
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kVecThreads / 32; w++)
      if (sm[w] > m) m = sm[w];
    atomicMax(out, (unsigned long long)__double_as_longlong(m));
  }
}
void launchMaxErr(idx_t n, const real_t* x, double* out, cudaStream_t s)
{
  SB_CUDA(cudaMemsetAsync(out, 0, sizeof(double), s));
  if (n == 0) return;
  maxErrKernel<<<vecGrid(n, 4), kVecThreads, 0, s>>>(n, x, reinterpret_cast<unsigned long long*>(out));
  SB_CUDA(cudaGetLastError());
  countLaunch();
}

} // namespace sb

using namespace sb;

extern "C" {

void waxpby(const CG_UINT n, const CG_FLOAT alpha, const CG_FLOAT* x, const CG_FLOAT beta, const CG_FLOAT* y,
    CG_FLOAT* w)
{
  ensureOnDevice(x);
  ensureOnDevice(y);
  ensureOnDevice(w);
  launchWaxpby(n, alpha, x, beta, y, w, ctx().stream);
}

void ddot(const CG_UINT n, const CG_FLOAT* x, const CG_FLOAT* y, CG_FLOAT* result)
{
  Context& c = ctx();
  ensureOnDevice(x);
  ensureOnDevice(y);
  launchDot(n, x, y, c.dScalar, 0, c.stream);
  SB_CUDA(cudaMemcpyAsync(c.hScalar, c.dScalar, sizeof(real_t), cudaMemcpyDeviceToHost, c.stream));
  SB_CUDA(cudaStreamSynchronize(c.stream));
  real_t sum = c.hScalar[0];
  commReduction(&sum, SB_SUM);      // solver.c:60
  *result = sum;
}

} // extern "C"
