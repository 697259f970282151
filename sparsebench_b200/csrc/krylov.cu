// The two solver types the reference names but never implemented (main.c:22 `enum { CG, SPMV, GMRES, CHEBFD }`; the
// GMRES case of main.c:217-222 prints its name and does nothing, CHEBFD has no case at all), built on the same SpMV
// kernels, halo exchange and deterministic reductions as the CG. There is no reference behaviour to be identical to:
// parity is pinned by a numpy restatement of the same algorithms (oracle/krylov_ref.py) and by the mathematical
// invariants (true residual, Chebyshev recurrence against dense T_k).
//
//   sbSolveGMRES        restarted GMRES(m): Arnoldi with classical Gram-Schmidt in TWO vector passes per step -- one
//                       kernel computes all j+1 projections v_i.w while reading w once, one kernel subtracts them and
//                       accumulates ||w||^2 -- Givens rotations of the (m+1) x m Hessenberg matrix on the host
//   sbChebyshevFilter   y = sum_k c_k T_k(A~) x and the moments mu_k = x . T_k(A~) x, A~ = (A - c I)/e mapped to [-1, 1]:
//                       the kernel of Chebyshev filter diagonalisation / the kernel polynomial method; per degree one
//                       SpMV and ONE fused vector pass (recurrence + filter accumulation + moment)
#include <math.h>
#include <string.h>

#include <vector>

#include "device_utils.cuh"
#include "sb_internal.h"

namespace sb {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBasis = 64;            // restart length limit (m + 1 basis vectors)
constexpr int kDotGroup = 8;             // projections computed per pass over w in registers

int gridFor(uint64_t n, int perThread)
{
  Context& c = ctx();
  uint64_t blocks = (n + (uint64_t)kThreads * perThread - 1) / ((uint64_t)kThreads * perThread);
  uint64_t cap = (uint64_t)c.numSMs * 8;
  if (blocks > cap) blocks = cap;
  if (blocks > (uint64_t)kMaxPartials) blocks = kMaxPartials;
  return (int)(blocks ? blocks : 1);
}

struct BasisPtrs {
  const real_t* v[kMaxBasis];
};

// 16 bytes of a vector (2 doubles / 4 floats): every pass below moves the vectors with 128-bit accesses, two per
// thread and vector in flight (the buffers come from sbAllocateDevice: 256-byte aligned)
constexpr int kPL = 16 / sizeof(real_t);
struct alignas(16) Pack {
  real_t v[kPL];
};

// acc[g] += V_{g0+g} . w for G basis vectors in one pass over w
template <int G>
__device__ __forceinline__ void dotGroup(idx_t n, const BasisPtrs& V, int g0, const real_t* __restrict__ w, real_t (&acc)[kDotGroup])
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t n2 = n / kPL;
  const Pack* w2 = reinterpret_cast<const Pack*>(w);
  const Pack* v2[G];
#pragma unroll
  for (int g = 0; g < G; g++) v2[g] = reinterpret_cast<const Pack*>(V.v[g0 + g]);
  for (uint64_t i = tid; i < n2; i += 2 * stride) {
    const bool second = i + stride < n2;
    const uint64_t i1 = second ? i + stride : i;
    const Pack wa = w2[i], wb = w2[i1];
    Pack va[G], vb[G];
#pragma unroll
    for (int g = 0; g < G; g++) {
      va[g] = v2[g][i];
      vb[g] = v2[g][i1];
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
#pragma unroll
      for (int c = 0; c < kPL; c++) acc[g] = fma(va[g].v[c], wa.v[c], acc[g]);
      if (second) {
#pragma unroll
        for (int c = 0; c < kPL; c++) acc[g] = fma(vb[g].v[c], wb.v[c], acc[g]);
      }
    }
  }
  if (tid == 0)
    for (uint64_t i = n2 * kPL; i < n; i++) {
#pragma unroll
      for (int g = 0; g < G; g++) acc[g] = fma(V.v[g0 + g][i], w[i], acc[g]);
    }
}

// out[i] = V_i . w for i in [0, count): w is read once per group of kDotGroup basis vectors; deterministic: every
// block deposits its `count` partial sums, the last block adds them in block order.
__global__ void __launch_bounds__(kThreads)
multiDotKernel(idx_t n, int count, BasisPtrs V, const real_t* __restrict__ w, real_t* partials, unsigned int* ticket, real_t* out)
{
  __shared__ real_t scratch[32];
  __shared__ bool amLast;
  for (int g0 = 0; g0 < count; g0 += kDotGroup) {
    real_t acc[kDotGroup];
#pragma unroll
    for (int g = 0; g < kDotGroup; g++) acc[g] = 0.0;
    switch (count - g0 < kDotGroup ? count - g0 : kDotGroup) {
      case 1: dotGroup<1>(n, V, g0, w, acc); break;
      case 2: dotGroup<2>(n, V, g0, w, acc); break;
      case 3: dotGroup<3>(n, V, g0, w, acc); break;
      case 4: dotGroup<4>(n, V, g0, w, acc); break;
      case 5: dotGroup<5>(n, V, g0, w, acc); break;
      case 6: dotGroup<6>(n, V, g0, w, acc); break;
      case 7: dotGroup<7>(n, V, g0, w, acc); break;
      default: dotGroup<8>(n, V, g0, w, acc); break;
    }
#pragma unroll
    for (int g = 0; g < kDotGroup; g++) {
      const real_t b = blockSum(acc[g], scratch);
      if (threadIdx.x == 0 && g0 + g < count) partials[(size_t)blockIdx.x * kMaxBasis + g0 + g] = b;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicInc(ticket, gridDim.x - 1);
    amLast = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (amLast) {
    // one warp per projection: lanes stride over the blocks, fixed shuffle tree (same bits for the same grid)
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = warp; i < count; i += kThreads / 32) {
      real_t s = 0.0;
      for (unsigned int b = lane; b < gridDim.x; b += 32) s += __ldcg(partials + (size_t)b * kMaxBasis + i);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (lane == 0) out[i] = s;
    }
  }
}

// w -= sum_i h[i] V_i  and  *norm2 = w . w  (second half of a classical Gram-Schmidt step); the projections are
// subtracted one after the other in index order, each product rounded (as the numpy restatement does)
__global__ void __launch_bounds__(kThreads)
projectOutKernel(idx_t n, int count, BasisPtrs V, const real_t* __restrict__ h, real_t* __restrict__ w, real_t* partials,
    unsigned int* ticket, real_t* norm2)
{
  __shared__ real_t scratch[32];
  __shared__ real_t hs[kMaxBasis];
  for (int i = threadIdx.x; i < count; i += blockDim.x) hs[i] = h[i];
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t n2 = n / kPL;
  Pack* w2 = reinterpret_cast<Pack*>(w);
  real_t acc = 0.0;
  for (uint64_t i = tid; i < n2; i += 2 * stride) {
    const bool second = i + stride < n2;
    const uint64_t i1 = second ? i + stride : i;
    Pack wa = w2[i], wb = w2[i1];
    int k = 0;
    for (; k + 4 <= count; k += 4) {                        // 8 independent 16-byte loads in flight
      Pack va[4], vb[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        va[u] = reinterpret_cast<const Pack*>(V.v[k + u])[i];
        vb[u] = reinterpret_cast<const Pack*>(V.v[k + u])[i1];
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int c = 0; c < kPL; c++) {
          wa.v[c] = addRn(wa.v[c], -mulRn(hs[k + u], va[u].v[c]));
          wb.v[c] = addRn(wb.v[c], -mulRn(hs[k + u], vb[u].v[c]));
        }
    }
    for (; k < count; k++) {
      const Pack va = reinterpret_cast<const Pack*>(V.v[k])[i], vb = reinterpret_cast<const Pack*>(V.v[k])[i1];
#pragma unroll
      for (int c = 0; c < kPL; c++) {
        wa.v[c] = addRn(wa.v[c], -mulRn(hs[k], va.v[c]));
        wb.v[c] = addRn(wb.v[c], -mulRn(hs[k], vb.v[c]));
      }
    }
    w2[i] = wa;
#pragma unroll
    for (int c = 0; c < kPL; c++) acc = fma(wa.v[c], wa.v[c], acc);
    if (second) {
      w2[i1] = wb;
#pragma unroll
      for (int c = 0; c < kPL; c++) acc = fma(wb.v[c], wb.v[c], acc);
    }
  }
  if (tid == 0)
    for (uint64_t i = n2 * kPL; i < n; i++) {
      real_t wi = w[i];
      for (int k = 0; k < count; k++) wi = addRn(wi, -mulRn(hs[k], V.v[k][i]));
      w[i] = wi;
      acc = fma(wi, wi, acc);
    }
  const real_t b = blockSum(acc, scratch);
  gridSum(b, partials, ticket, norm2, false, scratch);
}

// out = in * (1 / sqrt(*norm2))   (next basis vector); a zero norm (lucky breakdown) leaves zeros
__global__ void __launch_bounds__(kThreads)
normalizeKernel(idx_t n, const real_t* __restrict__ in, const real_t* __restrict__ norm2, real_t* __restrict__ out)
{
  const real_t nrm = sqrt(*norm2);
  const real_t inv = nrm > (real_t)0.0 ? (real_t)1.0 / nrm : (real_t)0.0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = in[i] * inv;
}

// x += sum_i y[i] V_i   (solution update at the end of a restart cycle)
__global__ void __launch_bounds__(kThreads)
combineKernel(idx_t n, int count, BasisPtrs V, const real_t* __restrict__ y, real_t* __restrict__ x)
{
  __shared__ real_t ys[kMaxBasis];
  for (int i = threadIdx.x; i < count; i += blockDim.x) ys[i] = y[i];
  __syncthreads();
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    real_t xi = x[i];
    for (int k = 0; k < count; k++) xi = fma(ys[k], V.v[k][i], xi);
    x[i] = xi;
  }
}

// One Chebyshev step after the SpMV q = A t:  tNext = a (q - c t) - tPrev  (a = 2/e, or 1/e for the first step with
// tPrev ignored),  y += coef * tNext,  *moment = x0 . tNext -- one pass, everything that needs tNext fused.
__global__ void __launch_bounds__(kThreads)
chebStepKernel(idx_t n, real_t a, real_t c, int first, const real_t* __restrict__ q, const real_t* t,
    const real_t* tPrev, real_t* tNext, real_t coef, real_t* __restrict__ y, const real_t* x0, real_t* partials,
    unsigned int* ticket, real_t* moment)
{
  __shared__ real_t scratch[32];
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t n2 = n / kPL;
  real_t acc = 0.0;
  auto step = [&](real_t qi, real_t ti, real_t pi) { return first ? a * (qi - c * ti) : a * (qi - c * ti) - pi; };
  for (uint64_t i = tid; i < n2; i += stride) {
    const Pack q2 = reinterpret_cast<const Pack*>(q)[i], t2 = reinterpret_cast<const Pack*>(t)[i];
    const Pack x2 = reinterpret_cast<const Pack*>(x0)[i];
    Pack p2 = t2, y2 = t2, o;
    if (!first) p2 = reinterpret_cast<const Pack*>(tPrev)[i];      // tNext may alias tPrev: read before written, same thread
    if (y) y2 = reinterpret_cast<const Pack*>(y)[i];
#pragma unroll
    for (int k = 0; k < kPL; k++) {
      o.v[k] = step(q2.v[k], t2.v[k], p2.v[k]);
      y2.v[k] = fma(coef, o.v[k], y2.v[k]);
      acc = fma(x2.v[k], o.v[k], acc);
    }
    reinterpret_cast<Pack*>(tNext)[i] = o;
    if (y) reinterpret_cast<Pack*>(y)[i] = y2;
  }
  if (tid == 0)
    for (uint64_t i = n2 * kPL; i < n; i++) {
      const real_t v = step(q[i], t[i], first ? (real_t)0.0 : tPrev[i]);
      tNext[i] = v;
      if (y) y[i] = fma(coef, v, y[i]);
      acc = fma(x0[i], v, acc);
    }
  const real_t b = blockSum(acc, scratch);
  gridSum(b, partials, ticket, moment, false, scratch);
}

// y = coef * x  and  *moment = x . x   (degree 0)
__global__ void __launch_bounds__(kThreads)
chebInitKernel(idx_t n, real_t coef, const real_t* __restrict__ x, real_t* __restrict__ y, real_t* partials, unsigned int* ticket,
    real_t* moment)
{
  __shared__ real_t scratch[32];
  real_t acc = 0.0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const real_t v = x[i];
    if (y) y[i] = coef * v;
    acc = fma(v, v, acc);
  }
  const real_t b = blockSum(acc, scratch);
  gridSum(b, partials, ticket, moment, false, scratch);
}

bool commActive(const Comm* c) { return c && c->size > 1; }

// caller vector (host or device, original row order) <-> device vector in solver order (SELL: permuted rows)
void importVector(const Operator& A, const real_t* src, real_t* dst, real_t* stage, cudaStream_t s)
{
  const size_t bytes = sizeof(real_t) * A.nr;
  const bool dev = isDevicePointer(src);
  if (dev) ensureOnDevice(src);
  if (!A.oldToNew) {
    SB_CUDA(cudaMemcpyAsync(dst, src, bytes, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  } else {
    const real_t* staged = src;
    if (!dev) {
      SB_CUDA(cudaMemcpyAsync(stage, src, bytes, cudaMemcpyHostToDevice, s));
      staged = stage;
    }
    launchScatter(A.nr, A.oldToNew, staged, dst, s);
  }
}

void exportVector(const Operator& A, const real_t* src, real_t* dst, real_t* stage, cudaStream_t s)
{
  const size_t bytes = sizeof(real_t) * A.nr;
  const bool dev = isDevicePointer(dst);
  if (!A.oldToNew) {
    SB_CUDA(cudaMemcpyAsync(dst, src, bytes, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  } else if (dev) {
    launchGather(A.nr, A.oldToNew, src, dst, s);
  } else {
    launchGather(A.nr, A.oldToNew, src, stage, s);
    SB_CUDA(cudaMemcpyAsync(dst, stage, bytes, cudaMemcpyDeviceToHost, s));
  }
}

// y = A v with the halo of v exchanged first (v has colSlots entries)
void applyOperator(Comm* comm, const Operator& A, const int* elems, real_t* v, real_t* y, cudaStream_t s)
{
  if (commActive(comm)) commExchangeOnStream(comm, A.nr, v, elems, s);
  launchSpmv(A, v, y, 0, spmvUnits(A), nullptr, s);
}

// sum over the ranks of `count` device scalars (a no-op on one rank)
void allreduce(Comm* comm, real_t* d, int count, cudaStream_t s)
{
  if (commActive(comm)) commAllreduceDevice(comm, d, count, SB_SUM, s);
}

} // namespace

} // namespace sb

using namespace sb;

extern "C" {

int sbSolveGMRES(Comm* comm, Parameter* param, void* matrix, int fmt, SbCGInfo* info, int restart)
{
  Context& c = ctx();
  cudaStream_t s = c.stream;
  const Operator A = makeOperator(matrix, fmt);
  const idx_t n = A.nr;
  const int m = restart < 1 ? 30 : restart;
  if (m + 1 > kMaxBasis) SB_FATAL("sbSolveGMRES: restart length %d exceeds the supported %d", m, kMaxBasis - 1);
  const int itermax = param->itermax;
  const double eps = (double)(real_t)param->eps;
  const bool print = info ? (info->flags & SB_CG_PRINT) != 0 && (!comm || comm->rank == 0) : (!comm || comm->rank == 0);
  const bool generated = param->filename && (strcmp(param->filename, "generate") == 0 || strcmp(param->filename, "generate7P") == 0);
  const size_t rowSlots = (size_t)(A.nrPadded > n ? A.nrPadded : n) + 2;
  const size_t colSlots = (A.nc > rowSlots ? (size_t)A.nc : rowSlots) + 2;
  const int* elems = commActive(comm) ? commSolverElements(comm, A.permKey, A.oldToNew, s) : nullptr;

  // basis V_0 .. V_m (each with a halo part: every one of them is multiplied by A), w, x, b, stage
  std::vector<real_t*> V((size_t)m + 1);
  for (auto& v : V) {
    v = (real_t*)sbAllocateDevice(64, sizeof(real_t) * colSlots);
    SB_CUDA(cudaMemsetAsync(v, 0, sizeof(real_t) * colSlots, s));
  }
  real_t* w = (real_t*)sbAllocateDevice(64, sizeof(real_t) * rowSlots);
  real_t* x = (real_t*)sbAllocateDevice(64, sizeof(real_t) * colSlots);
  real_t* b = (real_t*)sbAllocateDevice(64, sizeof(real_t) * rowSlots);
  real_t* stage = (real_t*)sbAllocateDevice(64, sizeof(real_t) * rowSlots);
  constexpr size_t kSlot = kMaxBasis + 2;                                  // projections + ||w||^2 of one step
  real_t* dH = (real_t*)sbAllocateDevice(64, sizeof(real_t) * kSlot * (size_t)(m + 1));
  real_t* hH = (real_t*)sbAllocateHost(sizeof(real_t) * kSlot * (size_t)(m + 1));
  std::vector<cudaEvent_t> stepDone((size_t)m);
  for (auto& e : stepDone) SB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  real_t* dY = (real_t*)sbAllocateDevice(64, sizeof(real_t) * kMaxBasis);
  SB_CUDA(cudaMemsetAsync(x, 0, sizeof(real_t) * colSlots, s));
  launchInitVectors(n, A.rowPtr, A.rowLen, generated, x, b, s);            // x = 0, b = the CG's right-hand side rule
  if (info && info->b) importVector(A, info->b, b, stage, s);
  if (info && info->x) importVector(A, info->x, x, stage, s);
  BasisPtrs ptrs;
  for (int i = 0; i < kMaxBasis; i++) ptrs.v[i] = V[(size_t)(i <= m ? i : m)];
  const int gridV = gridFor(n, 4);
  real_t* partials = c.partials;                                           // reduction slot 3 (one value per block)
  real_t* partialsK = (real_t*)sbAllocateDevice(64, sizeof(real_t) * (size_t)gridV * kMaxBasis);   // multi-dot: kMaxBasis values per block

  std::vector<double> hist;
  std::vector<double> H((size_t)(m + 1) * m), cs((size_t)m), sn((size_t)m), g((size_t)m + 1), y((size_t)m);
  cudaEvent_t t0, t1;
  SB_CUDA(cudaEventCreate(&t0));
  SB_CUDA(cudaEventCreate(&t1));
  SB_CUDA(cudaEventRecord(t0, s));
  int k = 0;                                                              // total Arnoldi steps (matrix-vector products)
  double resid = 0.0;
  bool done = false;
  while (!done) {
    // r0 = b - A x  ->  V_0 = r0 / beta
    applyOperator(comm, A, elems, x, w, s);
    launchWaxpby(n, 1.0, b, -1.0, w, w, s);
    launchDot(n, w, w, dH, 3, s);
    allreduce(comm, dH, 1, s);
    normalizeKernel<<<gridV, kThreads, 0, s>>>(n, w, dH, V[0]);
    countLaunch();
    SB_CUDA(cudaMemcpyAsync(hH, dH, sizeof(real_t), cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    const double beta = sqrt((double)hH[0]);
    if (hist.empty()) {
      hist.push_back(beta);
      if (print) printf("Initial Residual = %E\n", beta);
    }
    resid = beta;
    if (!(beta > eps) || k >= itermax - 1) break;
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    // The device never waits for the host inside a cycle: step e+1 is enqueued before the host reads the Hessenberg
    // column of step e (own slot of dH/hH, own event), applies the rotations and tests the residual. On convergence
    // the one step already enqueued is simply not used (x += V y takes the first j columns).
    auto enqueueStep = [&](int e) {
      real_t* d = dH + (size_t)e * kSlot;
      applyOperator(comm, A, elems, V[(size_t)e], w, s);                   // w = A v_e
      multiDotKernel<<<gridV, kThreads, 0, s>>>(n, e + 1, ptrs, w, partialsK, c.tickets + 0, d);
      countLaunch();
      allreduce(comm, d, e + 1, s);
      projectOutKernel<<<gridV, kThreads, 0, s>>>(n, e + 1, ptrs, d, w, partials + 3 * (size_t)kMaxPartials, c.tickets + 3, d + e + 1);
      countLaunch();
      allreduce(comm, d + e + 1, 1, s);
      normalizeKernel<<<gridV, kThreads, 0, s>>>(n, w, d + e + 1, V[(size_t)e + 1]);
      countLaunch();
      SB_CUDA(cudaMemcpyAsync(hH + (size_t)e * kSlot, d, sizeof(real_t) * (size_t)(e + 2), cudaMemcpyDeviceToHost, s));
      SB_CUDA(cudaEventRecord(stepDone[(size_t)e], s));
    };
    // column e of the Hessenberg matrix, previous rotations, new rotation (host, double); true: converged
    auto absorbStep = [&](int e) {
      SB_CUDA(cudaEventSynchronize(stepDone[(size_t)e]));
      const real_t* h = hH + (size_t)e * kSlot;
      k++;
      double* col = &H[(size_t)e * (m + 1)];
      for (int i = 0; i <= e; i++) col[i] = (double)h[i];
      col[e + 1] = sqrt((double)h[e + 1]);
      for (int i = 0; i < e; i++) {
        const double a = cs[(size_t)i] * col[i] + sn[(size_t)i] * col[i + 1];
        col[i + 1] = -sn[(size_t)i] * col[i] + cs[(size_t)i] * col[i + 1];
        col[i] = a;
      }
      const double d = hypot(col[e], col[e + 1]);
      cs[(size_t)e] = d > 0.0 ? col[e] / d : 1.0;
      sn[(size_t)e] = d > 0.0 ? col[e + 1] / d : 0.0;
      col[e] = d;
      col[e + 1] = 0.0;
      g[(size_t)e + 1] = -sn[(size_t)e] * g[(size_t)e];
      g[(size_t)e] = cs[(size_t)e] * g[(size_t)e];
      resid = fabs(g[(size_t)e + 1]);
      hist.push_back(resid);
      if (print) printf("Iteration = %d Residual = %E\n", k, resid);
      return !(resid > eps);
    };
    int j = 0;                                                            // columns absorbed in this cycle
    int enq = 0;                                                          // steps enqueued in this cycle
    while (!done && j < m && k < itermax - 1) {
      while (enq < m && enq <= j + 1 && k + (enq - j) < itermax - 1) enqueueStep(enq++);
      if (absorbStep(j++)) done = true;
    }
    if (k >= itermax - 1) done = true;
    // y = R^-1 g (back substitution), x += V y
    for (int i = j - 1; i >= 0; i--) {
      double acc = g[(size_t)i];
      for (int l = i + 1; l < j; l++) acc -= H[(size_t)l * (m + 1) + i] * y[(size_t)l];
      const double diag = H[(size_t)i * (m + 1) + i];
      y[(size_t)i] = diag != 0.0 ? acc / diag : 0.0;
    }
    if (j > 0) {
      for (int i = 0; i < j; i++) hH[i] = (real_t)y[(size_t)i];
      SB_CUDA(cudaMemcpyAsync(dY, hH, sizeof(real_t) * (size_t)j, cudaMemcpyHostToDevice, s));
      combineKernel<<<gridV, kThreads, 0, s>>>(n, j, ptrs, dY, x);
      countLaunch();
      SB_CUDA(cudaStreamSynchronize(s));
    }
  }
  SB_CUDA(cudaEventRecord(t1, s));
  SB_CUDA(cudaEventSynchronize(t1));
  float ms = 0.f;
  SB_CUDA(cudaEventElapsedTime(&ms, t0, t1));
  SB_CUDA(cudaEventDestroy(t0));
  SB_CUDA(cudaEventDestroy(t1));
  if (print) printf("Solution performed %d iterations and took %.2fs\n", k, ms * 1e-3);
  double maxErr = -1.0;
  if (generated) {
    launchMaxErr(n, x, c.dWide, s);
    SB_CUDA(cudaMemcpyAsync(c.hWide, c.dWide, sizeof(double), cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    CG_FLOAT worst = (CG_FLOAT)c.hWide[0];
    commReduction(&worst, SB_MAX);
    maxErr = (double)worst;
    if (print) printf("Difference between computed and exact  = %f\n", maxErr);
  }
  if (info) {
    if (info->x) exportVector(A, x, info->x, stage, s);
    SB_CUDA(cudaStreamSynchronize(s));
    info->nhist = (int)hist.size();
    if (info->history)
      for (int i = 0; i < info->nhist && i < info->historyCap; i++) info->history[i] = hist[(size_t)i];
    info->solveMs = ms;
    info->maxError = maxErr;
  }
  for (auto v : V) sbFree(v);
  sbFree(w); sbFree(x); sbFree(b); sbFree(stage); sbFree(dH); sbFree(dY); sbFree(partialsK);
  sbFreeHost(hH);
  for (auto e : stepDone) SB_CUDA(cudaEventDestroy(e));
  return k;
}

void sbChebyshevFilter(Comm* comm, void* matrix, int fmt, int degree, double lambdaMin, double lambdaMax, const CG_FLOAT* coef,
    const CG_FLOAT* xIn, CG_FLOAT* yOut, CG_FLOAT* moments)
{
  Context& c = ctx();
  cudaStream_t s = c.stream;
  const Operator A = makeOperator(matrix, fmt);
  const idx_t n = A.nr;
  if (degree < 0 || !(lambdaMax > lambdaMin)) SB_FATAL("sbChebyshevFilter: need degree >= 0 and lambdaMax > lambdaMin");
  const real_t cc = (real_t)(0.5 * (lambdaMax + lambdaMin)), e = (real_t)(0.5 * (lambdaMax - lambdaMin));
  const size_t rowSlots = (size_t)(A.nrPadded > n ? A.nrPadded : n) + 2;
  const size_t colSlots = (A.nc > rowSlots ? (size_t)A.nc : rowSlots) + 2;
  const int* elems = commActive(comm) ? commSolverElements(comm, A.permKey, A.oldToNew, s) : nullptr;
  // xs = x in solver order (kept: every moment is x . t_k), two work vectors for the three-term recurrence; all of
  // them get multiplied by A, so all have a halo part
  real_t* xs = (real_t*)sbAllocateDevice(64, sizeof(real_t) * colSlots);
  real_t* work[2];
  for (auto& v : work) v = (real_t*)sbAllocateDevice(64, sizeof(real_t) * colSlots);
  SB_CUDA(cudaMemsetAsync(xs, 0, sizeof(real_t) * colSlots, s));
  SB_CUDA(cudaMemsetAsync(work[0], 0, sizeof(real_t) * colSlots, s));
  SB_CUDA(cudaMemsetAsync(work[1], 0, sizeof(real_t) * colSlots, s));
  real_t* q = (real_t*)sbAllocateDevice(64, sizeof(real_t) * rowSlots);
  real_t* y = yOut ? (real_t*)sbAllocateDevice(64, sizeof(real_t) * rowSlots) : nullptr;
  real_t* stage = (real_t*)sbAllocateDevice(64, sizeof(real_t) * rowSlots);
  real_t* dMu = (real_t*)sbAllocateDevice(64, sizeof(real_t) * (size_t)(degree + 2));
  importVector(A, xIn, xs, stage, s);
  const int grid = gridFor(n, 4);
  real_t* partials = c.partials + 3 * (size_t)kMaxPartials;
  auto coefAt = [&](int k) { return coef ? (real_t)coef[k] : (k == degree ? (real_t)1.0 : (real_t)0.0); };   // default: y = T_degree(A~) x
  chebInitKernel<<<grid, kThreads, 0, s>>>(n, coefAt(0), xs, y, partials, c.tickets + 3, dMu);
  countLaunch();
  real_t* prev = xs;                   // t_{k-2}
  real_t* cur = xs;                    // t_{k-1}
  for (int k = 1; k <= degree; k++) {
    applyOperator(comm, A, elems, cur, q, s);
    // t_1 -> work[0]; t_2 -> work[1] (t_0 = xs must survive); from t_3 on in place over t_{k-2}
    real_t* dst = k == 1 ? work[0] : k == 2 ? work[1] : prev;
    chebStepKernel<<<grid, kThreads, 0, s>>>(n, k == 1 ? (real_t)1.0 / e : (real_t)2.0 / e, cc, k == 1 ? 1 : 0, q, cur, prev, dst, coefAt(k), y,
        xs, partials, c.tickets + 3, dMu + k);
    countLaunch();
    prev = cur;
    cur = dst;
  }
  allreduce(comm, dMu, degree + 1, s);
  if (moments) {
    if (isDevicePointer(moments)) SB_CUDA(cudaMemcpyAsync(moments, dMu, sizeof(real_t) * (size_t)(degree + 1), cudaMemcpyDeviceToDevice, s));
    else SB_CUDA(cudaMemcpyAsync(moments, dMu, sizeof(real_t) * (size_t)(degree + 1), cudaMemcpyDeviceToHost, s));
  }
  if (yOut) exportVector(A, y, yOut, stage, s);
  SB_CUDA(cudaStreamSynchronize(s));
  sbFree(xs); sbFree(work[0]); sbFree(work[1]);
  sbFree(q); sbFree(y); sbFree(stage); sbFree(dMu);
}

} // extern "C"
