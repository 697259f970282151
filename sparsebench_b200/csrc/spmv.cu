// Sparse matrix-vector kernels of the three format plugins (replace spMVM of matrix-CRS.c:46-65,
// matrix-SCS.c:198-228, matrix-CCRS.c:14-31). HBM-bound, no tensor cores: 12 B (16 B for CCRS) of matrix
// stream per non-zero against 2 flops; x is gathered through L1/L2, y written once.
//
// Common shape: persistent grids of numSMs * residentBlocks CTAs walk the row/chunk range in a
// round-robin so that, at any instant, the whole chip works on one contiguous window of the matrix (and
// of x). An optional fused epilogue accumulates sum_i x[i]*y[i] (the CG's p.Ap, CGSolver.c:125) with the
// deterministic one-kernel grid reduction from device_utils.cuh.
#include "device_utils.cuh"
#include "sb_internal.h"

namespace sb {

// ------------------------------------------------------------------------------------------- SELL-32-sigma
// One warp per chunk, lane = row of the chunk: the j-th column of a chunk is 32 consecutive values
// (256 B) and 32 consecutive column ids (128 B), i.e. perfectly coalesced. Each lane sums its row in
// stored order, exactly like tmp[k] += val*x of matrix-SCS.c:216-222.
constexpr int kSellUnroll = 8;

template <bool DOT>
__global__ void __launch_bounds__(256, 4)
spmvSell32Kernel(SellView A, const double* __restrict__ x, double* __restrict__ y, uint32_t lo, uint32_t hi,
    double* partials, unsigned int* ticket, double* dotOut, bool accumulate)
{
  __shared__ double scratch[32];
  const int lane = threadIdx.x & 31;
  const uint32_t warpsPerBlock = blockDim.x >> 5;
  double dotAcc = 0.0;
  for (uint64_t chunk = (uint64_t)lo + blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); chunk < hi;
       chunk += (uint64_t)gridDim.x * warpsPerBlock) {
    const uint64_t base = (uint64_t)A.chunkPtr[chunk] + lane;
    const uint32_t len = A.chunkLens[chunk];
    const double* __restrict__ v = A.val + base;
    const uint32_t* __restrict__ c = A.col + base;
    double sum = 0.0;
    uint32_t j = 0;
    for (; j + kSellUnroll <= len; j += kSellUnroll) {
      uint32_t cc[kSellUnroll];
      double vv[kSellUnroll], xx[kSellUnroll];
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
      uint32_t cc[kSellUnroll];
      double vv[kSellUnroll], xx[kSellUnroll];
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
    y[row] = sum;                                  // padded rows are stored too (matrix-SCS.c:224-226)
    if (DOT && row < A.nr) dotAcc = fma(sum, __ldg(x + row), dotAcc);
  }
  if (DOT) {
    const double b = blockSum(dotAcc, scratch);
    gridSum(b, partials, ticket, dotOut, accumulate, scratch);
  }
}

// Any other chunk height (the reference's tests use C = 1, 2, 4): one thread per padded row, same
// per-row summation order. Correctness path, not tuned.
template <bool DOT>
__global__ void __launch_bounds__(256)
spmvSellAnyCKernel(SellView A, const double* __restrict__ x, double* __restrict__ y, uint32_t lo, uint32_t hi,
    double* partials, unsigned int* ticket, double* dotOut, bool accumulate)
{
  __shared__ double scratch[32];
  double dotAcc = 0.0;
  const uint64_t rowLo = (uint64_t)lo * A.C, rowHi = (uint64_t)hi * A.C;
  for (uint64_t row = rowLo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; row < rowHi;
       row += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t chunk = row / A.C;
    const uint64_t base = (uint64_t)A.chunkPtr[chunk] + row % A.C;
    const uint32_t len = A.chunkLens[chunk];
    double sum = 0.0;
    for (uint32_t j = 0; j < len; j++) {
      const uint64_t e = base + (uint64_t)j * A.C;
      sum = mulAdd(sum, A.val[e], __ldg(x + A.col[e]));
    }
    y[row] = sum;
    if (DOT && row < A.nr) dotAcc = fma(sum, __ldg(x + row), dotAcc);
  }
  if (DOT) {
    const double b = blockSum(dotAcc, scratch);
    gridSum(b, partials, ticket, dotOut, accumulate, scratch);
  }
}

// ------------------------------------------------------------------------------------------- CRS / CCRS
// LANES consecutive lanes share a row (27-point rows: 8 lanes x 4 strided elements), partial sums are
// combined with xor-shuffles. Row-length independent; coalescing comes from neighbouring rows being
// neighbouring in memory.
struct CrsAccess {
  const uint32_t* col;
  const double* val;
  __device__ __forceinline__ void load(uint64_t j, uint32_t& c, double& v) const
  {
    c = ldStream(col + j);
    v = ldStream(val + j);
  }
};
struct CcrsAccess {
  const Entry* entries;
  __device__ __forceinline__ void load(uint64_t j, uint32_t& c, double& v) const
  {
    const double2 e = ldStream2(reinterpret_cast<const double*>(entries + j));   // one 16-byte record
    c = (uint32_t)__double_as_longlong(e.x);
    v = e.y;
  }
};

template <int LANES, bool DOT, typename Access>
__global__ void __launch_bounds__(256, 4)
spmvRowsKernel(Access acc, const uint32_t* __restrict__ rowPtr, uint32_t nrTotal, const double* __restrict__ x,
    double* __restrict__ y, uint32_t lo, uint32_t hi, double* partials, unsigned int* ticket, double* dotOut,
    bool accumulate)
{
  __shared__ double scratch[32];
  constexpr int kUnroll = 4;
  const int sub = threadIdx.x % LANES;
  const uint32_t groupsPerBlock = blockDim.x / LANES;
  double dotAcc = 0.0;
  for (uint64_t first = (uint64_t)lo + (uint64_t)blockIdx.x * groupsPerBlock; first < hi;
       first += (uint64_t)gridDim.x * groupsPerBlock) {
    const uint64_t row = first + threadIdx.x / LANES;
    const bool live = row < hi;
    uint64_t j = 0, end = 0;
    if (live) {
      j = (uint64_t)__ldg(rowPtr + row) + sub;
      end = __ldg(rowPtr + row + 1);
    }
    double sum = 0.0;
    while (j < end) {
      uint32_t cc[kUnroll];
      double vv[kUnroll], xx[kUnroll];
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
    const double b = blockSum(dotAcc, scratch);
    gridSum(b, partials, ticket, dotOut, accumulate, scratch);
  }
}

template <int LANES, typename Access>
static void launchRows(Access acc, const uint32_t* rowPtr, uint32_t nr, const double* x, double* y, uint32_t lo,
    uint32_t hi, const DotArgs* dot, cudaStream_t s)
{
  Context& c = ctx();
  const uint32_t groupsPerBlock = 256 / LANES;
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

template <typename Access>
static void launchRowsAuto(Access acc, const uint32_t* rowPtr, uint32_t nr, uint64_t nnz, const double* x, double* y,
    uint32_t lo, uint32_t hi, const DotArgs* dot, cudaStream_t s)
{
  const double avg = nr ? (double)nnz / (double)nr : 0.0;
  if (avg <= 6.0) launchRows<2, Access>(acc, rowPtr, nr, x, y, lo, hi, dot, s);
  else if (avg <= 12.0) launchRows<4, Access>(acc, rowPtr, nr, x, y, lo, hi, dot, s);
  else if (avg <= 40.0) launchRows<8, Access>(acc, rowPtr, nr, x, y, lo, hi, dot, s);
  else if (avg <= 96.0) launchRows<16, Access>(acc, rowPtr, nr, x, y, lo, hi, dot, s);
  else launchRows<32, Access>(acc, rowPtr, nr, x, y, lo, hi, dot, s);
}

uint32_t spmvUnits(const Operator& A) { return A.fmt == SB_FMT_SCS ? A.sell.nChunks : A.nr; }

void launchSpmv(const Operator& A, const double* x, double* y, uint32_t lo, uint32_t hi, const DotArgs* dot,
    cudaStream_t s)
{
  if (hi <= lo) {
    if (dot && !dot->accumulate) SB_CUDA(cudaMemsetAsync(dot->out, 0, sizeof(double), s));
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
    launchRowsAuto(CrsAccess { A.crs.col, A.crs.val }, A.crs.rowPtr, A.nr, A.nnzTrue, x, y, lo, hi, dot, s);
  } else {
    launchRowsAuto(CcrsAccess { A.ccrs.entries }, A.ccrs.rowPtr, A.nr, A.nnzTrue, x, y, lo, hi, dot, s);
  }
}

} // namespace sb

using namespace sb;

extern "C" {

void sbCRS_spMVM(SbCRSMatrix* m, const CG_FLOAT* x, CG_FLOAT* y)
{
  Operator A = makeOperator(m, SB_FMT_CRS);
  launchSpmv(A, x, y, 0, A.nr, nullptr, ctx().stream);
}

void sbCCRS_spMVM(SbCCRSMatrix* m, const CG_FLOAT* x, CG_FLOAT* y)
{
  Operator A = makeOperator(m, SB_FMT_CCRS);
  launchSpmv(A, x, y, 0, A.nr, nullptr, ctx().stream);
}

void sbSCS_spMVM(SbSCSMatrix* m, const CG_FLOAT* x, CG_FLOAT* y)
{
  // Reference semantics (matrix-SCS.c:198-228): x is indexed by the stored (un-permuted) column ids, y is
  // written in permuted row order and needs nrPadded slots.
  Operator A = makeOperator(m, SB_FMT_SCS);
  A.sell.col = m->colInd;
  launchSpmv(A, x, y, 0, A.sell.nChunks, nullptr, ctx().stream);
}

} // extern "C"
