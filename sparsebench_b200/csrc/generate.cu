// HPCG 27-point / 7-point stencil matrix source (replaces matrixGenerate, matrix.c:30-121).
// The matrix defines every measurement input of this package, so it is produced here in two ways with
// identical results: on the host (same calling convention as the reference: host GMatrix) and directly in
// device memory (what bench.py uses: 7.2 GB of entries at 256^3 never cross PCIe).
//
// Geometry (matrix.c:34-41, :63-96): each rank owns an nx*ny*nz block, blocks are stacked along z, row
// ids are lexicographic (x fastest); neighbours are enumerated dz,dy,dx ascending; x/y neighbours are
// clipped at the block faces, z neighbours only by the global row range; diagonal 27, off-diagonal -1.
#include <cub/device/device_scan.cuh>

#include "sb_internal.h"

namespace {

struct Geometry {
  int nx, ny, nz;
  long long plane, localRows, totalRows, firstRow;
  int use7pt;
};

__host__ __device__ inline Geometry makeGeometry(int nx, int ny, int nz, int rank, int size, int use7pt)
{
  Geometry g;
  g.nx = nx; g.ny = ny; g.nz = nz;
  g.plane = (long long)nx * ny;
  g.localRows = g.plane * nz;
  g.totalRows = g.localRows * size;
  g.firstRow = g.localRows * rank;
  g.use7pt = use7pt;
  return g;
}

// Visits the stored neighbours of local row `row` in the reference's order; f(globalCol, isDiagonal).
template <typename F>
__host__ __device__ inline int visitRow(const Geometry& g, long long row, F f)
{
  const int x = (int)(row % g.nx);
  const int y = (int)((row / g.nx) % g.ny);
  const long long self = g.firstRow + row;
  int count = 0;
  for (int dz = -1; dz <= 1; dz++)
    for (int dy = -1; dy <= 1; dy++) {
      if (y + dy < 0 || y + dy >= g.ny) continue;
      for (int dx = -1; dx <= 1; dx++) {
        if (x + dx < 0 || x + dx >= g.nx) continue;
        if (g.use7pt && dz * dz + dy * dy + dx * dx > 1) continue;
        const long long c = self + dz * g.plane + (long long)dy * g.nx + dx;
        if (c < 0 || c >= g.totalRows) continue;
        f(c, c == self);
        count++;
      }
    }
  return count;
}

__global__ void rowLengthKernel(Geometry g, idx_t* len)
{
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < g.localRows;
       row += (long long)gridDim.x * blockDim.x)
    len[row] = (idx_t)visitRow(g, row, [](long long, bool) {});
}

__global__ void fillKernel(Geometry g, const idx_t* __restrict__ rowPtr, Entry* __restrict__ entries)
{
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < g.localRows;
       row += (long long)gridDim.x * blockDim.x) {
    Entry* out = entries + rowPtr[row];
    visitRow(g, row, [&](long long c, bool diag) {
      Entry e;
      e.col = (CG_UINT)c;
      e.val = diag ? 27.0 : -1.0;
      *out++ = e;
    });
  }
}

void fillHeader(GMatrix* m, const Geometry& g)
{
  // matrix.c:114-120: nnz/totalNnz are the 27-per-row allocation bound, not the stored count
  m->startRow = (CG_UINT)g.firstRow;
  m->stopRow = (CG_UINT)(g.firstRow + g.localRows - 1);
  m->totalNr = (CG_UINT)g.totalRows;
  m->totalNnz = (CG_UINT)(27 * g.totalRows);
  m->nr = (CG_UINT)g.localRows;
  m->nc = (CG_UINT)g.localRows;
  m->nnz = (CG_UINT)(27 * g.localRows);
}

void checkSizes(const Geometry& g)
{
  if (g.nx < 1 || g.ny < 1 || g.nz < 1) SB_FATAL("matrixGenerate: grid dimensions must be positive");
  if (g.totalRows > 0xffffffffLL || 27 * g.localRows > 0xffffffffLL)
    SB_FATAL("matrixGenerate: %lld rows do not fit 32-bit CG_UINT indices", g.totalRows);
}

} // namespace

extern "C" {

void matrixGenerate(GMatrix* m, Parameter* p, int rank, int size, bool use_7pt_stencil)
{
  const Geometry g = makeGeometry(p->nx, p->ny, p->nz, rank, size, use_7pt_stencil ? 1 : 0);
  checkSizes(g);
  if (!rank) {   // matrix.c:43-52
    printf(use_7pt_stencil ? "Generate 7pt matrix with " : "Generate 27pt matrix with ");
    printf("%.2e total rows and %.2e nonzeros\n", (real_t)g.totalRows, (real_t)(27 * g.localRows));
  }
  void *rp = nullptr, *en = nullptr;
  if (posix_memalign(&rp, 64, sizeof(CG_UINT) * (size_t)(g.localRows + 1)) ||
      posix_memalign(&en, 64, sizeof(Entry) * (size_t)(27 * g.localRows)))
    SB_FATAL("matrixGenerate: out of host memory");
  m->rowPtr = (CG_UINT*)rp;
  m->entries = (Entry*)en;
  size_t cursor = 0;
  m->rowPtr[0] = 0;
  for (long long row = 0; row < g.localRows; row++) {
    visitRow(g, row, [&](long long c, bool diag) {
      Entry e;
      memset(&e, 0, sizeof(e));
      e.col = (CG_UINT)c;
      e.val = diag ? 27.0 : -1.0;
      m->entries[cursor++] = e;
    });
    m->rowPtr[row + 1] = (CG_UINT)cursor;
  }
  fillHeader(m, g);
}

void sbGenerateDevice(GMatrix* m, Parameter* p, int rank, int size, bool use_7pt_stencil)
{
  const Geometry g = makeGeometry(p->nx, p->ny, p->nz, rank, size, use_7pt_stencil ? 1 : 0);
  checkSizes(g);
  sb::Context& c = sb::ctx();
  const size_t n = (size_t)g.localRows;
  idx_t* len = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * n);
  m->rowPtr = (CG_UINT*)sbAllocateDevice(64, sizeof(CG_UINT) * (n + 1 + 8));   // +8: bulk-copy granularity of the SpMV kernels (CCRS shares this array)
  const int threads = 256;
  const int blocks = (int)((n + threads - 1) / threads < (size_t)c.numSMs * 16 ? (n + threads - 1) / threads : (size_t)c.numSMs * 16);
  rowLengthKernel<<<blocks, threads, 0, c.stream>>>(g, len);
  SB_CUDA(cudaGetLastError());
  SB_CUDA(cudaMemsetAsync(m->rowPtr, 0, sizeof(CG_UINT), c.stream));
  size_t tmpBytes = 0;
  SB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmpBytes, len, m->rowPtr + 1, (long long)n, c.stream));
  void* tmp = sbAllocateDevice(64, tmpBytes);
  SB_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmpBytes, len, m->rowPtr + 1, (long long)n, c.stream));
  CG_UINT stored = 0;
  SB_CUDA(cudaMemcpyAsync(&stored, m->rowPtr + n, sizeof(CG_UINT), cudaMemcpyDeviceToHost, c.stream));
  SB_CUDA(cudaStreamSynchronize(c.stream));
  // the reference allocates 27 entries per row (matrix.c:35,54); only rowPtr[nr] are ever valid, so the
  // device copy keeps just those
  m->entries = (Entry*)sbAllocateDevice(64, sizeof(Entry) * (size_t)(stored ? stored : 1));
  SB_CUDA(cudaMemsetAsync(m->entries, 0, sizeof(Entry) * (size_t)stored, c.stream));   // defined padding bytes
  fillKernel<<<blocks, threads, 0, c.stream>>>(g, m->rowPtr, m->entries);
  SB_CUDA(cudaGetLastError());
  SB_CUDA(cudaStreamSynchronize(c.stream));
  sbFree(tmp);
  sbFree(len);
  fillHeader(m, g);
}

void sbFreeGMatrix(GMatrix* m)
{
  if (sb::isDevicePointer(m->rowPtr)) {
    sbFree(m->rowPtr);
    sbFree(m->entries);
  } else {
    free(m->rowPtr);
    free(m->entries);
  }
  m->rowPtr = nullptr;
  m->entries = nullptr;
}

} // extern "C"
