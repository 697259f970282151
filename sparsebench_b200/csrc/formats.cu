// Matrix-format plugins, conversion half (replaces convertMatrix of matrix-CRS.c:12-44,
// matrix-SCS.c:31-196 and matrix-CCRS.c:12). All conversions run on the device: a host GMatrix (the
// reference's calling convention) is uploaded once, a device GMatrix (sbGenerateDevice) is used in place.
// Integer outputs (permutations, chunk tables, column ids) are bit-identical to the reference's.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include <vector>

#include "sb_internal.h"

namespace sb {

static std::mutex g_extMutex;
static std::unordered_map<const void*, ScsExt*> g_scs;
static std::unordered_map<const void*, CrsExt*> g_crs;

ScsExt* scsExt(const void* key, bool create)
{
  std::lock_guard<std::mutex> l(g_extMutex);
  auto it = g_scs.find(key);
  if (it != g_scs.end()) {
    if (create) *it->second = ScsExt();      // the key was freed behind our back and handed out again: start clean
    return it->second;
  }
  if (!create) return nullptr;
  return g_scs[key] = new ScsExt();
}
CrsExt* crsExt(const void* key, bool create)
{
  std::lock_guard<std::mutex> l(g_extMutex);
  auto it = g_crs.find(key);
  if (it != g_crs.end()) {
    if (create) {
      if (it->second->blocks.start) sbFree(it->second->blocks.start);
      *it->second = CrsExt();
    }
    return it->second;
  }
  if (!create) return nullptr;
  return g_crs[key] = new CrsExt();
}
void eraseExt(const void* key)
{
  std::lock_guard<std::mutex> l(g_extMutex);
  auto a = g_scs.find(key);
  if (a != g_scs.end()) { delete a->second; g_scs.erase(a); }
  auto b = g_crs.find(key);
  if (b != g_crs.end()) {
    if (b->second->blocks.start) sbFree(b->second->blocks.start);
    delete b->second;
    g_crs.erase(b);
  }
}

// Device view of the input GMatrix (uploads host arrays; owns what it uploaded).
struct DeviceInput {
  const idx_t* rowPtr = nullptr;
  const Entry* entries = nullptr;
  uint64_t stored = 0;      // rowPtr[nr]
  bool owned = false;
  void release()
  {
    if (owned) { sbFree((void*)rowPtr); sbFree((void*)entries); }
    owned = false;
  }
};

static DeviceInput stageInput(const GMatrix* im)
{
  Context& c = ctx();
  DeviceInput in;
  const size_t nr = im->nr;
  if (isDevicePointer(im->entries) || isDevicePointer(im->rowPtr)) {
    in.rowPtr = im->rowPtr;
    in.entries = im->entries;
    idx_t last = 0;
    SB_CUDA(cudaMemcpyAsync(&last, im->rowPtr + nr, sizeof(idx_t), cudaMemcpyDeviceToHost, c.stream));
    SB_CUDA(cudaStreamSynchronize(c.stream));
    in.stored = last;
  } else {
    in.stored = im->rowPtr[nr];
    idx_t* rp = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * (nr + 1 + 8));   // +8: bulk-copy granularity (CCRS shares it)
    Entry* en = (Entry*)sbAllocateDevice(64, sizeof(Entry) * (in.stored ? in.stored : 1));
    SB_CUDA(cudaMemcpyAsync(rp, im->rowPtr, sizeof(idx_t) * (nr + 1), cudaMemcpyHostToDevice, c.stream));
    SB_CUDA(cudaMemcpyAsync(en, im->entries, sizeof(Entry) * in.stored, cudaMemcpyHostToDevice, c.stream));
    SB_CUDA(cudaStreamSynchronize(c.stream));
    in.rowPtr = rp;
    in.entries = en;
    in.owned = true;
  }
  return in;
}

static inline int gridFor(uint64_t work, int threads, int perSM = 16)
{
  uint64_t b = (work + threads - 1) / threads;
  uint64_t cap = (uint64_t)ctx().numSMs * perSM;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}

// ------------------------------------------------------------------------------------------- CRS
__global__ void splitEntriesKernel(uint64_t n, const Entry* __restrict__ e, idx_t* __restrict__ col,
    real_t* __restrict__ val)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Entry t = e[i];
    col[i] = t.col;
    val[i] = t.val;
  }
}

// ------------------------------------------------------------------------------------------- SELL-C-sigma
__global__ void scsKeysKernel(idx_t nr, idx_t nrPadded, idx_t sigma, const idx_t* __restrict__ rowPtr,
    uint64_t* __restrict__ keys, idx_t* __restrict__ idx, idx_t* __restrict__ len)
{
  for (idx_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nrPadded; i += gridDim.x * blockDim.x) {
    const idx_t l = i < nr ? rowPtr[i + 1] - rowPtr[i] : 0u;     // padding rows have length 0 (matrix-SCS.c:49-58)
    // ascending sort of (window, ~length) == per-window descending length; radix sort is stable, so ties
    // keep ascending row order exactly like the reference's mergesort (matrix-SCS.c:20-29, :61-79)
    keys[i] = ((uint64_t)(i / sigma) << 32) | (uint64_t)(0xffffffffu - l);
    idx[i] = i;
    len[i] = l;
  }
}

__global__ void scsChunkLenKernel(idx_t nChunks, idx_t C, const uint64_t* __restrict__ sortedKeys,
    idx_t* __restrict__ chunkLens, uint64_t* __restrict__ chunkElems)
{
  for (idx_t ch = blockIdx.x * blockDim.x + threadIdx.x; ch < nChunks; ch += gridDim.x * blockDim.x) {
    idx_t longest = 0;
    for (idx_t k = 0; k < C; k++) {
      const idx_t l = 0xffffffffu - (idx_t)(sortedKeys[(uint64_t)ch * C + k] & 0xffffffffu);
      longest = l > longest ? l : longest;
    }
    chunkLens[ch] = longest;                       // matrix-SCS.c:100-108
    chunkElems[ch] = (uint64_t)longest * C;
  }
}

__global__ void scsNarrowKernel(idx_t n, const uint64_t* __restrict__ wide, idx_t* __restrict__ narrow)
{
  for (idx_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) narrow[i] = (idx_t)wide[i];
}

__global__ void scsPermKernel(idx_t nr, idx_t nrPadded, const idx_t* __restrict__ sortedIdx,
    idx_t* __restrict__ oldToNew, idx_t* __restrict__ newToOld, unsigned int* __restrict__ notIdentity)
{
  for (idx_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < nrPadded; pos += gridDim.x * blockDim.x) {
    const idx_t old = sortedIdx[pos];
    if (old < nr) {                                // matrix-SCS.c:120-143
      oldToNew[old] = pos;
      newToOld[pos] = old;                         // pos < nr: real rows always sort ahead of padding rows
      if (old != pos) *notIdentity = 1u;
    }
  }
}

__global__ void scsFillKernel(idx_t nr, idx_t C, const idx_t* __restrict__ rowPtr,
    const Entry* __restrict__ entries, const idx_t* __restrict__ oldToNew, const idx_t* __restrict__ chunkPtr,
    idx_t* __restrict__ col, real_t* __restrict__ val, idx_t* __restrict__ colPerm,
    idx_t* __restrict__ rowLenPerm)
{
  // one row per thread; j-th stored entry of row i goes to chunkPtr[r/C] + j*C + r%C (matrix-SCS.c:164-192)
  for (idx_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += gridDim.x * blockDim.x) {
    const idx_t r = oldToNew[i];
    const uint64_t base = (uint64_t)chunkPtr[r / C] + r % C;
    const idx_t lo = rowPtr[i], hi = rowPtr[i + 1];
    for (idx_t j = lo; j < hi; j++) {
      const Entry e = entries[j];
      const uint64_t at = base + (uint64_t)(j - lo) * C;
      col[at] = e.col;
      val[at] = e.val;
      colPerm[at] = e.col < nr ? oldToNew[e.col] : e.col;   // halo columns (>= nr) keep their slot
    }
    rowLenPerm[r] = hi - lo;
  }
}

__global__ void scsPadColPermKernel(uint64_t nElems, const real_t* __restrict__ val, idx_t* colPerm, idx_t zeroTarget)
{
  // padding elements carry col 0 / val 0 in the reference numbering (matrix-SCS.c:150-155); in the permuted
  // numbering they must keep pointing at the slot that holds original row 0
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nElems; i += (uint64_t)gridDim.x * blockDim.x)
    if (colPerm[i] == (idx_t) ~(idx_t)0) colPerm[i] = zeroTarget;     // still the 0xff fill: a padding element
}

__global__ void rowLenKernel(idx_t nr, const idx_t* __restrict__ rowPtr, idx_t* __restrict__ len)
{
  for (idx_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += gridDim.x * blockDim.x) len[i] = rowPtr[i + 1] - rowPtr[i];
}

Operator makeOperator(void* matrix, int fmt)
{
  Operator A;
  A.fmt = fmt;
  if (fmt == SB_FMT_CRS) {
    SbCRSMatrix* m = (SbCRSMatrix*)matrix;
    A.nr = m->nr; A.nc = m->nc; A.nrPadded = m->nr;
    A.rowPtr = m->rowPtr;
    A.crs = CrsView { m->nr, m->rowPtr, m->colInd, m->val };
    CrsExt* e = crsExt(m->val, false);
    A.nnzTrue = e ? e->nnzTrue : 0;
    A.split = e ? &e->split : nullptr;
    A.blocks = e ? &e->blocks : nullptr;
  } else if (fmt == SB_FMT_CCRS) {
    SbCCRSMatrix* m = (SbCCRSMatrix*)matrix;
    A.nr = m->nr; A.nc = m->nc; A.nrPadded = m->nr;
    A.rowPtr = m->rowPtr;
    A.ccrs = CcrsView { m->nr, m->rowPtr, m->entries };
    CrsExt* e = crsExt(m->entries, false);
    A.nnzTrue = e ? e->nnzTrue : 0;
    A.split = e ? &e->split : nullptr;
    A.blocks = e ? &e->blocks : nullptr;
  } else if (fmt == SB_FMT_SCS) {
    SbSCSMatrix* m = (SbSCSMatrix*)matrix;
    ScsExt* e = scsExt(m->val, false);
    if (!e) SB_FATAL("SCS matrix was not produced by sbSCS_convertMatrix");
    A.nr = m->nr; A.nc = e->nc; A.nrPadded = m->nrPadded;
    A.rowLen = e->rowLenPerm;
    A.nnzTrue = e->nnzTrue;
    A.split = &e->split;
    A.longc = (e->longc.count > 0 && m->C == 32) ? &e->longc : nullptr;
    if (!e->identityPerm) { A.oldToNew = m->oldToNewPerm; A.newToOld = m->newToOldPerm; A.permKey = e->id; }
    A.sell = SellView { m->nChunks, m->nr, m->C, m->chunkPtr, m->chunkLens, e->identityPerm ? m->colInd : e->colPerm, m->val };
  } else {
    SB_FATAL("unknown matrix format id %d", fmt);
  }
  return A;
}

} // namespace sb

using namespace sb;

static void copyHeader(CG_UINT* dst, const GMatrix* im)
{
  // all three Matrix structs start with the same seven CG_UINT fields (CRSMatrix.h:10-12 etc.)
  dst[0] = im->nr; dst[1] = im->nc; dst[2] = im->nnz; dst[3] = im->totalNr; dst[4] = im->totalNnz;
  dst[5] = im->startRow; dst[6] = im->stopRow;
}

extern "C" {

void sbCRS_convertMatrix(SbCRSMatrix* m, GMatrix* im)
{
  Context& c = ctx();
  copyHeader(&m->nr, im);
  DeviceInput in = stageInput(im);
  const size_t nr = im->nr;
  m->rowPtr = (CG_UINT*)sbAllocateDevice(64, sizeof(CG_UINT) * (nr + 1 + 8));   // +8: 16-byte granularity of the bulk copies
  // +8: the staged SpMV kernel widens its bulk copies to 16-byte granularity (up to 3 elements past the end)
  m->colInd = (CG_UINT*)sbAllocateDevice(64, sizeof(CG_UINT) * (in.stored + 8));
  m->val = (CG_FLOAT*)sbAllocateDevice(64, sizeof(CG_FLOAT) * (in.stored + 8));
  SB_CUDA(cudaMemsetAsync(m->colInd + in.stored, 0, sizeof(CG_UINT) * 8, c.stream));
  SB_CUDA(cudaMemsetAsync(m->val + in.stored, 0, sizeof(CG_FLOAT) * 8, c.stream));
  SB_CUDA(cudaMemcpyAsync(m->rowPtr, in.rowPtr, sizeof(CG_UINT) * (nr + 1), cudaMemcpyDeviceToDevice, c.stream));
  if (in.stored) {
    splitEntriesKernel<<<gridFor(in.stored, 256), 256, 0, c.stream>>>(in.stored, in.entries, m->colInd, m->val);
    SB_CUDA(cudaGetLastError());
  }
  SB_CUDA(cudaStreamSynchronize(c.stream));
  in.release();
  crsExt(m->val, true)->nnzTrue = in.stored;
}

void sbCRS_destroyMatrix(SbCRSMatrix* m)
{
  eraseExt(m->val);
  sbFree(m->rowPtr); sbFree(m->colInd); sbFree(m->val);
  m->rowPtr = m->colInd = nullptr; m->val = nullptr;
}

void sbCCRS_convertMatrix(SbCCRSMatrix* m, GMatrix* im)
{
  // matrix-CCRS.c:12 intends `Matrix` to alias the GMatrix (identical layouts, CCRSMatrix.h:14-20 vs
  // matrix.h:29-35); the device build needs the arrays in HBM, so host input is uploaded and device
  // input is shared without a copy.
  copyHeader(&m->nr, im);
  DeviceInput in = stageInput(im);
  m->rowPtr = (CG_UINT*)in.rowPtr;
  m->entries = (Entry*)in.entries;
  CrsExt* e = crsExt(m->entries, true);
  e->nnzTrue = in.stored;
  e->ownsArrays = in.owned;
}

void sbCCRS_destroyMatrix(SbCCRSMatrix* m)
{
  CrsExt* e = crsExt(m->entries, false);
  const bool owned = e && e->ownsArrays;
  eraseExt(m->entries);
  if (owned) { sbFree(m->rowPtr); sbFree(m->entries); }
  m->rowPtr = nullptr; m->entries = nullptr;
}

void sbSCS_convertMatrix(SbSCSMatrix* m, GMatrix* im)
{
  Context& c = ctx();
  cudaStream_t s = c.stream;
  if (m->C == 0 || m->sigma == 0) SB_FATAL("sbSCS_convertMatrix: Matrix.C and Matrix.sigma must be set by the caller (matrix-SCS.c:40)");
  copyHeader(&m->nr, im);
  m->nc = im->nr;                                        // matrix-SCS.c:38
  const idx_t nr = im->nr, C = m->C, sigma = m->sigma;
  const idx_t nChunks = (nr + C - 1) / C;             // :40
  const uint64_t nrPadded64 = (uint64_t)nChunks * C;     // :41
  if (nrPadded64 > 0xffffffffull) SB_FATAL("sbSCS_convertMatrix: more than 2^32 padded rows");   // the sort key packs (window, length) into 64 bits
  const idx_t nrPadded = (idx_t)nrPadded64;
  m->nChunks = nChunks;
  m->nrPadded = nrPadded;
  DeviceInput in = stageInput(im);

  const size_t np = nrPadded ? nrPadded : 1;
  uint64_t* keys = (uint64_t*)sbAllocateDevice(64, sizeof(uint64_t) * np);
  uint64_t* keysSorted = (uint64_t*)sbAllocateDevice(64, sizeof(uint64_t) * np);
  idx_t* idx = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * np);
  idx_t* idxSorted = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * np);
  ScsExt* ext = new ScsExt();
  ext->rowLenOrig = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * np);
  ext->rowLenPerm = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * np);
  ext->nnzTrue = in.stored;
  ext->nc = im->nc;
  static std::atomic<uint64_t> nextId { 1 };
  ext->id = nextId.fetch_add(1);
  SB_CUDA(cudaMemsetAsync(ext->rowLenPerm, 0, sizeof(idx_t) * np, s));
  scsKeysKernel<<<gridFor(nrPadded, 256), 256, 0, s>>>(nr, nrPadded, sigma, in.rowPtr, keys, idx, ext->rowLenOrig);
  SB_CUDA(cudaGetLastError());
  if (sigma > 1 && nrPadded > 1) {
    size_t tmpBytes = 0;
    SB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys, keysSorted, idx, idxSorted, (long long)nrPadded, 0, 64, s));
    void* tmp = sbAllocateDevice(64, tmpBytes);
    SB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keys, keysSorted, idx, idxSorted, (long long)nrPadded, 0, 64, s));
    SB_CUDA(cudaStreamSynchronize(s));
    sbFree(tmp);
  } else {
    SB_CUDA(cudaMemcpyAsync(keysSorted, keys, sizeof(uint64_t) * np, cudaMemcpyDeviceToDevice, s));
    SB_CUDA(cudaMemcpyAsync(idxSorted, idx, sizeof(idx_t) * np, cudaMemcpyDeviceToDevice, s));
  }

  m->chunkLens = (CG_UINT*)sbAllocateDevice(64, sizeof(CG_UINT) * (nChunks ? nChunks : 1));
  m->chunkPtr = (CG_UINT*)sbAllocateDevice(64, sizeof(CG_UINT) * ((size_t)nChunks + 1));
  uint64_t* chunkElems = (uint64_t*)sbAllocateDevice(64, sizeof(uint64_t) * ((size_t)nChunks + 1));
  uint64_t* chunkPtr64 = (uint64_t*)sbAllocateDevice(64, sizeof(uint64_t) * ((size_t)nChunks + 1));
  SB_CUDA(cudaMemsetAsync(chunkElems, 0, sizeof(uint64_t) * ((size_t)nChunks + 1), s));
  if (nChunks) {
    scsChunkLenKernel<<<gridFor(nChunks, 128), 128, 0, s>>>(nChunks, C, keysSorted, m->chunkLens, chunkElems);
    SB_CUDA(cudaGetLastError());
  }
  {
    size_t tmpBytes = 0;
    SB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, chunkElems, chunkPtr64, (long long)nChunks + 1, s));
    void* tmp = sbAllocateDevice(64, tmpBytes);
    SB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, chunkElems, chunkPtr64, (long long)nChunks + 1, s));
    SB_CUDA(cudaStreamSynchronize(s));
    sbFree(tmp);
  }
  uint64_t nElems = 0;
  SB_CUDA(cudaMemcpyAsync(&nElems, chunkPtr64 + nChunks, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));
  if (sizeof(idx_t) == 4 && nElems > 0xffffffffull)
    SB_FATAL("sbSCS_convertMatrix: %llu padded elements overflow CG_UINT chunkPtr (use the 64-bit index build)", (unsigned long long)nElems);
  m->nElems = (CG_UINT)nElems;                           // :110-114
  scsNarrowKernel<<<gridFor((uint64_t)nChunks + 1, 256), 256, 0, s>>>(nChunks + 1, chunkPtr64, m->chunkPtr);
  SB_CUDA(cudaGetLastError());

  m->oldToNewPerm = (CG_UINT*)sbAllocateDevice(64, sizeof(CG_UINT) * (nr ? nr : 1));
  m->newToOldPerm = (CG_UINT*)sbAllocateDevice(64, sizeof(CG_UINT) * (nr ? nr : 1));
  unsigned int* notIdentity = (unsigned int*)sbAllocateDevice(64, sizeof(unsigned int));
  SB_CUDA(cudaMemsetAsync(notIdentity, 0, sizeof(unsigned int), s));
  if (nrPadded) {
    scsPermKernel<<<gridFor(nrPadded, 256), 256, 0, s>>>(nr, nrPadded, idxSorted, m->oldToNewPerm, m->newToOldPerm, notIdentity);
    SB_CUDA(cudaGetLastError());
  }

  const size_t ne = nElems ? nElems : 1;
  m->colInd = (CG_UINT*)sbAllocateDevice(64, sizeof(CG_UINT) * ne);
  m->val = (CG_FLOAT*)sbAllocateDevice(64, sizeof(CG_FLOAT) * ne);
  ext->colPerm = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * ne);
  SB_CUDA(cudaMemsetAsync(m->colInd, 0, sizeof(CG_UINT) * ne, s));      // :150-155
  SB_CUDA(cudaMemsetAsync(m->val, 0, sizeof(CG_FLOAT) * ne, s));
  SB_CUDA(cudaMemsetAsync(ext->colPerm, 0xff, sizeof(idx_t) * ne, s));
  if (nr) {
    scsFillKernel<<<gridFor(nr, 128), 128, 0, s>>>(nr, C, in.rowPtr, in.entries, m->oldToNewPerm, m->chunkPtr, m->colInd,
        m->val, ext->colPerm, ext->rowLenPerm);
    SB_CUDA(cudaGetLastError());
  }
  unsigned int hostNotIdentity = 0;
  idx_t zeroTarget = 0;
  SB_CUDA(cudaMemcpyAsync(&hostNotIdentity, notIdentity, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
  if (nr) SB_CUDA(cudaMemcpyAsync(&zeroTarget, m->oldToNewPerm, sizeof(idx_t), cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));
  if (nElems) {
    scsPadColPermKernel<<<gridFor(nElems, 256), 256, 0, s>>>(nElems, m->val, ext->colPerm, zeroTarget);
    SB_CUDA(cudaGetLastError());
  }
  SB_CUDA(cudaStreamSynchronize(s));
  ext->identityPerm = hostNotIdentity == 0;
  if (ext->identityPerm) { sbFree(ext->colPerm); ext->colPerm = nullptr; }

  sbFree(keys); sbFree(keysSorted); sbFree(idx); sbFree(idxSorted); sbFree(chunkElems); sbFree(chunkPtr64); sbFree(notIdentity);
  in.release();
  if (C == 32 && nChunks > 0) {
    // a few very long chunks (heavy-tailed row lengths) would each keep one warp of the ring kernel busy long after
    // the rest is done: list them for the long-chunk kernel (spmv.cu)
    std::vector<idx_t> lens((size_t)nChunks);
    SB_CUDA(cudaMemcpyAsync(lens.data(), m->chunkLens, sizeof(idx_t) * (size_t)nChunks, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    idx_t longest = 0;
    for (idx_t l : lens) longest = l > longest ? l : longest;
    const double avg = (double)m->nElems / 32.0 / (double)nChunks;
    if (longest > 256 && (double)longest > 4.0 * avg) {
      std::vector<idx_t> list;
      for (size_t i = 0; i < lens.size(); i++)
        if (lens[i] > 256) {
          list.push_back((idx_t)i);
          lens[i] = 0;
        }
      ext->longc.count = (uint32_t)list.size();
      ext->longc.shortLens = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * lens.size());
      ext->longc.list = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * list.size());
      SB_CUDA(cudaMemcpyAsync(ext->longc.shortLens, lens.data(), sizeof(idx_t) * lens.size(), cudaMemcpyHostToDevice, s));
      SB_CUDA(cudaMemcpyAsync(ext->longc.list, list.data(), sizeof(idx_t) * list.size(), cudaMemcpyHostToDevice, s));
      SB_CUDA(cudaStreamSynchronize(s));
    }
  }
  {
    // register the side table under the val pointer
    ScsExt* slot = scsExt(m->val, true);
    *slot = *ext;
    delete ext;
  }
}

void sbSCS_destroyMatrix(SbSCSMatrix* m)
{
  ScsExt* e = scsExt(m->val, false);
  if (e) { sbFree(e->colPerm); sbFree(e->rowLenPerm); sbFree(e->rowLenOrig); sbFree(e->longc.shortLens); sbFree(e->longc.list); }
  eraseExt(m->val);
  sbFree(m->colInd); sbFree(m->val); sbFree(m->chunkPtr); sbFree(m->chunkLens); sbFree(m->oldToNewPerm); sbFree(m->newToOldPerm);
  m->colInd = m->chunkPtr = m->chunkLens = m->oldToNewPerm = m->newToOldPerm = nullptr; m->val = nullptr;
}

} // extern "C"
