// Internal declarations shared by the translation units of libsparsebench_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "peer_window.h"
#include "sb_types.h"
#include "sparsebench_b200.h"

// Error convention of the reference: message + exit(EXIT_FAILURE) (allocate.c:19-33).
#define SB_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "sparsebench_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_),  \
          __FILE__, __LINE__, cudaGetErrorString(e_));                                         \
      exit(EXIT_FAILURE);                                                                      \
    }                                                                                          \
  } while (0)

#define SB_FATAL(...)                                                                          \
  do {                                                                                         \
    fprintf(stderr, "sparsebench_b200: " __VA_ARGS__);                                         \
    fprintf(stderr, "\n");                                                                     \
    exit(EXIT_FAILURE);                                                                        \
  } while (0)

namespace sb {

constexpr int kMaxPartials = 4096;   // upper bound on reduction blocks of any kernel

// Per-process device context: one stream, reduction scratch, launch counter.
struct Context {
  int device = -1;
  int numSMs = 0;
  cudaStream_t stream = nullptr;     // blocking stream: ordered with the legacy default stream
  cudaStream_t commStream = nullptr; // halo exchange side stream
  real_t* partials = nullptr;        // kMaxPartials doubles per reduction slot, 4 slots
  unsigned int* tickets = nullptr;   // last-block tickets, one per reduction slot
  real_t* dScalar = nullptr;         // small device scalar block (64 doubles)
  real_t* hScalar = nullptr;         // pinned mirror
  double* dWide = nullptr;           // 8 doubles whatever the value type (maxErr is computed in double)
  double* hWide = nullptr;
  void* flushBuf = nullptr;
  size_t flushBytes = 0;
  size_t launches = 0;
};
Context& ctx();                      // lazily initialised on first use; exits if no CUDA device
inline void countLaunch(int n = 1) { ctx().launches += (size_t)n; }

bool isDevicePointer(const void* p);
// `p` may be unified memory from the ABI-level allocate() that the host has just filled: moved to the GPU once
void ensureOnDevice(const void* p);

// Kernel launch with the programmatic-dependent-launch attribute (device_utils.cuh: griddepWait). SB_NO_PDL=1 turns
// the attribute off (plain stream order) for A/B measurements.
bool pdlEnabled();
template <typename... KArgs, typename... Args>
inline void launchPdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdlEnabled() ? 1 : 0;
  SB_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

// ---- sparse formats: device-side views the kernels take
struct CrsView {
  idx_t nr;
  const idx_t* rowPtr;
  const idx_t* col;
  const real_t* val;
};
struct SellView {
  idx_t nChunks, nr, C;
  const idx_t* chunkPtr;
  const idx_t* chunkLens;
  const idx_t* col;
  const real_t* val;
};
struct CcrsView {
  idx_t nr;
  const idx_t* rowPtr;
  const Entry* entries;
};

// cached interior/boundary split of a converted matrix (spmvInteriorUnits)
struct HaloSplit {
  bool valid = false;
  idx_t lo = 0, hi = 0;
};
// CRS / CCRS with skewed row lengths: consecutive rows grouped into blocks of bounded non-zero count (built once per
// matrix, at its first SpMV); rows [start[b], start[b+1]) form block b
struct RowBlocks {
  int state = 0;                  // 0 not looked at yet, 1 near-uniform rows (tiles of a fixed row count), 2 skewed (this table)
  idx_t* start = nullptr;
  uint32_t count = 0;
};

// SELL-32 with a few very long chunks (longest > 256 columns and > 4x the average): those chunks are taken out of the
// ring kernel -- it is handed `shortLens`, where they have length 0 -- and multiplied by one CTA each.
struct LongChunks {
  idx_t* shortLens = nullptr;      // chunkLens with the long chunks zeroed
  idx_t* list = nullptr;           // their chunk ids, ascending
  uint32_t count = 0;
};

// A sparse operator as the CG driver sees it.
struct Operator {
  int fmt;
  HaloSplit* split = nullptr;              // cache slot in the matrix's side table
  RowBlocks* blocks = nullptr;             // CRS / CCRS: cache slot in the matrix's side table
  const LongChunks* longc = nullptr;       // SELL-32: set when the matrix has long chunks
  idx_t nr = 0, nc = 0, nrPadded = 0;   // vectors written by spmv need nrPadded slots
  uint64_t nnzTrue = 0;
  const idx_t* rowPtr = nullptr;        // CRS/CCRS: device rowPtr (b = 27-(len-1) rule); SCS: original-order rowLen
  const idx_t* rowLen = nullptr;        // SCS: row lengths in vector (permuted) order
  const idx_t* oldToNew = nullptr;      // SCS with sigma>1: vectors live in permuted order
  const idx_t* newToOld = nullptr;
  uint64_t permKey = 0;                    // unique id of the row permutation (0: none); cache key of derived lists
  CrsView crs{};
  SellView sell{};                         // col = symmetric-permuted columns when oldToNew != nullptr
  CcrsView ccrs{};
};

struct HaloGate;
struct FusedPut;

// Fused dot-product epilogue: *out = (accumulate ? *out : 0) + sum, reduced through scratch slot `slot` (0..3).
struct DotArgs {
  real_t* out;
  bool accumulate;
  int slot;
  const PeerReduce* push = nullptr;   // multi-GPU: the last block also stores the sum into every peer's window
};
// y = A x on units [lo,hi) (rows for CRS/CCRS, chunks for SELL); with `dot` also sum_i x[i]*y[i] over them.
idx_t spmvUnits(const Operator& A);
// largest unit range [lo, hi) around the middle whose rows reference no halo column (col >= nr); blocks the stream
void spmvInteriorUnits(const Operator& A, idx_t* lo, idx_t* hi, cudaStream_t s);
void launchSpmv(const Operator& A, const real_t* x, real_t* y, idx_t lo, idx_t hi, const DotArgs* dot,
    cudaStream_t s);
// y = A x over all units in ONE launch, ordered interior [intLo,intHi) first; the kernel waits on `gate` before the
// first unit outside that range (those reference halo columns, which a peer is storing while the interior runs)
bool spmvGatedAvailable(const Operator& A);
void launchSpmvGated(const Operator& A, const real_t* x, real_t* y, idx_t intLo, idx_t intHi, const HaloGate& gate,
    const DotArgs* dot, cudaStream_t s);

// ---- vector kernels (vecops.cu)
void launchWaxpby(idx_t n, real_t alpha, const real_t* x, real_t beta, const real_t* y, real_t* w, cudaStream_t s);
// *dResult (device) = sum x[i]*y[i], deterministic one-kernel grid reduction through scratch slot `slot`
void launchDot(idx_t n, const real_t* x, const real_t* y, real_t* dResult, int slot, cudaStream_t s);
// fused CG passes: rho[j] = r_j.r_j, pAp[k] = p_k.Ap_k live on the device, k is the 1-based iteration
// collect*: the scalar this kernel needs (rho[k-1] resp. pAp[k]) is still spread over the peer window and is summed
// in the kernel's prologue; pushRho: rho[k] is pushed to the peers instead of being all-reduced by a separate kernel
// hostRho: mapped pinned mirror of rho[] -- the kernel that produces the GLOBAL rho[j] also stores it there
void launchCgUpdateP(idx_t n, int k, real_t* rho, const real_t* r, real_t* p, const PeerReduce* collectRho,
    const FusedPut* put, real_t* hostRho, cudaStream_t s);
void launchCgUpdateXR(idx_t n, int k, real_t* rho, real_t* pAp, real_t* x, real_t* r, const real_t* p,
    const real_t* Ap, int slot, const PeerReduce* collectPAp, const PeerReduce* pushRho, real_t* hostRho, cudaStream_t s);

void launchInitVectors(idx_t n, const idx_t* rowPtr, const idx_t* rowLen, bool generated, real_t* x, real_t* b,
    cudaStream_t s);
void launchScatter(idx_t n, const idx_t* map, const real_t* in, real_t* out, cudaStream_t s);   // out[map[i]] = in[i]
void launchGather(idx_t n, const idx_t* map, const real_t* in, real_t* out, cudaStream_t s);    // out[i] = in[map[i]]
void launchMaxErr(idx_t n, const real_t* x, double* out, cudaStream_t s);
void launchPermuteIndices(idx_t n, const idx_t* map, const int* in, int* out, cudaStream_t s);   // out[i] = map[in[i]]

// ---- communication (comm.cu)
// `elements` overrides the device copy of Comm.elementsToSend (the CG passes row-permuted indices for SELL)
void commExchangeOnStream(Comm* c, idx_t numRows, real_t* x, const int* elements, cudaStream_t s);
void commAllreduceDevice(Comm* c, real_t* d, int count, int op, cudaStream_t s);
// NVLink peer-window transport
bool commPeerMode(const Comm* c);
const int* commDeviceElements(Comm* c);                    // device copy of Comm.elementsToSend
PeerReduce commBeginReduce(Comm* c);                       // next all-reduce epoch, for fused push / collect
// Direct halo delivery for one registered vector (the CG's p): the sender stores straight behind the receiver's
// local rows, the receiver's SpMV kernel itself waits on the arrival counters before it touches the first row that
// references a halo column (HaloGate). No acknowledge: the caller guarantees that a new exchange only starts after
// the previous SpMV has completed on every rank (in CG the two dot-product all-reduces in between do).
constexpr int kMaxGateSources = 8;
struct HaloGate {
  int nsrc = 0;                                            // 0: no gate
  unsigned long long target[kMaxGateSources] = {};         // elements received from source i since the vector was attached
  const unsigned long long* flag[kMaxGateSources] = {};
  unsigned long long* trace = nullptr;                     // SB_SYNC_TRACE: [0] += ns waited, [1] = max ns, [2] += 1 (per waiting CTA / warp)
};
// The same delivery fused into the kernel that produces the values (the CG's p update): element e of the vector
// goes to position inv[d][e - lo[d]] of destination d's halo (negative: not sent there).
constexpr int kMaxFusedDests = 8;   // up to 8 neighbours per rank: every rank of an 8-GPU box may neighbour all others
struct FusedPut {
  int ndest = 0;                                           // 0: not in use
  idx_t lo[kMaxFusedDests] = {}, hi[kMaxFusedDests] = {};   // inclusive element range that holds everything sent to d
  const int* inv[kMaxFusedDests] = {};
  real_t* remote[kMaxFusedDests] = {};
  unsigned long long* remoteFlag[kMaxFusedDests] = {};
};
// Collective. Borrows the Comm's persistent halo vector (>= slots doubles, zero-filled when first registered; peers
// map it once per partition, not per solve) for one solver; nullptr (on every rank) -> allocate your own vector and
// use commExchangeOnStream. The borrower must not clear slots >= numRows: a faster peer may already have stored the
// next exchange's halo there.
real_t* commAcquireHaloVector(Comm* c, idx_t numRows, size_t slots, bool localOk);
void commReleaseHaloVector(Comm* c);
// send list in the solver's row numbering (device; SELL keeps vectors in permuted order), cached per permutation key
const int* commSolverElements(Comm* c, uint64_t key, const idx_t* oldToNew, cudaStream_t s);
bool commPrepareFusedPut(Comm* c, uint64_t key, const int* elements);    // false: too many destinations for FusedPut
HaloGate commFusedPutBegin(Comm* c, FusedPut* fp);                   // next exchange, performed by the caller's kernel
HaloGate commHaloPutDirect(Comm* c, const real_t* x, const int* elements, cudaStream_t s);   // returns the gate to wait on

// ---- side tables keyed by the device array a Matrix struct points to
struct ScsExt {
  HaloSplit split;
  idx_t* colPerm = nullptr;   // symmetric-permuted column ids (CG keeps vectors in permuted order)
  idx_t* rowLenPerm = nullptr;
  idx_t* rowLenOrig = nullptr;
  bool identityPerm = true;
  uint64_t nnzTrue = 0;
  idx_t nc = 0;               // columns incl. halo (the reference's SCS struct drops it, matrix-SCS.c:38)
  LongChunks longc;
  uint64_t id = 0;               // unique per conversion, never reused
};
struct CrsExt {
  HaloSplit split;
  RowBlocks blocks;
  uint64_t nnzTrue = 0;
  bool ownsArrays = true;
};
ScsExt* scsExt(const void* key, bool create);
CrsExt* crsExt(const void* key, bool create);
void eraseExt(const void* key);

Operator makeOperator(void* matrix, int fmt);

} // namespace sb
