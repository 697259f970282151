// Communication layer of the hot path (replaces commInit/commFinalize/commPartition/commExchange/
// commReduction of comm.c). One process per GPU; the MPI-3 neighbourhood collective and the scalar
// MPI_Allreduce of the reference become NCCL operations enqueued on CUDA streams (NVLink 5 / NVSwitch
// inside one B200 box). Row-block partitioning and the halo index lists are the reference's own
// (partition.cpp), bit for bit.
#include <arpa/inet.h>
#include <cub/device/device_select.cuh>
#include <nccl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <vector>

#include "device_utils.cuh"
#include "peer_window.h"
#include "sb_internal.h"
#include "sb_partition.h"

#define SB_NCCL(call)                                                                            \
  do {                                                                                           \
    ncclResult_t r_ = (call);                                                                    \
    if (r_ != ncclSuccess) {                                                                     \
      fprintf(stderr, "sparsebench_b200: NCCL error at %s:%d: %s\n", __FILE__, __LINE__,         \
          ncclGetErrorString(r_));                                                               \
      exit(EXIT_FAILURE);                                                                        \
    }                                                                                            \
  } while (0)

static const ncclDataType_t kNcclReal = sizeof(real_t) == 8 ? ncclDouble : ncclFloat;

struct SbPartitionPlan {
  sb::PartitionPlan plan;
};

namespace sb {

// ------------------------------------------------------------------------------------------- peer windows
// Inside one NVSwitch box every GPU can store straight into every other GPU's memory. Each rank owns two
// windows that its peers map through CUDA IPC:
//   control window : halo arrival / acknowledge counters and the slots of the scalar all-reduce
//   halo window    : two receive slots (double buffering by exchange parity) of externalCount values, laid
//                    out like the halo part of x (grouped by source in rdispls order)
// Halo exchange = the SENDER's kernel gathers x[elementsToSend[i]] and stores the values directly into the
// receivers' slots over NVLink, then publishes the exchange number (release, system scope); the receiver's
// kernel spins on its own counters (acquire), copies the slot behind its x vector and acknowledges. No
// host involvement, no staging buffer, no receive-side collective: the pack kernel (K6) and the
// MPI_Neighbor_alltoallv (C1) of comm.c:627-651 are one launch. The all-reduce of a dot product is one
// single-CTA kernel: every rank stores its partial into slot [rank] of every peer and sums the slots of its
// own window in rank order, so all ranks obtain bit-identical results.

struct PutPlan {
  int ndest;
  int sdispl[kMaxRanks + 1];                           // element offsets per destination (comm.c:150)
  real_t* remote[2][kMaxRanks];                        // destination slot (per parity), already offset to my segment
  unsigned long long* remoteFlag[kMaxRanks];           // &destCtrl->haloFlag[myRank]
  const unsigned long long* ack[kMaxRanks];            // &myCtrl->haloAck[destRank]
};

struct WaitPlan {
  int nsrc;
  int externalCount;
  const unsigned long long* flag[kMaxRanks];           // &myCtrl->haloFlag[sourceRank]
  unsigned long long* remoteAck[kMaxRanks];            // &sourceCtrl->haloAck[myRank]
  const real_t* slot[2];
};

enum { COMM_NCCL = 0, COMM_PEER = 1 };

struct CommExt {
  ncclComm_t nccl = nullptr;
  int rank = 0, size = 1;
  int mode = COMM_NCCL;
  int* dElementsToSend = nullptr;
  real_t* dScalar = nullptr;      // staging for host-scalar reductions
  real_t* hScalar = nullptr;
  std::vector<int> wantMatrix;    // [requester][owner] halo counts of the current partition
  bool installed = false;         // device lists / windows match the Comm lists
  // peer mode
  CtrlWindow* ctrl = nullptr;
  std::vector<CtrlWindow*> peerCtrl;
  CtrlWindow** dPeerCtrl = nullptr;
  real_t* halo = nullptr;
  std::vector<real_t*> peerHalo;
  PutPlan* dPut = nullptr;
  WaitPlan* dWait = nullptr;
  unsigned long long haloSeq = 0, redEpoch = 0;
  // persistent halo vector (the CG's p): allocated, zeroed and mapped by every peer ONCE per partition
  // (commAcquireHaloVector); solves only borrow it, the arrival counters keep running across solves
  real_t* arena = nullptr;
  size_t arenaSlots = 0;
  idx_t arenaRows = 0;
  bool arenaBusy = false;
  std::vector<real_t*> peerVec;
  PutPlan putDirect;               // host copy of *dPutDirect
  PutPlan* dPutDirect = nullptr;
  unsigned long long directSeq = 0;
  // send list in solver numbering (SELL row permutation `elemsKey`) and the FusedPut maps derived from it
  uint64_t elemsKey = 0;
  int* dElemsSolver = nullptr;
  bool fusedValid = false;
  uint64_t fusedKey = 0;
  FusedPut fused;                  // usable (ndest > 0) when the sends fit kMaxFusedDests
  std::vector<int*> fusedInv;      // device inverse maps owned by `fused`
};

static CommExt* g_world = nullptr;   // commReduction has no Comm* argument (it used MPI_COMM_WORLD, comm.c:653-662)

static CommExt* ext(const Comm* c) { return c ? (CommExt*)c->communicator : nullptr; }

static void freeHostLists(Comm* c)
{
  free(c->sources); free(c->recvCounts); free(c->rdispls);          // comm.c:896-903
  free(c->destinations); free(c->sendCounts); free(c->sdispls);
  free(c->elementsToSend);
}

static void resetLists(Comm* c)
{
  c->externalCount = 0; c->totalSendCount = 0; c->elementsToSend = nullptr;
  c->indegree = 0; c->outdegree = 0;
  c->sources = c->recvCounts = c->rdispls = nullptr;
  c->destinations = c->sendCounts = c->sdispls = nullptr;
  c->sendBuffer = nullptr;
}

// all-gather of `bytes` bytes per rank through NCCL (setup-time, tiny)
static void allGatherBytes(CommExt* e, const void* mine, size_t bytes, void* all)
{
  Context& c = ctx();
  char* d = (char*)sbAllocateDevice(64, bytes * (size_t)(e->size + 1));
  SB_CUDA(cudaMemcpyAsync(d, mine, bytes, cudaMemcpyHostToDevice, c.stream));
  SB_NCCL(ncclAllGather(d, d + bytes, bytes, ncclChar, e->nccl, c.stream));
  SB_CUDA(cudaMemcpyAsync(all, d + bytes, bytes * (size_t)e->size, cudaMemcpyDeviceToHost, c.stream));
  SB_CUDA(cudaStreamSynchronize(c.stream));
  sbFree(d);
}

static void allGatherInts(CommExt* e, const int* mine, int count, int* all)
{
  allGatherBytes(e, mine, sizeof(int) * (size_t)count, all);
}

static void ncclBarrier(CommExt* e)
{
  Context& c = ctx();
  SB_CUDA(cudaMemsetAsync(e->dScalar + 4, 0, sizeof(real_t), c.stream));
  SB_NCCL(ncclAllReduce(e->dScalar + 4, e->dScalar + 4, 1, kNcclReal, ncclSum, e->nccl, c.stream));
  SB_CUDA(cudaStreamSynchronize(c.stream));
}

// Maps every peer's copy of a window; returns false (on every rank) if any rank could not map any peer.
static bool mapPeers(CommExt* e, void* mine, std::vector<void*>& peers)
{
  cudaIpcMemHandle_t h;
  memset(&h, 0, sizeof(h));
  int ok = cudaIpcGetMemHandle(&h, mine) == cudaSuccess ? 1 : 0;
  if (!ok) cudaGetLastError();
  std::vector<cudaIpcMemHandle_t> all((size_t)e->size);
  allGatherBytes(e, &h, sizeof(h), all.data());
  std::vector<int> oks((size_t)e->size);
  allGatherInts(e, &ok, 1, oks.data());
  for (int r = 0; r < e->size; r++) ok &= oks[(size_t)r];
  peers.assign((size_t)e->size, nullptr);
  if (ok) {
    for (int r = 0; r < e->size; r++) {
      if (r == e->rank) { peers[(size_t)r] = mine; continue; }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        break;
      }
      peers[(size_t)r] = p;
    }
  }
  allGatherInts(e, &ok, 1, oks.data());
  int allOk = 1;
  for (int r = 0; r < e->size; r++) allOk &= oks[(size_t)r];
  if (!allOk) {
    for (int r = 0; r < e->size; r++)
      if (r != e->rank && peers[(size_t)r]) cudaIpcCloseMemHandle(peers[(size_t)r]);
    peers.assign((size_t)e->size, nullptr);
  }
  return allOk != 0;
}

static void unmapPeers(CommExt* e, std::vector<void*>& peers)
{
  for (int r = 0; r < (int)peers.size(); r++)
    if (r != e->rank && peers[(size_t)r]) cudaIpcCloseMemHandle(peers[(size_t)r]);
  peers.clear();
}

static void attach(Comm* c, int rank, int size, int device, const ncclUniqueId* id)
{
  c->rank = rank;
  c->size = size;
  c->logFile = nullptr;
  resetLists(c);
  c->communicator = nullptr;
  if (size <= 1) return;
  if (size > kMaxRanks) SB_FATAL("commInit: %d ranks exceed the supported maximum of %d", size, kMaxRanks);
  sbSetDevice(device);
  CommExt* e = new CommExt();
  e->rank = rank;
  e->size = size;
  SB_NCCL(ncclCommInitRank(&e->nccl, size, *id, rank));
  e->dScalar = (real_t*)sbAllocateDevice(64, sizeof(real_t) * 8);
  e->hScalar = (real_t*)sbAllocateHost(sizeof(real_t) * 8);
  c->communicator = e;
  g_world = e;
  // SB_COMM=nccl keeps every exchange on NCCL (the measured baseline); default: NVLink peer windows
  const char* modeEnv = getenv("SB_COMM");
  const bool wantPeer = !(modeEnv && strcmp(modeEnv, "nccl") == 0);
  if (wantPeer) {
    e->ctrl = (CtrlWindow*)sbAllocateDevice(256, sizeof(CtrlWindow));
    SB_CUDA(cudaMemset(e->ctrl, 0, sizeof(CtrlWindow)));
    SB_CUDA(cudaDeviceSynchronize());
    std::vector<void*> peers;
    if (mapPeers(e, e->ctrl, peers)) {
      e->peerCtrl.resize((size_t)size);
      for (int r = 0; r < size; r++) e->peerCtrl[(size_t)r] = (CtrlWindow*)peers[(size_t)r];
      e->dPeerCtrl = (CtrlWindow**)sbAllocateDevice(64, sizeof(CtrlWindow*) * (size_t)size);
      sbCopyToDevice(e->dPeerCtrl, e->peerCtrl.data(), sizeof(CtrlWindow*) * (size_t)size);
      e->mode = COMM_PEER;
    } else {
      if (rank == 0) fprintf(stderr, "sparsebench_b200: CUDA IPC peer mapping unavailable, using NCCL for halo exchange and reductions\n");
      sbFree(e->ctrl);
      e->ctrl = nullptr;
    }
  }
}

// ---- unique-id rendezvous for launchers that only export RANK/WORLD_SIZE/MASTER_ADDR/MASTER_PORT
static void tcpBroadcastId(int rank, int size, ncclUniqueId* id)
{
  const char* addr = getenv("MASTER_ADDR");
  const char* portEnv = getenv("SB_BOOTSTRAP_PORT");
  int port = portEnv ? atoi(portEnv) : (getenv("MASTER_PORT") ? atoi(getenv("MASTER_PORT")) + 1 : 29511);
  if (!addr) addr = "127.0.0.1";
  if (rank == 0) {
    int ls = socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in sa;
    memset(&sa, 0, sizeof(sa));
    sa.sin_family = AF_INET;
    sa.sin_addr.s_addr = htonl(INADDR_ANY);
    sa.sin_port = htons((uint16_t)port);
    if (bind(ls, (sockaddr*)&sa, sizeof(sa)) != 0 || listen(ls, size) != 0) SB_FATAL("commInit: cannot listen on port %d", port);
    for (int i = 1; i < size; i++) {
      int cs = accept(ls, nullptr, nullptr);
      if (cs < 0 || write(cs, id, sizeof(*id)) != (ssize_t)sizeof(*id)) SB_FATAL("commInit: rendezvous send failed");
      close(cs);
    }
    close(ls);
  } else {
    addrinfo hints, *res = nullptr;
    memset(&hints, 0, sizeof(hints));
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    char portStr[16];
    snprintf(portStr, sizeof(portStr), "%d", port);
    if (getaddrinfo(addr, portStr, &hints, &res) != 0 || !res) SB_FATAL("commInit: cannot resolve %s", addr);
    int fd = -1;
    for (int attempt = 0; attempt < 600; attempt++) {   // rank 0 may not be listening yet
      fd = socket(AF_INET, SOCK_STREAM, 0);
      if (connect(fd, res->ai_addr, res->ai_addrlen) == 0) break;
      close(fd);
      fd = -1;
      usleep(100000);
    }
    if (fd < 0) SB_FATAL("commInit: cannot reach rank 0 at %s:%d", addr, port);
    size_t got = 0;
    while (got < sizeof(*id)) {
      ssize_t r = read(fd, (char*)id + got, sizeof(*id) - got);
      if (r <= 0) SB_FATAL("commInit: rendezvous receive failed");
      got += (size_t)r;
    }
    close(fd);
    freeaddrinfo(res);
  }
}

// ---- kernels
__global__ void packKernel(int n, const int* __restrict__ elements, const real_t* __restrict__ x, real_t* __restrict__ out)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = x[elements[i]];   // comm.c:635-638
}

__device__ __forceinline__ unsigned long long ldAcquireSys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stReleaseSys(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// A dead peer must not hang the GPU: give up (trap -> the host sees a launch failure and exits) after ~20 s.
__device__ __forceinline__ void spinUntilAtLeast(const unsigned long long* p, unsigned long long target)
{
  if (ldAcquireSys(p) >= target) return;
  const unsigned long long start = globalTimerNs();
  while (ldAcquireSys(p) < target) {
    __nanosleep(40);
    if (globalTimerNs() - start > kPeerTimeoutNs) __trap();
  }
}

// Sender side of the halo exchange: gather + direct stores into the receivers' slots. Every block publishes its
// part with one system-scope reduction on the receiver's arrival counter (no grid-wide rendezvous): exchange
// number `seq` is complete at a receiver when its counter for this source has reached seq * kPutBlocks.
constexpr int kPutBlocks = 32, kWaitBlocks = 32, kHaloThreads = 512;

__device__ __forceinline__ void signalAddSys(unsigned long long* p, unsigned long long n = 1)
{
  asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(n) : "memory");
}

__global__ void __launch_bounds__(kHaloThreads)
haloPutKernel(const PutPlan* __restrict__ plan, const int* __restrict__ elements, const real_t* __restrict__ x,
    unsigned long long seq, bool direct)
{
  __shared__ unsigned int stored[kMaxRanks];            // direct mode: elements this block delivered per destination
  const int nd = plan->ndest;
  if ((int)threadIdx.x < kMaxRanks) stored[threadIdx.x] = 0;
  // the slot of parity seq&1 was last filled by exchange seq-2: wait until every receiver has drained that one
  if (!direct && (int)threadIdx.x < nd && seq > 2) spinUntilAtLeast(plan->ack[threadIdx.x], (seq - 2) * kWaitBlocks);
  __syncthreads();
  const int total = plan->sdispl[nd];
  const int par = direct ? 0 : (int)(seq & 1ull);
  const int stride = gridDim.x * blockDim.x;
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {   // 4 independent gathers in flight
    int idx[4];
    real_t v[4];
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i0 + u * stride < total) idx[u] = elements[i0 + u * stride];
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i0 + u * stride < total) v[u] = x[idx[u]];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * stride;
      if (i < total) {
        int d = 0;
        while (i >= plan->sdispl[d + 1]) d++;
        plan->remote[par][d][i - plan->sdispl[d]] = v[u];
        if (direct) atomicAdd(&stored[d], 1u);
      }
    }
  }
  // the barrier orders every thread's stores before the signalling threads' release (cumulative, system scope)
  __syncthreads();
  if ((int)threadIdx.x < nd) {
    if (!direct) signalAddSys(plan->remoteFlag[threadIdx.x]);
    else if (stored[threadIdx.x]) signalAddSys(plan->remoteFlag[threadIdx.x], stored[threadIdx.x]);   // receivers count elements
  }
}

// Receiver side: wait for every source's `seq`, copy the slot behind the local part of x, acknowledge.
__global__ void __launch_bounds__(kHaloThreads)
haloWaitKernel(const WaitPlan* __restrict__ plan, real_t* __restrict__ xHalo, unsigned long long seq)
{
  const int ns = plan->nsrc;
  if ((int)threadIdx.x < ns) spinUntilAtLeast(plan->flag[threadIdx.x], seq * kPutBlocks);
  __syncthreads();
  const real_t* src = plan->slot[seq & 1ull];
  const int n = plan->externalCount;
  const int stride = gridDim.x * blockDim.x;
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
    real_t v[4];
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i0 + u * stride < n) v[u] = __ldcg(src + i0 + u * stride);
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i0 + u * stride < n) xHalo[i0 + u * stride] = v[u];
  }
  __syncthreads();
  if ((int)threadIdx.x < ns) signalAddSys(plan->remoteAck[threadIdx.x]);   // release: ordered after this block's reads
}

// All-reduce of one value over the peer windows (replaces MPI_Allreduce of comm.c:657,659).
__global__ void __launch_bounds__(kMaxRanks)
peerAllreduceKernel(CtrlWindow* mine, CtrlWindow* const* __restrict__ peers, int rank, int size, unsigned long long epoch,
    real_t* d, int op)
{
  __shared__ real_t vals[kMaxRanks];
  const int t = threadIdx.x;
  const int slot = (int)(epoch % kRedDepth);
  if (t < size) {
    const real_t v = *d;
    CtrlWindow* w = peers[t];
    *(volatile double*)&w->redVal[slot][rank] = (double)v;   // the window's slots are doubles in every build
    __threadfence_system();
    stReleaseSys(&w->redFlag[slot][rank], epoch);
    spinUntilAtLeast(&mine->redFlag[slot][t], epoch);
    vals[t] = (real_t) * (volatile double*)&mine->redVal[slot][t];
  }
  __syncthreads();
  if (t == 0) {
    real_t acc = vals[0];
    for (int r = 1; r < size; r++) acc = (op == SB_MAX) ? (vals[r] > acc ? vals[r] : acc) : acc + vals[r];   // rank order: same bits on every rank
    *d = acc;
  }
}

__global__ void flagExternalKernel(uint64_t n, const Entry* __restrict__ e, idx_t startRow, idx_t stopRow,
    unsigned char* __restrict__ flag, idx_t* __restrict__ col)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const idx_t c = e[i].col;
    col[i] = c;
    flag[i] = (c < startRow || c > stopRow) ? 1 : 0;       // comm.c:457 (stopRow inclusive)
  }
}

__global__ void renumberKernel(uint64_t n, Entry* __restrict__ e, idx_t startRow, idx_t stopRow, int nExt,
    const idx_t* __restrict__ sortedGlobal, const idx_t* __restrict__ sortedLocal)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const idx_t c = e[i].col;
    if (c >= startRow && c <= stopRow) {
      e[i].col = c - startRow;                             // comm.c:100-101
    } else {
      int lo = 0, hi = nExt - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sortedGlobal[mid] < c) lo = mid + 1; else hi = mid;
      }
      e[i].col = sortedLocal[lo];                          // comm.c:102-104
    }
  }
}

static void releaseArena(Comm* c);

// Releases the device-side state of the current partition (collective in peer mode: peers may still store into
// the halo window until everybody has reached this point).
static void uninstallPartition(Comm* c)
{
  CommExt* e = ext(c);
  if (!e || !e->installed) return;
  SB_CUDA(cudaDeviceSynchronize());
  releaseArena(c);
  sbFree(e->dElemsSolver); e->dElemsSolver = nullptr;
  e->elemsKey = 0;
  if (e->mode == COMM_PEER) {
    ncclBarrier(e);
    std::vector<void*> peers(e->peerHalo.begin(), e->peerHalo.end());
    unmapPeers(e, peers);
    e->peerHalo.clear();
    ncclBarrier(e);
    sbFree(e->halo); e->halo = nullptr;
    sbFree(e->dPut); e->dPut = nullptr;
    sbFree(e->dWait); e->dWait = nullptr;
  }
  sbFree(e->dElementsToSend); e->dElementsToSend = nullptr;
  if (c->sendBuffer) { sbFree(c->sendBuffer); c->sendBuffer = nullptr; }
  e->installed = false;
}

// Device-side state for the Comm lists of the current partition. Collective.
static void installPartition(Comm* c)
{
  CommExt* e = ext(c);
  if (!e || e->installed) return;
  const int size = e->size, rank = e->rank;
  if (c->totalSendCount > 0) {
    e->dElementsToSend = (int*)sbAllocateDevice(64, sizeof(int) * (size_t)c->totalSendCount);
    sbCopyToDevice(e->dElementsToSend, c->elementsToSend, sizeof(int) * (size_t)c->totalSendCount);
    c->sendBuffer = (CG_FLOAT*)sbAllocateDevice(64, sizeof(CG_FLOAT) * (size_t)c->totalSendCount);   // comm.c:124-125
  }
  if (e->mode == COMM_PEER) {
    if ((int)e->wantMatrix.size() != size * size) SB_FATAL("commPartition: missing halo count matrix");
    const int* W = e->wantMatrix.data();
    const size_t ext0 = (size_t)c->externalCount;
    e->halo = (real_t*)sbAllocateDevice(256, sizeof(real_t) * (2 * ext0 + 2));
    SB_CUDA(cudaMemset(e->halo, 0, sizeof(real_t) * (2 * ext0 + 2)));
    SB_CUDA(cudaDeviceSynchronize());
    std::vector<void*> peers;
    if (!mapPeers(e, e->halo, peers)) SB_FATAL("commPartition: CUDA IPC mapping of the halo windows failed");
    e->peerHalo.resize((size_t)size);
    for (int r = 0; r < size; r++) e->peerHalo[(size_t)r] = (real_t*)peers[(size_t)r];
    PutPlan put;
    memset(&put, 0, sizeof(put));
    put.ndest = c->outdegree;
    for (int i = 0; i < c->outdegree; i++) {
      const int d = c->destinations[i];
      put.sdispl[i] = c->sdispls[i];
      // my segment inside d's halo: after everything d receives from lower-ranked sources (its rdispls, comm.c:135)
      size_t off = 0, extD = 0;
      for (int s2 = 0; s2 < size; s2++) {
        if (s2 < rank) off += (size_t)W[(size_t)d * size + s2];
        extD += (size_t)W[(size_t)d * size + s2];
      }
      put.remote[0][i] = e->peerHalo[(size_t)d] + off;
      put.remote[1][i] = e->peerHalo[(size_t)d] + extD + off;
      put.remoteFlag[i] = &e->peerCtrl[(size_t)d]->haloFlag[rank];
      put.ack[i] = &e->ctrl->haloAck[d];
    }
    put.sdispl[c->outdegree] = c->totalSendCount;
    WaitPlan wait;
    memset(&wait, 0, sizeof(wait));
    wait.nsrc = c->indegree;
    wait.externalCount = c->externalCount;
    for (int i = 0; i < c->indegree; i++) {
      const int s2 = c->sources[i];
      wait.flag[i] = &e->ctrl->haloFlag[s2];
      wait.remoteAck[i] = &e->peerCtrl[(size_t)s2]->haloAck[rank];
    }
    wait.slot[0] = e->halo;
    wait.slot[1] = e->halo + ext0;
    // arrival / acknowledge counters restart with every partition (the neighbour sets may have changed)
    e->haloSeq = 0;
    ncclBarrier(e);
    SB_CUDA(cudaMemset(e->ctrl->haloFlag, 0, sizeof(e->ctrl->haloFlag)));
    SB_CUDA(cudaMemset(e->ctrl->haloAck, 0, sizeof(e->ctrl->haloAck)));
    SB_CUDA(cudaDeviceSynchronize());
    e->dPut = (PutPlan*)sbAllocateDevice(64, sizeof(PutPlan));
    e->dWait = (WaitPlan*)sbAllocateDevice(64, sizeof(WaitPlan));
    sbCopyToDevice(e->dPut, &put, sizeof(put));
    sbCopyToDevice(e->dWait, &wait, sizeof(wait));
    ncclBarrier(e);
  }
  e->installed = true;
}

bool commPeerMode(const Comm* c)
{
  CommExt* e = ext(c);
  return e && e->mode == COMM_PEER;
}

// First half of an exchange: after this call the neighbours' copies of my boundary values are on their way.
void commHaloPut(Comm* c, const real_t* x, const int* elements, cudaStream_t s)
{
  CommExt* e = ext(c);
  if (!e || (c->indegree == 0 && c->outdegree == 0)) return;
  installPartition(c);
  if (e->mode != COMM_PEER) SB_FATAL("commHaloPut needs the peer-window transport");
  e->haloSeq++;
  if (c->outdegree > 0) {
    haloPutKernel<<<kPutBlocks, kHaloThreads, 0, s>>>(e->dPut, elements ? elements : e->dElementsToSend, x, e->haloSeq, false);
    SB_CUDA(cudaGetLastError());
    countLaunch();
  }
}

// Second half: x[numRows .. numRows+externalCount) holds the neighbours' values when this kernel has run.
void commHaloWait(Comm* c, idx_t numRows, real_t* x, cudaStream_t s)
{
  CommExt* e = ext(c);
  if (!e || (c->indegree == 0 && c->outdegree == 0)) return;
  if (c->indegree > 0) {
    haloWaitKernel<<<kWaitBlocks, kHaloThreads, 0, s>>>(e->dWait, x + numRows, e->haloSeq);
    SB_CUDA(cudaGetLastError());
    countLaunch();
  }
}

// All-reduce of one host scalar over the ranks of `e` (the body of commReduction).
static void hostAllreduce(CommExt* e, real_t* v, int op)
{
  Context& c = ctx();
  e->hScalar[0] = *v;
  SB_CUDA(cudaMemcpyAsync(e->dScalar, e->hScalar, sizeof(real_t), cudaMemcpyHostToDevice, c.stream));
  Comm tmp;
  memset(&tmp, 0, sizeof(tmp));
  tmp.communicator = e;
  commAllreduceDevice(&tmp, e->dScalar, 1, op, c.stream);
  SB_CUDA(cudaMemcpyAsync(e->hScalar, e->dScalar, sizeof(real_t), cudaMemcpyDeviceToHost, c.stream));
  SB_CUDA(cudaStreamSynchronize(c.stream));
  *v = e->hScalar[0];
}

static void dropFusedMaps(CommExt* e)
{
  for (int* p : e->fusedInv) sbFree(p);
  e->fusedInv.clear();
  e->fused = FusedPut();
  e->fusedValid = false;
}

// Collective. Frees the persistent halo vector: nobody may store into it any more when it goes.
static void releaseArena(Comm* c)
{
  CommExt* e = ext(c);
  if (!e || !e->arena) return;
  SB_CUDA(cudaDeviceSynchronize());
  ncclBarrier(e);
  std::vector<void*> peers(e->peerVec.begin(), e->peerVec.end());
  unmapPeers(e, peers);
  e->peerVec.clear();
  ncclBarrier(e);
  sbFree(e->arena);
  e->arena = nullptr;
  e->arenaSlots = 0;
  e->arenaRows = 0;
  e->arenaBusy = false;
  sbFree(e->dPutDirect);
  e->dPutDirect = nullptr;
  dropFusedMaps(e);
}

// Collective. Allocates the persistent halo vector, lets every peer map it and builds the put plan: my boundary
// values go straight behind destination d's local rows, at d's rdispl for me (comm.c:135).
static bool registerArena(Comm* c, size_t slots, idx_t numRows)
{
  CommExt* e = ext(c);
  releaseArena(c);
  const int size = e->size, rank = e->rank;
  e->arena = (real_t*)sbAllocateDevice(256, sizeof(real_t) * slots);
  SB_CUDA(cudaMemset(e->arena, 0, sizeof(real_t) * slots));
  SB_CUDA(cudaDeviceSynchronize());
  std::vector<void*> peers;
  if (!mapPeers(e, e->arena, peers)) {
    sbFree(e->arena);
    e->arena = nullptr;
    return false;
  }
  e->arenaSlots = slots;
  e->arenaRows = numRows;
  e->peerVec.resize((size_t)size);
  for (int r = 0; r < size; r++) e->peerVec[(size_t)r] = (real_t*)peers[(size_t)r];
  std::vector<int> rows((size_t)size);
  int mine = (int)numRows;
  allGatherInts(e, &mine, 1, rows.data());
  const int* W = e->wantMatrix.data();
  PutPlan& put = e->putDirect;
  memset(&put, 0, sizeof(put));
  put.ndest = c->outdegree;
  for (int i = 0; i < c->outdegree; i++) {
    const int d = c->destinations[i];
    put.sdispl[i] = c->sdispls[i];
    size_t off = 0;
    for (int s2 = 0; s2 < rank; s2++) off += (size_t)W[(size_t)d * size + s2];
    put.remote[0][i] = put.remote[1][i] = e->peerVec[(size_t)d] + (size_t)rows[(size_t)d] + off;   // x_d + nr_d + rdispl
    put.remoteFlag[i] = &e->peerCtrl[(size_t)d]->directFlag[rank];
    put.ack[i] = nullptr;
  }
  put.sdispl[c->outdegree] = c->totalSendCount;
  e->dPutDirect = (PutPlan*)sbAllocateDevice(64, sizeof(PutPlan));
  sbCopyToDevice(e->dPutDirect, &put, sizeof(put));
  // the arrival counters count elements since registration and keep running from solve to solve
  e->directSeq = 0;
  SB_CUDA(cudaMemset(e->ctrl->directFlag, 0, sizeof(e->ctrl->directFlag)));
  SB_CUDA(cudaDeviceSynchronize());
  ncclBarrier(e);
  return true;
}

// Borrows the persistent halo vector for one solve (collective; one scalar all-reduce when nothing has to be
// registered). Returns nullptr -- on every rank -- when some rank cannot run the gated SpMV or the windows cannot be
// mapped; the caller then allocates its own vector and exchanges through commExchangeOnStream.
// Why the halo part may be reused from solve to solve without any further synchronisation: a rank can only start
// storing the next solve's first halo after it has collected the previous solve's last p.Ap, i.e. after every
// rank's last SpMV -- the only reader of the halo -- has completed. The caller must not clear slots >= numRows.
real_t* commAcquireHaloVector(Comm* c, idx_t numRows, size_t slots, bool localOk)
{
  CommExt* e = ext(c);
  if (!e || e->mode != COMM_PEER) return nullptr;
  installPartition(c);
  if (e->arenaBusy) SB_FATAL("commAcquireHaloVector: the halo vector is in use by another solver");
  const bool fits = e->arena && e->arenaSlots >= slots && e->arenaRows == numRows;
  const bool ok = localOk && c->indegree <= kMaxGateSources;
  real_t v = (ok ? 1.0 : 0.0) + (fits ? 0.0 : 1024.0);       // both collective decisions in one sum
  hostAllreduce(e, &v, SB_SUM);
  const long long sum = (long long)(v + 0.5);
  if (sum % 1024 != e->size) return nullptr;
  if (sum / 1024 > 0 && !registerArena(c, slots > e->arenaSlots ? slots : e->arenaSlots, numRows)) return nullptr;
  e->arenaBusy = true;
  return e->arena;
}

void commReleaseHaloVector(Comm* c)
{
  CommExt* e = ext(c);
  if (e) e->arenaBusy = false;
}

// Send list in the solver's row numbering: the CG keeps SELL vectors in permuted order, so it sends
// p[oldToNew[element]]. Cached per permutation (`key`: unique id of the converted matrix, 0 = no permutation).
const int* commSolverElements(Comm* c, uint64_t key, const idx_t* oldToNew, cudaStream_t s)
{
  CommExt* e = ext(c);
  if (!e) return nullptr;
  installPartition(c);
  if (key == 0 || !oldToNew || c->totalSendCount == 0) return e->dElementsToSend;
  if (e->dElemsSolver && e->elemsKey == key) return e->dElemsSolver;
  if (!e->dElemsSolver) e->dElemsSolver = (int*)sbAllocateDevice(64, sizeof(int) * (size_t)c->totalSendCount);
  launchPermuteIndices((idx_t)c->totalSendCount, oldToNew, e->dElementsToSend, e->dElemsSolver, s);
  e->elemsKey = key;
  return e->dElemsSolver;
}

// Inverse maps for the delivery fused into the producing kernel (FusedPut): element -> position in the
// destination's halo. Local work, cached per permutation; false when there are too many destinations.
bool commPrepareFusedPut(Comm* c, uint64_t key, const int* elements)
{
  CommExt* e = ext(c);
  if (!e || !e->arena) return false;
  if (c->outdegree > kMaxFusedDests) return false;
  if (e->fusedValid && e->fusedKey == key) return true;
  dropFusedMaps(e);
  if (c->outdegree > 0) {
    std::vector<int> elems((size_t)c->totalSendCount);
    if (elements && elements != e->dElementsToSend) sbCopyToHost(elems.data(), elements, sizeof(int) * elems.size());
    else memcpy(elems.data(), c->elementsToSend, sizeof(int) * elems.size());
    e->fused.ndest = c->outdegree;
    for (int i = 0; i < c->outdegree; i++) {
      const int b = c->sdispls[i], cnt = c->sendCounts[i];
      int lo = elems[(size_t)b], hi = elems[(size_t)b];
      for (int j = 1; j < cnt; j++) {
        lo = std::min(lo, elems[(size_t)(b + j)]);
        hi = std::max(hi, elems[(size_t)(b + j)]);
      }
      std::vector<int> inv((size_t)(hi - lo + 1), -1);
      for (int j = 0; j < cnt; j++) inv[(size_t)(elems[(size_t)(b + j)] - lo)] = j;
      int* dInv = (int*)sbAllocateDevice(64, sizeof(int) * inv.size());
      sbCopyToDevice(dInv, inv.data(), sizeof(int) * inv.size());
      e->fusedInv.push_back(dInv);
      e->fused.lo[i] = (idx_t)lo;
      e->fused.hi[i] = (idx_t)hi;
      e->fused.inv[i] = dInv;
      e->fused.remote[i] = e->putDirect.remote[0][i];
      e->fused.remoteFlag[i] = e->putDirect.remoteFlag[i];
    }
  }
  e->fusedValid = true;
  e->fusedKey = key;
  return true;
}

static HaloGate directGate(Comm* c, CommExt* e)
{
  HaloGate g;
  g.nsrc = c->indegree;
  for (int i = 0; i < c->indegree; i++) {
    g.flag[i] = &e->ctrl->directFlag[c->sources[i]];
    g.target[i] = e->directSeq * (unsigned long long)c->recvCounts[i];     // senders signal element counts
  }
  return g;
}

HaloGate commHaloPutDirect(Comm* c, const real_t* x, const int* elements, cudaStream_t s)
{
  CommExt* e = ext(c);
  if (!e || !e->arenaBusy || x != e->arena) SB_FATAL("commHaloPutDirect: not the registered halo vector");
  e->directSeq++;
  if (c->outdegree > 0) {
    haloPutKernel<<<kPutBlocks, kHaloThreads, 0, s>>>(e->dPutDirect, elements ? elements : e->dElementsToSend, x, e->directSeq, true);
    SB_CUDA(cudaGetLastError());
    countLaunch();
  }
  return directGate(c, e);
}

// The next exchange of the registered vector is performed by the caller's own kernel (FusedPut): returns what that
// kernel needs and the gate the receiving SpMV waits on.
HaloGate commFusedPutBegin(Comm* c, FusedPut* fp)
{
  CommExt* e = ext(c);
  if (!e || !e->arenaBusy || !e->fusedValid) SB_FATAL("commFusedPutBegin: no registered halo vector");
  e->directSeq++;
  *fp = e->fused;
  return directGate(c, e);
}

void commExchangeOnStream(Comm* c, idx_t numRows, real_t* x, const int* elements, cudaStream_t s)
{
  CommExt* e = ext(c);
  if (!e || (c->indegree == 0 && c->outdegree == 0)) return;
  installPartition(c);
  if (e->mode == COMM_PEER) {
    commHaloPut(c, x, elements, s);
    commHaloWait(c, numRows, x, s);
    return;
  }
  if (c->totalSendCount > 0) {
    const int threads = 256;
    int blocks = (c->totalSendCount + threads - 1) / threads;
    if (blocks > ctx().numSMs * 4) blocks = ctx().numSMs * 4;
    packKernel<<<blocks, threads, 0, s>>>(c->totalSendCount, elements ? elements : e->dElementsToSend, x, c->sendBuffer);
    SB_CUDA(cudaGetLastError());
    countLaunch();
  }
  // MPI_Neighbor_alltoallv (comm.c:640-648): straight into the halo part of x, no receive staging
  SB_NCCL(ncclGroupStart());
  for (int i = 0; i < c->outdegree; i++)
    SB_NCCL(ncclSend(c->sendBuffer + c->sdispls[i], (size_t)c->sendCounts[i], kNcclReal, c->destinations[i], e->nccl, s));
  for (int i = 0; i < c->indegree; i++)
    SB_NCCL(ncclRecv(x + numRows + c->rdispls[i], (size_t)c->recvCounts[i], kNcclReal, c->sources[i], e->nccl, s));
  SB_NCCL(ncclGroupEnd());
}

void commAllreduceDevice(Comm* c, real_t* d, int count, int op, cudaStream_t s)
{
  CommExt* e = ext(c);
  if (!e) return;
  // peer windows: one value per launch (the CG's fused kernels carry their own); a longer vector (the GMRES
  // projections) goes through one NCCL all-reduce instead of `count` launches
  if (e->mode == COMM_PEER && count <= 2) {
    for (int i = 0; i < count; i++) {
      e->redEpoch++;
      peerAllreduceKernel<<<1, kMaxRanks, 0, s>>>(e->ctrl, e->dPeerCtrl, e->rank, e->size, e->redEpoch, d + i, op);
      SB_CUDA(cudaGetLastError());
      countLaunch();
    }
    return;
  }
  SB_NCCL(ncclAllReduce(d, d, (size_t)count, kNcclReal, op == SB_MAX ? ncclMax : ncclSum, e->nccl, s));
}

// Next epoch of the peer-window all-reduce, for kernels that fuse the push / collect halves (peer mode only).
PeerReduce commBeginReduce(Comm* c)
{
  CommExt* e = ext(c);
  PeerReduce pr;
  if (!e || e->mode != COMM_PEER) SB_FATAL("commBeginReduce needs the peer-window transport");
  pr.size = e->size;
  pr.rank = e->rank;
  pr.epoch = ++e->redEpoch;
  pr.mine = e->ctrl;
  pr.peers = e->dPeerCtrl;
  return pr;
}

const int* commDeviceElements(Comm* c)
{
  CommExt* e = ext(c);
  if (!e) return nullptr;
  installPartition(c);
  return e->dElementsToSend;
}

} // namespace sb

using namespace sb;

extern "C" {

int sbCommUniqueIdBytes(void) { return (int)sizeof(ncclUniqueId); }

void sbCommGetUniqueId(void* id) { SB_NCCL(ncclGetUniqueId((ncclUniqueId*)id)); }

void sbCommInitRank(Comm* c, int rank, int size, int device, const void* id)
{
  attach(c, rank, size, device, (const ncclUniqueId*)id);
}

void commInit(Comm* c, int argc, char** argv)
{
  (void)argc; (void)argv;
  const char* ws = getenv("WORLD_SIZE");
  const int size = ws ? atoi(ws) : 1;
  const int rank = getenv("RANK") ? atoi(getenv("RANK")) : 0;
  if (size <= 1) {                        // comm.c:869-872
    attach(c, 0, 1, 0, nullptr);
    return;
  }
  const int ndev = sbDeviceCount();
  if (ndev == 0) SB_FATAL("commInit: no CUDA device");
  const int local = getenv("LOCAL_RANK") ? atoi(getenv("LOCAL_RANK")) : rank;
  ncclUniqueId id;
  if (rank == 0) SB_NCCL(ncclGetUniqueId(&id));
  tcpBroadcastId(rank, size, &id);
  attach(c, rank, size, local % ndev, &id);
}

void commPrintBanner(Comm* c)
{
  // comm.c:185-274 prints the build configuration and where every rank runs (host, pid, CPU affinity mask); here
  // the interesting placement is the GPU of every rank
  cudaDeviceProp prop;
  int dev = 0;
  SB_CUDA(cudaGetDevice(&dev));
  SB_CUDA(cudaGetDeviceProperties(&prop, dev));
  char host[256] = "";
  gethostname(host, sizeof(host) - 1);
  if (c->rank == 0) {
    printf("sparsebench_b200: CUDA hot path for sm_100a, %s precision floats and integer type %s\n", sizeof(real_t) == 8 ? "double" : "single",
        sizeof(idx_t) == 4 ? "unsigned int" : "unsigned long long int");   // comm.c:196-199 (PRECISION_STRING, UINT_STRING)
    if (c->size > 1) printf("One process per GPU using %d ranks, transport: %s\n", c->size, commPeerMode(c) ? "NVLink peer windows" : "NCCL");
  }
  for (int i = 0; i < c->size; i++) {
    if (i == c->rank) {
      printf("Process with rank %d running on Node %s with pid %d on GPU %d (%s, %d SMs)\n", c->rank, host, (int)getpid(), dev,
          prop.name, prop.multiProcessorCount);
      fflush(stdout);
    }
    if (c->size > 1) {
      real_t z = 0.0;
      commReduction(&z, SB_SUM);                          // commBarrier() of comm.c:211,219
    }
  }
}

void commAbort(Comm* c, char* msg)
{
  if (c->rank == 0 && msg) printf("%s\n", msg);           // comm.c:880-891
  commFinalize(c);
  exit(EXIT_SUCCESS);
}

void commFinalize(Comm* c)
{
  freeHostLists(c);
  CommExt* e = ext(c);
  if (e) {
    SB_CUDA(cudaDeviceSynchronize());
    uninstallPartition(c);
    if (e->mode == COMM_PEER) {
      ncclBarrier(e);                                                // nobody stores into a window that is about to go
      std::vector<void*> peers(e->peerCtrl.begin(), e->peerCtrl.end());
      unmapPeers(e, peers);
      ncclBarrier(e);
      sbFree(e->ctrl);
      sbFree(e->dPeerCtrl);
    }
    sbFree(e->dScalar);
    sbFreeHost(e->hScalar);
    ncclCommDestroy(e->nccl);
    if (g_world == e) g_world = nullptr;
    delete e;
  }
  resetLists(c);
  c->communicator = nullptr;
}

void commReduction(CG_FLOAT* v, int op)
{
  if (g_world) hostAllreduce(g_world, v, op);   // single rank: no-op (comm.c:655,661)
}

void sbCommAllreduceDevice(Comm* c, CG_FLOAT* dev, int count, int op) { commAllreduceDevice(c, dev, count, op, ctx().stream); }

void commExchange(Comm* c, CG_UINT numRows, CG_FLOAT* x)
{
  ensureOnDevice(x);
  commExchangeOnStream(c, numRows, x, nullptr, ctx().stream);
}

void commDistributeMatrix(Comm* c, MMMatrix* m, MMMatrix* mLocal)
{
  CommExt* e = ext(c);
  if (c->size <= 1 || !e) {
    mLocal->startRow = 0;                  // comm.c:404-410
    mLocal->stopRow = m->nr - 1;
    mLocal->count = m->count;
    mLocal->nr = m->nr;
    mLocal->nnz = m->nnz;
    mLocal->entries = m->entries;
    mLocal->totalNr = m->nr;               // left unset by the reference's single-rank branch
    mLocal->totalNnz = m->nnz;
    return;
  }
  // comm.c:311-402: rank 0 holds the sorted entry list and scatters contiguous row blocks of
  // N/size (+1 for the first N%size ranks) rows (sizeOfRank, comm.c:35-38). Setup-time transfer through NCCL.
  Context& cx = ctx();
  const int size = c->size, rank = c->rank;
  std::vector<int> table((size_t)2 * size + 2, 0);         // totals, then (count, displ) per rank
  if (rank == 0) {
    table[0] = m->nr;
    table[1] = m->nnz;
    int cursor = 0;
    size_t at = 0;
    for (int i = 0; i < size; i++) {
      const int numRows = m->nr / size + ((m->nr % size > i) ? 1 : 0);
      const int stopRow = cursor + numRows - 1;
      cursor += numRows;
      const size_t begin = at;                              // scanMM (comm.c:276-296) for rows that all hold entries
      while (at < m->count && m->entries[at].row <= stopRow) at++;
      table[(size_t)2 + 2 * i] = (int)(at - begin);
      table[(size_t)3 + 2 * i] = (int)begin;
      printf("Rank %d count %d displ %d start %d stop %d\n", i, (int)(at - begin), (int)begin, stopRow - numRows + 1, stopRow);
    }
  }
  int* dTable = (int*)sbAllocateDevice(64, sizeof(int) * table.size());
  if (rank == 0) SB_CUDA(cudaMemcpyAsync(dTable, table.data(), sizeof(int) * table.size(), cudaMemcpyHostToDevice, cx.stream));
  SB_NCCL(ncclBroadcast(dTable, dTable, table.size(), ncclInt32, 0, e->nccl, cx.stream));
  SB_CUDA(cudaMemcpyAsync(table.data(), dTable, sizeof(int) * table.size(), cudaMemcpyDeviceToHost, cx.stream));
  SB_CUDA(cudaStreamSynchronize(cx.stream));
  sbFree(dTable);
  const int count = table[(size_t)2 + 2 * rank];
  mLocal->count = (size_t)count;
  mLocal->totalNr = table[0];
  mLocal->totalNnz = table[1];
  void* host = nullptr;
  if (posix_memalign(&host, 64, sizeof(MMEntry) * (size_t)(count ? count : 1)) != 0) SB_FATAL("commDistributeMatrix: out of host memory");
  mLocal->entries = (MMEntry*)host;
  // MPI_Scatterv (comm.c:373-381)
  const size_t allBytes = rank == 0 ? sizeof(MMEntry) * m->count : 0;
  char* dAll = rank == 0 ? (char*)sbAllocateDevice(64, allBytes) : nullptr;
  char* dMine = (char*)sbAllocateDevice(64, sizeof(MMEntry) * (size_t)(count ? count : 1));
  if (rank == 0) SB_CUDA(cudaMemcpyAsync(dAll, m->entries, allBytes, cudaMemcpyHostToDevice, cx.stream));
  SB_NCCL(ncclGroupStart());
  if (rank == 0)
    for (int i = 1; i < size; i++) {
      const size_t cnt = (size_t)table[(size_t)2 + 2 * i], off = (size_t)table[(size_t)3 + 2 * i];
      if (cnt) SB_NCCL(ncclSend(dAll + off * sizeof(MMEntry), cnt * sizeof(MMEntry), ncclChar, i, e->nccl, cx.stream));
    }
  else if (count)
    SB_NCCL(ncclRecv(dMine, (size_t)count * sizeof(MMEntry), ncclChar, 0, e->nccl, cx.stream));
  SB_NCCL(ncclGroupEnd());
  if (rank == 0) {
    memcpy(mLocal->entries, m->entries + table[3], sizeof(MMEntry) * (size_t)count);
    SB_CUDA(cudaStreamSynchronize(cx.stream));
  } else {
    SB_CUDA(cudaMemcpyAsync(mLocal->entries, dMine, sizeof(MMEntry) * (size_t)count, cudaMemcpyDeviceToHost, cx.stream));
    SB_CUDA(cudaStreamSynchronize(cx.stream));
  }
  sbFree(dAll);
  sbFree(dMine);
  if (count == 0) SB_FATAL("commDistributeMatrix: rank %d received no matrix entries (fewer rows than ranks?)", rank);
  mLocal->startRow = mLocal->entries[0].row;               // comm.c:383-386
  mLocal->stopRow = mLocal->entries[count - 1].row;
  mLocal->nr = mLocal->stopRow - mLocal->startRow + 1;
  mLocal->nnz = count;
  printf("Rank %d count %zu start %d stop %d\n", rank, mLocal->count, mLocal->startRow, mLocal->stopRow);
}

SbPartitionPlan* sbPartitionLocal(GMatrix* m, int rank, int size, const CG_UINT* startRows, int* wantCounts)
{
  SbPartitionPlan* P = new SbPartitionPlan();
  const idx_t startRow = m->startRow, stopRow = m->stopRow, nr = m->nr;
  if (isDevicePointer(m->entries)) {
    // device GMatrix: compact the external references in entry order, number them on the host, rewrite on the device
    Context& c = ctx();
    cudaStream_t s = c.stream;
    idx_t stored = 0;
    SB_CUDA(cudaMemcpyAsync(&stored, m->rowPtr + nr, sizeof(idx_t), cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    const uint64_t n = stored;
    unsigned char* flag = (unsigned char*)sbAllocateDevice(64, n ? n : 1);
    idx_t* col = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * (n ? n : 1));
    idx_t* sel = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * (n ? n : 1));
    int* dCount = (int*)sbAllocateDevice(64, sizeof(int));
    const int threads = 256;
    const int blocks = (int)std::min<uint64_t>((n + threads - 1) / threads + 1, (uint64_t)c.numSMs * 16);
    flagExternalKernel<<<blocks, threads, 0, s>>>(n, m->entries, startRow, stopRow, flag, col);
    SB_CUDA(cudaGetLastError());
    size_t tmpBytes = 0;
    SB_CUDA(cub::DeviceSelect::Flagged(nullptr, tmpBytes, col, flag, sel, dCount, (long long)n, s));
    void* tmp = sbAllocateDevice(64, tmpBytes);
    SB_CUDA(cub::DeviceSelect::Flagged(tmp, tmpBytes, col, flag, sel, dCount, (long long)n, s));   // stable: entry order kept
    int nRefs = 0;
    SB_CUDA(cudaMemcpyAsync(&nRefs, dCount, sizeof(int), cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    std::vector<idx_t> refs((size_t)nRefs);
    if (nRefs) sbCopyToHost(refs.data(), sel, sizeof(idx_t) * (size_t)nRefs);
    P->plan.build(refs.data(), refs.size(), rank, size, nr, startRow, startRows);
    const int nExt = (int)P->plan.extGlobal.size();
    std::vector<std::pair<idx_t, idx_t>> table((size_t)nExt);
    for (int i = 0; i < nExt; i++) table[(size_t)i] = { P->plan.extGlobal[(size_t)i], P->plan.localId[(size_t)i] };
    std::sort(table.begin(), table.end());
    std::vector<idx_t> g((size_t)nExt), l((size_t)nExt);
    for (int i = 0; i < nExt; i++) { g[(size_t)i] = table[(size_t)i].first; l[(size_t)i] = table[(size_t)i].second; }
    idx_t* dg = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * (size_t)(nExt ? nExt : 1));
    idx_t* dl = (idx_t*)sbAllocateDevice(64, sizeof(idx_t) * (size_t)(nExt ? nExt : 1));
    if (nExt) {
      sbCopyToDevice(dg, g.data(), sizeof(idx_t) * (size_t)nExt);
      sbCopyToDevice(dl, l.data(), sizeof(idx_t) * (size_t)nExt);
    }
    renumberKernel<<<blocks, threads, 0, s>>>(n, m->entries, startRow, stopRow, nExt, dg, dl);
    SB_CUDA(cudaGetLastError());
    SB_CUDA(cudaStreamSynchronize(s));
    sbFree(flag); sbFree(col); sbFree(sel); sbFree(dCount); sbFree(tmp); sbFree(dg); sbFree(dl);
  } else {
    const idx_t stored = m->rowPtr[nr];
    std::vector<idx_t> refs;
    for (idx_t j = 0; j < stored; j++) {
      const idx_t col = m->entries[j].col;
      if (col < startRow || col > stopRow) refs.push_back(col);
    }
    P->plan.build(refs.data(), refs.size(), rank, size, nr, startRow, startRows);
    for (idx_t j = 0; j < stored; j++) m->entries[j].col = P->plan.renumber(m->entries[j].col, stopRow);
  }
  m->nc = m->nc + (CG_UINT)P->plan.extGlobal.size();        // comm.c:616
  for (int o = 0; o < size; o++) wantCounts[o] = P->plan.want[(size_t)o];
  return P;
}

const int* sbPartitionRequestSlice(SbPartitionPlan* P, const int* wantMatrix, int source, int* count)
{
  const int* ptr = nullptr;
  P->plan.requestSlice(wantMatrix, source, &ptr, count);
  return ptr;
}

static int* dupInts(const std::vector<int>& v)
{
  int* p = (int*)malloc(sizeof(int) * (v.size() ? v.size() : 1));
  if (!v.empty()) memcpy(p, v.data(), sizeof(int) * v.size());
  return p;
}

void sbPartitionFinish(SbPartitionPlan* P, Comm* c, const int* wantMatrix, const int* received)
{
  CommLists L;
  P->plan.finish(L, wantMatrix, received);
  if (CommExt* e = ext(c)) {
    uninstallPartition(c);                                   // device state of a previous partition (collective)
    e->wantMatrix.assign(wantMatrix, wantMatrix + (size_t)c->size * c->size);
  }
  freeHostLists(c);                                          // lists of a previous partition on this Comm
  c->externalCount = L.externalCount;
  c->totalSendCount = L.totalSendCount;
  c->indegree = (int)L.sources.size();
  c->outdegree = (int)L.destinations.size();
  c->sources = dupInts(L.sources); c->recvCounts = dupInts(L.recvCounts); c->rdispls = dupInts(L.rdispls);
  c->destinations = dupInts(L.destinations); c->sendCounts = dupInts(L.sendCounts); c->sdispls = dupInts(L.sdispls);
  c->elementsToSend = dupInts(L.elementsToSend);
  c->sendBuffer = nullptr;                                   // device buffer, created with the device lists
  delete P;
}

void commPartition(Comm* c, GMatrix* m)
{
  const int size = c->size, rank = c->rank;
  CommExt* e = ext(c);
  if (size <= 1) {   // one row block: every column is local already, all lists stay empty (startRow is 0)
    freeHostLists(c);
    resetLists(c);
    c->sources = dupInts({}); c->recvCounts = dupInts({}); c->rdispls = dupInts({});
    c->destinations = dupInts({}); c->sendCounts = dupInts({}); c->sdispls = dupInts({});
    c->elementsToSend = dupInts({});
    return;
  }
  if (size > 1 && !e) SB_FATAL("commPartition: Comm has %d ranks but no communicator (call commInit first)", size);
  std::vector<CG_UINT> startRows((size_t)size);
  std::vector<int> wantMatrix((size_t)size * size), mine((size_t)size);
  if (size > 1) {
    int my = (int)m->startRow;
    std::vector<int> all((size_t)size);
    allGatherInts(e, &my, 1, all.data());                   // comm.c:496-502
    for (int i = 0; i < size; i++) startRows[(size_t)i] = (CG_UINT)all[(size_t)i];
  } else {
    startRows[0] = m->startRow;
  }
  SbPartitionPlan* P = sbPartitionLocal(m, rank, size, startRows.data(), mine.data());
  if (size > 1) allGatherInts(e, mine.data(), size, wantMatrix.data());
  else wantMatrix[0] = mine[0];
  // request lists: requester -> owner (comm.c:130-161)
  int total = 0;
  std::vector<int> recvOff((size_t)size, 0);
  for (int s = 0; s < size; s++) {
    recvOff[(size_t)s] = total;
    total += wantMatrix[(size_t)s * size + rank];
  }
  std::vector<int> received((size_t)(total ? total : 1));
  if (size > 1) {
    Context& cx = ctx();
    int nReq = 0;
    for (int s = 0; s < size; s++) nReq += wantMatrix[(size_t)rank * size + s];
    int* dSend = (int*)sbAllocateDevice(64, sizeof(int) * (size_t)(nReq ? nReq : 1));
    int* dRecv = (int*)sbAllocateDevice(64, sizeof(int) * (size_t)(total ? total : 1));
    if (nReq) sbCopyToDevice(dSend, P->plan.requests.data(), sizeof(int) * (size_t)nReq);
    SB_NCCL(ncclGroupStart());
    for (int s = 0; s < size; s++) {
      int cnt = 0;
      const int* slice = sbPartitionRequestSlice(P, wantMatrix.data(), s, &cnt);
      if (cnt > 0) SB_NCCL(ncclSend(dSend + (slice - P->plan.requests.data()), (size_t)cnt, ncclInt32, s, e->nccl, cx.stream));
      const int in = wantMatrix[(size_t)s * size + rank];
      if (in > 0) SB_NCCL(ncclRecv(dRecv + recvOff[(size_t)s], (size_t)in, ncclInt32, s, e->nccl, cx.stream));
    }
    SB_NCCL(ncclGroupEnd());
    if (total) sbCopyToHost(received.data(), dRecv, sizeof(int) * (size_t)total);
    else SB_CUDA(cudaStreamSynchronize(cx.stream));
    sbFree(dSend); sbFree(dRecv);
  }
  sbPartitionFinish(P, c, wantMatrix.data(), received.data());
  if (size > 1) installPartition(c);
}

} // extern "C"
