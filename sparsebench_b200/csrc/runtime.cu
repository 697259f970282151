// Runtime utilities of the hot path: device allocation (replaces allocate.c:12-36), timing (replaces
// timing.c:8-20 / the PROFILE macro's time source, profiler.h:18-21) and the per-process device context.
#include <time.h>

#include <map>
#include <mutex>
#include <unordered_map>

#include "sb_internal.h"

namespace sb {

static Context g_ctx;
static std::mutex g_poolMutex;
static std::once_flag g_once;
static int g_requestedDevice = -1;

static void initContext()
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    SB_FATAL("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
  int dev = g_requestedDevice >= 0 ? g_requestedDevice : 0;
  if (g_requestedDevice < 0) {
    const char* lr = getenv("LOCAL_RANK");   // one process per GPU under torchrun
    if (lr) dev = atoi(lr) % n;
  }
  // getTimeStamp() drains the device around every PROFILE()d call of the reference's drivers: let those waits spin
  // instead of sleeping (tens of microseconds less per call). Refused when another library already created the
  // context with other flags: not an error.
  if (cudaSetDeviceFlags(cudaDeviceScheduleSpin) != cudaSuccess) cudaGetLastError();
  SB_CUDA(cudaSetDevice(dev));
  g_ctx.device = dev;
  SB_CUDA(cudaDeviceGetAttribute(&g_ctx.numSMs, cudaDevAttrMultiProcessorCount, dev));
  SB_CUDA(cudaStreamCreate(&g_ctx.stream));
  SB_CUDA(cudaStreamCreateWithFlags(&g_ctx.commStream, cudaStreamNonBlocking));
  SB_CUDA(cudaMalloc(&g_ctx.partials, sizeof(double) * kMaxPartials * 4));
  SB_CUDA(cudaMalloc(&g_ctx.tickets, sizeof(unsigned int) * 16));
  SB_CUDA(cudaMemset(g_ctx.tickets, 0, sizeof(unsigned int) * 16));
  SB_CUDA(cudaMalloc(&g_ctx.dScalar, sizeof(double) * 64));
  SB_CUDA(cudaMemset(g_ctx.dScalar, 0, sizeof(double) * 64));
  SB_CUDA(cudaMallocHost(&g_ctx.hScalar, sizeof(double) * 64));
  SB_CUDA(cudaMalloc(&g_ctx.dWide, sizeof(double) * 8));
  SB_CUDA(cudaMallocHost(&g_ctx.hWide, sizeof(double) * 8));
}

Context& ctx()
{
  std::call_once(g_once, initContext);
  return g_ctx;
}

} // namespace sb
extern "C" int sbPrefetchManaged(const void* p);
namespace sb {
void ensureOnDevice(const void* p)
{
  if (p) sbPrefetchManaged(p);
}

// Read-only streaming kernel: the second roofline denominator. The driver's MEASURED_PEAKS.json number is a COPY
// (half reads, half writes); the SpMV is a ~98 % read stream, which HBM3e serves faster than a 50/50 mix.
__global__ void __launch_bounds__(512) readStreamKernel(const double2* __restrict__ src, uint64_t n2, double* sink)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n2; i += 4 * stride) {
    double2 v0, v1, v2, v3;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v0.x), "=d"(v0.y) : "l"(src + i));
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v1.x), "=d"(v1.y) : "l"(src + i + stride));
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v2.x), "=d"(v2.y) : "l"(src + i + 2 * stride));
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v3.x), "=d"(v3.y) : "l"(src + i + 3 * stride));
    a0 += v0.x + v0.y; a1 += v1.x + v1.y; a2 += v2.x + v2.y; a3 += v3.x + v3.y;
  }
  for (; i < n2; i += stride) a0 += src[i].x + src[i].y;
  const double s = (a0 + a1) + (a2 + a3);
  if (s == 1.2345e300) *sink = s;                           // never true: keeps the loads alive
}

bool pdlEnabled()
{
  static const bool on = getenv("SB_NO_PDL") == nullptr;
  return on;
}

bool isDevicePointer(const void* p)
{
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

} // namespace sb

using namespace sb;

extern "C" {

// cudaMalloc / cudaFree synchronise the device and cost up to milliseconds each; a solver allocates and releases
// the same nine vectors per solve (CGSolver.c:69-79 never frees, a library must). Released blocks are therefore
// parked in a small exact-size cache and handed out again (contents undefined, exactly like fresh cudaMalloc memory).
struct Block {
  size_t bytes;
  bool parked;
};
static std::unordered_map<void*, Block> g_blockSize;                // every live or parked block
static std::unordered_multimap<size_t, void*> g_parked;
static size_t g_parkedBytes = 0;
constexpr size_t kParkMaxBlock = (size_t)2 << 30;                   // per block
// total; SB_POOL_MB overrides (0 disables parking). The parked bytes are invisible to other allocators in the
// process (PyTorch, NCCL): sbTrimPool() hands them back.
static size_t parkLimitBytes()
{
  static const size_t v = [] {
    const char* e = getenv("SB_POOL_MB");
    return e && *e ? (size_t)strtoull(e, nullptr, 10) << 20 : (size_t)8 << 30;
  }();
  return v;
}

static void releaseParked()
{
  for (auto& kv : g_parked) {
    cudaFree(kv.second);
    g_blockSize.erase(kv.second);
  }
  g_parked.clear();
  g_parkedBytes = 0;
}

static void checkAlignment(size_t alignment)
{
  // cudaMalloc / cudaMallocManaged return >= 256-byte aligned blocks; the reference only ever asks for 64 (ARRAY_ALIGNMENT).
  if (alignment == 0 || (alignment & (alignment - 1)) != 0) {
    fprintf(stderr, "Error: Alignment parameter is not a power of two\n");          // allocate.c:19-22
    exit(EXIT_FAILURE);
  }
  if (alignment > 256) SB_FATAL("allocate: alignment %zu > 256 not supported on the device", alignment);
}

// allocate() of the reference's ABI (allocate.h:9). Its callers are host programs that go on to WRITE the array with
// plain stores (main.c:208-211 fills x and y of `-t spmv` that way) and then hand it to spMVM / waxpby / ddot: the
// block is unified (managed) memory. The entry points that receive such a vector move it to the GPU once, before the
// first kernel that uses it (sb::ensureOnDevice), after which it behaves like device memory. Everything the library
// allocates for itself -- matrices, solver vectors, the peer-mapped windows, which CUDA IPC could not export from
// managed memory -- comes from sbAllocateDevice (cudaMalloc).
struct ManagedBlock {
  size_t bytes;
  bool onDevice;          // prefetched since it was allocated (a later host write migrates pages back: still correct)
};
static std::map<uintptr_t, ManagedBlock> g_managed;

void* allocate(size_t alignment, size_t bytesize)
{
  checkAlignment(alignment);
  ctx();
  // + 256: matrix-SCS.c:224-226 stores nrPadded entries into a y the caller sized with nr (main.c:206)
  const size_t bytes = (((bytesize ? bytesize : 1) + 255) & ~(size_t)255) + 256;
  void* p = nullptr;
  cudaError_t e = cudaMallocManaged(&p, bytes, cudaMemAttachGlobal);
  if (e != cudaSuccess || p == nullptr) {
    fprintf(stderr, "Error: Insufficient memory to fulfill the request (%zu bytes of unified memory: %s)\n", bytesize,
        cudaGetErrorString(e));                                                        // allocate.c:24-33
    exit(EXIT_FAILURE);
  }
  std::lock_guard<std::mutex> lock(g_poolMutex);
  g_managed[(uintptr_t)p] = ManagedBlock { bytes, false };
  return p;
}

void* sbAllocateDevice(size_t alignment, size_t bytesize)
{
  checkAlignment(alignment);
  ctx();
  const size_t bytes = ((bytesize ? bytesize : 1) + 255) & ~(size_t)255;
  std::lock_guard<std::mutex> lock(g_poolMutex);
  auto it = g_parked.find(bytes);
  if (it != g_parked.end()) {
    void* p = it->second;
    g_parked.erase(it);
    g_parkedBytes -= bytes;
    g_blockSize[p].parked = false;
    return p;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {                  // make room: give the parked blocks back first
    cudaGetLastError();
    releaseParked();
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess || p == nullptr) {
    fprintf(stderr, "Error: Insufficient memory to fulfill the request (%zu bytes on device: %s)\n", bytesize,
        cudaGetErrorString(e));
    exit(EXIT_FAILURE);
  }
  g_blockSize[p] = Block { bytes, false };
  return p;
}

// A vector from allocate() (unified memory) that the host has filled: bring the whole block to the GPU before the first
// kernel touches it instead of paying for page faults there. Returns 1 if `p` lies in such a block.
int sbPrefetchManaged(const void* p)
{
  std::lock_guard<std::mutex> lock(g_poolMutex);
  if (g_managed.empty()) return 0;
  auto it = g_managed.upper_bound((uintptr_t)p);
  if (it == g_managed.begin()) return 0;
  --it;
  if ((uintptr_t)p >= it->first + it->second.bytes) return 0;
  if (!it->second.onDevice) {
    static const int mode = getenv("SB_MANAGED_MODE") ? atoi(getenv("SB_MANAGED_MODE")) : 0;
    if (mode & 1) {
      SB_CUDA(cudaMemAdvise((const void*)it->first, it->second.bytes, cudaMemAdviseSetPreferredLocation, g_ctx.device));
      SB_CUDA(cudaMemAdvise((const void*)it->first, it->second.bytes, cudaMemAdviseSetAccessedBy, g_ctx.device));
    }
    if (mode & 2) SB_CUDA(cudaStreamAttachMemAsync(g_ctx.stream, (void*)it->first, 0, cudaMemAttachSingle));
    static const bool trace = getenv("SB_TRACE_MANAGED") != nullptr;
    struct timespec t0, t1;
    if (trace) clock_gettime(CLOCK_MONOTONIC, &t0);
    SB_CUDA(cudaMemPrefetchAsync((const void*)it->first, it->second.bytes, g_ctx.device, g_ctx.stream));
    if (trace) {
      SB_CUDA(cudaStreamSynchronize(g_ctx.stream));
      clock_gettime(CLOCK_MONOTONIC, &t1);
      fprintf(stderr, "[sbPrefetchManaged] %zu bytes -> GPU in %.2f ms\n", it->second.bytes,
          (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
    }
    it->second.onDevice = true;
  }
  return 1;
}

double sbMeasureReadBandwidth(size_t bytes, int reps)
{
  Context& c = ctx();
  bytes &= ~(size_t)15;
  double2* buf = nullptr;
  SB_CUDA(cudaMalloc(&buf, bytes));
  SB_CUDA(cudaMemsetAsync(buf, 0, bytes, c.stream));
  cudaEvent_t a, b;
  SB_CUDA(cudaEventCreate(&a));
  SB_CUDA(cudaEventCreate(&b));
  float best = 1e30f;
  for (int r = 0; r < reps + 2; r++) {
    SB_CUDA(cudaEventRecord(a, c.stream));
    readStreamKernel<<<c.numSMs * 4, 512, 0, c.stream>>>(buf, bytes / 16, c.dWide + 4);
    SB_CUDA(cudaEventRecord(b, c.stream));
    SB_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    SB_CUDA(cudaEventElapsedTime(&ms, a, b));
    if (r >= 2 && ms < best) best = ms;
  }
  SB_CUDA(cudaGetLastError());
  countLaunch(reps + 2);
  SB_CUDA(cudaEventDestroy(a));
  SB_CUDA(cudaEventDestroy(b));
  SB_CUDA(cudaFree(buf));
  return (double)bytes / ((double)best * 1e-3) / 1e9;
}

void sbTrimPool(void)
{
  std::lock_guard<std::mutex> lock(g_poolMutex);
  if (g_ctx.stream) SB_CUDA(cudaStreamSynchronize(g_ctx.stream));
  releaseParked();
}

void sbFree(void* p)
{
  if (!p) return;
  std::lock_guard<std::mutex> lock(g_poolMutex);
  auto mit = g_managed.find((uintptr_t)p);
  if (mit != g_managed.end()) {
    g_managed.erase(mit);
    SB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    SB_CUDA(cudaFree(p));
    return;
  }
  auto it = g_blockSize.find(p);
  if (it == g_blockSize.end()) {           // not ours (or already released): plain free
    SB_CUDA(cudaFree(p));
    return;
  }
  if (it->second.parked) SB_FATAL("sbFree: block %p released twice", p);
  const size_t bytes = it->second.bytes;
  if (bytes <= kParkMaxBlock && g_parkedBytes + bytes <= parkLimitBytes()) {
    // work queued on the stream may still use the block: it may only be reused once that work has drained
    SB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    it->second.parked = true;
    g_parked.emplace(bytes, p);
    g_parkedBytes += bytes;
    return;
  }
  g_blockSize.erase(it);
  SB_CUDA(cudaFree(p));
}

// pinned host memory; small blocks (the solver's scalar mirror) are parked like device blocks
static std::unordered_map<void*, size_t> g_hostSize;
static std::unordered_multimap<size_t, void*> g_hostParked;
constexpr size_t kHostParkMaxBlock = (size_t)1 << 20;

void* sbAllocateHost(size_t bytesize)
{
  ctx();
  const size_t bytes = ((bytesize ? bytesize : 1) + 4095) & ~(size_t)4095;
  {
    std::lock_guard<std::mutex> lock(g_poolMutex);
    auto it = g_hostParked.find(bytes);
    if (it != g_hostParked.end()) {
      void* p = it->second;
      g_hostParked.erase(it);
      return p;
    }
  }
  void* p = nullptr;
  SB_CUDA(cudaMallocHost(&p, bytes));
  std::lock_guard<std::mutex> lock(g_poolMutex);
  g_hostSize[p] = bytes;
  return p;
}

void sbFreeHost(void* p)
{
  if (!p) return;
  {
    std::lock_guard<std::mutex> lock(g_poolMutex);
    auto it = g_hostSize.find(p);
    if (it != g_hostSize.end() && it->second <= kHostParkMaxBlock && g_hostParked.size() < 64) {
      g_hostParked.emplace(it->second, p);
      return;
    }
    if (it != g_hostSize.end()) g_hostSize.erase(it);
  }
  SB_CUDA(cudaFreeHost(p));
}

void sbCopyToDevice(void* dev, const void* host, size_t bytes)
{
  SB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx().stream));
  SB_CUDA(cudaStreamSynchronize(ctx().stream));
}

void sbCopyToHost(void* host, const void* dev, size_t bytes)
{
  SB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx().stream));
  SB_CUDA(cudaStreamSynchronize(ctx().stream));
}

void sbDeviceSynchronize(void) { SB_CUDA(cudaDeviceSynchronize()); }

double getTimeStamp(void)
{
  // The reference brackets every kernel with getTimeStamp() pairs (profiler.h:18-21); with asynchronous
  // launches the stamp is only meaningful once the device has drained.
  SB_CUDA(cudaDeviceSynchronize());
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + (double)ts.tv_nsec * 1.e-9;
}

double getTimeResolution(void)
{
  struct timespec ts;
  clock_getres(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + (double)ts.tv_nsec * 1.e-9;
}

struct SbTimer {
  cudaEvent_t a, b;
};

void* sbTimerCreate(void)
{
  ctx();
  SbTimer* t = new SbTimer;
  SB_CUDA(cudaEventCreate(&t->a));
  SB_CUDA(cudaEventCreate(&t->b));
  return t;
}

void sbTimerStart(void* timer) { SB_CUDA(cudaEventRecord(((SbTimer*)timer)->a, ctx().stream)); }

double sbTimerStopMs(void* timer)
{
  SbTimer* t = (SbTimer*)timer;
  SB_CUDA(cudaEventRecord(t->b, ctx().stream));
  SB_CUDA(cudaEventSynchronize(t->b));
  float ms = 0.f;
  SB_CUDA(cudaEventElapsedTime(&ms, t->a, t->b));
  return (double)ms;
}

void sbTimerDestroy(void* timer)
{
  SbTimer* t = (SbTimer*)timer;
  cudaEventDestroy(t->a);
  cudaEventDestroy(t->b);
  delete t;
}

int sbDeviceCount(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void sbSetDevice(int device)
{
  if (g_ctx.device >= 0 && g_ctx.device != device) SB_FATAL("sbSetDevice(%d) after the context was bound to device %d", device, g_ctx.device);
  g_requestedDevice = device;
  ctx();
}

void sbFlushL2(void)
{
  Context& c = ctx();
  if (!c.flushBuf) {
    c.flushBytes = (size_t)256 << 20;   // 2x the 126 MB L2
    SB_CUDA(cudaMalloc(&c.flushBuf, c.flushBytes));
  }
  SB_CUDA(cudaMemsetAsync(c.flushBuf, 0, c.flushBytes, c.stream));
}

size_t sbKernelLaunchCount(void) { return g_ctx.launches; }

} // extern "C"
