// Runtime utilities of the hot path: device allocation (replaces allocate.c:12-36), timing (replaces
// timing.c:8-20 / the PROFILE macro's time source, profiler.h:18-21) and the per-process device context.
#include <time.h>

#include <mutex>
#include <unordered_map>

#include "sb_internal.h"

namespace sb {

static Context g_ctx;
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
}

Context& ctx()
{
  std::call_once(g_once, initContext);
  return g_ctx;
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

void* allocate(size_t alignment, size_t bytesize)
{
  // cudaMalloc returns >= 256-byte aligned blocks; the reference only ever asks for 64 (ARRAY_ALIGNMENT).
  if (alignment == 0 || (alignment & (alignment - 1)) != 0) {
    fprintf(stderr, "Error: Alignment parameter is not a power of two\n");
    exit(EXIT_FAILURE);
  }
  if (alignment > 256) SB_FATAL("allocate: alignment %zu > 256 not supported on the device", alignment);
  ctx();
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytesize ? bytesize : 1);
  if (e != cudaSuccess || p == nullptr) {
    fprintf(stderr, "Error: Insufficient memory to fulfill the request (%zu bytes on device: %s)\n", bytesize,
        cudaGetErrorString(e));
    exit(EXIT_FAILURE);
  }
  return p;
}

void sbFree(void* p)
{
  if (p) SB_CUDA(cudaFree(p));
}

void* sbAllocateHost(size_t bytesize)
{
  ctx();
  void* p = nullptr;
  SB_CUDA(cudaMallocHost(&p, bytesize ? bytesize : 1));
  return p;
}

void sbFreeHost(void* p)
{
  if (p) SB_CUDA(cudaFreeHost(p));
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
