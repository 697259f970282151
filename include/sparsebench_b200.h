/* sparsebench_b200 -- C ABI of the B200-native SparseBench CG/SpMV hot path.
 *
 * Every entry point below replaces one reference interface (cited file:line, paths relative to the
 * SparseBench tree). Struct layouts are the reference's for CG_UINT = unsigned int, CG_FLOAT = double
 * (util.h:35-53), so host code that only forwards the structs keeps working. What changes:
 *   - array members filled by convertMatrix() and everything returned by allocate() are DEVICE pointers;
 *   - x / y / w / p vectors handed to spMVM, waxpby, ddot, commExchange are DEVICE pointers
 *     (obtained from allocate()); scalars (alpha, beta, *result, Parameter) stay on the host;
 *   - errors follow the reference convention: message on stderr + exit(EXIT_FAILURE)
 *     (allocate.c:19-33, comm.c:462-468). A missing/failed CUDA device is such an error: there is
 *     no CPU fallback anywhere in this library.
 * The reference selects the matrix format at link time (Makefile:20,32-34): one object defines
 * convertMatrix/spMVM for the -D<FMT> Matrix typedef (matrix.h:14-22). Here the three formats live in
 * one library under sbCRS_/sbSCS_/sbCCRS_ prefixes; libsparsebench_b200_<FMT>.so re-exports them
 * under the reference's bare names (convertMatrix, spMVM, solveCG) for link-time substitution, and
 * defining CRS, SCS or CCRS before including this header gives the same mapping at compile time.
 */
#ifndef SPARSEBENCH_B200_H
#define SPARSEBENCH_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* util.h:35-53: the reference's compile-time type switches. PRECISION=2 / UINT_TYPE=1 (double, unsigned int) are the
 * defaults and what libsparsebench_b200.so is built with; the variant libraries libsparsebench_b200_f32.so
 * (-DPRECISION=1), _u64.so (-DUINT_TYPE=2) and _f32u64.so are the same sources compiled with the other settings --
 * compile the caller with the same -DPRECISION / -DUINT_TYPE and link the matching library. When the reference's util.h
 * was included first (a reference translation unit compiled against this header, see integration/), its CG_UINT /
 * CG_FLOAT macros are used. */
#ifndef PRECISION
#define PRECISION 2
#endif
#ifndef UINT_TYPE
#define UINT_TYPE 1
#endif
#ifndef CG_UINT
#if UINT_TYPE == 1
typedef unsigned int CG_UINT;
#else
typedef unsigned long long int CG_UINT;
#endif
#endif
#ifndef CG_FLOAT
#if PRECISION == 1
typedef float CG_FLOAT;
#else
typedef double CG_FLOAT;
#endif
#endif

/* ---------------------------------------------------------------- data structures */
typedef struct {                /* matrix.h:24-27 (16 bytes: 4 B padding after col) */
  CG_UINT col;
  CG_FLOAT val;
} Entry;

typedef struct {                /* matrix.h:29-35 */
  CG_UINT nr, nc, nnz;
  CG_UINT totalNr, totalNnz;
  CG_UINT startRow, stopRow;
  CG_UINT* rowPtr;              /* host (reference generator/reader) or device (sbGenerateDevice) */
  Entry* entries;
} GMatrix;

typedef struct {                /* CRSMatrix.h:9-16 */
  CG_UINT nr, nc, nnz;
  CG_UINT totalNr, totalNnz;
  CG_UINT startRow, stopRow;
  CG_UINT* rowPtr;              /* device, nr+1 */
  CG_UINT* colInd;              /* device, rowPtr[nr] valid entries */
  CG_FLOAT* val;                /* device */
} SbCRSMatrix;

typedef struct {                /* SCSMatrix.h:13-27 */
  CG_UINT nr, nc, nnz;
  CG_UINT totalNr, totalNnz;
  CG_UINT startRow, stopRow;
  CG_UINT* colInd;              /* device, nElems, reference numbering (un-permuted columns) */
  CG_FLOAT* val;                /* device, nElems */
  CG_UINT C, sigma;             /* INPUTS: set by the caller before convertMatrix (matrix-SCS.c:40) */
  CG_UINT nrPadded, nChunks;
  CG_UINT nElems;
  CG_UINT* chunkPtr;            /* device, nChunks+1 */
  CG_UINT* chunkLens;           /* device, nChunks */
  CG_UINT* oldToNewPerm;        /* device, nr */
  CG_UINT* newToOldPerm;        /* device, nr */
} SbSCSMatrix;

typedef struct {                /* CCRSMatrix.h:9-20: same field order as GMatrix */
  CG_UINT nr, nc, nnz;
  CG_UINT totalNr, totalNnz;
  CG_UINT startRow, stopRow;
  CG_UINT* rowPtr;              /* device */
  Entry* entries;               /* device, 16-byte {col,val} records */
} SbCCRSMatrix;

#if defined(CRS)
typedef SbCRSMatrix Matrix;
#define convertMatrix sbCRS_convertMatrix
#define spMVM sbCRS_spMVM
#define solveCG sbCRS_solveCG
#elif defined(SCS)
typedef SbSCSMatrix Matrix;
#define convertMatrix sbSCS_convertMatrix
#define spMVM sbSCS_spMVM
#define solveCG sbSCS_solveCG
#elif defined(CCRS)
typedef SbCCRSMatrix Matrix;
#define convertMatrix sbCCRS_convertMatrix
#define spMVM sbCCRS_spMVM
#define solveCG sbCCRS_solveCG
#endif

typedef struct {                /* matrix.h:37-41 */
  int row;
  int col;
  double val;
} MMEntry;

typedef struct {                /* matrix.h:43-49 */
  size_t count;
  int nr, nnz;
  int totalNr, totalNnz;
  int startRow, stopRow;
  MMEntry* entries;
} MMMatrix;

#ifndef __PARAMETER_H_           /* the reference's own parameter.h (identical layout) wins when it was included first */
typedef struct {                /* parameter.h:8-13 */
  char* filename;
  int nx, ny, nz;
  int itermax;
  double eps;
} Parameter;
#endif

enum { SB_MAX = 0, SB_SUM = 1 };   /* comm.h:25 `enum op { MAX = 0, SUM }` */

typedef struct {                /* comm.h:27-46, the _MPI member set is always present */
  int rank;
  int size;
  FILE* logFile;
  int externalCount;
  int totalSendCount;
  int* elementsToSend;          /* host copy (bit-exact vs comm.c:116-182); device copy kept internally */
  int indegree;
  int outdegree;
  int* sources;
  int* recvCounts;
  int* rdispls;
  int* destinations;
  int* sendCounts;
  int* sdispls;
  CG_FLOAT* sendBuffer;         /* device, totalSendCount */
  void* communicator;           /* replaces MPI_Comm: opaque handle (NCCL communicator, peer windows, streams) */
} Comm;

/* ---------------------------------------------------------------- runtime (allocate.c, timing.c) */
/* allocate.h:9 -- what a reference caller gets for its vectors: UNIFIED memory, so host code that fills the array with
 * plain stores (main.c:208-211, `-t spmv`) keeps working; spMVM / waxpby / ddot / commExchange move such a block to the
 * GPU once, before the first kernel that touches it, after which it is device-resident. Alignment honoured up to 256 B;
 * exits on failure (allocate.c:19-33). */
void* allocate(size_t alignment, size_t bytesize);
/* plain device memory (cudaMalloc, parked-block cache): what the library uses for itself and what a caller that never
 * dereferences the pointer on the host should use */
void* sbAllocateDevice(size_t alignment, size_t bytesize);
int sbPrefetchManaged(const void* p);                /* allocate()d block containing p -> GPU now; returns 1 if p is in one */
void sbFree(void* devPtr);                           /* released blocks are parked for reuse (SB_POOL_MB caps the cache, default 8192) */
void sbTrimPool(void);                               /* returns every parked block to the driver (before handing the GPU to another allocator) */
void* sbAllocateHost(size_t bytesize);               /* pinned host memory for staging */
void sbFreeHost(void* hostPtr);
void sbCopyToDevice(void* dev, const void* host, size_t bytes);
void sbCopyToHost(void* host, const void* dev, size_t bytes);
void sbDeviceSynchronize(void);
/* timing.h:8-9 -- monotonic seconds AFTER draining the device, so PROFILE(tag, call) (profiler.h:18-21)
 * around asynchronous launches still measures the call */
double getTimeStamp(void);
double getTimeResolution(void);
/* CUDA-event timing on the library's stream (replaces getTimeStamp pairs in measurement code) */
void* sbTimerCreate(void);
void sbTimerStart(void* timer);
double sbTimerStopMs(void* timer);                   /* records stop, waits for it, returns milliseconds */
void sbTimerDestroy(void* timer);
int sbDeviceCount(void);
void sbSetDevice(int device);
void sbFlushL2(void);                                /* overwrites a buffer larger than L2 */
/* GB/s of a read-only streaming kernel over a fresh buffer of `bytes` (>> L2), best of `reps`: the read-stream peak
 * quoted beside the copy peak of MEASURED_PEAKS.json (the SpMV is a ~98 % read stream) */
double sbMeasureReadBandwidth(size_t bytes, int reps);
size_t sbKernelLaunchCount(void);                    /* number of this library's kernel launches so far */

/* ---------------------------------------------------------------- matrix sources */
/* matrix.h:52-53 / matrix.c:30-121 -- HPCG 27-pt / 7-pt block of rank `rank` of `size`, host arrays */
void matrixGenerate(GMatrix* m, Parameter* p, int rank, int size, bool use_7pt_stencil);
/* same matrix, generated directly in device memory (rowPtr/entries are device pointers) */
void sbGenerateDevice(GMatrix* m, Parameter* p, int rank, int size, bool use_7pt_stencil);
void sbFreeGMatrix(GMatrix* m);                      /* releases host or device arrays of a GMatrix */
/* matrix.h:50-51 / matrix.c:123-269 -- MatrixMarket coordinate files (real/integer/pattern, general/symmetric);
 * host arrays, sorted by row then column exactly like the reference (stable sorts) */
void MMMatrixRead(MMMatrix* m, char* filename);
void matrixConvertfromMM(MMMatrix* mm, GMatrix* m);

/* ---------------------------------------------------------------- format plugins */
/* matrix.h:57 convertMatrix / solver.h:13 spMVM, one pair per format. `im` may hold host or device arrays. */
void sbCRS_convertMatrix(SbCRSMatrix* m, GMatrix* im);                        /* matrix-CRS.c:12-44 */
void sbCRS_spMVM(SbCRSMatrix* m, const CG_FLOAT* x, CG_FLOAT* y);             /* matrix-CRS.c:46-65 */
void sbSCS_convertMatrix(SbSCSMatrix* m, GMatrix* im);                        /* matrix-SCS.c:31-196 (w/o :42-43) */
void sbSCS_spMVM(SbSCSMatrix* m, const CG_FLOAT* x, CG_FLOAT* y);             /* matrix-SCS.c:198-228 */
void sbCCRS_convertMatrix(SbCCRSMatrix* m, GMatrix* im);                      /* matrix-CCRS.c:12 (alias intent) */
void sbCCRS_spMVM(SbCCRSMatrix* m, const CG_FLOAT* x, CG_FLOAT* y);           /* matrix-CCRS.c:14-31 */
/* y = A x in ONE launch that takes the units [intLo, intHi) first (rows for CRS/CCRS, chunks for SCS) and the rest
 * afterwards: the kernel the multi-GPU CG uses to overlap the halo exchange, here without a
 * halo to wait for. Returns 0 if the matrix has no pipelined kernel (then nothing was launched). */
int sbSpmvOrdered(void* matrix, int fmt, const CG_FLOAT* x, CG_FLOAT* y, CG_UINT intLo, CG_UINT intHi);
/* y = A x fused with *dDot = x . y (dDot: device scalar) -- spMVM + ddot of CGSolver.c:123-125 in one pass, asynchronous.
 * Vectors in solver order (SCS with sigma > 1: permuted row order for x and y, as inside sbSolveCG). */
void sbSpmvDot(void* matrix, int fmt, const CG_FLOAT* x, CG_FLOAT* y, CG_FLOAT* dDot);
/* Which SpMV kernel family takes this matrix (decided once per matrix from its row lengths; blocks the stream the
 * first time): 0 SELL-32 bulk-copy rings, 1 CRS/CCRS tiles of a fixed row count through the bulk-copy pipeline
 * (near-uniform rows: the stencil), 2 CRS/CCRS blocks of bounded non-zero count (uneven rows: longest row >
 * 1.2 avg + 4), 3 register-staged kernels (SELL with C != 32, rows too long for a tile, SB_SPMV_LEGACY), 4 SELL-32 rings for
 * the ordinary chunks + one CTA per chunk longer than 256 columns (heavy-tailed matrices: longest chunk > 4x the average). */
int sbSpmvKernelFamily(void* matrix, int fmt);
void sbCRS_destroyMatrix(SbCRSMatrix* m);
void sbSCS_destroyMatrix(SbSCSMatrix* m);
void sbCCRS_destroyMatrix(SbCCRSMatrix* m);

/* ---------------------------------------------------------------- Krylov vector kernels (solver.c) */
void waxpby(const CG_UINT n, const CG_FLOAT alpha, const CG_FLOAT* x, const CG_FLOAT beta,
    const CG_FLOAT* y, CG_FLOAT* w);                                          /* solver.h:15-20, solver.c:16-39 */
void ddot(const CG_UINT n, const CG_FLOAT* x, const CG_FLOAT* y, CG_FLOAT* result); /* solver.h:22-25, solver.c:41-62 */

/* ---------------------------------------------------------------- CG driver (CGSolver.c) */
int sbCRS_solveCG(Comm* comm, Parameter* param, SbCRSMatrix* m);              /* solver.h:11, CGSolver.c:62-141 */
int sbSCS_solveCG(Comm* comm, Parameter* param, SbSCSMatrix* m);
int sbCCRS_solveCG(Comm* comm, Parameter* param, SbCCRSMatrix* m);

enum { SB_FMT_CRS = 0, SB_FMT_SCS = 1, SB_FMT_CCRS = 2 };
enum {
  SB_CG_FUSED = 1,        /* single-pass fused kernels + device-resident scalars (default product path) */
  SB_CG_PRINT = 2,        /* print the reference's stdout lines (CGSolver.c:102,119,133,58) */
  SB_CG_HOST_VECTORS = 4, /* b / x are HOST buffers: upload b,x0 and download x inside the call */
  SB_CG_NO_OVERLAP = 8,   /* multi-GPU: do not overlap the halo exchange with interior rows */
  SB_CG_PROFILE = 16      /* CUDA events around every kernel of the loop -> regionMs (measurement runs only) */
};
enum { SB_REGION_UPDATE_P = 0, SB_REGION_EXCHANGE, SB_REGION_SPMV, SB_REGION_ALLREDUCE, SB_REGION_UPDATE_XR,
       SB_REGION_HALO_WAIT, SB_REGION_SPMV_BOUNDARY, SB_REGION_COUNT };
typedef struct {
  int flags;
  const CG_FLOAT* b;      /* right-hand side, nr entries; NULL -> initVectors rule (CGSolver.c:19-38) */
  CG_FLOAT* x;            /* in: start vector, out: solution, nr entries (original row order); NULL -> x0 = 0, discarded */
  double* history;        /* host, capacity historyCap: history[0] initial ||r||, history[k] = normr of iteration k */
  int historyCap;
  int nhist;              /* out */
  double solveMs;         /* out: device time of the iteration loop (CUDA events), CGSolver.c:106,130 */
  double maxError;        /* out: max|x - 1| for generated matrices (CGSolver.c:40-60), else -1 */
  double regionMs[SB_REGION_COUNT]; /* out (SB_CG_PROFILE): device time per kernel class, replaces _t[] of profiler.c:17 */
  double createMs, finishMs; /* out (sbSolveCG): host wall time before / after the loop: allocation, uploads, pre-loop (CGSolver.c:69-102) / check, download, free */
} SbCGInfo;
/* solveCG with explicit right-hand side / history capture; returns the reference's k */
int sbSolveCG(Comm* comm, Parameter* param, void* matrix, int fmt, SbCGInfo* info);
/* the same solve in three steps, so that a harness can time exactly K iterations after W warm-up iterations:
 * create = allocate + initVectors + pre-loop (CGSolver.c:69-102); iterate = run the loop while
 * k < min(untilK, itermax) && normr > eps, asynchronously (returns k); finish = drain, report, free. */
void* sbCGCreate(Comm* comm, Parameter* param, void* matrix, int fmt, SbCGInfo* info);
int sbCGIterate(void* solver, int untilK);
int sbCGFinish(void* solver, SbCGInfo* info, double loopMs);

/* ---------------------------------------------------------------- the solver types main.c:22 names but never implements
 * (`-t gmres` prints its name and returns, main.c:217-222; `-t cheb` has no case). No reference behaviour exists; both are
 * built on the same SpMV kernels, halo exchange and deterministic reductions as the CG.
 * sbSolveGMRES: restarted GMRES(restart) (classical Gram-Schmidt with fused projections, Givens rotations on the host);
 *   b / x / history / flags as in sbSolveCG; history[k] = |g_{k+1}| (the Arnoldi residual estimate) after k
 *   matrix-vector products; stops when it is <= param->eps or after param->itermax - 1 products; returns their number.
 * sbChebyshevFilter: y = sum_{k<=degree} coef[k] T_k(A~) x (coef == NULL: y = T_degree(A~) x; y == NULL: moments only) and
 *   moments[k] = x . T_k(A~) x, A~ = (A - c I) / e with c, e from [lambdaMin, lambdaMax] -- the kernel of Chebyshev filter
 *   diagonalisation / the kernel polynomial method. x, y: nr entries, host or device; moments: degree + 1, host or device. */
int sbSolveGMRES(Comm* comm, Parameter* param, void* matrix, int fmt, SbCGInfo* info, int restart);
void sbChebyshevFilter(Comm* comm, void* matrix, int fmt, int degree, double lambdaMin, double lambdaMax, const CG_FLOAT* coef,
    const CG_FLOAT* x, CG_FLOAT* y, CG_FLOAT* moments);

/* ---------------------------------------------------------------- communication (comm.c) */
void commInit(Comm* c, int argc, char** argv);                                /* comm.h:48, comm.c:863-878 */
void commPrintBanner(Comm* c);                                                /* comm.h:59, comm.c:185-274 (one line per rank: GPU instead of CPU affinity) */
void commAbort(Comm* c, char* msg);                                           /* comm.h:60, comm.c:880-891: finalize + exit(EXIT_SUCCESS) */
void commFinalize(Comm* c);                                                   /* comm.h:49, comm.c:893-910 */
void commPartition(Comm* c, GMatrix* m);                                      /* comm.h:51, comm.c:414-625 */
void commExchange(Comm* c, CG_UINT numRows, CG_FLOAT* x);                     /* comm.h:57, comm.c:627-651 */
void commReduction(CG_FLOAT* v, int op);                                      /* comm.h:58, comm.c:653-662 (host scalar) */
void sbCommAllreduceDevice(Comm* c, CG_FLOAT* dev, int count, int op);        /* the same reduction on device scalars, asynchronous */
void commDistributeMatrix(Comm* c, MMMatrix* m, MMMatrix* mLocal);            /* comm.h:50, comm.c:311-412 */
/* matrixBinfile.h:22-23 / matrixBinfile.c:38-236 -- .bmx files (24-byte header, u32 sizes and row offsets, {u32 col;
 * f32 val} records); plain POSIX I/O instead of MPI-IO, every rank reads its own row block; host arrays */
void matrixBinWrite(GMatrix* m, Comm* c, char* filename);
void matrixBinRead(GMatrix* m, Comm* c, char* filename);
/* bootstrap pieces used when another launcher (torchrun) already owns the rendezvous */
int sbCommUniqueIdBytes(void);
void sbCommGetUniqueId(void* id);
void sbCommInitRank(Comm* c, int rank, int size, int device, const void* id);
/* commPartition split at its two small exchanges, for launchers that move them over their own transport
 * (commPartition itself uses NCCL; the CPU tests use torch.distributed/gloo). Pure host integer work when
 * `m` holds host arrays; a device GMatrix is scanned and renumbered by kernels.
 *   1. all-gather every rank's m->startRow                      -> startRows[size]        (comm.c:496)
 *   2. plan = sbPartitionLocal(...): renumbers m->entries[].col in place, m->nc += externalCount,
 *      fills wantCounts[owner] = number of halo entries owned by `owner`                  (comm.c:452-520, :40-114)
 *   3. all-gather wantCounts                                    -> wantMatrix[size*size], row = requester
 *   4. for every source s with wantMatrix[rank][s] > 0 send sbPartitionRequestSlice(plan, wantMatrix, s)
 *      to s; concatenate what arrives by ascending requester    -> received             (comm.c:130-161)
 *   5. sbPartitionFinish fills the Comm lists (malloc'ed, released by commFinalize) and frees the plan. */
typedef struct SbPartitionPlan SbPartitionPlan;
SbPartitionPlan* sbPartitionLocal(GMatrix* m, int rank, int size, const CG_UINT* startRows, int* wantCounts);
const int* sbPartitionRequestSlice(SbPartitionPlan* plan, const int* wantMatrix, int source, int* count);
void sbPartitionFinish(SbPartitionPlan* plan, Comm* c, const int* wantMatrix, const int* received);

#ifdef __cplusplus
}
#endif
#endif /* SPARSEBENCH_B200_H */
