/* TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference sources as P in-process "MPI ranks"
 * (pthreads, see mpi.h / mpi_shim.c) and hands the resulting integer lists, renumbered matrices
 * and full-precision CG histories back to the Python test-suite.
 *
 * Compiled together with /root/reference/src/{comm,solver,CGSolver,matrix,matrix-CRS,...}.c
 * with -D_MPI -DCRS into oracle/_ref/libref_mpi_CRS.so by oracle/Makefile. It only CALLS the
 * reference API (matrixGenerate, commPartition, convertMatrix, commExchange, spMVM, waxpby,
 * ddot, solveCG); the operation order of the re-driven CG below follows CGSolver.c:94-128 so
 * that every iteration's residual is available at full precision (the reference prints %E).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "allocate.h"
#include "comm.h"
#include "matrix.h"
#include "matrixBinfile.h"
#include "parameter.h"
#include "solver.h"

typedef struct {
  int nr, nc, externalCount, totalSendCount, indegree, outdegree;
  int startRow, stopRow;
  int* sources; int* recvCounts; int* rdispls;
  int* destinations; int* sendCounts; int* sdispls;
  int* elementsToSend;
  unsigned int* rowPtr;   /* nr+1 */
  unsigned int* cols;     /* rowPtr[nr] renumbered column ids */
  double* vals;           /* rowPtr[nr] */
  double* haloProbe;      /* externalCount: halo part after commExchange of x[i]=startRow+i */
  int k_solveCG;          /* return value of the reference's own solveCG */
  int k_redriven;         /* loop counter of the re-driven CG (must equal k_solveCG) */
  int nhist;              /* entries in hist: hist[0]=initial, hist[k]=normr set in iteration k */
  double* hist;
  double* x;              /* nr: solution of the re-driven CG */
} RefRankOut;

typedef struct {
  int nx, ny, nz, use7pt, itermax, do_cg;
  double eps;
  RefRankOut* out;
} RunArgs;

static int* dup_int(const int* p, int n)
{
  int* q = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  if (n > 0) memcpy(q, p, sizeof(int) * (size_t)n);
  return q;
}

static void rank_main(int rank, int size, void* argp)
{
  RunArgs* a = (RunArgs*)argp;
  RefRankOut* o = &a->out[rank];
  Comm c;
  memset(&c, 0, sizeof(c));
  commInit(&c, 0, NULL);
  Parameter p;
  p.filename = a->use7pt ? "generate7P" : "generate";
  p.nx = a->nx; p.ny = a->ny; p.nz = a->nz; p.itermax = a->itermax; p.eps = a->eps;

  GMatrix m;
  matrixGenerate(&m, &p, rank, size, a->use7pt != 0);
  commPartition(&c, &m);

  o->nr = (int)m.nr; o->nc = (int)m.nc; o->startRow = (int)m.startRow; o->stopRow = (int)m.stopRow;
  o->externalCount = c.externalCount; o->totalSendCount = c.totalSendCount;
  o->indegree = c.indegree; o->outdegree = c.outdegree;
  o->sources = dup_int(c.sources, c.indegree);
  o->recvCounts = dup_int(c.recvCounts, c.indegree);
  o->rdispls = dup_int(c.rdispls, c.indegree);
  o->destinations = dup_int(c.destinations, c.outdegree);
  o->sendCounts = dup_int(c.sendCounts, c.outdegree);
  o->sdispls = dup_int(c.sdispls, c.outdegree);
  o->elementsToSend = dup_int(c.elementsToSend, c.totalSendCount);
  size_t nnz = m.rowPtr[m.nr];
  o->rowPtr = (unsigned int*)malloc(sizeof(unsigned int) * (m.nr + 1));
  memcpy(o->rowPtr, m.rowPtr, sizeof(unsigned int) * (m.nr + 1));
  o->cols = (unsigned int*)malloc(sizeof(unsigned int) * (nnz ? nnz : 1));
  o->vals = (double*)malloc(sizeof(double) * (nnz ? nnz : 1));
  for (size_t j = 0; j < nnz; j++) { o->cols[j] = m.entries[j].col; o->vals[j] = m.entries[j].val; }

  /* halo probe: exchange a vector that carries global row ids */
  double* probe = (double*)allocate(64, sizeof(double) * (m.nc ? m.nc : 1));
  for (unsigned i = 0; i < m.nc; i++) probe[i] = -1.0;
  for (unsigned i = 0; i < m.nr; i++) probe[i] = (double)(m.startRow + i);
  commExchange(&c, m.nr, probe);
  o->haloProbe = (double*)malloc(sizeof(double) * (size_t)(c.externalCount > 0 ? c.externalCount : 1));
  for (int i = 0; i < c.externalCount; i++) o->haloProbe[i] = probe[m.nr + i];
  free(probe);

  o->k_solveCG = -1; o->k_redriven = -1; o->nhist = 0; o->hist = NULL; o->x = NULL;
  if (a->do_cg) {
    Matrix A;
    convertMatrix(&A, &m);
    unsigned nrow = A.nr, ncol = A.nc;
    double* r  = (double*)allocate(64, sizeof(double) * nrow);
    double* pv = (double*)allocate(64, sizeof(double) * ncol);
    double* Ap = (double*)allocate(64, sizeof(double) * nrow);
    double* x  = (double*)allocate(64, sizeof(double) * nrow);
    double* b  = (double*)allocate(64, sizeof(double) * nrow);
    for (unsigned i = 0; i < nrow; i++) { /* CGSolver.c:19-38, generated-matrix branch */
      int nnzrow = (int)(A.rowPtr[i + 1] - A.rowPtr[i]);
      x[i] = 0.0; b[i] = 27.0 - ((double)(nnzrow - 1));
    }
    for (unsigned i = 0; i < ncol; i++) pv[i] = 0.0;
    o->hist = (double*)malloc(sizeof(double) * (size_t)(a->itermax + 2));
    double eps = a->eps, normr, rtrans = 0.0, oldrtrans = 0.0;
    waxpby(nrow, 1.0, x, 0.0, x, pv);
    commExchange(&c, A.nr, pv);
    spMVM(&A, pv, Ap);
    waxpby(nrow, 1.0, b, -1.0, Ap, r);
    ddot(nrow, r, r, &rtrans);
    normr = sqrt(rtrans);
    o->hist[0] = normr; o->nhist = 1;
    int k;
    for (k = 1; k < a->itermax && normr > eps; k++) {
      if (k == 1) {
        waxpby(nrow, 1.0, r, 0.0, r, pv);
      } else {
        oldrtrans = rtrans;
        ddot(nrow, r, r, &rtrans);
        double beta = rtrans / oldrtrans;
        waxpby(nrow, 1.0, r, beta, pv, pv);
      }
      normr = sqrt(rtrans);
      o->hist[k] = normr; o->nhist = k + 1;
      commExchange(&c, A.nr, pv);
      spMVM(&A, pv, Ap);
      double alpha = 0.0;
      ddot(nrow, pv, Ap, &alpha);
      alpha = rtrans / alpha;
      waxpby(nrow, 1.0, x, alpha, pv, x);
      waxpby(nrow, 1.0, r, -alpha, Ap, r);
    }
    o->k_redriven = k;
    o->x = (double*)malloc(sizeof(double) * (nrow ? nrow : 1));
    memcpy(o->x, x, sizeof(double) * nrow);
    /* and the reference's own driver, for its return value */
    o->k_solveCG = solveCG(&c, &p, &A);
    fflush(stdout);
  }
}

void refdrv_run(int P, int nx, int ny, int nz, int use7pt, int itermax, double eps, int do_cg, RefRankOut* out)
{
  RunArgs a = { nx, ny, nz, use7pt, itermax, do_cg, eps, out };
  shim_run(P, rank_main, &a);
}

void refdrv_free(int P, RefRankOut* out)
{
  for (int r = 0; r < P; r++) {
    RefRankOut* o = &out[r];
    free(o->sources); free(o->recvCounts); free(o->rdispls);
    free(o->destinations); free(o->sendCounts); free(o->sdispls);
    free(o->elementsToSend); free(o->rowPtr); free(o->cols); free(o->vals);
    free(o->haloProbe); free(o->hist); free(o->x);
  }
}


/* ---- file input paths of main.c on P shim ranks: MatrixMarket (main.c:64-71, comm.c:311-402) and .bmx
 * (main.c:36-47, :72-76, matrixBinfile.c:38-236). Hands every rank's GMatrix back. */
typedef struct {
  int nr, nc, nnz, totalNr, totalNnz, startRow, stopRow;
  unsigned int* rowPtr;   /* nr+1 */
  unsigned int* cols;     /* rowPtr[nr] */
  double* vals;
} RefGmOut;

typedef struct { const char* path; const char* path2; RefGmOut* out; int mode; } FileArgs;

static void keep_gm(const GMatrix* m, RefGmOut* o)
{
  o->nr = (int)m->nr; o->nc = (int)m->nc; o->nnz = (int)m->nnz; o->totalNr = (int)m->totalNr; o->totalNnz = (int)m->totalNnz;
  o->startRow = (int)m->startRow; o->stopRow = (int)m->stopRow;
  size_t stored = m->rowPtr[m->nr];
  o->rowPtr = (unsigned int*)malloc(sizeof(unsigned int) * (m->nr + 1));
  memcpy(o->rowPtr, m->rowPtr, sizeof(unsigned int) * (m->nr + 1));
  o->cols = (unsigned int*)malloc(sizeof(unsigned int) * (stored ? stored : 1));
  o->vals = (double*)malloc(sizeof(double) * (stored ? stored : 1));
  for (size_t j = 0; j < stored; j++) { o->cols[j] = m->entries[j].col; o->vals[j] = m->entries[j].val; }
}

static void file_rank_main(int rank, int size, void* argp)
{
  FileArgs* a = (FileArgs*)argp;
  Comm c;
  memset(&c, 0, sizeof(c));
  commInit(&c, 0, NULL);
  GMatrix m;
  if (a->mode == 0 || a->mode == 1) {          /* MatrixMarket: main.c:64-71 (and :36-47 for mode 1) */
    MMMatrix mm, mmLocal;
    if (commIsMaster(&c)) MMMatrixRead(&mm, (char*)a->path);
    commDistributeMatrix(&c, &mm, &mmLocal);
    matrixConvertfromMM(&mmLocal, &m);
    if (a->mode == 1) matrixBinWrite(&m, &c, (char*)a->path2);
  } else {                                      /* .bmx: main.c:72-76 */
    matrixBinRead(&m, &c, (char*)a->path);
  }
  keep_gm(&m, &a->out[rank]);
  fflush(stdout);
}

void refdrv_mm_read(int P, const char* mtx, RefGmOut* out)
{
  FileArgs a = { mtx, NULL, out, 0 };
  shim_run(P, file_rank_main, &a);
}

void refdrv_bmx_write(const char* mtx, const char* bmx, RefGmOut* out)
{
  FileArgs a = { mtx, bmx, out, 1 };
  shim_run(1, file_rank_main, &a);
}

void refdrv_bmx_read(int P, const char* bmx, RefGmOut* out)
{
  FileArgs a = { bmx, NULL, out, 2 };
  shim_run(P, file_rank_main, &a);
}

void refdrv_gm_free(int P, RefGmOut* out)
{
  for (int r = 0; r < P; r++) { free(out[r].rowPtr); free(out[r].cols); free(out[r].vals); }
}
