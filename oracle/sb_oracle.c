/* TEST INFRASTRUCTURE ONLY -- see sb_oracle.h.
 *
 * Plain C, single thread, IEEE double without contraction or reassociation (built with
 * -O2 -fno-fast-math -ffp-contract=off), so every sum is the left-to-right sum the reference's
 * source text describes. Citations are into /root/reference/src/.
 */
#include "sb_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ generator: matrix.c:30-121 */
int64_t orc_generate(int nx, int ny, int nz, int rank, int size, int use7pt,
    uint32_t* rowPtr, uint32_t* col, double* val, int64_t cap)
{
  const int64_t plane = (int64_t)nx * ny;
  const int64_t localRows = plane * nz;
  const int64_t totalRows = localRows * size;     /* ranks are stacked along z (matrix.c:34-41) */
  const int64_t first = localRows * rank;
  int64_t n = 0, row = 0;
  if (rowPtr) rowPtr[0] = 0;
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++, row++) {
        const int64_t g = first + row;
        for (int dz = -1; dz <= 1; dz++)
          for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
              /* x and y are clipped per block, z only by the global row range (matrix.c:76-83) */
              if (x + dx < 0 || x + dx >= nx || y + dy < 0 || y + dy >= ny) continue;
              const int64_t c = g + dz * plane + (int64_t)dy * nx + dx;
              if (c < 0 || c >= totalRows) continue;
              if (use7pt && dz * dz + dy * dy + dx * dx > 1) continue;
              if (col) {
                if (n >= cap) return -1;
                col[n] = (uint32_t)c;
                val[n] = (c == g) ? 27.0 : -1.0;   /* matrix.c:87-91 */
              }
              n++;
            }
        if (rowPtr) rowPtr[row + 1] = (uint32_t)n;
      }
  return n;
}

/* ------------------------------------------------------------------ CGSolver.c:19-38 */
void orc_init_vectors(uint32_t nr, const uint32_t* rowPtr, int generated, double* x, double* b, double* xexact)
{
  for (uint32_t i = 0; i < nr; i++) {
    int len = (int)(rowPtr[i + 1] - rowPtr[i]);
    x[i] = 0.0;
    if (generated) {
      b[i] = 27.0 - (double)(len - 1);
      if (xexact) xexact[i] = 1.0;
    } else {
      b[i] = 1.0;
    }
  }
}

/* ------------------------------------------------------------------ SpMV kernels */
void orc_spmv_crs(uint32_t nr, const uint32_t* rowPtr, const uint32_t* col, const double* val,
    const double* x, double* y)
{ /* matrix-CRS.c:54-64 */
  for (uint32_t i = 0; i < nr; i++) {
    double s = 0.0;
    for (uint32_t j = rowPtr[i]; j < rowPtr[i + 1]; j++) s += val[j] * x[col[j]];
    y[i] = s;
  }
}

void orc_spmv_ccrs(uint32_t nr, const uint32_t* rowPtr, const OrcEntry* e, const double* x, double* y)
{ /* matrix-CCRS.c:20-30 */
  for (uint32_t i = 0; i < nr; i++) {
    double s = 0.0;
    for (uint32_t j = rowPtr[i]; j < rowPtr[i + 1]; j++) s += e[j].val * x[e[j].col];
    y[i] = s;
  }
}

void orc_spmv_scs(uint32_t nChunks, uint32_t C, const uint32_t* chunkPtr, const uint32_t* chunkLens,
    const uint32_t* col, const double* val, const double* x, double* y)
{ /* matrix-SCS.c:208-227: every lane of the chunk (padding rows included) is accumulated and stored */
  double* acc = (double*)malloc(sizeof(double) * C);
  for (uint32_t ch = 0; ch < nChunks; ch++) {
    for (uint32_t k = 0; k < C; k++) acc[k] = 0.0;
    const uint64_t base = chunkPtr[ch];
    for (uint32_t j = 0; j < chunkLens[ch]; j++)
      for (uint32_t k = 0; k < C; k++) {
        const uint64_t e = base + (uint64_t)j * C + k;
        acc[k] += val[e] * x[col[e]];
      }
    for (uint32_t k = 0; k < C; k++) y[(uint64_t)ch * C + k] = acc[k];
  }
  free(acc);
}

/* ------------------------------------------------------------------ SELL-C-sigma: matrix-SCS.c:31-196 */
typedef struct { int index; int count; } RowLen;

/* stable merge sort, descending count (the reference relies on glibc qsort being a mergesort, :67-78) */
static void sort_window(RowLen* a, RowLen* tmp, int n)
{
  for (int w = 1; w < n; w *= 2) {
    for (int lo = 0; lo < n; lo += 2 * w) {
      int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      int i = lo, j = mid, k = lo;
      while (i < mid && j < hi) tmp[k++] = (a[j].count > a[i].count) ? a[j++] : a[i++];
      while (i < mid) tmp[k++] = a[i++];
      while (j < hi) tmp[k++] = a[j++];
    }
    memcpy(a, tmp, sizeof(RowLen) * (size_t)n);
  }
}

int64_t orc_scs_structure(uint32_t nr, uint32_t C, uint32_t sigma, const uint32_t* rowPtr,
    uint32_t* oldToNew, uint32_t* newToOld, uint32_t* chunkLens, uint32_t* chunkPtr)
{
  const uint32_t nChunks = (nr + C - 1) / C;        /* :40 */
  const uint32_t nrPadded = nChunks * C;            /* :41 */
  RowLen* rows = (RowLen*)malloc(sizeof(RowLen) * (nrPadded ? nrPadded : 1));
  RowLen* tmp = (RowLen*)malloc(sizeof(RowLen) * (sigma ? sigma : 1));
  for (uint32_t i = 0; i < nrPadded; i++) {         /* :49-58: padding rows have length 0 */
    rows[i].index = (int)i;
    rows[i].count = i < nr ? (int)(rowPtr[i + 1] - rowPtr[i]) : 0;
  }
  for (uint32_t lo = 0; lo < nrPadded; lo += sigma) {   /* :61-79 */
    uint32_t n = (lo + sigma < nrPadded) ? sigma : nrPadded - lo;
    sort_window(rows + lo, tmp, (int)n);
  }
  uint64_t cursor = 0;
  for (uint32_t ch = 0; ch < nChunks; ch++) {       /* :88-117 */
    uint32_t longest = 0;
    for (uint32_t k = 0; k < C; k++) {
      uint32_t len = (uint32_t)rows[ch * C + k].count;
      if (len > longest) longest = len;
    }
    chunkLens[ch] = longest;
    chunkPtr[ch] = (uint32_t)cursor;
    cursor += (uint64_t)longest * C;
  }
  chunkPtr[nChunks] = (uint32_t)cursor;             /* :112-114 */
  for (uint32_t pos = 0; pos < nrPadded; pos++)     /* :120-125 */
    if ((uint32_t)rows[pos].index < nr) oldToNew[rows[pos].index] = pos;
  for (uint32_t i = 0; i < nr; i++) newToOld[oldToNew[i]] = i;   /* :128-143 */
  free(rows);
  free(tmp);
  return (int64_t)cursor;
}

void orc_scs_fill(uint32_t nr, uint32_t C, const uint32_t* rowPtr, const uint32_t* col, const double* val,
    const uint32_t* oldToNew, const uint32_t* chunkPtr, int64_t nElems, uint32_t* colOut, double* valOut)
{
  for (int64_t e = 0; e < nElems; e++) { valOut[e] = 0.0; colOut[e] = 0; }   /* :150-155 */
  for (uint32_t i = 0; i < nr; i++) {                                       /* :164-192 */
    const uint32_t r = oldToNew[i];
    const uint64_t base = (uint64_t)chunkPtr[r / C] + r % C;
    uint32_t seen = 0;
    for (uint32_t j = rowPtr[i]; j < rowPtr[i + 1]; j++, seen++) {
      colOut[base + (uint64_t)seen * C] = col[j];
      valOut[base + (uint64_t)seen * C] = val[j];
    }
  }
}

/* ------------------------------------------------------------------ vector kernels: solver.c */
void orc_waxpby(uint32_t n, double alpha, const double* x, double beta, const double* y, double* w)
{ /* solver.c:23-38: the three branches differ in rounding (no multiply by the unit factor) */
  if (alpha == 1.0)      for (uint32_t i = 0; i < n; i++) w[i] = x[i] + beta * y[i];
  else if (beta == 1.0)  for (uint32_t i = 0; i < n; i++) w[i] = alpha * x[i] + y[i];
  else                   for (uint32_t i = 0; i < n; i++) w[i] = alpha * x[i] + beta * y[i];
}

/* solver.c:46-61: `#pragma omp parallel for reduction(+ : sum) schedule(static)`. The summation order depends on the
 * thread count T of the run: every thread sums one contiguous chunk left to right (libgomp's static schedule: n / T
 * elements each, the first n % T threads one more), the partial sums are combined afterwards (thread order here; OpenMP
 * leaves it open). T = 1 is the sequential loop of the strict single-thread build. T = 0 is not a reference mode: it
 * accumulates in long double and serves as the rounding-free yardstick when two summation orders are compared. */
static int g_dot_threads = 1;
void orc_set_dot_threads(int t) { g_dot_threads = t < 0 ? 1 : t; }
int orc_get_dot_threads(void) { return g_dot_threads; }

double orc_ddot(uint32_t n, const double* x, const double* y)
{
  if (g_dot_threads == 0) {
    long double s = 0.0L;
    for (uint32_t i = 0; i < n; i++) s += (long double)x[i] * (long double)y[i];
    return (double)s;
  }
  const uint32_t T = (uint32_t)g_dot_threads;
  double total = 0.0;
  uint32_t at = 0;
  for (uint32_t t = 0; t < T; t++) {
    const uint32_t len = n / T + (t < n % T ? 1u : 0u);
    double s = 0.0;
    for (uint32_t i = at; i < at + len; i++) s += x[i] * y[i];
    total = (t == 0) ? s : total + s;
    at += len;
  }
  return total;
}

/* ------------------------------------------------------------------ CG: CGSolver.c:62-141 */
int orc_cg_crs(uint32_t nr, uint32_t nc, const uint32_t* rowPtr, const uint32_t* col, const double* val,
    const double* b, double* x, int itermax, double eps, double* hist, int* nhist)
{
  double* r = (double*)calloc(nr ? nr : 1, sizeof(double));
  double* p = (double*)calloc(nc ? nc : 1, sizeof(double));
  double* Ap = (double*)calloc(nr ? nr : 1, sizeof(double));
  double rtrans = 0.0, oldrtrans = 0.0, normr;
  orc_waxpby(nr, 1.0, x, 0.0, x, p);                 /* :94 */
  orc_spmv_crs(nr, rowPtr, col, val, p, Ap);         /* :96 */
  orc_waxpby(nr, 1.0, b, -1.0, Ap, r);               /* :97 */
  rtrans = orc_ddot(nr, r, r);                       /* :98 */
  normr = sqrt(rtrans);                              /* :100 */
  hist[0] = normr; *nhist = 1;
  int k;
  for (k = 1; k < itermax && normr > eps; k++) {     /* :107 lagging test */
    if (k == 1) {
      orc_waxpby(nr, 1.0, r, 0.0, r, p);             /* :109 */
    } else {
      oldrtrans = rtrans;
      rtrans = orc_ddot(nr, r, r);                   /* :112 */
      double beta = rtrans / oldrtrans;
      orc_waxpby(nr, 1.0, r, beta, p, p);            /* :114 */
    }
    normr = sqrt(rtrans);                            /* :116 */
    hist[k] = normr; *nhist = k + 1;
    orc_spmv_crs(nr, rowPtr, col, val, p, Ap);       /* :123 */
    double alpha = orc_ddot(nr, p, Ap);              /* :125 */
    alpha = rtrans / alpha;                          /* :126 */
    orc_waxpby(nr, 1.0, x, alpha, p, x);             /* :127 */
    orc_waxpby(nr, 1.0, r, -alpha, Ap, r);           /* :128 */
  }
  free(r); free(p); free(Ap);
  return k;
}

/* ------------------------------------------------------------------ partition: comm.c:414-625 */
typedef struct { uint32_t* key; int* ord; uint32_t mask; } ColSet;

static void set_init(ColSet* s, size_t expect)
{
  uint32_t cap = 16;
  while (cap < 2 * expect + 8) cap *= 2;
  s->key = (uint32_t*)malloc(sizeof(uint32_t) * cap);
  s->ord = (int*)malloc(sizeof(int) * cap);
  for (uint32_t i = 0; i < cap; i++) s->ord[i] = -1;
  s->mask = cap - 1;
}
static inline uint32_t set_slot(const ColSet* s, uint32_t k)
{
  uint32_t h = (k * 2654435761u) & s->mask;
  while (s->ord[h] >= 0 && s->key[h] != k) h = (h + 1) & s->mask;
  return h;
}
static void set_free(ColSet* s) { free(s->key); free(s->ord); }

int orc_partition_all(int P, OrcRank* R)
{
  uint32_t* starts = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)P);   /* :496 Allgather of startRow */
  int** want = (int**)malloc(sizeof(int*) * (size_t)P);                 /* want[r][owner] = #externals */
  for (int r = 0; r < P; r++) starts[r] = R[r].startRow;

  for (int r = 0; r < P; r++) {
    OrcRank* q = &R[r];
    const uint32_t nnz = q->rowPtr[q->nr];
    /* step 1 (:452-473): externals in first-encounter order; stopRow is inclusive */
    size_t nExtRefs = 0;
    for (uint32_t j = 0; j < nnz; j++) nExtRefs += (q->col[j] < q->startRow || q->col[j] > q->stopRow);
    ColSet seen;
    set_init(&seen, nExtRefs);
    int* extGlobal = (int*)malloc(sizeof(int) * (nExtRefs ? nExtRefs : 1));
    int nExt = 0;
    for (uint32_t j = 0; j < nnz; j++) {
      uint32_t c = q->col[j];
      if (c < q->startRow || c > q->stopRow) {
        uint32_t h = set_slot(&seen, c);
        if (seen.ord[h] < 0) { seen.key[h] = c; seen.ord[h] = nExt; extGlobal[nExt++] = (int)c; }
      }
    }
    /* step 2 (:496-520): owner = largest rank whose startRow <= id */
    int* owner = (int*)malloc(sizeof(int) * (size_t)(nExt ? nExt : 1));
    want[r] = (int*)calloc((size_t)P, sizeof(int));
    for (int i = 0; i < nExt; i++) {
      int o = P - 1;
      while (o > 0 && starts[o] > (uint32_t)extGlobal[i]) o--;
      owner[i] = o;
      want[r][o]++;
    }
    /* step 3 (:40-114): halo slots grouped by owner, owners in order of first appearance,
       first-encounter order inside a group */
    int* localId = (int*)malloc(sizeof(int) * (size_t)(nExt ? nExt : 1));
    for (int i = 0; i < nExt; i++) localId[i] = -1;
    int next = (int)q->nr;
    for (int i = 0; i < nExt; i++) {
      if (localId[i] >= 0) continue;
      localId[i] = next++;
      for (int j = i + 1; j < nExt; j++)
        if (localId[j] < 0 && owner[j] == owner[i]) localId[j] = next++;
    }
    for (uint32_t j = 0; j < nnz; j++) {            /* :96-106 */
      uint32_t c = q->col[j];
      if (c >= q->startRow && c <= q->stopRow) q->col[j] = c - q->startRow;
      else q->col[j] = (uint32_t)localId[seen.ord[set_slot(&seen, c)]];
    }
    q->externalCount = nExt;
    q->externalsReordered = (int*)malloc(sizeof(int) * (size_t)(nExt ? nExt : 1));
    for (int i = 0; i < nExt; i++) q->externalsReordered[localId[i] - (int)q->nr] = extGlobal[i];   /* :108-110 */
    free(localId); free(owner); free(extGlobal);
    set_free(&seen);
  }

  /* topology (:522-580): in-neighbours = owners I need, out-neighbours = ranks that need me,
     both ascending; displacements are running sums in that order (:135,:150) */
  for (int r = 0; r < P; r++) {
    OrcRank* q = &R[r];
    q->indegree = q->outdegree = 0;
    for (int s = 0; s < P; s++) { q->indegree += want[r][s] > 0; q->outdegree += want[s][r] > 0; }
    q->sources = (int*)malloc(sizeof(int) * (size_t)(q->indegree + 1));
    q->recvCounts = (int*)malloc(sizeof(int) * (size_t)(q->indegree + 1));
    q->rdispls = (int*)malloc(sizeof(int) * (size_t)(q->indegree + 1));
    q->destinations = (int*)malloc(sizeof(int) * (size_t)(q->outdegree + 1));
    q->sendCounts = (int*)malloc(sizeof(int) * (size_t)(q->outdegree + 1));
    q->sdispls = (int*)malloc(sizeof(int) * (size_t)(q->outdegree + 1));
    int i = 0, o = 0, racc = 0, sacc = 0;
    for (int s = 0; s < P; s++) {
      if (want[r][s] > 0) { q->sources[i] = s; q->recvCounts[i] = want[r][s]; q->rdispls[i] = racc; racc += want[r][s]; i++; }
      if (want[s][r] > 0) { q->destinations[o] = s; q->sendCounts[o] = want[s][r]; q->sdispls[o] = sacc; sacc += want[s][r]; o++; }
    }
    q->totalSendCount = sacc;
  }
  /* send lists (:116-182): the requester ships the slice of its reordered externals that starts at
     its rdispls for that source; the owner subtracts its startRow */
  for (int r = 0; r < P; r++) {
    OrcRank* q = &R[r];
    q->elementsToSend = (int*)malloc(sizeof(int) * (size_t)(q->totalSendCount ? q->totalSendCount : 1));
    for (int o = 0; o < q->outdegree; o++) {
      const OrcRank* d = &R[q->destinations[o]];
      int slot = 0;
      while (d->sources[slot] != r) slot++;
      for (int t = 0; t < q->sendCounts[o]; t++)
        q->elementsToSend[q->sdispls[o] + t] = d->externalsReordered[d->rdispls[slot] + t] - (int)q->startRow;
    }
  }
  for (int r = 0; r < P; r++) free(want[r]);
  free(want); free(starts);
  return 0;
}

void orc_partition_free(int P, OrcRank* R)
{
  for (int r = 0; r < P; r++) {
    free(R[r].sources); free(R[r].recvCounts); free(R[r].rdispls);
    free(R[r].destinations); free(R[r].sendCounts); free(R[r].sdispls);
    free(R[r].elementsToSend); free(R[r].externalsReordered);
  }
}

void orc_exchange_all(int P, const OrcRank* R, double** x)
{ /* comm.c:635-648: pack by elementsToSend, deliver into x + nr in source order */
  for (int r = 0; r < P; r++) {
    const OrcRank* q = &R[r];
    for (int i = 0; i < q->indegree; i++) {
      const int s = q->sources[i];
      const OrcRank* src = &R[s];
      int o = 0;
      while (src->destinations[o] != r) o++;
      for (int t = 0; t < q->recvCounts[i]; t++)
        x[r][q->nr + (uint32_t)(q->rdispls[i] + t)] = x[s][src->elementsToSend[src->sdispls[o] + t]];
    }
  }
}

static double dot_all(int P, const OrcRank* R, double** a, double** b)
{ /* solver.c:46-61 + comm.c:657-659: local sums, then a SUM over ranks (ascending rank order here) */
  double total = 0.0;
  for (int r = 0; r < P; r++) {
    double s = orc_ddot(R[r].nr, a[r], b[r]);
    total = (r == 0) ? s : total + s;
  }
  return total;
}

int orc_cg_multi(int P, const OrcRank* R, double** vals, double** b, double** x,
    int itermax, double eps, double* hist, int* nhist)
{
  double** r = (double**)malloc(sizeof(double*) * (size_t)P);
  double** p = (double**)malloc(sizeof(double*) * (size_t)P);
  double** Ap = (double**)malloc(sizeof(double*) * (size_t)P);
  for (int q = 0; q < P; q++) {
    r[q] = (double*)calloc(R[q].nr ? R[q].nr : 1, sizeof(double));
    p[q] = (double*)calloc(R[q].nr + (uint32_t)R[q].externalCount + 1, sizeof(double));
    Ap[q] = (double*)calloc(R[q].nr ? R[q].nr : 1, sizeof(double));
  }
#define EACH for (int q = 0; q < P; q++)
  double rtrans, oldrtrans = 0.0, normr;
  EACH orc_waxpby(R[q].nr, 1.0, x[q], 0.0, x[q], p[q]);
  orc_exchange_all(P, R, p);
  EACH orc_spmv_crs(R[q].nr, R[q].rowPtr, R[q].col, vals[q], p[q], Ap[q]);
  EACH orc_waxpby(R[q].nr, 1.0, b[q], -1.0, Ap[q], r[q]);
  rtrans = dot_all(P, R, r, r);
  normr = sqrt(rtrans);
  hist[0] = normr; *nhist = 1;
  int k;
  for (k = 1; k < itermax && normr > eps; k++) {
    if (k == 1) {
      EACH orc_waxpby(R[q].nr, 1.0, r[q], 0.0, r[q], p[q]);
    } else {
      oldrtrans = rtrans;
      rtrans = dot_all(P, R, r, r);
      double beta = rtrans / oldrtrans;
      EACH orc_waxpby(R[q].nr, 1.0, r[q], beta, p[q], p[q]);
    }
    normr = sqrt(rtrans);
    hist[k] = normr; *nhist = k + 1;
    orc_exchange_all(P, R, p);
    EACH orc_spmv_crs(R[q].nr, R[q].rowPtr, R[q].col, vals[q], p[q], Ap[q]);
    double alpha = dot_all(P, R, p, Ap);
    alpha = rtrans / alpha;
    EACH orc_waxpby(R[q].nr, 1.0, x[q], alpha, p[q], x[q]);
    EACH orc_waxpby(R[q].nr, 1.0, r[q], -alpha, Ap[q], r[q]);
  }
#undef EACH
  for (int q = 0; q < P; q++) { free(r[q]); free(p[q]); free(Ap[q]); }
  free(r); free(p); free(Ap);
  return k;
}
