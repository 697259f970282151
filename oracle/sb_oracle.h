/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the SparseBench CG/SpMV hot path.
 *
 * This is the parity oracle: a plain-C, single-threaded restatement of what the reference
 * computes on this path, written from the reference's behaviour (each function cites the
 * reference file:line it follows). It is pinned against
 *   - the reference's 7 golden files and the klein anchor (tests/golden/, tests/test_oracle_*.py),
 *   - the reference's own sources compiled under oracle/_ref/ (same tests, when _ref is present).
 * Two parts have no pin in the reference's own test suite ("parity unpinned" there): the multi-rank halo lists
 * (pinned here by the unmodified comm.c driven through the in-process MPI shim, oracle/mpi_shim) and CG on the SCS
 * format (the reference's initVectors is CRS-only; pinned to the CRS history, which a permutation leaves unchanged
 * up to rounding).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it. The product (sparsebench_b200/) never links, imports or calls anything in oracle/.
 */
#ifndef SB_ORACLE_H
#define SB_ORACLE_H
#include <stdint.h>

typedef struct { uint32_t col; uint32_t pad_; double val; } OrcEntry; /* matrix.h:24-27, 16 bytes */

/* matrix.c:30-121 */
int64_t orc_generate(int nx, int ny, int nz, int rank, int size, int use7pt,
    uint32_t* rowPtr, uint32_t* col, double* val, int64_t cap);
/* CGSolver.c:19-38 */
void orc_init_vectors(uint32_t nr, const uint32_t* rowPtr, int generated, double* x, double* b, double* xexact);
/* matrix-CRS.c:46-65 */
void orc_spmv_crs(uint32_t nr, const uint32_t* rowPtr, const uint32_t* col, const double* val,
    const double* x, double* y);
/* matrix-CCRS.c:14-31 */
void orc_spmv_ccrs(uint32_t nr, const uint32_t* rowPtr, const OrcEntry* e, const double* x, double* y);
/* matrix-SCS.c:31-196 (without the :42-43 overwrite) */
int64_t orc_scs_structure(uint32_t nr, uint32_t C, uint32_t sigma, const uint32_t* rowPtr,
    uint32_t* oldToNew, uint32_t* newToOld, uint32_t* chunkLens, uint32_t* chunkPtr);
void orc_scs_fill(uint32_t nr, uint32_t C, const uint32_t* rowPtr, const uint32_t* col, const double* val,
    const uint32_t* oldToNew, const uint32_t* chunkPtr, int64_t nElems, uint32_t* colOut, double* valOut);
/* matrix-SCS.c:198-228 */
void orc_spmv_scs(uint32_t nChunks, uint32_t C, const uint32_t* chunkPtr, const uint32_t* chunkLens,
    const uint32_t* col, const double* val, const double* x, double* y);
/* solver.c:16-39, :41-62 */
void orc_waxpby(uint32_t n, double alpha, const double* x, double beta, const double* y, double* w);
double orc_ddot(uint32_t n, const double* x, const double* y);
/* summation order of orc_ddot: T >= 1 = the reference's OpenMP static schedule with T threads (1 = sequential, the
 * default); 0 = long double accumulation (yardstick, not a reference mode) */
void orc_set_dot_threads(int t);
int orc_get_dot_threads(void);
/* CGSolver.c:62-141 (single rank, CRS operator) */
int orc_cg_crs(uint32_t nr, uint32_t nc, const uint32_t* rowPtr, const uint32_t* col, const double* val,
    const double* b, double* x, int itermax, double eps, double* hist, int* nhist);

/* comm.c:414-625 (+ :40-114, :116-182), serial over all ranks */
typedef struct {
  /* in */
  uint32_t nr, startRow, stopRow;
  const uint32_t* rowPtr;
  uint32_t* col;              /* global ids in, local+halo ids out */
  /* out (allocated by orc_partition_all, released by orc_partition_free) */
  int externalCount, totalSendCount, indegree, outdegree;
  int *sources, *recvCounts, *rdispls, *destinations, *sendCounts, *sdispls, *elementsToSend;
  int* externalsReordered;    /* global id held by halo slot j */
} OrcRank;
int orc_partition_all(int P, OrcRank* ranks);
void orc_partition_free(int P, OrcRank* ranks);
/* comm.c:627-651, serial over all ranks: x[r] has nr+externalCount entries */
void orc_exchange_all(int P, const OrcRank* ranks, double** x);
/* CGSolver.c:62-141 run in lock-step over P row-block partitions (dots summed in rank order) */
int orc_cg_multi(int P, const OrcRank* ranks, double** vals, double** b, double** x,
    int itermax, double eps, double* hist, int* nhist);

#endif
