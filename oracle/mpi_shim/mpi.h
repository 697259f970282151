/* TEST INFRASTRUCTURE ONLY -- in-process MPI shim for the parity oracle.
 *
 * There is no MPI in this image. To run the UNMODIFIED reference comm.c
 * (commPartition / commExchange / commReduction, /root/reference/src/comm.c:414-662)
 * as the multi-rank oracle, "ranks" are pthreads of one process and this header
 * + mpi_shim.c provide exactly the MPI-3 subset those functions call.
 * It is never linked into the product library.
 */
#ifndef SB_ORACLE_MPI_SHIM_H
#define SB_ORACLE_MPI_SHIM_H
#include <stddef.h>

typedef struct shim_comm* MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef long MPI_Aint;
typedef int MPI_Info;
typedef struct shim_request* MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef int MPI_File;

#define MPI_COMM_WORLD ((MPI_Comm)0)
#define MPI_INFO_NULL 0
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE ((void*)1)
#define MPI_SUCCESS 0

enum { MPI_INT = 1, MPI_UNSIGNED, MPI_UNSIGNED_LONG_LONG, MPI_FLOAT, MPI_DOUBLE, MPI_BYTE, MPI_SHIM_USERTYPE };
enum { MPI_SUM = 1, MPI_MAX, MPI_MIN };

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm c, int code);
int MPI_Comm_rank(MPI_Comm c, int* rank);
int MPI_Comm_size(MPI_Comm c, int* size);
int MPI_Barrier(MPI_Comm c);
int MPI_Allgather(const void* sb, int sc, MPI_Datatype st, void* rb, int rc, MPI_Datatype rt, MPI_Comm c);
int MPI_Allreduce(const void* sb, void* rb, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Dist_graph_create(MPI_Comm old, int n, const int sources[], const int degrees[],
    const int destinations[], const int weights[], MPI_Info info, int reorder, MPI_Comm* newc);
int MPI_Dist_graph_neighbors_count(MPI_Comm c, int* indeg, int* outdeg, int* weighted);
int MPI_Dist_graph_neighbors(MPI_Comm c, int maxin, int sources[], int sourceweights[],
    int maxout, int destinations[], int destweights[]);
int MPI_Irecv(void* buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* r);
int MPI_Send(const void* buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c);
int MPI_Waitall(int n, MPI_Request reqs[], MPI_Status st[]);
int MPI_Neighbor_alltoallv(const void* sb, const int sc[], const int sd[], MPI_Datatype st,
    void* rb, const int rc[], const int rd[], MPI_Datatype rt, MPI_Comm c);

/* link-only stubs (MatrixMarket distribution + profiler printing; never reached by the oracle) */
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Get_address(const void* p, MPI_Aint* a);
MPI_Aint MPI_Aint_diff(MPI_Aint a, MPI_Aint b);
int MPI_Type_create_struct(int n, const int bl[], const MPI_Aint d[], const MPI_Datatype t[], MPI_Datatype* nt);
int MPI_Type_commit(MPI_Datatype* t);
int MPI_Type_free(MPI_Datatype* t);
int MPI_Scatter(const void* sb, int sc, MPI_Datatype st, void* rb, int rc, MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Scatterv(const void* sb, const int sc[], const int sd[], MPI_Datatype st, void* rb, int rc,
    MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Gather(const void* sb, int sc, MPI_Datatype st, void* rb, int rc, MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Reduce(const void* sb, void* rb, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);

/* shim control (called by ref_driver.c) */
typedef void (*shim_rank_fn)(int rank, int size, void* arg);
void shim_run(int nranks, shim_rank_fn fn, void* arg);

#endif
