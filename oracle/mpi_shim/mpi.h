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
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; long long shim_bytes; } MPI_Status;
typedef int MPI_File;
typedef long long MPI_Offset;

#define MPI_COMM_WORLD ((MPI_Comm)0)
#define MPI_INFO_NULL 0
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE ((void*)1)
#define MPI_SUCCESS 0

enum { MPI_INT = 1, MPI_UNSIGNED, MPI_UNSIGNED_LONG_LONG, MPI_FLOAT, MPI_DOUBLE, MPI_BYTE, MPI_CHAR, MPI_SHIM_USERTYPE };
enum { MPI_MODE_RDONLY = 1, MPI_MODE_WRONLY = 2, MPI_MODE_CREATE = 4 };
enum { MPI_SEEK_SET = 0, MPI_SEEK_CUR = 1 };
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

/* MatrixMarket distribution (comm.c:311-402) and .bmx files (matrixBinfile.c:38-236): struct datatypes, rooted
 * collectives, and the MPI-IO subset those two files use, on POSIX files (view = byte displacement + element type) */
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Get_count(const MPI_Status* st, MPI_Datatype t, int* count);
int MPI_File_open(MPI_Comm c, const char* filename, int amode, MPI_Info info, MPI_File* fh);
int MPI_File_close(MPI_File* fh);
int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype etype, MPI_Datatype filetype, const char* datarep, MPI_Info info);
int MPI_File_write(MPI_File fh, const void* buf, int count, MPI_Datatype t, MPI_Status* st);
int MPI_File_read(MPI_File fh, void* buf, int count, MPI_Datatype t, MPI_Status* st);
int MPI_File_sync(MPI_File fh);
int MPI_File_get_size(MPI_File fh, MPI_Offset* size);
int MPI_File_get_position(MPI_File fh, MPI_Offset* offset);
int MPI_File_get_byte_offset(MPI_File fh, MPI_Offset offset, MPI_Offset* disp);
int MPI_File_seek(MPI_File fh, MPI_Offset offset, int whence);
int MPI_Get_address(const void* p, MPI_Aint* a);
MPI_Aint MPI_Aint_diff(MPI_Aint a, MPI_Aint b);
int MPI_Type_create_struct(int n, const int bl[], const MPI_Aint d[], const MPI_Datatype t[], MPI_Datatype* nt);
int MPI_Type_commit(MPI_Datatype* t);
int MPI_Type_free(MPI_Datatype* t);
int MPI_Scatter(const void* sb, int sc, MPI_Datatype st, void* rb, int rc, MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Scatterv(const void* sb, const int sc[], const int sd[], MPI_Datatype st, void* rb, int rc,
    MPI_Datatype rt, int root, MPI_Comm c);
/* link-only stubs (profiler printing; never reached by the oracle) */
int MPI_Gather(const void* sb, int sc, MPI_Datatype st, void* rb, int rc, MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Reduce(const void* sb, void* rb, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);

/* shim control (called by ref_driver.c) */
typedef void (*shim_rank_fn)(int rank, int size, void* arg);
void shim_run(int nranks, shim_rank_fn fn, void* arg);

#endif
