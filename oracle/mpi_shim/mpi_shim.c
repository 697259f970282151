/* TEST INFRASTRUCTURE ONLY -- in-process MPI shim (see mpi.h).
 * Ranks are pthreads; collectives are "publish pointer, barrier, copy, barrier".
 * Reductions combine in ascending rank order so every rank gets the same bits.
 */
#define _GNU_SOURCE
#include "mpi.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SHIM_MAX_RANKS 64

struct shim_comm {
  int indeg, outdeg;
  int* src; int* srcw;
  int* dst; int* dstw;
};

typedef struct shim_msg {
  int src, tag; size_t bytes; void* data; struct shim_msg* next;
} shim_msg;

struct shim_request { void* buf; size_t bytes; int src, tag; };

typedef struct { int src, dst, w; } shim_edge;

static struct {
  int size;
  pthread_barrier_t bar;
  pthread_mutex_t mtx;
  pthread_cond_t cv;
  const void* slot[SHIM_MAX_RANKS];
  const int* slot_cnt[SHIM_MAX_RANKS];
  const int* slot_dsp[SHIM_MAX_RANKS];
  struct shim_comm* slot_comm[SHIM_MAX_RANKS];
  double red[SHIM_MAX_RANKS];
  shim_msg* mbox[SHIM_MAX_RANKS];
  shim_edge* edges; int nedges, capedges;
} W;

static __thread int t_rank = 0;

#define SHIM_MAX_TYPES 64
static size_t g_usertype[SHIM_MAX_TYPES];      /* extent of struct datatypes, MPI_SHIM_USERTYPE + i */
static int g_nusertypes = 0;

static size_t tsize(MPI_Datatype t)
{
  switch (t) {
  case MPI_INT: case MPI_UNSIGNED: case MPI_FLOAT: return 4;
  case MPI_UNSIGNED_LONG_LONG: case MPI_DOUBLE: return 8;
  case MPI_BYTE: case MPI_CHAR: return 1;
  default:
    if (t >= MPI_SHIM_USERTYPE && t < MPI_SHIM_USERTYPE + g_nusertypes) return g_usertype[t - MPI_SHIM_USERTYPE];
    fprintf(stderr, "mpi_shim: unsupported datatype %d\n", t); abort();
  }
}

static void bar(void) { pthread_barrier_wait(&W.bar); }

int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm c, int code) { (void)c; fprintf(stderr, "mpi_shim: MPI_Abort(%d)\n", code); abort(); }
int MPI_Comm_rank(MPI_Comm c, int* rank) { (void)c; *rank = t_rank; return 0; }
int MPI_Comm_size(MPI_Comm c, int* size) { (void)c; *size = W.size; return 0; }
int MPI_Barrier(MPI_Comm c) { (void)c; bar(); return 0; }

int MPI_Allgather(const void* sb, int sc, MPI_Datatype st, void* rb, int rc, MPI_Datatype rt, MPI_Comm c)
{
  (void)c; (void)rc; (void)rt;
  size_t b = (size_t)sc * tsize(st);
  W.slot[t_rank] = sb;
  bar();
  for (int r = 0; r < W.size; r++) memcpy((char*)rb + (size_t)r * b, W.slot[r], b);
  bar();
  return 0;
}

int MPI_Allreduce(const void* sb, void* rb, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
  (void)c;
  if (n != 1 || t != MPI_DOUBLE) { fprintf(stderr, "mpi_shim: Allreduce supports 1 double only\n"); abort(); }
  double mine = (sb == MPI_IN_PLACE) ? *(double*)rb : *(const double*)sb;
  W.red[t_rank] = mine;
  bar();
  double acc = W.red[0];
  for (int r = 1; r < W.size; r++) {
    double v = W.red[r];
    if (op == MPI_SUM) acc += v;
    else if (op == MPI_MAX) acc = (v > acc) ? v : acc;
    else acc = (v < acc) ? v : acc;
  }
  bar();
  *(double*)rb = acc;
  return 0;
}

static int cmp_edge_src(const void* a, const void* b) { return ((const shim_edge*)a)->src - ((const shim_edge*)b)->src; }
static int cmp_edge_dst(const void* a, const void* b) { return ((const shim_edge*)a)->dst - ((const shim_edge*)b)->dst; }

int MPI_Dist_graph_create(MPI_Comm old, int n, const int sources[], const int degrees[],
    const int destinations[], const int weights[], MPI_Info info, int reorder, MPI_Comm* newc)
{
  (void)old; (void)info; (void)reorder;
  if (t_rank == 0) W.nedges = 0;
  bar();
  pthread_mutex_lock(&W.mtx);
  int d = 0;
  for (int i = 0; i < n; i++)
    for (int k = 0; k < degrees[i]; k++, d++) {
      if (W.nedges == W.capedges) {
        W.capedges = W.capedges ? 2 * W.capedges : 256;
        W.edges = (shim_edge*)realloc(W.edges, (size_t)W.capedges * sizeof(shim_edge));
      }
      W.edges[W.nedges++] = (shim_edge) { sources[i], destinations[d], weights[d] };
    }
  pthread_mutex_unlock(&W.mtx);
  bar();
  struct shim_comm* g = (struct shim_comm*)calloc(1, sizeof(*g));
  shim_edge* in  = (shim_edge*)malloc((size_t)(W.nedges + 1) * sizeof(shim_edge));
  shim_edge* out = (shim_edge*)malloc((size_t)(W.nedges + 1) * sizeof(shim_edge));
  for (int e = 0; e < W.nedges; e++) {
    if (W.edges[e].dst == t_rank) in[g->indeg++] = W.edges[e];
    if (W.edges[e].src == t_rank) out[g->outdeg++] = W.edges[e];
  }
  /* neighbours are reported in ascending rank order (what MPICH/OpenMPI do for this call pattern) */
  qsort(in, (size_t)g->indeg, sizeof(shim_edge), cmp_edge_src);
  qsort(out, (size_t)g->outdeg, sizeof(shim_edge), cmp_edge_dst);
  g->src = (int*)malloc(sizeof(int) * (size_t)(g->indeg + 1));
  g->srcw = (int*)malloc(sizeof(int) * (size_t)(g->indeg + 1));
  g->dst = (int*)malloc(sizeof(int) * (size_t)(g->outdeg + 1));
  g->dstw = (int*)malloc(sizeof(int) * (size_t)(g->outdeg + 1));
  for (int i = 0; i < g->indeg; i++) { g->src[i] = in[i].src; g->srcw[i] = in[i].w; }
  for (int i = 0; i < g->outdeg; i++) { g->dst[i] = out[i].dst; g->dstw[i] = out[i].w; }
  free(in); free(out);
  bar();
  *newc = g;
  return 0;
}

int MPI_Dist_graph_neighbors_count(MPI_Comm c, int* indeg, int* outdeg, int* weighted)
{
  *indeg = c->indeg; *outdeg = c->outdeg; *weighted = 1; return 0;
}

int MPI_Dist_graph_neighbors(MPI_Comm c, int maxin, int sources[], int sourceweights[],
    int maxout, int destinations[], int destweights[])
{
  for (int i = 0; i < c->indeg && i < maxin; i++) { sources[i] = c->src[i]; sourceweights[i] = c->srcw[i]; }
  for (int i = 0; i < c->outdeg && i < maxout; i++) { destinations[i] = c->dst[i]; destweights[i] = c->dstw[i]; }
  return 0;
}

int MPI_Irecv(void* buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* r)
{
  (void)c;
  struct shim_request* q = (struct shim_request*)malloc(sizeof(*q));
  q->buf = buf; q->bytes = (size_t)n * tsize(t); q->src = src; q->tag = tag;
  *r = q;
  return 0;
}

int MPI_Send(const void* buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c)
{
  (void)c;
  shim_msg* m = (shim_msg*)malloc(sizeof(*m));
  m->src = t_rank; m->tag = tag; m->bytes = (size_t)n * tsize(t);
  m->data = malloc(m->bytes ? m->bytes : 1);
  memcpy(m->data, buf, m->bytes);
  pthread_mutex_lock(&W.mtx);
  m->next = NULL;
  shim_msg** tail = &W.mbox[dst];
  while (*tail) tail = &(*tail)->next;
  *tail = m;
  pthread_cond_broadcast(&W.cv);
  pthread_mutex_unlock(&W.mtx);
  return 0;
}

int MPI_Waitall(int n, MPI_Request reqs[], MPI_Status st[])
{
  (void)st;
  for (int i = 0; i < n; i++) {
    struct shim_request* q = reqs[i];
    pthread_mutex_lock(&W.mtx);
    for (;;) {
      shim_msg** pp = &W.mbox[t_rank];
      while (*pp && !((*pp)->src == q->src && (*pp)->tag == q->tag)) pp = &(*pp)->next;
      if (*pp) {
        shim_msg* m = *pp; *pp = m->next;
        memcpy(q->buf, m->data, m->bytes < q->bytes ? m->bytes : q->bytes);
        free(m->data); free(m);
        break;
      }
      pthread_cond_wait(&W.cv, &W.mtx);
    }
    pthread_mutex_unlock(&W.mtx);
    free(q);
  }
  return 0;
}

int MPI_Neighbor_alltoallv(const void* sb, const int sc[], const int sd[], MPI_Datatype st,
    void* rb, const int rc[], const int rd[], MPI_Datatype rt, MPI_Comm c)
{
  size_t es = tsize(st); (void)rt;
  W.slot[t_rank] = sb; W.slot_cnt[t_rank] = sc; W.slot_dsp[t_rank] = sd; W.slot_comm[t_rank] = c;
  bar();
  for (int j = 0; j < c->indeg; j++) {
    int s = c->src[j];
    struct shim_comm* sc_ = W.slot_comm[s];
    int k = -1;
    for (int i = 0; i < sc_->outdeg; i++) if (sc_->dst[i] == t_rank) { k = i; break; }
    if (k < 0 || W.slot_cnt[s][k] != rc[j]) { fprintf(stderr, "mpi_shim: neighbour count mismatch\n"); abort(); }
    memcpy((char*)rb + (size_t)rd[j] * es, (const char*)W.slot[s] + (size_t)W.slot_dsp[s][k] * es, (size_t)rc[j] * es);
  }
  bar();
  return 0;
}

#define STUB(name) { fprintf(stderr, "mpi_shim: " name " is a link-only stub\n"); abort(); return 0; }
int MPI_Get_address(const void* p, MPI_Aint* a) { *a = (MPI_Aint)p; return 0; }
MPI_Aint MPI_Aint_diff(MPI_Aint a, MPI_Aint b) { return a - b; }

/* struct datatype = its extent: last member's end, rounded up to the widest member (C layout of the structs the
   reference describes: MMEntry {int,int,double} = 16, FEntry {unsigned,float} = 8) */
int MPI_Type_create_struct(int n, const int bl[], const MPI_Aint d[], const MPI_Datatype t[], MPI_Datatype* nt)
{
  size_t end = 0, align = 1;
  for (int i = 0; i < n; i++) {
    size_t sz = tsize(t[i]);
    if ((size_t)d[i] + sz * (size_t)bl[i] > end) end = (size_t)d[i] + sz * (size_t)bl[i];
    if (sz > align) align = sz;
  }
  end = (end + align - 1) / align * align;
  pthread_mutex_lock(&W.mtx);
  if (g_nusertypes >= SHIM_MAX_TYPES) g_nusertypes = 0;          /* test code: recycle */
  g_usertype[g_nusertypes] = end;
  *nt = MPI_SHIM_USERTYPE + g_nusertypes++;
  pthread_mutex_unlock(&W.mtx);
  return 0;
}
int MPI_Type_commit(MPI_Datatype* t) { (void)t; return 0; }
int MPI_Type_free(MPI_Datatype* t) { (void)t; return 0; }

/* rooted collectives: the root publishes its buffers, everybody copies its share */
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c)
{
  (void)c;
  W.slot[t_rank] = b;
  bar();
  if (t_rank != root) memcpy(b, W.slot[root], (size_t)n * tsize(t));
  bar();
  return 0;
}
int MPI_Scatter(const void* sb, int sc, MPI_Datatype st, void* rb, int rc, MPI_Datatype rt, int root, MPI_Comm c)
{
  (void)c; (void)rc; (void)rt;
  W.slot[t_rank] = sb;
  bar();
  memcpy(rb, (const char*)W.slot[root] + (size_t)t_rank * (size_t)sc * tsize(st), (size_t)sc * tsize(st));
  bar();
  return 0;
}
int MPI_Scatterv(const void* sb, const int sc[], const int sd[], MPI_Datatype st, void* rb, int rc, MPI_Datatype rt, int root, MPI_Comm c)
{
  (void)c; (void)rt;
  W.slot[t_rank] = sb; W.slot_cnt[t_rank] = sc; W.slot_dsp[t_rank] = sd;
  bar();
  if (W.slot_cnt[root][t_rank] != rc) { fprintf(stderr, "mpi_shim: Scatterv count mismatch\n"); abort(); }
  memcpy(rb, (const char*)W.slot[root] + (size_t)W.slot_dsp[root][t_rank] * tsize(st), (size_t)rc * tsize(st));
  bar();
  return 0;
}
int MPI_Gather(const void* sb, int sc, MPI_Datatype st, void* rb, int rc, MPI_Datatype rt, int root, MPI_Comm c) STUB("MPI_Gather")
int MPI_Reduce(const void* sb, void* rb, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) STUB("MPI_Reduce")

/* ---- MPI-IO subset on POSIX files. A view is (byte displacement, element type); the individual file pointer counts
   elements of that type from the displacement (MPI-3.1 section 13.3, 13.4.3), which is all matrixBinfile.c relies on. */
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#define SHIM_MAX_FILES 256
static struct { int fd; long long disp; size_t esize; long long pos; } g_file[SHIM_MAX_FILES];
static int g_nfiles = 0;

int MPI_File_open(MPI_Comm c, const char* filename, int amode, MPI_Info info, MPI_File* fh)
{
  (void)c; (void)info;
  int flags = (amode & MPI_MODE_WRONLY) ? O_WRONLY : O_RDONLY;
  if (amode & MPI_MODE_CREATE) flags |= O_CREAT;
  int fd = open(filename, flags, 0644);
  if (fd < 0) { fprintf(stderr, "mpi_shim: cannot open %s\n", filename); abort(); }
  pthread_mutex_lock(&W.mtx);
  if (g_nfiles >= SHIM_MAX_FILES) g_nfiles = 0;
  int id = g_nfiles++;
  g_file[id].fd = fd; g_file[id].disp = 0; g_file[id].esize = 1; g_file[id].pos = 0;
  pthread_mutex_unlock(&W.mtx);
  *fh = id;
  return 0;
}
int MPI_File_close(MPI_File* fh) { close(g_file[*fh].fd); g_file[*fh].fd = -1; return 0; }
int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype etype, MPI_Datatype filetype, const char* datarep, MPI_Info info)
{
  (void)filetype; (void)datarep; (void)info;
  g_file[fh].disp = disp; g_file[fh].esize = tsize(etype); g_file[fh].pos = 0;
  return 0;
}
int MPI_File_write(MPI_File fh, const void* buf, int count, MPI_Datatype t, MPI_Status* st)
{
  size_t bytes = (size_t)count * tsize(t);
  ssize_t w = pwrite(g_file[fh].fd, buf, bytes, (off_t)(g_file[fh].disp + g_file[fh].pos * (long long)g_file[fh].esize));
  if (w != (ssize_t)bytes) { fprintf(stderr, "mpi_shim: short write\n"); abort(); }
  g_file[fh].pos += (long long)(bytes / g_file[fh].esize);
  if (st) st->shim_bytes = (long long)w;
  return 0;
}
int MPI_File_read(MPI_File fh, void* buf, int count, MPI_Datatype t, MPI_Status* st)
{
  size_t bytes = (size_t)count * tsize(t);
  ssize_t r = pread(g_file[fh].fd, buf, bytes, (off_t)(g_file[fh].disp + g_file[fh].pos * (long long)g_file[fh].esize));
  if (r < 0) r = 0;
  g_file[fh].pos += (long long)((size_t)r / g_file[fh].esize);
  if (st) st->shim_bytes = (long long)r;
  return 0;
}
int MPI_Get_count(const MPI_Status* st, MPI_Datatype t, int* count) { *count = (int)((size_t)st->shim_bytes / tsize(t)); return 0; }
int MPI_File_sync(MPI_File fh) { fsync(g_file[fh].fd); return 0; }
int MPI_File_get_size(MPI_File fh, MPI_Offset* size)
{
  struct stat sb;
  fstat(g_file[fh].fd, &sb);
  *size = (MPI_Offset)sb.st_size;
  return 0;
}
int MPI_File_get_position(MPI_File fh, MPI_Offset* offset) { *offset = g_file[fh].pos; return 0; }
int MPI_File_get_byte_offset(MPI_File fh, MPI_Offset offset, MPI_Offset* disp)
{
  *disp = g_file[fh].disp + offset * (long long)g_file[fh].esize;
  return 0;
}
int MPI_File_seek(MPI_File fh, MPI_Offset offset, int whence)
{
  g_file[fh].pos = (whence == MPI_SEEK_CUR) ? g_file[fh].pos + offset : offset;
  return 0;
}

typedef struct { int rank, size; shim_rank_fn fn; void* arg; } shim_thread_arg;

static void* shim_thread(void* p)
{
  shim_thread_arg* a = (shim_thread_arg*)p;
  t_rank = a->rank;
  a->fn(a->rank, a->size, a->arg);
  return NULL;
}

void shim_run(int nranks, shim_rank_fn fn, void* arg)
{
  if (nranks < 1 || nranks > SHIM_MAX_RANKS) { fprintf(stderr, "mpi_shim: bad rank count\n"); abort(); }
  W.size = nranks;
  pthread_barrier_init(&W.bar, NULL, (unsigned)nranks);
  pthread_mutex_init(&W.mtx, NULL);
  pthread_cond_init(&W.cv, NULL);
  memset(W.mbox, 0, sizeof(W.mbox));
  pthread_t th[SHIM_MAX_RANKS];
  shim_thread_arg ta[SHIM_MAX_RANKS];
  for (int r = 0; r < nranks; r++) {
    ta[r] = (shim_thread_arg) { r, nranks, fn, arg };
    pthread_create(&th[r], NULL, shim_thread, &ta[r]);
  }
  for (int r = 0; r < nranks; r++) pthread_join(th[r], NULL);
  pthread_barrier_destroy(&W.bar);
}
