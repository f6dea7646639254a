/* TEST INFRASTRUCTURE ONLY -- thread-based MPI emulation for the oracle build of
 * the reference (oracle/_ref/libpnfft_ref.so).  One POSIX thread per "rank",
 * all in one address space; collectives are barrier + direct memory access.
 * Not a general MPI: only what the reference PNFFT path calls.
 */
#define _GNU_SOURCE
#include "mpi.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct shim_world_s {
  int size;
  pthread_barrier_t bar;
  void **slots;
} shim_world;

struct shim_comm_s {
  shim_world *w;
  int rank, size;
  int ndims;            /* 0: no Cartesian topology */
  int dims[3], periods[3], coords[3];
};

static __thread struct shim_comm_s *tls_world_comm = NULL;
static struct shim_comm_s single_world_comm;
static shim_world single_world;
static int single_init = 0;

static void init_single(void)
{
  if (single_init) return;
  single_world.size = 1;
  pthread_barrier_init(&single_world.bar, NULL, 1);
  single_world.slots = (void **)calloc(1, sizeof(void *));
  memset(&single_world_comm, 0, sizeof(single_world_comm));
  single_world_comm.w = &single_world;
  single_world_comm.rank = 0;
  single_world_comm.size = 1;
  single_init = 1;
}

MPI_Comm shim_comm_world(void)
{
  if (tls_world_comm) return tls_world_comm;
  init_single();
  return &single_world_comm;
}

typedef struct {
  shim_world *w;
  int rank;
  void (*fn)(int, void *);
  void *arg;
} thread_arg;

static void *thread_main(void *p)
{
  thread_arg *ta = (thread_arg *)p;
  struct shim_comm_s world;
  memset(&world, 0, sizeof(world));
  world.w = ta->w;
  world.rank = ta->rank;
  world.size = ta->w->size;
  tls_world_comm = &world;
  ta->fn(ta->rank, ta->arg);
  tls_world_comm = NULL;
  return NULL;
}

void shim_mpi_run(int nranks, void (*fn)(int rank, void *arg), void *arg)
{
  shim_world w;
  w.size = nranks;
  pthread_barrier_init(&w.bar, NULL, (unsigned)nranks);
  w.slots = (void **)calloc((size_t)nranks, sizeof(void *));
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nranks);
  thread_arg *ta = (thread_arg *)malloc(sizeof(thread_arg) * (size_t)nranks);
  for (int r = 0; r < nranks; r++) {
    ta[r].w = &w; ta[r].rank = r; ta[r].fn = fn; ta[r].arg = arg;
    pthread_create(&th[r], NULL, thread_main, &ta[r]);
  }
  for (int r = 0; r < nranks; r++) pthread_join(th[r], NULL);
  pthread_barrier_destroy(&w.bar);
  free(w.slots); free(th); free(ta);
}

void **shim_publish(MPI_Comm comm, void *mine)
{
  comm->w->slots[comm->rank] = mine;
  pthread_barrier_wait(&comm->w->bar);
  return comm->w->slots;
}

void shim_unpublish(MPI_Comm comm)
{
  pthread_barrier_wait(&comm->w->bar);
}

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { *rank = comm->rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { *size = comm->size; return MPI_SUCCESS; }

int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm)
{
  struct shim_comm_s *c = (struct shim_comm_s *)malloc(sizeof(*c));
  *c = *comm;
  *newcomm = c;
  return MPI_SUCCESS;
}

int MPI_Comm_free(MPI_Comm *comm)
{
  if (comm && *comm && *comm != &single_world_comm && *comm != tls_world_comm) free(*comm);
  if (comm) *comm = MPI_COMM_NULL;
  return MPI_SUCCESS;
}

static void rank_to_coords(int ndims, const int *dims, int rank, int *coords)
{
  /* row-major: last dimension varies fastest (MPI standard) */
  for (int t = ndims - 1; t >= 0; t--) { coords[t] = rank % dims[t]; rank /= dims[t]; }
}

int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods,
                    int reorder, MPI_Comm *comm_cart)
{
  (void)reorder;
  int prod = 1;
  for (int t = 0; t < ndims; t++) prod *= dims[t];
  if (prod != comm->size || ndims > 3) { *comm_cart = MPI_COMM_NULL; return 1; }
  struct shim_comm_s *c = (struct shim_comm_s *)malloc(sizeof(*c));
  *c = *comm;
  c->ndims = ndims;
  for (int t = 0; t < 3; t++) { c->dims[t] = 1; c->periods[t] = 1; c->coords[t] = 0; }
  for (int t = 0; t < ndims; t++) { c->dims[t] = dims[t]; c->periods[t] = periods ? periods[t] : 1; }
  rank_to_coords(ndims, c->dims, c->rank, c->coords);
  *comm_cart = c;
  return MPI_SUCCESS;
}

int MPI_Cartdim_get(MPI_Comm comm, int *ndims) { *ndims = comm->ndims; return MPI_SUCCESS; }

int MPI_Cart_get(MPI_Comm comm, int maxdims, int *dims, int *periods, int *coords)
{
  for (int t = 0; t < maxdims && t < comm->ndims; t++) {
    dims[t] = comm->dims[t]; periods[t] = comm->periods[t]; coords[t] = comm->coords[t];
  }
  return MPI_SUCCESS;
}

int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int *coords)
{
  int c[3] = {0, 0, 0};
  rank_to_coords(comm->ndims, comm->dims, rank, c);
  for (int t = 0; t < maxdims && t < comm->ndims; t++) coords[t] = c[t];
  return MPI_SUCCESS;
}

int shim_comm_dims(MPI_Comm comm, int *ndims, int dims[3], int coords[3])
{
  *ndims = comm->ndims;
  for (int t = 0; t < 3; t++) { dims[t] = comm->ndims ? comm->dims[t] : 1; coords[t] = comm->ndims ? comm->coords[t] : 0; }
  return 0;
}

int shim_rank_coords(MPI_Comm comm, int rank, int coords[3])
{
  coords[0] = coords[1] = coords[2] = 0;
  if (comm->ndims) rank_to_coords(comm->ndims, comm->dims, rank, coords);
  return 0;
}

int MPI_Barrier(MPI_Comm comm)
{
  pthread_barrier_wait(&comm->w->bar);
  return MPI_SUCCESS;
}

static size_t type_size(MPI_Datatype t)
{
  switch (t) {
    case MPI_CHAR: case MPI_BYTE: return 1;
    case MPI_INT: case MPI_UNSIGNED: return sizeof(int);
    case MPI_LONG: return sizeof(long);
    case MPI_FLOAT: return sizeof(float);
    case MPI_DOUBLE: return sizeof(double);
    case MPI_LONG_DOUBLE: return sizeof(long double);
  }
  return 0;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm)
{
  void **tab = shim_publish(comm, buf);
  if (comm->rank != root) memcpy(buf, tab[root], type_size(type) * (size_t)count);
  shim_unpublish(comm);
  return MPI_SUCCESS;
}

#define REDUCE_LOOP(T)                                                              \
  for (int r = 0; r < comm->size; r++) {                                            \
    const T *src = (const T *)tab[r];                                               \
    T *dst = (T *)acc;                                                              \
    for (int i = 0; i < count; i++) {                                               \
      if (r == 0) dst[i] = src[i];                                                  \
      else if (op == MPI_SUM) dst[i] += src[i];                                     \
      else if (op == MPI_MAX) dst[i] = (src[i] > dst[i]) ? src[i] : dst[i];         \
      else dst[i] = (src[i] < dst[i]) ? src[i] : dst[i];                            \
    }                                                                               \
  }

static void reduce_into(void **tab, void *acc, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm)
{
  switch (type) {
    case MPI_INT: REDUCE_LOOP(int) break;
    case MPI_UNSIGNED: REDUCE_LOOP(unsigned) break;
    case MPI_LONG: REDUCE_LOOP(long) break;
    case MPI_FLOAT: REDUCE_LOOP(float) break;
    case MPI_DOUBLE: REDUCE_LOOP(double) break;
    case MPI_LONG_DOUBLE: REDUCE_LOOP(long double) break;
    default: fprintf(stderr, "shim MPI: unsupported reduce type %d\n", type); abort();
  }
}

int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
               MPI_Op op, int root, MPI_Comm comm)
{
  void **tab = shim_publish(comm, (void *)sendbuf);
  if (comm->rank == root) {
    void *tmp = malloc(type_size(type) * (size_t)count);
    reduce_into(tab, tmp, count, type, op, comm);
    memcpy(recvbuf, tmp, type_size(type) * (size_t)count);
    free(tmp);
  }
  shim_unpublish(comm);
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
                  MPI_Op op, MPI_Comm comm)
{
  void **tab = shim_publish(comm, (void *)sendbuf);
  void *tmp = malloc(type_size(type) * (size_t)count);
  reduce_into(tab, tmp, count, type, op, comm);
  shim_unpublish(comm);
  memcpy(recvbuf, tmp, type_size(type) * (size_t)count);
  free(tmp);
  return MPI_SUCCESS;
}

double MPI_Wtime(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
