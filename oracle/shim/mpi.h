/* TEST INFRASTRUCTURE ONLY -- part of the oracle, never linked into the product.
 *
 * Minimal stand-in for <mpi.h> so that the unmodified reference sources under
 * /root/reference compile in a container without MPI.  "Ranks" are POSIX
 * threads of one process (see shim_mpi.c); only the calls the reference's
 * window-convolution path and its test drivers use are provided
 * (SURVEY.md section 8c lists them).
 */
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H 1

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

struct shim_comm_s;
typedef struct shim_comm_s *MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Fint;

#define MPI_SUCCESS 0
#define MPI_COMM_NULL ((MPI_Comm)0)

#define MPI_CHAR        1
#define MPI_INT         2
#define MPI_UNSIGNED    3
#define MPI_LONG        4
#define MPI_FLOAT       5
#define MPI_DOUBLE      6
#define MPI_LONG_DOUBLE 7
#define MPI_BYTE        8

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

MPI_Comm shim_comm_world(void);
#define MPI_COMM_WORLD (shim_comm_world())

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods,
                    int reorder, MPI_Comm *comm_cart);
int MPI_Cartdim_get(MPI_Comm comm, int *ndims);
int MPI_Cart_get(MPI_Comm comm, int maxdims, int *dims, int *periods, int *coords);
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int *coords);
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
               MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
                  MPI_Op op, MPI_Comm comm);
double MPI_Wtime(void);

/* ---- shim-only entry points (used by oracle/ref_driver.c) ---- */
/* Run fn(rank, arg) on nranks threads, each seeing its own MPI_COMM_WORLD. */
void shim_mpi_run(int nranks, void (*fn)(int rank, void *arg), void *arg);
/* Collective helper: every rank publishes a pointer, all ranks see all pointers
 * between the two internal barriers.  Returns the table (valid until shim_unpublish). */
void **shim_publish(MPI_Comm comm, void *mine);
void shim_unpublish(MPI_Comm comm);
/* per-world scratch pointer shared by all ranks (rank 0 allocates inside a publish section) */
int shim_comm_dims(MPI_Comm comm, int *ndims, int dims[3], int coords[3]);
int shim_rank_coords(MPI_Comm comm, int rank, int coords[3]);

#ifdef __cplusplus
}
#endif
#endif
