/* pnfft-b200 mini-MPI: the control-plane subset of MPI that PNFFT and its callers use.
 *
 * PNFFT's public header includes <mpi.h> (reference api/pnfft.h:28) and every entry point
 * takes an MPI_Comm.  The target image has no MPI, so the library ships this stand-in:
 * one process per GPU, started by any launcher that exports RANK / WORLD_SIZE /
 * LOCAL_RANK / MASTER_ADDR / MASTER_PORT (torchrun does), rendezvous over TCP on
 * MASTER_PORT+PNFFT_B200_PORT_OFFSET, host-side collectives through rank 0, and the
 * data plane (halo exchange, FFT transposes) on NCCL over NVLink.  A build against a
 * real MPI only has to replace this header and pnfft_b200/csrc/minimpi.cpp.
 *
 * Handles are small integers (MPICH style) so they pass through ctypes/FFI unchanged.
 */
#ifndef PNFFT_B200_MINI_MPI_H
#define PNFFT_B200_MINI_MPI_H 1

#include <stddef.h>

/* The stand-in never exports the names of a real MPI library: every entry point is compiled and called as pnb_MPI_*, so
 * that libpnfft_b200.so can live in a process that also loaded libmpi (no interposition, no clash).  Callers built
 * against THIS header bind to the prefixed symbols automatically.  A build of the library against a real MPI
 * (make -C pnfft_b200/csrc REAL_MPI=1 with the MPI compiler wrappers' include path) does not use this header at all. */
#define MPI_Init pnb_MPI_Init
#define MPI_Initialized pnb_MPI_Initialized
#define MPI_Finalize pnb_MPI_Finalize
#define MPI_Abort pnb_MPI_Abort
#define MPI_Comm_rank pnb_MPI_Comm_rank
#define MPI_Comm_size pnb_MPI_Comm_size
#define MPI_Comm_dup pnb_MPI_Comm_dup
#define MPI_Comm_free pnb_MPI_Comm_free
#define MPI_Cart_create pnb_MPI_Cart_create
#define MPI_Cartdim_get pnb_MPI_Cartdim_get
#define MPI_Cart_get pnb_MPI_Cart_get
#define MPI_Cart_coords pnb_MPI_Cart_coords
#define MPI_Cart_rank pnb_MPI_Cart_rank
#define MPI_Barrier pnb_MPI_Barrier
#define MPI_Bcast pnb_MPI_Bcast
#define MPI_Reduce pnb_MPI_Reduce
#define MPI_Allreduce pnb_MPI_Allreduce
#define MPI_Wtime pnb_MPI_Wtime

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Fint;

#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15

#define MPI_COMM_NULL  0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF  2

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

int MPI_Init(int *argc, char ***argv);
int MPI_Initialized(int *flag);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *newcomm);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods,
                    int reorder, MPI_Comm *comm_cart);
int MPI_Cartdim_get(MPI_Comm comm, int *ndims);
int MPI_Cart_get(MPI_Comm comm, int maxdims, int *dims, int *periods, int *coords);
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int *coords);
int MPI_Cart_rank(MPI_Comm comm, const int *coords, int *rank);
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
               MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type,
                  MPI_Op op, MPI_Comm comm);
double MPI_Wtime(void);

#ifdef __cplusplus
}
#endif
#endif
