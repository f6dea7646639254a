// NCCL data plane bootstrap: one communicator over all ranks of the job, created lazily the first
// time a multi-rank plan needs it.  The unique id travels over the mini-MPI control plane.
#include <cuda_runtime.h>
#include <mpi.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>

namespace pnb {

static ncclComm_t g_comm = nullptr;

void select_device_for_rank() {
  static bool done = false;
  if (done) return;
  done = true;
  int size = 1;
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  if (size <= 1) return;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return;
  const char *lr = getenv("LOCAL_RANK");
  int rank = 0;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  const int dev = (lr && *lr) ? atoi(lr) : rank;
  cudaSetDevice(dev % ndev);
}

ncclComm_t world_nccl() {
  if (g_comm) return g_comm;
  int rank = 0, size = 1;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  select_device_for_rank();
  ncclUniqueId id;
  if (rank == 0) {
    if (ncclGetUniqueId(&id) != ncclSuccess) { fprintf(stderr, "pnfft-b200: ncclGetUniqueId failed\n"); abort(); }
  }
  MPI_Bcast(&id, (int)sizeof(id), MPI_BYTE, 0, MPI_COMM_WORLD);
  ncclResult_t r = ncclCommInitRank(&g_comm, size, id, rank);
  if (r != ncclSuccess) { fprintf(stderr, "pnfft-b200: ncclCommInitRank failed: %s\n", ncclGetErrorString(r)); abort(); }
  return g_comm;
}

void destroy_world_nccl() {
  if (g_comm) { ncclCommDestroy(g_comm); g_comm = nullptr; }
}

}  // namespace pnb
