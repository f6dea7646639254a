// NCCL data plane bootstrap: one communicator over all ranks of the job, created lazily the first
// time a multi-rank plan needs it.  The unique id travels over the mini-MPI control plane.
// NCCL itself is bound with dlopen (see NcclApi in fftpipe.cuh).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <mpi.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>

#include "fftpipe.cuh"

namespace pnb {

static ncclComm_t g_comm = nullptr;

const NcclApi &nccl_api() {
  static NcclApi api;
  static bool ready = false;
  if (ready) return api;
  void *h = nullptr;
  const char *names[] = {getenv("PNFFT_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; i < 3 && !h; i++)
    if (names[i] && *names[i]) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "pnfft-b200: cannot load NCCL (%s); set PNFFT_B200_NCCL_LIB\n", dlerror()); abort(); }
  auto sym = [&](const char *n) {
    void *p = dlsym(h, n);
    if (!p) { fprintf(stderr, "pnfft-b200: NCCL symbol %s not found\n", n); abort(); }
    return p;
  };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.Send = (decltype(api.Send))sym("ncclSend");
  api.Recv = (decltype(api.Recv))sym("ncclRecv");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  ready = true;
  return api;
}

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
    if (nccl_api().GetUniqueId(&id) != ncclSuccess) { fprintf(stderr, "pnfft-b200: ncclGetUniqueId failed\n"); abort(); }
  }
  MPI_Bcast(&id, (int)sizeof(id), MPI_BYTE, 0, MPI_COMM_WORLD);
  ncclResult_t r = nccl_api().CommInitRank(&g_comm, size, id, rank);
  if (r != ncclSuccess) { fprintf(stderr, "pnfft-b200: ncclCommInitRank failed: %s\n", nccl_api().GetErrorString(r)); abort(); }
  return g_comm;
}

void destroy_world_nccl() {
  if (g_comm) { nccl_api().CommDestroy(g_comm); g_comm = nullptr; }
}

}  // namespace pnb
