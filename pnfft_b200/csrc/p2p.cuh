// Peer-memory data plane: the pencil FFT's re-distributions and the ghost-cell exchange as ONE strided-copy kernel per
// stage that writes straight into (or reads straight out of) the neighbours' arrays over NVLink, instead of
// pack -> ncclSend/ncclRecv group -> unpack (reference: PFFT's global transposes and pfft_exchange / pfft_reduce,
// kernel/ndft-parallel.c:1524-1546, :2558, :2679).
//
//  * every rank maps the work buffers, the padded grid and a flag array of every other rank of the plan with CUDA IPC
//    (one process per GPU on one NVSwitch box); handles travel over the mini-MPI control plane at plan time
//  * a stage is a list of BoxJobs (destination base + strides, source base + strides, extents, sign / add): the maps of
//    fftpipe.cuh composed rank-to-rank at plan time (my send map o the peer's receive map), so a cell moves once
//  * ranks order their accesses with a flag barrier in peer memory: one tiny kernel that releases this rank's epoch into
//    every peer's flag array and spins (bounded) until every peer's epoch has arrived
// NCCL stays as the fall-back path (PNFFT_B200_P2P=0, or when IPC mapping is not available).
#pragma once
#include <vector>

#include "boxcopy.cuh"

namespace pnb {

struct BoxJob {
  long long dims[3];
  long long d_off, d_str[3];     // destination element offset / strides
  long long s_off, s_str[3];     // source
  void *dst;
  const void *src;
  int parity;                    // sign variant: factor (-1)^(i0+i1+i2+parity)
  int sign;                      // apply the sign factor
  int add;                       // dst += src instead of dst = src
  int first_block;               // prefix of blocks over the job list
  int bx, by;                    // blocks along the contiguous dim / over the rows
};

constexpr int kMaxJobs = 40;
struct JobList {
  BoxJob j[kMaxJobs];
  int n;
  int nblocks;
};

// one launch for a whole list of strided box copies (innermost box dimension contiguous on both sides)
template <class T>
__global__ void __launch_bounds__(256) k_box_jobs(const JobList *__restrict__ jl) {
  __shared__ int s_job;
  if (threadIdx.x == 0) {
    int k = 0;
    while (k + 1 < jl->n && (int)blockIdx.x >= jl->j[k + 1].first_block) k++;
    s_job = k;
  }
  __syncthreads();
  const BoxJob &J = jl->j[s_job];
  const int b = (int)blockIdx.x - J.first_block;
  const int ix = b % J.bx, iy = b / J.bx;
  const long long i2 = (long long)ix * blockDim.x + threadIdx.x;
  if (i2 < J.dims[2]) {
    T *dst = reinterpret_cast<T *>(J.dst);
    const T *src = reinterpret_cast<const T *>(J.src);
    const long long nrows = J.dims[0] * J.dims[1];
    for (long long r = iy; r < nrows; r += J.by) {
      const long long i0 = r / J.dims[1], i1 = r - i0 * J.dims[1];
      const long long id = J.d_off + i0 * J.d_str[0] + i1 * J.d_str[1] + i2 * J.d_str[2];
      const long long is = J.s_off + i0 * J.s_str[0] + i1 * J.s_str[1] + i2 * J.s_str[2];
      T v = src[is];
      if (J.sign && ((i0 + i1 + i2 + J.parity) & 1)) v = BoxArith<T>::neg(v);
      if (J.add) v = BoxArith<T>::add(dst[id], v);
      dst[id] = v;
    }
  }
  __threadfence_system();      // peer stores of this thread are performed before the grid completes and the epoch is released
}

// finish a host-built job list: block layout (rows folded so that a list stays below ~16 waves of the GPU)
inline void finish_jobs(JobList &L) {
  int total = 0;
  for (int k = 0; k < L.n; k++) {
    BoxJob &J = L.j[k];
    J.bx = (int)((J.dims[2] + 255) / 256);
    const long long nrows = J.dims[0] * J.dims[1];
    long long by = nrows;
    const long long cap = std::max<long long>(1, (148LL * 16) / std::max(1, J.bx));
    if (by > cap) by = cap;
    J.by = (int)by;
    J.first_block = total;
    total += J.bx * J.by;
  }
  L.nblocks = total;
}
inline bool push_job(JobList &L, const BoxMap &bm, void *dst, bool dst_is_a, const void *src, bool sign, bool add) {
  if (bm.dims[0] <= 0 || bm.dims[1] <= 0 || bm.dims[2] <= 0) return true;
  if (L.n >= kMaxJobs) return false;
  BoxJob &J = L.j[L.n++];
  for (int t = 0; t < 3; t++) {
    J.dims[t] = bm.dims[t];
    J.d_str[t] = dst_is_a ? bm.a_str[t] : bm.c_str[t];
    J.s_str[t] = dst_is_a ? bm.c_str[t] : bm.a_str[t];
  }
  J.d_off = dst_is_a ? bm.a_off : bm.c_off;
  J.s_off = dst_is_a ? bm.c_off : bm.a_off;
  J.dst = dst; J.src = src; J.parity = bm.parity; J.sign = sign ? 1 : 0; J.add = add ? 1 : 0;
  J.first_block = 0; J.bx = J.by = 1;
  return true;
}

// release `epoch` into every peer's flag slot of this rank, then wait (bounded) for every peer's epoch
static __global__ void k_peer_barrier(int *const *peer_flags, int *my_flags, int my_rank, int nranks, int epoch, int *error) {
  const int q = threadIdx.x;
  if (q >= nranks) return;
  __threadfence_system();
  volatile int *out = peer_flags[q] + my_rank;
  *out = epoch;
  __threadfence_system();
  volatile int *in = my_flags + q;
  long long spins = 0;
  while (*in - epoch < 0) {
    __nanosleep(64);
    if (++spins > (1LL << 26)) { *error = epoch; break; }     // ~ seconds: a peer died; do not hang the GPU
  }
  __threadfence_system();
}

}  // namespace pnb
