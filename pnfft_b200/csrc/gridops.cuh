// Element-wise Fourier-side kernels (D, D^H, ik scaling: reference kernel/matrix_D.c:227-451,
// kernel/ndft-parallel.c:3012-3054) and the ghost-cell fill / reduce of the padded grid
// (PFFT's pfft_exchange / pfft_reduce, reference call sites kernel/ndft-parallel.c:2558,2679).
#pragma once
#include "fftpipe.cuh"

namespace pnb {

// g1[k] = f_hat[k] * c0[k0] c1[k1] c2[k2]   (overwrite; reference matrix_D.c:319-356 / :397-423)
template <class R, class C>
__global__ void k_deconv_fwd(const C *__restrict__ f_hat, C *__restrict__ g1, const R *__restrict__ c0,
                             const R *__restrict__ c1, const R *__restrict__ c2, int l0, int l1, int l2) {
  const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i2 >= l2) return;
  const R w2 = c2[i2];
  for (int i0 = blockIdx.z; i0 < l0; i0 += gridDim.z)
    for (int i1 = blockIdx.y; i1 < l1; i1 += gridDim.y) {
      const size_t i = ((size_t)i0 * l1 + i1) * l2 + i2;
      const R w = c0[i0] * c1[i1] * w2;
      C v = f_hat[i];
      v.x *= w; v.y *= w;
      g1[i] = v;
    }
}

// f_hat[k] += g1[k] * c0 c1 c2   (accumulate; reference matrix_D.c:358-395 / :425-451)
template <class R, class C>
__global__ void k_deconv_adj(C *__restrict__ f_hat, const C *__restrict__ g1, const R *__restrict__ c0,
                             const R *__restrict__ c1, const R *__restrict__ c2, int l0, int l1, int l2) {
  const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i2 >= l2) return;
  const R w2 = c2[i2];
  for (int i0 = blockIdx.z; i0 < l0; i0 += gridDim.z)
    for (int i1 = blockIdx.y; i1 < l1; i1 += gridDim.y) {
      const size_t i = ((size_t)i0 * l1 + i1) * l2 + i2;
      const R w = c0[i0] * c1[i1] * w2;
      C v = f_hat[i];
      const C g = g1[i];
      v.x += g.x * w; v.y += g.y * w;
      f_hat[i] = v;
    }
}

// mode 0 (trafo): out = -2 pi i k_dim * in          (reference ndft-parallel.c:3034-3054)
// mode 1 (adj)  : out += +2 pi i k_dim * in         (reference ndft-parallel.c:3012-3032)
// mode 2        : out += in
template <class R, class C>
__global__ void k_ik_scale(const C *__restrict__ in, C *__restrict__ out, int mode, int dim, int s0, int s1, int s2,
                           int l0, int l1, int l2) {
  const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i2 >= l2) return;
  const R twopi = (R)(2.0 * 3.14159265358979323846);
  for (int i0 = blockIdx.z; i0 < l0; i0 += gridDim.z)
    for (int i1 = blockIdx.y; i1 < l1; i1 += gridDim.y) {
      const size_t i = ((size_t)i0 * l1 + i1) * l2 + i2;
      const int k = dim == 0 ? s0 + i0 : (dim == 1 ? s1 + i1 : s2 + i2);
      const R w = twopi * (R)k;
      const C v = in[i];
      C o;
      if (mode == 0) { o.x = w * v.y; o.y = -w * v.x; }
      else if (mode == 1) { o = out[i]; o.x += -w * v.y; o.y += w * v.x; }
      else { o = out[i]; o.x += v.x; o.y += v.y; }
      out[i] = o;
    }
}

inline dim3 grid3(int l0, int l1, int l2, int bs) {
  return dim3((unsigned)((l2 + bs - 1) / bs), (unsigned)(l1 < 65535 ? l1 : 65535), (unsigned)(l0 < 1024 ? l0 : 1024));
}

// ---------------------------------------------------------------------------------------------
// periodic ghost cells along an axis the process mesh does not split
// ---------------------------------------------------------------------------------------------
struct HaloGeom {
  long long pitch0, pitch1;      // element strides of dims 0, 1 (dim 2 is contiguous)
  int lo[3], hi[3];              // index ranges covered in the other two dims (the axis' own entries are ignored)
  int axis, gcb, gca, lno;       // halo widths and interior extent along the axis
};

__device__ __forceinline__ int halo_slot_index(int h, int gcb, int lno) { return h < gcb ? h : lno + h; }
__device__ __forceinline__ int halo_src_index(int i, int gcb, int lno) {
  int r = (i - gcb) % lno;
  if (r < 0) r += lno;
  return gcb + r;
}

// fill: halo <- wrapped interior.  One thread per halo cell.
template <class T> __global__ void k_halo_fill(T *__restrict__ grid, HaloGeom hg) {
  int ext[3];
  for (int t = 0; t < 3; t++) ext[t] = hg.hi[t] - hg.lo[t];
  ext[hg.axis] = hg.gcb + hg.gca;
  const long long total = (long long)ext[0] * ext[1] * ext[2];
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
    int r[3];
    r[2] = (int)(id % ext[2]);
    r[1] = (int)((id / ext[2]) % ext[1]);
    r[0] = (int)(id / ((long long)ext[2] * ext[1]));
    int dst[3], src[3];
    for (int t = 0; t < 3; t++) dst[t] = src[t] = hg.lo[t] + r[t];
    dst[hg.axis] = halo_slot_index(r[hg.axis], hg.gcb, hg.lno);
    src[hg.axis] = halo_src_index(dst[hg.axis], hg.gcb, hg.lno);
    grid[dst[0] * hg.pitch0 + dst[1] * hg.pitch1 + dst[2]] = grid[src[0] * hg.pitch0 + src[1] * hg.pitch1 + src[2]];
  }
}

// reduce: wrapped interior += halo.  One thread per line along the axis, halo slots visited in order
// (several slots may fold onto the same interior cell when the block is narrower than the halo).
template <class T> __global__ void k_halo_reduce(T *__restrict__ grid, HaloGeom hg) {
  int ext[3];
  for (int t = 0; t < 3; t++) ext[t] = hg.hi[t] - hg.lo[t];
  ext[hg.axis] = 1;
  const long long total = (long long)ext[0] * ext[1] * ext[2];
  for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
    int r[3];
    r[2] = (int)(id % ext[2]);
    r[1] = (int)((id / ext[2]) % ext[1]);
    r[0] = (int)(id / ((long long)ext[2] * ext[1]));
    int dst[3], src[3];
    for (int t = 0; t < 3; t++) dst[t] = src[t] = hg.lo[t] + r[t];
    for (int h = 0; h < hg.gcb + hg.gca; h++) {
      src[hg.axis] = halo_slot_index(h, hg.gcb, hg.lno);
      dst[hg.axis] = halo_src_index(src[hg.axis], hg.gcb, hg.lno);
      T *d = &grid[dst[0] * hg.pitch0 + dst[1] * hg.pitch1 + dst[2]];
      *d = BoxArith<T>::add(*d, grid[src[0] * hg.pitch0 + src[1] * hg.pitch1 + src[2]]);
    }
  }
}

}  // namespace pnb
