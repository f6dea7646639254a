// Element-wise Fourier-side kernels (D, D^H, ik scaling: reference kernel/matrix_D.c:227-451,
// kernel/ndft-parallel.c:3012-3054) and the ghost-cell fill / reduce of the padded grid
// (PFFT's pfft_exchange / pfft_reduce, reference call sites kernel/ndft-parallel.c:2558,2679).
#pragma once
#include "fftpipe.cuh"
#include "p2p.cuh"

namespace pnb {

// All three kernels walk the local f_hat block in MEMORY order (axes A, B, C; C contiguous): k0, k1, k2 for the plain
// layout, k1, k2, k0 for PNFFT_TRANSPOSED_F_HAT (reference kernel/matrix_D.c:331-341) -- the host passes the tables,
// the first frequencies and the FFT sizes in that order.
struct FhatGeom {
  int l[3];        // local extents, memory order
  int s[3];        // first frequency index of each axis, memory order
  double n[3];     // FFT sizes n_t, memory order (interlacing twiddle)
  int il_sign;     // 0: none; +1 / -1: multiply by exp(+- pi i (kA/nA + kB/nB + kC/nC)) (reference matrix_D.c:282-317)
};

// exp(-sign pi i h), h accumulated outer -> inner exactly like the reference's loops (matrix_D.c:296-312)
template <class R, class C> __device__ __forceinline__ C il_twiddle(const FhatGeom &fg, int iA, int iB, int iC) {
  const R hA = (R)(fg.s[0] + iA) / (R)fg.n[0];
  const R hB = hA + (R)(fg.s[1] + iB) / (R)fg.n[1];
  const R hC = hB + (R)(fg.s[2] + iC) / (R)fg.n[2];
  const R ang = (R)fg.il_sign * (R)3.14159265358979323846264338327950288 * hC;
  C w; w.x = m_cos(ang); w.y = m_sin(ang);
  return w;
}

// g1[k] = f_hat[k] * cA cB cC (* interlacing twiddle)   (overwrite; reference matrix_D.c:227-252, :319-356 / :397-423)
template <class R, class C>
__global__ void k_deconv_fwd(const C *__restrict__ f_hat, C *__restrict__ g1, const R *__restrict__ cA,
                             const R *__restrict__ cB, const R *__restrict__ cC, FhatGeom fg) {
  const int iC = blockIdx.x * blockDim.x + threadIdx.x;
  if (iC >= fg.l[2]) return;
  const R wC = cC[iC];
  for (int iA = blockIdx.z; iA < fg.l[0]; iA += gridDim.z)
    for (int iB = blockIdx.y; iB < fg.l[1]; iB += gridDim.y) {
      const size_t i = ((size_t)iA * fg.l[1] + iB) * fg.l[2] + iC;
      const R w = cA[iA] * cB[iB] * wC;
      C v = f_hat[i];
      v.x *= w; v.y *= w;
      if (fg.il_sign) {
        const C t = il_twiddle<R, C>(fg, iA, iB, iC);
        const R re = v.x * t.x - v.y * t.y, im = v.x * t.y + v.y * t.x;
        v.x = re; v.y = im;
      }
      g1[i] = v;
    }
}

// f_hat[k] += (g1[k] * interlacing twiddle) * cA cB cC   (accumulate; reference matrix_D.c:255-275, :358-395 / :425-451)
template <class R, class C>
__global__ void k_deconv_adj(C *__restrict__ f_hat, const C *__restrict__ g1, const R *__restrict__ cA,
                             const R *__restrict__ cB, const R *__restrict__ cC, FhatGeom fg) {
  const int iC = blockIdx.x * blockDim.x + threadIdx.x;
  if (iC >= fg.l[2]) return;
  const R wC = cC[iC];
  for (int iA = blockIdx.z; iA < fg.l[0]; iA += gridDim.z)
    for (int iB = blockIdx.y; iB < fg.l[1]; iB += gridDim.y) {
      const size_t i = ((size_t)iA * fg.l[1] + iB) * fg.l[2] + iC;
      const R w = cA[iA] * cB[iB] * wC;
      C v = f_hat[i];
      C g = g1[i];
      if (fg.il_sign) {
        const C t = il_twiddle<R, C>(fg, iA, iB, iC);
        const R re = g.x * t.x - g.y * t.y, im = g.x * t.y + g.y * t.x;
        g.x = re; g.y = im;
      }
      v.x += g.x * w; v.y += g.y * w;
      f_hat[i] = v;
    }
}

// mode 0 (trafo): out = -2 pi i k_dim * in          (reference ndft-parallel.c:3034-3054)
// mode 1 (adj)  : out += +2 pi i k_dim * in         (reference ndft-parallel.c:3012-3032)
// mode 2        : out += in
// axis: the memory axis that carries k_dim
template <class R, class C>
__global__ void k_ik_scale(const C *__restrict__ in, C *__restrict__ out, int mode, int axis, FhatGeom fg) {
  const int iC = blockIdx.x * blockDim.x + threadIdx.x;
  if (iC >= fg.l[2]) return;
  const R twopi = (R)(2.0 * 3.14159265358979323846);
  for (int iA = blockIdx.z; iA < fg.l[0]; iA += gridDim.z)
    for (int iB = blockIdx.y; iB < fg.l[1]; iB += gridDim.y) {
      const size_t i = ((size_t)iA * fg.l[1] + iB) * fg.l[2] + iC;
      const int k = axis == 0 ? fg.s[0] + iA : (axis == 1 ? fg.s[1] + iB : fg.s[2] + iC);
      const R w = twopi * (R)k;
      const C v = in[i];
      C o;
      if (mode == 0) { o.x = w * v.y; o.y = -w * v.x; }
      else if (mode == 1) { o = out[i]; o.x += -w * v.y; o.y += w * v.x; }
      else { o = out[i]; o.x += v.x; o.y += v.y; }
      out[i] = o;
    }
}

// Hessian by ik differentiation (reference kernel/ndft-parallel.c:3056-3087): out = -4 pi^2 k_a k_b in, a / b the MEMORY
// axes that carry the two differentiated dimensions
template <class R, class C>
__global__ void k_ik_scale2(const C *__restrict__ in, C *__restrict__ out, int axis_a, int axis_b, FhatGeom fg) {
  const int iC = blockIdx.x * blockDim.x + threadIdx.x;
  if (iC >= fg.l[2]) return;
  const R m4pi2 = (R)(-4.0 * 3.14159265358979323846 * 3.14159265358979323846);
  for (int iA = blockIdx.z; iA < fg.l[0]; iA += gridDim.z)
    for (int iB = blockIdx.y; iB < fg.l[1]; iB += gridDim.y) {
      const size_t i = ((size_t)iA * fg.l[1] + iB) * fg.l[2] + iC;
      const int k[3] = {fg.s[0] + iA, fg.s[1] + iB, fg.s[2] + iC};
      const R w = m4pi2 * (R)k[axis_a] * (R)k[axis_b];
      const C v = in[i];
      C o; o.x = w * v.x; o.y = w * v.y;
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
