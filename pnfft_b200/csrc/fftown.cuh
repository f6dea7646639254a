// Own 1-d FFT kernels of the pencil FFT (the F / F^H matrices, reference kernel/ndft-parallel.c:1524-1546), used for
// power-of-two lengths of complex transforms; other lengths and the real (c2r / r2c) z pass stay with cuFFT.
//
// One CTA transforms TB pencils at once in shared memory: element (position j, pencil tb) lives at buf[j * (TB + 1) + tb], so
// the TB pencils that are neighbours in memory (the x and y passes run along the OUTERMOST axis of their array) are read
// and written as TB * 16-byte segments, and a butterfly pass touches shared memory in whole (TB+1)-pitched rows, which is
// bank-conflict free for 16-byte elements.  A pass is Stockham's auto-sort step of radix 8, 4 or 2 (no bit reversal):
//   v[r] = x[j + r n/R] * w^(r k),  k = j mod Ns;   X = DFT_R(v);   y[(j div Ns) Ns R + k + r Ns] = X[r]
// done in place with the butterflies of a pass held in registers between two CTA barriers.  Twiddles come from a table
// exp(-2 pi i q / n), q < n, built on the host in long double.
// The z pass (contiguous axis) is fused with what surrounds it: forward it writes the kept part of every transformed row
// straight into the padded grid the gridding kernels read (no separate crop / embed pass), backward it reads the rows out
// of the padded grid (zero filling what a truncated torus leaves out).
#pragma once
#include <cmath>
#include <vector>

#include "plan.h"

namespace pnb {

template <class C> __device__ __forceinline__ C cx_mul(C a, C b) { C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <class C> __device__ __forceinline__ C cx_add(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <class C> __device__ __forceinline__ C cx_sub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
// multiplication by -i (forward, DIR = -1) or +i (backward)
template <int DIR, class C> __device__ __forceinline__ C cx_rot(C a) { C r; if (DIR < 0) { r.x = a.y; r.y = -a.x; } else { r.x = -a.y; r.y = a.x; } return r; }

template <int DIR, class C> __device__ __forceinline__ void dft2(C &a, C &b) { const C t = cx_sub(a, b); a = cx_add(a, b); b = t; }
template <int DIR, class C> __device__ __forceinline__ void dft4(C &v0, C &v1, C &v2, C &v3) {
  const C t0 = cx_add(v0, v2), t1 = cx_sub(v0, v2), t2 = cx_add(v1, v3), t3 = cx_rot<DIR>(cx_sub(v1, v3));
  v0 = cx_add(t0, t2); v2 = cx_sub(t0, t2); v1 = cx_add(t1, t3); v3 = cx_sub(t1, t3);
}
template <int DIR, class C> __device__ __forceinline__ void dft8(C (&v)[8]) {
  typedef decltype(v[0].x) R;
  const R h = (R)0.70710678118654752440084436210484903928;
  C a[4], b[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { a[k] = cx_add(v[k], v[k + 4]); b[k] = cx_sub(v[k], v[k + 4]); }
  // b[k] *= w8^k, w8 = exp(DIR * 2 pi i / 8)
  { C t; t.x = h * (b[1].x - (R)DIR * b[1].y); t.y = h * (b[1].y + (R)DIR * b[1].x); b[1] = t; }
  b[2] = cx_rot<DIR>(b[2]);      // w8^2 = DIR * i
  { C t; t.x = h * (-b[3].x - (R)DIR * b[3].y); t.y = h * (-b[3].y + (R)DIR * b[3].x); b[3] = t; }
  dft4<DIR>(a[0], a[1], a[2], a[3]);
  dft4<DIR>(b[0], b[1], b[2], b[3]);
#pragma unroll
  for (int q = 0; q < 4; q++) { v[2 * q] = a[q]; v[2 * q + 1] = b[q]; }
}

// one Stockham pass of radix RAD over the TB pencils in shared memory, in place
template <int RAD, int DIR, int NIT, class C>
__device__ __forceinline__ void fft_pass(C *buf, int n, int tbn, int pitch, int ns, const C *__restrict__ tw) {
  const int items = (n / RAD) * tbn;
  C v[NIT][RAD];
  const int tstep = n / (ns * RAD);        // twiddle table step: w^(r k) = tw[r k tstep]
#pragma unroll
  for (int it = 0; it < NIT; it++) {
    const int i = threadIdx.x + it * blockDim.x;
    if (i < items) {
      const int j = i / tbn, tb = i - j * tbn;
      const int k = j % ns;
#pragma unroll
      for (int r = 0; r < RAD; r++) {
        C x = buf[(j + r * (n / RAD)) * pitch + tb];
        if (r > 0 && ns > 1) {
          C w = tw[r * k * tstep];
          if (DIR > 0) w.y = -w.y;
          x = cx_mul(x, w);
        }
        v[it][r] = x;
      }
      if constexpr (RAD == 8) dft8<DIR>(v[it]);
      else if constexpr (RAD == 4) dft4<DIR>(v[it][0], v[it][1], v[it][2], v[it][3]);
      else dft2<DIR>(v[it][0], v[it][1]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < NIT; it++) {
    const int i = threadIdx.x + it * blockDim.x;
    if (i < items) {
      const int j = i / tbn, tb = i - j * tbn;
      const int k = j % ns, j0 = (j / ns) * ns * RAD + k;
#pragma unroll
      for (int r = 0; r < RAD; r++) buf[(j0 + r * ns) * pitch + tb] = v[it][r];
    }
  }
  __syncthreads();
}

// all passes: n = 8^a * {1, 2, 4}
template <int DIR, int NIT, class C> __device__ __forceinline__ void fft_smem(C *buf, int n, int tbn, int pitch, const C *__restrict__ tw) {
  int ns = 1;
  while ((n / ns) % 8 == 0) { fft_pass<8, DIR, NIT>(buf, n, tbn, pitch, ns, tw); ns *= 8; }
  if (n / ns == 4) fft_pass<4, DIR, 2 * NIT>(buf, n, tbn, pitch, ns, tw);
  else if (n / ns == 2) fft_pass<2, DIR, 4 * NIT>(buf, n, tbn, pitch, ns, tw);
}

constexpr int kFftThreads = 256;
template <class C> struct FftTB { static constexpr int value = sizeof(C) == 16 ? 4 : 8; };    // pencils per CTA at full width

// pencils per CTA for length n: NIT = 2 radix-8 butterflies per thread at most
inline int fft_tb(int n, int full) {
  int tb = full;
  while (tb > 1 && (long long)n * tb > 16LL * kFftThreads) tb /= 2;
  return tb;
}
inline bool fft_own_ok(long long n) { return n >= 8 && n <= 4096 && (n & (n - 1)) == 0; }

// FFT along a strided axis, in place: element (j, b) at data[j * stride + b], b in [0, batch)
template <class C, int DIR, int NIT>
__global__ void __launch_bounds__(kFftThreads, NIT == 1 ? 3 : 2) k_fft_strided(C *__restrict__ data, int n, long long stride, long long batch, int tbn, const C *__restrict__ twg) {
  extern __shared__ __align__(16) unsigned char fft_raw[];
  C *buf = reinterpret_cast<C *>(fft_raw);
  const int pitch = tbn + 1;
  C *tw = buf + (size_t)n * pitch;
  for (int i = threadIdx.x; i < n; i += blockDim.x) tw[i] = twg[i];
  const long long b0 = (long long)blockIdx.x * tbn;
  const int nb = (int)min((long long)tbn, batch - b0);
  C zero; zero.x = 0; zero.y = 0;
  for (int i = threadIdx.x; i < n * tbn; i += blockDim.x) {
    const int j = i / tbn, tb = i - j * tbn;
    buf[j * pitch + tb] = tb < nb ? data[(long long)j * stride + b0 + tb] : zero;
  }
  __syncthreads();
  fft_smem<DIR, NIT>(buf, n, tbn, pitch, tw);
  for (int i = threadIdx.x; i < n * tbn; i += blockDim.x) {
    const int j = i / tbn, tb = i - j * tbn;
    if (tb < nb) data[(long long)j * stride + b0 + tb] = buf[j * pitch + tb];
  }
}

// z pass, forward: rows of n contiguous cells of `in` (row r at in + r * n) are transformed and the cells
// [o_off, o_off + no) of every row land in the padded grid: row r = (i0, i1) -> grid + ((i0 + gcb0) * ngc1 + i1 + gcb1) * pitch2 + gcb2
struct FftGridMap {
  long long rows, rows1;     // number of rows, rows per i0 (= local_no[1])
  long long ngc1, pitch2;
  int gcb0, gcb1, gcb2, o_off, no;
};
template <class C, int DIR, int NIT>
__global__ void __launch_bounds__(kFftThreads, NIT == 1 ? 3 : 2) k_fft_z_grid(C *__restrict__ lin, C *__restrict__ grid, int n, FftGridMap gm, int tbn, const C *__restrict__ twg) {
  extern __shared__ __align__(16) unsigned char fft_raw[];
  C *buf = reinterpret_cast<C *>(fft_raw);
  const int pitch = tbn + 1;
  C *tw = buf + (size_t)n * pitch;
  for (int i = threadIdx.x; i < n; i += blockDim.x) tw[i] = twg[i];
  const long long r0 = (long long)blockIdx.x * tbn;
  const int nb = (int)min((long long)tbn, gm.rows - r0);
  C zero; zero.x = 0; zero.y = 0;
  if (DIR < 0) {
    // coalesced along the row, transposed into the (TB+1)-pitched layout
    for (int i = threadIdx.x; i < n * tbn; i += blockDim.x) {
      const int tb = i / n, j = i - tb * n;
      buf[j * pitch + tb] = tb < nb ? lin[(r0 + tb) * n + j] : zero;
    }
  } else {
    for (int i = threadIdx.x; i < n * tbn; i += blockDim.x) {
      const int tb = i / n, j = i - tb * n;
      C v = zero;
      const int c = j - gm.o_off;
      if (tb < nb && c >= 0 && c < gm.no) {
        const long long r = r0 + tb, i0 = r / gm.rows1, i1 = r - i0 * gm.rows1;
        v = grid[((i0 + gm.gcb0) * gm.ngc1 + i1 + gm.gcb1) * gm.pitch2 + gm.gcb2 + c];
      }
      buf[j * pitch + tb] = v;
    }
  }
  __syncthreads();
  fft_smem<DIR, NIT>(buf, n, tbn, pitch, tw);
  if (DIR < 0) {
    for (int i = threadIdx.x; i < gm.no * tbn; i += blockDim.x) {
      const int tb = i / gm.no, c = i - tb * gm.no;
      if (tb < nb) {
        const long long r = r0 + tb, i0 = r / gm.rows1, i1 = r - i0 * gm.rows1;
        grid[((i0 + gm.gcb0) * gm.ngc1 + i1 + gm.gcb1) * gm.pitch2 + gm.gcb2 + c] = buf[(c + gm.o_off) * pitch + tb];
      }
    }
  } else {
    for (int i = threadIdx.x; i < n * tbn; i += blockDim.x) {
      const int tb = i / n, j = i - tb * n;
      if (tb < nb) lin[(r0 + tb) * n + j] = buf[j * pitch + tb];
    }
  }
}

// host side: twiddle table exp(-2 pi i q / n), q in [0, n)
template <class C> inline void fft_make_twiddles(int n, C **d_tw) {
  std::vector<C> h((size_t)n);
  const long double tau = 6.283185307179586476925286766559005768394L;
  for (int q = 0; q < n; q++) {
    const long double a = -tau * (long double)q / (long double)n;
    h[(size_t)q].x = (decltype(h[0].x))cosl(a);
    h[(size_t)q].y = (decltype(h[0].x))sinl(a);
  }
  PNB_CUDA(cudaMalloc((void **)d_tw, sizeof(C) * (size_t)n));
  PNB_CUDA(cudaMemcpy(*d_tw, h.data(), sizeof(C) * (size_t)n, cudaMemcpyHostToDevice));
}

template <class C> inline void fft_strided_launch(C *data, int n, long long stride, long long batch, int dir, const C *tw, cudaStream_t st) {
  if (batch <= 0) return;
  const int tbn = fft_tb(n, FftTB<C>::value);
  const size_t sm = sizeof(C) * ((size_t)n * (size_t)(tbn + 1) + (size_t)n);
  const unsigned nblk = (unsigned)((batch + tbn - 1) / tbn);
  const bool one = (long long)(n / 8) * tbn <= kFftThreads;      // one radix-8 butterfly per thread and pass
  auto go = [&](auto kern) {
    PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<nblk, kFftThreads, sm, st>>>(data, n, stride, batch, tbn, tw);
  };
  if (dir < 0) { if (one) go(k_fft_strided<C, -1, 1>); else go(k_fft_strided<C, -1, 2>); }
  else { if (one) go(k_fft_strided<C, 1, 1>); else go(k_fft_strided<C, 1, 2>); }
}
template <class C> inline void fft_z_grid_launch(C *lin, C *grid, int n, const FftGridMap &gm, int dir, const C *tw, cudaStream_t st) {
  if (gm.rows <= 0) return;
  const int tbn = fft_tb(n, FftTB<C>::value);
  const size_t sm = sizeof(C) * ((size_t)n * (size_t)(tbn + 1) + (size_t)n);
  const unsigned nblk = (unsigned)((gm.rows + tbn - 1) / tbn);
  const bool one = (long long)(n / 8) * tbn <= kFftThreads;
  auto go = [&](auto kern) {
    PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<nblk, kFftThreads, sm, st>>>(lin, grid, n, gm, tbn, tw);
  };
  if (dir < 0) { if (one) go(k_fft_z_grid<C, -1, 1>); else go(k_fft_z_grid<C, -1, 2>); }
  else { if (one) go(k_fft_z_grid<C, 1, 1>); else go(k_fft_z_grid<C, 1, 2>); }
}

}  // namespace pnb
