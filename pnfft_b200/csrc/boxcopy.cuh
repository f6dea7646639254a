// Strided 3-d box copies: the one data-movement primitive of the FFT pipeline (zero-padding,
// cropping, pencil re-distribution pack/unpack, halo pack/unpack/add).  The innermost box
// dimension is contiguous (stride +1 or -1) on both sides in every layout this library uses, so
// every copy is coalesced.
#pragma once
#include "plan.h"

namespace pnb {

enum BoxOp { BOX_A2C = 0, BOX_C2A = 1, BOX_C2A_ADD = 2 };

template <class T> struct BoxArith;
template <> struct BoxArith<double> {
  static __device__ __forceinline__ double neg(double v) { return -v; }
  static __device__ __forceinline__ double add(double a, double b) { return a + b; }
};
template <> struct BoxArith<float> {
  static __device__ __forceinline__ float neg(float v) { return -v; }
  static __device__ __forceinline__ float add(float a, float b) { return a + b; }
};
template <> struct BoxArith<double2> {
  static __device__ __forceinline__ double2 neg(double2 v) { return make_double2(-v.x, -v.y); }
  static __device__ __forceinline__ double2 add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
};
template <> struct BoxArith<float2> {
  static __device__ __forceinline__ float2 neg(float2 v) { return make_float2(-v.x, -v.y); }
  static __device__ __forceinline__ float2 add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
};

// grid: x = ceil(dims[2]/blockDim.x), y = dims[1] (folded), z = dims[0] (folded)
template <class T, int OP, bool SIGN>
__global__ void k_box_copy(T *__restrict__ A, T *__restrict__ Cb, BoxMap bm) {
  const long long i2 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i2 >= bm.dims[2]) return;
  for (long long i0 = blockIdx.z; i0 < bm.dims[0]; i0 += gridDim.z)
    for (long long i1 = blockIdx.y; i1 < bm.dims[1]; i1 += gridDim.y) {
      const long long ia = bm.a_off + i0 * bm.a_str[0] + i1 * bm.a_str[1] + i2 * bm.a_str[2];
      const long long ic = bm.c_off + i0 * bm.c_str[0] + i1 * bm.c_str[1] + i2 * bm.c_str[2];
      if (OP == BOX_A2C) {
        T v = A[ia];
        if (SIGN && ((i0 + i1 + i2 + bm.parity) & 1)) v = BoxArith<T>::neg(v);
        Cb[ic] = v;
      } else {
        T v = Cb[ic];
        if (SIGN && ((i0 + i1 + i2 + bm.parity) & 1)) v = BoxArith<T>::neg(v);
        if (OP == BOX_C2A_ADD) v = BoxArith<T>::add(A[ia], v);
        A[ia] = v;
      }
    }
}

template <class T>
inline void box_copy(cudaStream_t st, T *A, T *Cb, const BoxMap &bm, int op, bool sign, long long *launches) {
  if (bm.dims[0] <= 0 || bm.dims[1] <= 0 || bm.dims[2] <= 0) return;
  const int bs = bm.dims[2] >= 256 ? 256 : (bm.dims[2] >= 128 ? 128 : (bm.dims[2] >= 64 ? 64 : 32));
  dim3 grid((unsigned)((bm.dims[2] + bs - 1) / bs), (unsigned)(bm.dims[1] < 65535 ? bm.dims[1] : 65535),
            (unsigned)(bm.dims[0] < 65535 ? bm.dims[0] : 65535));
  // keep the grid from exploding for huge boxes: fold dim0
  while ((long long)grid.x * grid.y * grid.z > (1LL << 22) && grid.z > 1) grid.z = (grid.z + 1) / 2;
  if (op == BOX_A2C) {
    if (sign) k_box_copy<T, BOX_A2C, true><<<grid, bs, 0, st>>>(A, Cb, bm);
    else k_box_copy<T, BOX_A2C, false><<<grid, bs, 0, st>>>(A, Cb, bm);
  } else if (op == BOX_C2A) {
    if (sign) k_box_copy<T, BOX_C2A, true><<<grid, bs, 0, st>>>(A, Cb, bm);
    else k_box_copy<T, BOX_C2A, false><<<grid, bs, 0, st>>>(A, Cb, bm);
  } else {
    k_box_copy<T, BOX_C2A_ADD, false><<<grid, bs, 0, st>>>(A, Cb, bm);
  }
  if (launches) (*launches)++;
}

// dense row-major box map helper: A is [*][e1][e2] starting at (s0,s1,s2); chunk is dense [d0][d1][d2] at c_off
inline BoxMap dense_map(long long d0, long long d1, long long d2, long long aE1, long long aE2,
                        long long s0, long long s1, long long s2, long long c_off) {
  BoxMap b;
  b.dims[0] = d0; b.dims[1] = d1; b.dims[2] = d2;
  b.a_str[0] = aE1 * aE2; b.a_str[1] = aE2; b.a_str[2] = 1;
  b.a_off = (s0 * aE1 + s1) * aE2 + s2;
  b.c_str[0] = d1 * d2; b.c_str[1] = d2; b.c_str[2] = 1;
  b.c_off = c_off;
  b.parity = 0;
  return b;
}

}  // namespace pnb
