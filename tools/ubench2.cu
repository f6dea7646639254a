// Microbenchmarks (development aid) behind the round-2 restructuring of the z-march kernels.
//  1. dependent-issue latency of DFMA (one warp, one chain) and throughput against chains / warps
//  2. cost of shared-memory loads by address pattern: warp-uniform, half-warp-uniform, quarter, distinct; 64 / 128 bit
//  3. node-loop models: v2 (one row per thread, 16 taps, warp-uniform weight quads) against v3 (two rows per thread,
//     8 taps, the two lane halves read different weight quads), F and F+grad, 4 / 8 / 12 warps per SM
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/ubench2.cu -o tools/ubench2.bin
#include <cuda_runtime.h>
#include <cstdio>

static double g_clk = 1.965e9;

template <int CH> __global__ void k_dfma(double *out, int iters) {
  double a[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) a[i] = threadIdx.x + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = fma(a[i], b, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float best_ms(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
  return best;
}

// ---- 2. LDS patterns ----
// PAT: 0 uniform, 1 two addresses (lane halves), 2 four addresses (quarters), 3 distinct per lane (conflict free)
// the loaded words are consumed by ONE 32-bit xor per load (alu pipe), so the load rate is what is measured
template <int BITS, int PAT> __global__ void k_lds(double *out, int iters) {
  extern __shared__ __align__(16) unsigned char sm[];
  for (int i = threadIdx.x; i < 16384 / 8; i += blockDim.x) reinterpret_cast<double *>(sm)[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = PAT == 0 ? 0 : (PAT == 1 ? lane >> 4 : (PAT == 2 ? lane >> 3 : lane));
  unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)(grp * (BITS / 8)) + (unsigned)(warp & 3) * 2048u;
  unsigned acc = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const unsigned a = base + (unsigned)(u * 64) + (unsigned)((it & 1) * 1024);
      unsigned x, y, z, w;
      if (BITS == 128) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a));
      else if (BITS == 64) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a));
      else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(a));
      acc ^= x;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// SHFL rate: one xor-shuffle per iteration step, consumed by one xor
__global__ void k_shfl(double *out, int iters) {
  unsigned v = threadIdx.x * 2654435761u, acc = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) { v = __shfl_xor_sync(0xffffffffu, v + u, 16); acc ^= v; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---- 3. node loop models ----
// Row table in shared memory: 16 rows of 64 doubles, filled from global memory at run time (opaque to the compiler):
//   [0..16) psi_z, [16..32) dpsi_z, [32..40) psi_x, [40..60) psi_y, [60..62) f, [62..64) header ints
constexpr int RL = 64, NR = 16;
__device__ __forceinline__ const double *model_rows(double *rows, const double *src, int nthreads) {
  for (int i = threadIdx.x; i < NR * RL; i += nthreads) rows[i] = src[i % 64] * (1.0 + 1e-3 * (i / 64));
  __syncthreads();
  return rows;
}
// v2: one row per thread, W = 16 window cells, warp-uniform weight quads; GRAD: t, td (4 chains); !GRAD: 2+2 chains
template <bool GRAD> __global__ void __launch_bounds__(384, 1) k_g2(double *out, const double *src, int iters) {
  __shared__ __align__(16) double rows[NR * RL];
  __shared__ __align__(16) double2 part[12][2][32];
  model_rows(rows, src, blockDim.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double2 win[16];
#pragma unroll
  for (int i = 0; i < 16; i++) win[i] = make_double2(out[threadIdx.x * 16 + i], out[8192 + threadIdx.x * 16 + i]);
  for (int it = 0; it < iters; it++) {
    const double *row = rows + ((it * 7 + warp) & (NR - 1)) * RL;
    const int4 hd = *reinterpret_cast<const int4 *>(row + 62);
    double2 t = make_double2(0, 0), td = t, t1 = t, td1 = t;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      const double2 w = *reinterpret_cast<const double2 *>(row + k);
      double2 dw = w;
      if (GRAD) dw = *reinterpret_cast<const double2 *>(row + 16 + k);
      if (GRAD) {
        t.x = fma(w.x, win[k].x, t.x); t.y = fma(w.x, win[k].y, t.y); td.x = fma(dw.x, win[k].x, td.x); td.y = fma(dw.x, win[k].y, td.y);
        t.x = fma(w.y, win[k + 1].x, t.x); t.y = fma(w.y, win[k + 1].y, t.y); td.x = fma(dw.y, win[k + 1].x, td.x); td.y = fma(dw.y, win[k + 1].y, td.y);
      } else {
        t.x = fma(w.x, win[k].x, t.x); t.y = fma(w.x, win[k].y, t.y);
        t1.x = fma(w.y, win[k + 1].x, t1.x); t1.y = fma(w.y, win[k + 1].y, t1.y);
      }
    }
    if (!GRAD) { t.x += t1.x; t.y += t1.y; }
    part[warp][0][(lane + hd.x) & 31] = t;
    if (GRAD) part[warp][1][(lane + hd.y) & 31] = td;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = part[warp][0][lane].x + part[warp][1][lane].y;
}
// v4: two x rows per thread, W = 13 cells each, the two lane halves work on two different nodes (two-address quads);
// the window advances by one cell every ADV passes (register shift + one new cell per row from a staging box)
template <bool GRAD, int ADV> __global__ void __launch_bounds__(320, 1) k_g4(double *out, const double *src, int iters) {
  __shared__ __align__(16) double rows[NR * RL];
  __shared__ __align__(16) double2 part[12][4][32];
  __shared__ __align__(16) double2 stage[12][2][32];
  model_rows(rows, src, blockDim.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hlf = lane >> 4;
  for (int i = lane; i < 64; i += 32) stage[warp][i / 32][i % 32] = make_double2(src[i], src[63 - i]);
  double2 win[2][13];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int i = 0; i < 13; i++) win[r][i] = make_double2(out[threadIdx.x * 32 + r * 16 + i], out[16384 + threadIdx.x * 32 + r * 16 + i]);
  for (int it = 0; it < iters; it++) {
    const double *row = rows + ((it * 14 + warp * 2 + hlf) & (NR - 1)) * RL;      // the halves read different nodes' rows
    const int4 hd = *reinterpret_cast<const int4 *>(row + 62);
    double2 t[2], td[2];
#pragma unroll
    for (int r = 0; r < 2; r++) { t[r] = make_double2(0, 0); td[r] = t[r]; }
#pragma unroll
    for (int k = 0; k < 13; k += 2) {
      const double2 w = *reinterpret_cast<const double2 *>(row + k);
      double2 dw = w;
      if (GRAD) dw = *reinterpret_cast<const double2 *>(row + 16 + k);
#pragma unroll
      for (int r = 0; r < 2; r++) {
        t[r].x = fma(w.x, win[r][k].x, t[r].x); t[r].y = fma(w.x, win[r][k].y, t[r].y);
        if (GRAD) { td[r].x = fma(dw.x, win[r][k].x, td[r].x); td[r].y = fma(dw.x, win[r][k].y, td[r].y); }
        if (k + 1 < 13) {
          t[r].x = fma(w.y, win[r][k + 1].x, t[r].x); t[r].y = fma(w.y, win[r][k + 1].y, t[r].y);
          if (GRAD) { td[r].x = fma(dw.y, win[r][k + 1].x, td[r].x); td[r].y = fma(dw.y, win[r][k + 1].y, td[r].y); }
        }
      }
    }
    part[warp][0][(lane + hd.x) & 31] = t[0]; part[warp][1][(lane + hd.x) & 31] = t[1];
    if (GRAD) { part[warp][2][(lane + hd.y) & 31] = td[0]; part[warp][3][(lane + hd.y) & 31] = td[1]; }
    if (it % ADV == ADV - 1) {
#pragma unroll
      for (int r = 0; r < 2; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) win[r][i] = win[r][i + 1];
        win[r][12] = stage[warp][r][(lane + it) & 31];
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = part[warp][0][lane].x + part[warp][3][lane].y + win[0][3].x + win[1][7].y;
}
// variants of the v2 gather model isolating what throttles the FP64 pipe.  MODE 0: as k_g2<true>; 1: weights in registers
// (no LDS in the loop); 2: LDS.64 instead of LDS.128; 3: no STS; 4: 8 chains (two per sum); 5: LDS but one row for all iterations
template <int MODE, int NWARP> __global__ void __launch_bounds__(NWARP * 32, 1) k_gx(double *out, const double *src, int iters) {
  __shared__ __align__(16) double rows[NR * RL];
  __shared__ __align__(16) double2 part[16][2][32];
  model_rows(rows, src, blockDim.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double2 win[16];
#pragma unroll
  for (int i = 0; i < 16; i++) win[i] = make_double2(out[threadIdx.x * 16 + i], out[8192 + threadIdx.x * 16 + i]);
  double wr[32];
  if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < 32; i++) wr[i] = rows[i + (warp & 1)];
  }
  for (int it = 0; it < iters; it++) {
    const double *row = rows + (MODE == 5 ? 0 : ((it * 7 + warp) & (NR - 1)) * RL);
    const int4 hd = *reinterpret_cast<const int4 *>(row + 62);
    double2 t = make_double2(0, 0), td = t, t1 = t, td1 = t;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      double2 w, dw;
      if (MODE == 1) { w = make_double2(wr[k], wr[k + 1]); dw = make_double2(wr[16 + k], wr[17 + k]); }
      else if (MODE == 2) { w.x = row[k]; w.y = row[k + 1]; dw.x = row[16 + k]; dw.y = row[17 + k]; }
      else { w = *reinterpret_cast<const double2 *>(row + k); dw = *reinterpret_cast<const double2 *>(row + 16 + k); }
      t.x = fma(w.x, win[k].x, t.x); t.y = fma(w.x, win[k].y, t.y); td.x = fma(dw.x, win[k].x, td.x); td.y = fma(dw.x, win[k].y, td.y);
      if (MODE == 4) { t1.x = fma(w.y, win[k + 1].x, t1.x); t1.y = fma(w.y, win[k + 1].y, t1.y); td1.x = fma(dw.y, win[k + 1].x, td1.x); td1.y = fma(dw.y, win[k + 1].y, td1.y); }
      else { t.x = fma(w.y, win[k + 1].x, t.x); t.y = fma(w.y, win[k + 1].y, t.y); td.x = fma(dw.y, win[k + 1].x, td.x); td.y = fma(dw.y, win[k + 1].y, td.y); }
    }
    if (MODE == 4) { t.x += t1.x; t.y += t1.y; td.x += td1.x; td.y += td1.y; }
    if (MODE == 1) { wr[it & 31] += 1e-9; }
    if (MODE == 3) { if (t.x == 1.2345 && td.y == 5.4321) part[warp][0][lane] = t; }
    else { part[warp][0][(lane + hd.x) & 31] = t; part[warp][1][(lane + hd.y) & 31] = td; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = part[warp][0][lane].x + part[warp][1][lane].y;
}

// ROWS x rows per thread x 16 window cells, warp-uniform weight quads (v2 with register blocking), NWARP warps per CTA
template <bool GRAD, int ROWS, int NWARP> __global__ void __launch_bounds__(NWARP * 32, 1) k_gr(double *out, const double *src, int iters) {
  __shared__ __align__(16) double rows[NR * RL];
  __shared__ __align__(16) double2 part[NWARP][2 * ROWS][32];
  model_rows(rows, src, blockDim.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double2 win[ROWS][16];
#pragma unroll
  for (int r = 0; r < ROWS; r++)
#pragma unroll
    for (int i = 0; i < 16; i++) win[r][i] = make_double2(out[threadIdx.x * 64 + r * 16 + i], out[32768 + threadIdx.x * 64 + r * 16 + i]);
  for (int it = 0; it < iters; it++) {
    const double *row = rows + ((it * 7 + warp) & (NR - 1)) * RL;
    const int4 hd = *reinterpret_cast<const int4 *>(row + 62);
    double2 t[ROWS], td[ROWS], t1[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) { t[r] = make_double2(0, 0); td[r] = t[r]; t1[r] = t[r]; }
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      const double2 w = *reinterpret_cast<const double2 *>(row + k);
      double2 dw = w;
      if (GRAD) dw = *reinterpret_cast<const double2 *>(row + 16 + k);
#pragma unroll
      for (int r = 0; r < ROWS; r++) {
        if (GRAD) {
          t[r].x = fma(w.x, win[r][k].x, t[r].x); t[r].y = fma(w.x, win[r][k].y, t[r].y); td[r].x = fma(dw.x, win[r][k].x, td[r].x); td[r].y = fma(dw.x, win[r][k].y, td[r].y);
          t[r].x = fma(w.y, win[r][k + 1].x, t[r].x); t[r].y = fma(w.y, win[r][k + 1].y, t[r].y); td[r].x = fma(dw.y, win[r][k + 1].x, td[r].x); td[r].y = fma(dw.y, win[r][k + 1].y, td[r].y);
        } else {
          t[r].x = fma(w.x, win[r][k].x, t[r].x); t[r].y = fma(w.x, win[r][k].y, t[r].y);
          t1[r].x = fma(w.y, win[r][k + 1].x, t1[r].x); t1[r].y = fma(w.y, win[r][k + 1].y, t1[r].y);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      if (!GRAD) { t[r].x += t1[r].x; t[r].y += t1[r].y; }
      part[warp][r][(lane + hd.x) & 31] = t[r];
      if (GRAD) part[warp][ROWS + r][(lane + hd.y) & 31] = td[r];
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = part[warp][0][lane].x + part[warp][2 * ROWS - 1][lane].y;
}
// scatter with ROWS x rows per thread
template <int ROWS, int NWARP> __global__ void __launch_bounds__(NWARP * 32, 1) k_sr(double *out, const double *src, int iters) {
  __shared__ __align__(16) double rows[NR * RL];
  model_rows(rows, src, blockDim.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hlf = lane >> 4, r1 = lane & 15;
  double2 win[ROWS][16];
#pragma unroll
  for (int r = 0; r < ROWS; r++)
#pragma unroll
    for (int i = 0; i < 16; i++) win[r][i] = make_double2(0, 0);
  for (int it = 0; it < iters; it++) {
    const double *row = rows + ((it * 7 + warp) & (NR - 1)) * RL;
    const int4 hd = *reinterpret_cast<const int4 *>(row + 62);
    const double w1 = row[40 + r1 + (hd.y & 3)];
    const double2 f = *reinterpret_cast<const double2 *>(row + 60);
    const double ux = w1 * f.x, uy = w1 * f.y;
    double2 A[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) { const double w0 = row[32 + ROWS * hlf + r + (hd.x & 3)]; A[r] = make_double2(w0 * ux, w0 * uy); }
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      const double2 w = *reinterpret_cast<const double2 *>(row + k);
#pragma unroll
      for (int r = 0; r < ROWS; r++) {
        win[r][k].x = fma(w.x, A[r].x, win[r][k].x); win[r][k].y = fma(w.x, A[r].y, win[r][k].y);
        win[r][k + 1].x = fma(w.y, A[r].x, win[r][k + 1].x); win[r][k + 1].y = fma(w.y, A[r].y, win[r][k + 1].y);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int r = 0; r < ROWS; r++)
#pragma unroll
    for (int i = 0; i < 16; i++) s += win[r][i].x + win[r][i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// scatter v2: one row, 16 cells, uniform quads
__global__ void __launch_bounds__(384, 1) k_s2(double *out, const double *src, int iters) {
  __shared__ __align__(16) double rows[NR * RL];
  model_rows(rows, src, blockDim.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hlf = lane >> 4, r1 = lane & 15;
  double2 win[16];
#pragma unroll
  for (int i = 0; i < 16; i++) win[i] = make_double2(0, 0);
  for (int it = 0; it < iters; it++) {
    const double *row = rows + ((it * 7 + warp) & (NR - 1)) * RL;
    const int4 hd = *reinterpret_cast<const int4 *>(row + 62);
    const double w0 = row[32 + hlf + (hd.x & 3)], w1 = row[40 + r1 + (hd.y & 3)];
    const double2 f = *reinterpret_cast<const double2 *>(row + 60);
    const double2 A = make_double2(w0 * (w1 * f.x), w0 * (w1 * f.y));
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      const double2 w = *reinterpret_cast<const double2 *>(row + k);
      win[k].x = fma(w.x, A.x, win[k].x); win[k].y = fma(w.x, A.y, win[k].y);
      win[k + 1].x = fma(w.y, A.x, win[k + 1].x); win[k + 1].y = fma(w.y, A.y, win[k + 1].y);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += win[i].x + win[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// scatter v4: two x rows, 13 cells, two nodes per pass in the lane halves; every ADV passes the finished cell of each row
// is combined across the halves (shuffles), staged, and the windows shift by one cell
template <int ADV> __global__ void __launch_bounds__(320, 1) k_s4(double *out, const double *src, int iters) {
  __shared__ __align__(16) double rows[NR * RL];
  __shared__ __align__(16) double2 stage[12][2][16];
  model_rows(rows, src, blockDim.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hlf = lane >> 4, r1 = lane & 15;
  double2 win[2][13];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int i = 0; i < 13; i++) win[r][i] = make_double2(0, 0);
  for (int it = 0; it < iters; it++) {
    const double *row = rows + ((it * 14 + warp * 2 + hlf) & (NR - 1)) * RL;
    const int4 hd = *reinterpret_cast<const int4 *>(row + 62);
    const double2 w0 = *reinterpret_cast<const double2 *>(row + 32 + 2 * (hd.x & 1));
    const double w1 = row[40 + r1 + (hd.y & 3)];
    const double2 f = *reinterpret_cast<const double2 *>(row + 60);
    const double ux = w1 * f.x, uy = w1 * f.y;
    const double2 A0 = make_double2(w0.x * ux, w0.x * uy), A1 = make_double2(w0.y * ux, w0.y * uy);
#pragma unroll
    for (int k = 0; k < 13; k += 2) {
      const double2 w = *reinterpret_cast<const double2 *>(row + k);
      win[0][k].x = fma(w.x, A0.x, win[0][k].x); win[0][k].y = fma(w.x, A0.y, win[0][k].y);
      win[1][k].x = fma(w.x, A1.x, win[1][k].x); win[1][k].y = fma(w.x, A1.y, win[1][k].y);
      if (k + 1 < 13) {
        win[0][k + 1].x = fma(w.y, A0.x, win[0][k + 1].x); win[0][k + 1].y = fma(w.y, A0.y, win[0][k + 1].y);
        win[1][k + 1].x = fma(w.y, A1.x, win[1][k + 1].x); win[1][k + 1].y = fma(w.y, A1.y, win[1][k + 1].y);
      }
    }
    if (it % ADV == ADV - 1) {
#pragma unroll
      for (int r = 0; r < 2; r++) {
        double2 c = win[r][0];
        c.x += __shfl_xor_sync(0xffffffffu, c.x, 16); c.y += __shfl_xor_sync(0xffffffffu, c.y, 16);
        if (hlf == 0) stage[warp][r][r1] = c;
#pragma unroll
        for (int i = 0; i < 12; i++) win[r][i] = win[r][i + 1];
        win[r][12] = make_double2(0, 0);
      }
    }
  }
  double s = stage[warp][0][r1].x;
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int i = 0; i < 13; i++) s += win[r][i].x + win[r][i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  g_clk = p.clockRate * 1e3;
  const int nsm = p.multiProcessorCount;
  printf("device %s, %d SMs, clock %.0f MHz\n", p.name, nsm, g_clk * 1e-6);
  double *out, *src; cudaMalloc(&out, 8 * 1024 * nsm); cudaMalloc(&src, 8 * 64);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(src, h, sizeof h, cudaMemcpyHostToDevice);

  printf("-- 1. DFMA: cycles per DFMA per warp, DFMA/clk/SM\n");
  {
    const int iters = 100000;
#define DF(CH, WARPS) { float ms = best_ms([&] { k_dfma<CH><<<nsm, WARPS * 32>>>(out, iters); }); \
      printf("warps/SM %2d chains %2d : %6.2f cycles per DFMA per warp, %5.1f DFMA/clk/SM\n", WARPS, CH, ms * 1e-3 * g_clk / ((double)iters * CH), \
             (double)CH * iters * WARPS * 32 / (ms * 1e-3 * g_clk)); }
    DF(1, 1) DF(2, 1) DF(4, 1) DF(8, 1) DF(16, 1) DF(1, 4) DF(2, 4) DF(4, 4) DF(8, 4) DF(16, 4) DF(4, 8) DF(8, 8) DF(4, 12) DF(8, 12) DF(4, 16)
  }
  printf("-- 2. LDS: cycles per LDS instruction per SM (all warps hammering), by pattern\n");
  {
    const int iters = 20000;
#define LD(BITS, PAT, WARPS, NAME) { float ms = best_ms([&] { k_lds<BITS, PAT><<<nsm, WARPS * 32, 16384>>>(out, iters); }); \
      printf("LDS.%-3d %-10s warps/SM %2d : %6.2f cycles per LDS per SM\n", BITS, NAME, WARPS, ms * 1e-3 * g_clk / ((double)iters * 16 * WARPS)); }
    LD(128, 0, 16, "uniform") LD(128, 1, 16, "2 addrs") LD(128, 2, 16, "4 addrs") LD(128, 3, 16, "distinct")
    LD(64, 0, 16, "uniform") LD(64, 1, 16, "2 addrs") LD(64, 2, 16, "4 addrs") LD(64, 3, 16, "distinct")
    LD(32, 0, 16, "uniform") LD(32, 3, 16, "distinct")
    { float ms = best_ms([&] { k_shfl<<<nsm, 16 * 32>>>(out, iters); });
      printf("SHFL.32 xor        warps/SM 16 : %6.2f cycles per SHFL per SM\n", ms * 1e-3 * g_clk / ((double)iters * 16 * 16)); }
  }
  printf("-- 3. node-loop models, 12 (v2) / 10 (v4) warps per SM all busy: cycles per SM per node visit (v4 pass = 2 visits); FP64 pipe cycles needed\n");
  {
    const int iters = 20000;
    cudaMemset(out, 0, 8 * 1024 * nsm);
#define RUNM(KERN, VIS, DFMA, NAME) { const int NW = (VIS) == 2 ? 10 : 12; float ms = best_ms([&] { KERN<<<nsm, NW * 32>>>(out, src, iters); }); \
      const double cyc = ms * 1e-3 * g_clk / iters / NW / (VIS); \
      printf("%-58s : %6.1f cycles per visit per SM (FP64 floor %5.1f) -> pipe %4.1f %%\n", NAME, cyc, (DFMA) * 2.0 / 4.0 / (VIS), 100.0 * (DFMA) * 2.0 / 4.0 / (VIS) / cyc); }
#define RUNX(MODE, NW, NAME) { float ms = best_ms([&] { k_gx<MODE, NW><<<nsm, NW * 32>>>(out, src, iters); }); \
      const double cyc = ms * 1e-3 * g_clk / iters / NW; \
      printf("gx mode %d %-44s warps %2d : %6.1f cycles per visit per SM -> pipe %4.1f %%\n", MODE, NAME, NW, cyc, 100.0 * 32.0 / cyc); }
    RUNX(0, 12, "baseline (16 LDS.128, 2 STS.128)") RUNX(0, 8, "baseline") RUNX(0, 4, "baseline") RUNX(0, 16, "baseline")
    RUNX(1, 12, "weights in registers") RUNX(1, 4, "weights in registers")
    RUNX(2, 12, "LDS.64") RUNX(3, 12, "no STS") RUNX(4, 12, "8 chains") RUNX(4, 4, "8 chains") RUNX(5, 12, "same row every iteration")
#define RUNR(KERN, DFMA, NW, NAME) { float ms = best_ms([&] { KERN<<<nsm, NW * 32>>>(out, src, iters); }); \
      const double cyc = ms * 1e-3 * g_clk / iters / NW; \
      printf("%-50s warps %2d : %6.1f cycles per visit per SM (floor %5.1f) -> pipe %4.1f %%\n", NAME, NW, cyc, (DFMA) / 2.0, 100.0 * (DFMA) / 2.0 / cyc); }
    RUNR((k_gr<true, 1, 12>), 64, 12, "gather F+grad RPT=1") RUNR((k_gr<true, 2, 8>), 128, 8, "gather F+grad RPT=2") RUNR((k_gr<true, 2, 4>), 128, 4, "gather F+grad RPT=2")
    RUNR((k_gr<false, 1, 12>), 32, 12, "gather F RPT=1") RUNR((k_gr<false, 2, 8>), 64, 8, "gather F RPT=2") RUNR((k_gr<false, 2, 4>), 64, 4, "gather F RPT=2")
    RUNR((k_sr<1, 12>), 32, 12, "scatter F RPT=1") RUNR((k_sr<2, 8>), 64, 8, "scatter F RPT=2") RUNR((k_sr<2, 4>), 64, 4, "scatter F RPT=2")
    RUNR((k_gr<false, 3, 8>), 96, 8, "gather F RPT=3") RUNR((k_sr<3, 8>), 96, 8, "scatter F RPT=3")
    RUNM(k_g2<true>, 1, 64, "gather F+grad v2 (1 row x16, uniform)")
    RUNM((k_g4<true, 2>), 2, 104, "gather F+grad v4 (2 rows x13, 2 nodes/pass, adv/2 passes)")
    RUNM((k_g4<true, 1000000>), 2, 104, "gather F+grad v4 (no window advance)")
    RUNM(k_g2<false>, 1, 32, "gather F v2")
    RUNM((k_g4<false, 2>), 2, 52, "gather F v4")
    RUNM(k_s2, 1, 32, "scatter F v2 (1 row x16, uniform)")
    RUNM(k_s4<2>, 2, 52, "scatter F v4 (2 rows x13, 2 nodes/pass, adv/2 passes)")
    RUNM(k_s4<1000000>, 2, 52, "scatter F v4 (no window advance)")
  }
  return 0;
}
