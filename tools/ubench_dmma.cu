// Microbenchmark (development aid): FP64 mma.sync (DMMA) rate and latency on sm_100a by shape, against the DFMA pipe.
// FMA/clk/SM is counted as M*N*K per instruction; the DFMA pipe peaks at 64 FMA/clk/SM (tools/ubench3.cu).
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void mma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// SHAPE 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16.  NA independent accumulator tiles per warp.
// MIX > 0: MIX plain DFMAs (independent chains) issued per MMA, to see whether the two share a pipe.
template <int SHAPE, int NA, int MIX> __global__ void k(double *out, const double *src, int iters) {
  double a[8], b[4], c[NA][4], d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = src[i + (threadIdx.x & 3)];
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = src[8 + i + (threadIdx.x & 7)];
#pragma unroll
  for (int j = 0; j < NA; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) c[j][i] = src[16 + i + j];
#pragma unroll
  for (int i = 0; i < 8; i++) d[i] = src[32 + i];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < NA; j++) {
      if (SHAPE == 0) { double (&cc)[2] = *reinterpret_cast<double (*)[2]>(&c[j][0]); mma884(cc, a[j & 7], b[j & 3]); }
      else if (SHAPE == 1) { const double aa[2] = {a[j & 7], a[(j + 1) & 7]}; mma1684(c[j], aa, b[j & 3]); }
      else if (SHAPE == 2) { const double aa[4] = {a[0], a[1], a[2], a[3]}; const double bb[2] = {b[0], b[1]}; mma1688(c[j], aa, bb); }
      else mma16816(c[j], a, b);
#pragma unroll
      for (int q = 0; q < MIX; q++) d[(j * MIX + q) & 7] = fma(d[(j * MIX + q) & 7], a[q & 7], b[q & 3]);
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < NA; j++) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
#pragma unroll
  for (int i = 0; i < 8; i++) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int NA, int MIX> void run(int warps, double *out, double *src, int nsm) {
  const int iters = 4000;
  static const char *names[] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
  static const int fma_per[] = {256, 512, 1024, 2048};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0); k<SHAPE, NA, MIX><<<nsm, warps * 32>>>(out, src, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  const double clk = best * 1e-3 * 1.965e9;
  const double n_mma = (double)iters * NA * warps;     // per SM
  printf("%-9s tiles/warp %2d warps/SM %2d dfma/mma %d : %6.1f cycles per MMA per warp, %6.1f MMA-FMA/clk/SM, +%5.1f DFMA/clk/SM\n",
         names[SHAPE], NA, warps, MIX, clk / ((double)iters * NA), n_mma * fma_per[SHAPE] / clk, n_mma * MIX * 32 / clk);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  double *out, *src; cudaMalloc(&out, 8 * 1024 * p.multiProcessorCount); cudaMalloc(&src, 8 * 64);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(src, h, sizeof h, cudaMemcpyHostToDevice);
  const int n = p.multiProcessorCount;
  printf("-- latency: one dependent tile per warp, one warp per SM\n");
  run<0, 1, 0>(1, out, src, n); run<1, 1, 0>(1, out, src, n); run<2, 1, 0>(1, out, src, n); run<3, 1, 0>(1, out, src, n);
  printf("-- throughput by independent tiles per warp and warps per SM\n");
  for (int w : {1, 4, 8, 12, 16}) {
    run<0, 2, 0>(w, out, src, n); run<0, 4, 0>(w, out, src, n); run<0, 8, 0>(w, out, src, n);
    run<1, 2, 0>(w, out, src, n); run<1, 4, 0>(w, out, src, n); run<1, 8, 0>(w, out, src, n);
    run<2, 2, 0>(w, out, src, n); run<2, 4, 0>(w, out, src, n); run<2, 8, 0>(w, out, src, n);
    run<3, 2, 0>(w, out, src, n); run<3, 4, 0>(w, out, src, n); run<3, 8, 0>(w, out, src, n);
  }
  printf("-- mixed with plain DFMA (same pipe?)\n");
  for (int w : {4, 12}) {
    run<1, 8, 2>(w, out, src, n); run<1, 8, 8>(w, out, src, n); run<3, 8, 8>(w, out, src, n); run<3, 8, 32>(w, out, src, n);
  }
  return 0;
}
