// Microbenchmark (development aid): FP64 FMA throughput of one SM as a function of resident warps and independent
// accumulator chains per thread.  Answers: can the 2-3 DFMA-issuing warps per scheduler of the z-march kernels keep the
// FP64 pipe busy, or does a single in-order warp cap it?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
template <int CH> __global__ void k(double *out, int iters) {
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
template <int CH> void run(int warps, double *out, int nsm) {
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0);
    k<CH><<<nsm, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  const double dfma_per_clk_sm = (double)CH * iters * warps * 32 / (best * 1e-3 * 1.965e9);
  printf("warps/SM %2d chains %2d : %.1f DFMA/clk/SM (of 64)\n", warps, CH, dfma_per_clk_sm);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double *out; cudaMalloc(&out, 8 * 1024 * p.multiProcessorCount);
  for (int w : {1, 2, 4, 8, 12, 16, 32}) { run<4>(w, out, p.multiProcessorCount); run<8>(w, out, p.multiProcessorCount); run<26>(w, out, p.multiProcessorCount); }
  return 0;
}
