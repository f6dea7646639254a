// Microbenchmark (development aid): cadence of DFMA with one warp per scheduler when the three 64-bit source operands
// are (a) two loop constants + accumulator, (b) one fresh register pair + constant + accumulator, (c) two fresh register
// pairs + accumulator (the z-march gather: t += w[k] * win[k]).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE> __global__ void __launch_bounds__(128) k(double *out, const double *src, int iters) {
  constexpr int NA = 8, NB = 16;
  double a[NA], b[NB], c[NB];
#pragma unroll
  for (int i = 0; i < NA; i++) a[i] = src[i];
#pragma unroll
  for (int i = 0; i < NB; i++) { b[i] = src[8 + i]; c[i] = src[24 + i + (threadIdx.x & 1)]; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NB; i++) {
      if (MODE == 0) a[i % NA] = fma(a[i % NA], b[0], c[0]);
      else if (MODE == 1) a[i % NA] = fma(b[i], c[0], a[i % NA]);
      else a[i % NA] = fma(b[i], c[(i * 5 + 3) % NB], a[i % NA]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, double *out, double *src, int nsm) {
  const int iters = 100000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0); k<MODE><<<nsm, 128>>>(out, src, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  printf("%-52s : %.2f cycles per DFMA (one warp per scheduler, 8 chains)\n", name, best * 1e-3 * 1.965e9 / ((double)iters * 16));
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double *out, *src; cudaMalloc(&out, 8 * 128 * p.multiProcessorCount); cudaMalloc(&src, 8 * 64);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(src, h, sizeof h, cudaMemcpyHostToDevice);
  run<0>("acc = acc * const + const", out, src, p.multiProcessorCount);
  run<1>("acc = fresh * const + acc", out, src, p.multiProcessorCount);
  run<2>("acc = fresh * fresh + acc (gather pattern)", out, src, p.multiProcessorCount);
  return 0;
}
