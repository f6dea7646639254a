// Microbenchmark (development aid): DFMA cadence against operand pattern AND warps per scheduler.
// MODE 0: acc = acc * const + const; 1: acc = fresh * const + acc; 2: acc = fresh_b * fresh_c + acc (gather pattern:
// weight x window cell + accumulator); 3: as 2 with the weight shared by two consecutive DFMAs (t.x, t.y)
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE, int NA> __global__ void k(double *out, const double *src, int iters) {
  constexpr int NB = 16;
  double a[NA], b[NB], c[2 * NB];
#pragma unroll
  for (int i = 0; i < NA; i++) a[i] = src[i];
#pragma unroll
  for (int i = 0; i < NB; i++) b[i] = src[8 + i];
#pragma unroll
  for (int i = 0; i < 2 * NB; i++) c[i] = src[24 + i + (threadIdx.x & 1)];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 2 * NB; i++) {
      if (MODE == 0) a[i % NA] = fma(a[i % NA], b[0], c[0]);
      else if (MODE == 1) a[i % NA] = fma(b[i % NB], c[0], a[i % NA]);
      else if (MODE == 2) a[i % NA] = fma(b[i % NB], c[(i * 5 + 3) % (2 * NB)], a[i % NA]);
      else a[i % NA] = fma(b[(i / 2) % NB], c[i], a[i % NA]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int NA> void run(const char *name, int warps, double *out, double *src, int nsm) {
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0); k<MODE, NA><<<nsm, warps * 32>>>(out, src, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  printf("%-44s chains %2d warps/SM %2d : %5.1f DFMA/clk/SM (of 64)\n", name, NA, warps, (double)iters * 32 * warps * 32 / (best * 1e-3 * 1.965e9));
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double *out, *src; cudaMalloc(&out, 8 * 1024 * p.multiProcessorCount); cudaMalloc(&src, 8 * 64);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(src, h, sizeof h, cudaMemcpyHostToDevice);
  const int n = p.multiProcessorCount;
  for (int w : {4, 8, 12}) {
    run<0, 8>("acc = acc * const + const", w, out, src, n);
    run<1, 8>("acc = fresh * const + acc", w, out, src, n);
    run<2, 8>("acc = fresh * fresh + acc", w, out, src, n);
    run<3, 8>("acc = w(shared by 2) * fresh + acc", w, out, src, n);
    run<2, 4>("acc = fresh * fresh + acc", w, out, src, n);
    run<3, 4>("acc = w(shared by 2) * fresh + acc", w, out, src, n);
  }
  return 0;
}
