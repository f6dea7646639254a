// Micro-benchmarks for the roofline denominators that MEASURED_PEAKS.json does not carry:
// FP64 FMA peak (independent DFMA chains on every SM) and shared-memory LDS.128 bandwidth.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *out, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_lds(double *out, int iters) {
  extern __shared__ double2 sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = make_double2(i, -i);
  __syncthreads();
  double2 acc = make_double2(0, 0);
  int idx = threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) { double2 v = sm[(idx + u * 1024) & 8191]; acc.x += v.x; acc.y += v.y; }
    idx = (idx + 37) & 8191;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nb = p.multiProcessorCount * 2, nt = 1024, iters = 20000;
  double *out; cudaMalloc(&out, sizeof(double) * nb * nt);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0); k_dfma<<<nb, nt>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double flops = 2.0 * 8 * (double)iters * nb * nt;
  printf("{\"fp64_fma_tflops\": %.2f, \"sms\": %d, \"ms\": %.3f}\n", flops / best * 1e-9, p.multiProcessorCount, best);
  cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
  best = 1e30f;
  const int it2 = 4000;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0); k_lds<<<p.multiProcessorCount, nt, 8192 * 16>>>(out, it2); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double bytes = 16.0 * 8 * (double)it2 * p.multiProcessorCount * nt;
  printf("{\"smem_lds128_tbs\": %.2f, \"ms\": %.3f}\n", bytes / best * 1e-9, best);
  return 0;
}
