// Roofline denominators the driver-written MEASURED_PEAKS.json does not carry: the FP64 FMA rate of this GPU,
// measured with independent DFMA chains on every SM (the gridding kernels are FP64-pipe work).
#include <cuda_runtime.h>

#include <cstdio>

namespace {
__global__ void __launch_bounds__(1024) k_dfma_chains(double *out, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
}  // namespace

extern "C" double pnfft_b200_measure_fp64_tflops(void) {
  cudaDeviceProp p;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) return -1.0;
  const int nb = p.multiProcessorCount * 2, nt = 1024, iters = 20000;
  double *out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * nb * nt) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 6; r++) {
    cudaEventRecord(e0);
    k_dfma_chains<<<nb, nt>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  const double flops = 2.0 * 8 * (double)iters * nb * nt;
  return flops / best * 1e-9;
}
