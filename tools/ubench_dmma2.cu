// Microbenchmark (development aid): what overlaps with DMMA on one scheduler?  Per DMMA.8x8x4, K extra instructions of one
// kind (independent of the MMAs): integer ALU, FP32 FMA, DFMA, shared-memory loads, shuffles.  8 warps per SM, 8 tiles.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void mma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// KIND 0: none, 1: IADD/LOP chains, 2: FFMA chains, 3: DFMA chains, 4: LDS.64 (conflict free), 5: SHFL, 6: DFMA dependent on MMA result
template <int KIND, int K> __global__ void k(double *out, const double *src, int iters) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = src[i & 63];
  __syncthreads();
  double a[4], b[4], c[8][2], d[8];
  int xi[8]; float xf[8];
#pragma unroll
  for (int i = 0; i < 4; i++) { a[i] = src[i + (threadIdx.x & 3)]; b[i] = src[8 + i + (threadIdx.x & 7)]; }
#pragma unroll
  for (int j = 0; j < 8; j++) { c[j][0] = src[16 + j]; c[j][1] = src[17 + j]; d[j] = src[32 + j]; xi[j] = threadIdx.x + j; xf[j] = (float)src[j]; }
  const double *lp = sm + (threadIdx.x & 31);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      mma884(c[j], a[j & 3], b[j & 3]);
#pragma unroll
      for (int q = 0; q < K; q++) {
        const int s = (j * K + q) & 7;
        if (KIND == 1) xi[s] = (xi[s] ^ it) + q;
        else if (KIND == 2) xf[s] = fmaf(xf[s], 1.0001f, 0.5f);
        else if (KIND == 3) d[s] = fma(d[s], a[q & 3], b[q & 3]);
        else if (KIND == 4) d[s] += lp[((it + q) & 15) * 32];
        else if (KIND == 5) xi[s] = __shfl_xor_sync(0xffffffffu, xi[s], 1 + (q & 3));
        else if (KIND == 6) d[s] = fma(c[(j + 5) & 7][q & 1], a[q & 3], d[s]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1] + d[j] + xi[j] + xf[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND, int K> void run(int warps, double *out, double *src, int nsm) {
  static const char *names[] = {"none", "int ALU", "FFMA", "DFMA", "LDS.64 (+DADD)", "SHFL", "DFMA on MMA result"};
  const int iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0); k<KIND, K><<<nsm, warps * 32>>>(out, src, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  const double clk = best * 1e-3 * 1.965e9;
  const double per_smsp = (double)iters * 8 * warps / 4;   // DMMAs per scheduler
  printf("%-20s %2d per DMMA, warps/SM %2d : %6.2f cycles per DMMA per scheduler (16 = pipe bound)\n", names[KIND], K, warps, clk / per_smsp);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double *out, *src; cudaMalloc(&out, 8 * 1024 * p.multiProcessorCount); cudaMalloc(&src, 8 * 64);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(src, h, sizeof h, cudaMemcpyHostToDevice);
  const int n = p.multiProcessorCount;
  for (int w : {4, 12}) {
    run<0, 0>(w, out, src, n);
    run<1, 2>(w, out, src, n); run<1, 4>(w, out, src, n); run<1, 8>(w, out, src, n); run<1, 16>(w, out, src, n);
    run<2, 4>(w, out, src, n); run<2, 8>(w, out, src, n);
    run<3, 1>(w, out, src, n); run<3, 2>(w, out, src, n); run<3, 4>(w, out, src, n);
    run<6, 1>(w, out, src, n); run<6, 2>(w, out, src, n);
    run<4, 1>(w, out, src, n); run<4, 2>(w, out, src, n); run<4, 4>(w, out, src, n);
    run<5, 1>(w, out, src, n); run<5, 2>(w, out, src, n);
  }
  return 0;
}
