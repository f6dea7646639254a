// Microbenchmark (development aid): the z-march gather node loop in isolation, ONE warp per scheduler, nothing else on
// the SM.  Per iteration: 8 (F) or 16 (F+grad) LDS.128 of warp-uniform weights, 32 / 64 DFMA  t += w[k] * win[k]  on 16
// register-resident cells, NCH accumulation chains per sum.  Separates chain latency, LDS issue cost and DFMA issue.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/nodeloop_model.cu
#include <cuda_runtime.h>
#include <cstdio>
constexpr int W = 16;
template <bool GRAD, int NCH, int LOADS> __global__ void __launch_bounds__(128) k(double *out, const double *src, int iters) {
  __shared__ __align__(16) double rows[64][32];
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) rows[i / 32][i % 32] = src[i % 64];
  __syncthreads();
  double win[W][2];
#pragma unroll
  for (int i = 0; i < W; i++) { win[i][0] = src[i] + threadIdx.x; win[i][1] = src[i + 16] - threadIdx.x; }
  double acc = 0;
  double wreg[W], dwreg[W];
#pragma unroll
  for (int i = 0; i < W; i++) { wreg[i] = src[32 + i]; dwreg[i] = src[48 + i]; }
  for (int it = 0; it < iters; it++) {
    const double *row = rows[(it * 7) & 63] + (LOADS == 2 ? 0 : 0);
    double t[NCH][2], td[NCH][2];
#pragma unroll
    for (int c = 0; c < NCH; c++) { t[c][0] = t[c][1] = td[c][0] = td[c][1] = 0; }
#pragma unroll
    for (int k = 0; k < W; k += 2) {
      double2 w, dw;
      if (LOADS == 0) { w = make_double2(wreg[k], wreg[k + 1]); dw = make_double2(dwreg[k], dwreg[k + 1]); }
      else {
        w = *reinterpret_cast<const double2 *>(row + k);
        if (GRAD) dw = *reinterpret_cast<const double2 *>(row + 16 + k);
      }
      const int c0 = k % NCH, c1 = (k + 1) % NCH;
      t[c0][0] = fma(w.x, win[k][0], t[c0][0]); t[c0][1] = fma(w.x, win[k][1], t[c0][1]);
      t[c1][0] = fma(w.y, win[k + 1][0], t[c1][0]); t[c1][1] = fma(w.y, win[k + 1][1], t[c1][1]);
      if (GRAD) {
        td[c0][0] = fma(dw.x, win[k][0], td[c0][0]); td[c0][1] = fma(dw.x, win[k][1], td[c0][1]);
        td[c1][0] = fma(dw.y, win[k + 1][0], td[c1][0]); td[c1][1] = fma(dw.y, win[k + 1][1], td[c1][1]);
      }
    }
    // tree sum of the chains, ONE dependent add on acc per iteration (the kernel stores the sums instead)
    double s0 = 0, s1 = 0;
#pragma unroll
    for (int c = 0; c < NCH; c++) { s0 += t[c][0] + t[c][1]; if (GRAD) s1 += td[c][0] + td[c][1]; }
    if (NCH > 1) { rows[63][(it & 1) * 16 + (threadIdx.x & 15)] = s0 + s1; } else acc += s0 + s1;
    if (LOADS == 0) { wreg[it & 15] += 1e-9; dwreg[(it + 3) & 15] += 1e-9; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <bool GRAD, int NCH, int LOADS> void run(const char *name, double *out, double *src, int nsm) {
  const int iters = 50000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0); k<GRAD, NCH, LOADS><<<nsm, 128>>>(out, src, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  printf("%-64s : %.0f cycles per node iteration\n", name, best * 1e-3 * 1.965e9 / iters);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double *out, *src; cudaMalloc(&out, 8 * 128 * p.multiProcessorCount); cudaMalloc(&src, 8 * 64);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(src, h, sizeof h, cudaMemcpyHostToDevice);
  const int n = p.multiProcessorCount;
  run<false, 1, 1>("F      32 DFMA,  8 LDS.128, 1 chain per sum (16 deep)", out, src, n);
  run<false, 2, 1>("F      32 DFMA,  8 LDS.128, 2 chains per sum", out, src, n);
  run<false, 4, 1>("F      32 DFMA,  8 LDS.128, 4 chains per sum", out, src, n);
  run<false, 2, 0>("F      32 DFMA,  weights in registers, 2 chains", out, src, n);
  run<true, 1, 1>("F+grad 64 DFMA, 16 LDS.128, 1 chain per sum (16 deep)", out, src, n);
  run<true, 2, 1>("F+grad 64 DFMA, 16 LDS.128, 2 chains per sum", out, src, n);
  run<true, 4, 1>("F+grad 64 DFMA, 16 LDS.128, 4 chains per sum", out, src, n);
  run<true, 2, 0>("F+grad 64 DFMA, weights in registers, 2 chains", out, src, n);
  return 0;
}
