// PNFFT_COMPUTE_DIRECT: the direct NDFT A / A^H -- the slow O(M N^3) truth the reference's own drivers compare the fast
// transform against (reference kernel/ndft-parallel.c:377-615 trafo_A, :617-722 adj_A, dispatched from
// api/api-basic.c:224-231, 358-365).  Not part of the accelerated path and not tuned: one thread per node (trafo) or per
// Fourier coefficient (adjoint), phases exp(-/+ 2 pi i k x) from exactly reduced products instead of the reference's
// three nested recurrences, sums kept in double for both precisions.
//
//   trafo : f_j = sum_k f_hat_k e^{-2 pi i k x_j},  grad_f_j = -2 pi i sum_k k f_hat_k e^{..},  hessian = -4 pi^2 sum_k k k^T ..
//           outputs are always overwritten (trafo_A zeroes them whatever PNFFT_COMPUTE_ACCUMULATED says, :404-418)
//   adj   : f_hat_k += sum_j (f_j + 2 pi i k . grad_f_j) e^{+2 pi i k x_j}   (adj_A adds to what f_hat holds, :648-651)
//   c2r   : f_hat is the half spectrum k2 in [-N2/2, 0]; every coefficient counts twice except the self-conjugate and the
//           redundant ones of the planes k2 = 0, -N2/2, which the reference skips (is_hermitian, :356-374; applied, as
//           there, to the loop variables in MEMORY order, also with PNFFT_TRANSPOSED_F_HAT).
//           Defect of the reference NOT replicated: its c2r gradient has the wrong sign (:517-519 add 2 k Im(f_hat e), :597
//           multiplies by -2 pi; d/dx Re(f_hat e^{-2 pi i k x}) = +2 pi k Im(f_hat e^{..})), so that its own NDFT and NFFT
//           c2r gradients differ by -1; here the direct gradient is the derivative of the direct f.
// Multi-rank: as in the reference every rank's block is broadcast in turn (trafo) / reduced to its owner (adjoint), through
// host memory (MPI_Bcast / MPI_Reduce of the plan's communicator).
#pragma once
#include <chrono>
#include <vector>

namespace pnb {

struct DirectBlock {
  int len[3];      // extents of the f_hat block, memory order
  int start[3];    // first frequency along each memory-order axis
  int axis[3];     // the coordinate axis (0, 1, 2) each memory-order axis belongs to
  int Nax[3];      // N of each memory-order axis
};

// exp(sign 2 pi i k x): the product k x is split into its rounded value and the exact residual (fma) before the integer
// part is dropped, so the phase is good to an ulp of the fraction whatever k
__device__ __forceinline__ double direct_turns(double k, double x) {
  const double p = k * x, r = fma(k, x, -p);
  return (p - rint(p)) + r;
}
__device__ __forceinline__ void direct_phase(double k, double x, double sign, double &c, double &s) {
  sincospi(2.0 * sign * direct_turns(k, x), &s, &c);
}

// weight of coefficient (k0, k1, k2) of a c2r half spectrum in the real-valued sums: 1 for the origin, 0 for the
// self-conjugate / redundant coefficients the reference leaves out (kernel/ndft-parallel.c:356-374), 2 otherwise
__device__ __forceinline__ double direct_c2r_weight(int k0, int k1, int k2, const int *N) {
  if (k0 == 0 && k1 == 0 && k2 == 0) return 1.0;
  const bool e0 = (k0 == 0 || k0 == -N[0] / 2), e1 = (k1 == 0 || k1 == -N[1] / 2), e2 = (k2 == 0 || k2 == -N[2] / 2);
  if (e0 && e1 && e2) return 0.0;
  if (e2 && k1 > 0) return 0.0;
  if (e2 && e1 && k0 > 0) return 0.0;
  return 2.0;
}

// acc[j][20] += { S, G_0..2, H_00 H_01 H_02 H_11 H_12 H_22 } (complex each, indices in MEMORY order of the block)
template <class R, bool CPLX>
__global__ void k_direct_trafo(DirectBlock B, const R *__restrict__ fh, const R *__restrict__ x, int M, double *__restrict__ acc) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const double x0 = (double)x[3 * (size_t)j + B.axis[0]], x1 = (double)x[3 * (size_t)j + B.axis[1]], x2 = (double)x[3 * (size_t)j + B.axis[2]];
  double s[20];
#pragma unroll
  for (int q = 0; q < 20; q++) s[q] = 0.0;
  double stc, sts;
  direct_phase(1.0, x2, -1.0, stc, sts);
  size_t m = 0;
  for (int i0 = 0; i0 < B.len[0]; i0++) {
    const int k0 = B.start[0] + i0;
    double c0, s0;
    direct_phase((double)k0, x0, -1.0, c0, s0);
    for (int i1 = 0; i1 < B.len[1]; i1++) {
      const int k1 = B.start[1] + i1;
      double c1, s1;
      direct_phase((double)k1, x1, -1.0, c1, s1);
      const double c01 = c0 * c1 - s0 * s1, s01 = c0 * s1 + s0 * c1;
      double ar = 0, ai = 0, br = 0, bi = 0, cr = 0, ci = 0, er = 0, ei = 0;
      for (int i2 = 0; i2 < B.len[2]; i2++, m++) {
        const int k2 = B.start[2] + i2;
        if ((i2 & 15) == 0) {            // fresh phase every 16 coefficients, a short recurrence in between
          double c2, s2;
          direct_phase((double)k2, x2, -1.0, c2, s2);
          er = c01 * c2 - s01 * s2; ei = c01 * s2 + s01 * c2;
        } else {
          const double t = er * stc - ei * sts;
          ei = er * sts + ei * stc; er = t;
        }
        const double fr = (double)fh[2 * m], fi = (double)fh[2 * m + 1];
        double vr = fr * er - fi * ei, vi = fr * ei + fi * er;
        if (!CPLX) { const double w = direct_c2r_weight(k0, k1, k2, B.Nax); vr *= w; vi *= w; }
        const double kk = (double)k2;
        ar += vr; ai += vi; br += kk * vr; bi += kk * vi; cr += kk * kk * vr; ci += kk * kk * vi;
      }
      const double a0 = (double)k0, a1 = (double)k1;
      s[0] += ar; s[1] += ai;
      s[2] += a0 * ar; s[3] += a0 * ai; s[4] += a1 * ar; s[5] += a1 * ai; s[6] += br; s[7] += bi;
      s[8] += a0 * a0 * ar; s[9] += a0 * a0 * ai; s[10] += a0 * a1 * ar; s[11] += a0 * a1 * ai; s[12] += a0 * br; s[13] += a0 * bi;
      s[14] += a1 * a1 * ar; s[15] += a1 * a1 * ai; s[16] += a1 * br; s[17] += a1 * bi; s[18] += cr; s[19] += ci;
    }
  }
  double *a = acc + 20 * (size_t)j;
#pragma unroll
  for (int q = 0; q < 20; q++) a[q] += s[q];
}

// sums -> user arrays (component order of the coordinate axes; Hessian xx, xy, xz, yy, yz, zz)
template <class R, bool CPLX>
__global__ void k_direct_trafo_store(DirectBlock B, const double *__restrict__ acc, int M, R *f, R *grad, R *hess) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const double *a = acc + 20 * (size_t)j;
  const double two_pi = 6.283185307179586476925286766559, four_pi2 = 39.478417604357434475337963999505;
  if (f) {
    if (CPLX) { f[2 * (size_t)j] = (R)a[0]; f[2 * (size_t)j + 1] = (R)a[1]; }
    else f[j] = (R)a[0];
  }
  if (grad)
    for (int q = 0; q < 3; q++) {
      const double gr = a[2 + 2 * q], gi = a[3 + 2 * q];
      const size_t o = 3 * (size_t)j + B.axis[q];
      if (CPLX) { grad[2 * o] = (R)(two_pi * gi); grad[2 * o + 1] = (R)(-two_pi * gr); }   // -2 pi i (gr + i gi)
      else grad[o] = (R)(two_pi * gi);
    }
  if (hess) {
    const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {0, 1, 2, 1, 2, 2};
    for (int q = 0; q < 6; q++) {
      int u = B.axis[pa[q]], v = B.axis[pb[q]];
      if (u > v) { const int t = u; u = v; v = t; }
      const int comp = u == 0 ? v : (u == 1 ? 2 + v : 5);      // (0,0) (0,1) (0,2) (1,1) (1,2) (2,2) -> 0..5
      const size_t o = 6 * (size_t)j + comp;
      if (CPLX) { hess[2 * o] = (R)(-four_pi2 * a[8 + 2 * q]); hess[2 * o + 1] = (R)(-four_pi2 * a[9 + 2 * q]); }
      else hess[o] = (R)(-four_pi2 * a[8 + 2 * q]);
    }
  }
}

// out[m] = sum_j (f_j + 2 pi i k . grad_j) e^{+2 pi i k x_j} for the coefficients of one block (memory order), in double
template <class R, bool CPLX>
__global__ void k_direct_adj(DirectBlock B, const R *__restrict__ x, const R *__restrict__ f, const R *__restrict__ grad, int M,
                             double *__restrict__ out) {
  const size_t total = (size_t)B.len[0] * B.len[1] * B.len[2];
  const size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = m < total;
  const int i2 = live ? (int)(m % B.len[2]) : 0, i1 = live ? (int)((m / B.len[2]) % B.len[1]) : 0, i0 = live ? (int)(m / ((size_t)B.len[2] * B.len[1])) : 0;
  const double k0 = (double)(B.start[0] + i0), k1 = (double)(B.start[1] + i1), k2 = (double)(B.start[2] + i2);
  const double two_pi = 6.283185307179586476925286766559;
  __shared__ double sx[128][3], sf[128][2], sgr[128][6];   // a tile of nodes: coordinates and gradient in memory order of the block, f
  double accr = 0, acci = 0;
  for (int base = 0; base < M; base += 128) {
    const int cnt = M - base < 128 ? M - base : 128;
    __syncthreads();
    for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
      const size_t j = (size_t)base + t;
      for (int q = 0; q < 3; q++) sx[t][q] = (double)x[3 * j + B.axis[q]];
      sf[t][0] = sf[t][1] = 0.0;
      if (f) { if (CPLX) { sf[t][0] = (double)f[2 * j]; sf[t][1] = (double)f[2 * j + 1]; } else sf[t][0] = (double)f[j]; }
      for (int q = 0; q < 3; q++) {
        sgr[t][2 * q] = sgr[t][2 * q + 1] = 0.0;
        if (grad) {
          const size_t o = 3 * j + B.axis[q];
          if (CPLX) { sgr[t][2 * q] = (double)grad[2 * o]; sgr[t][2 * q + 1] = (double)grad[2 * o + 1]; } else sgr[t][2 * q] = (double)grad[o];
        }
      }
    }
    __syncthreads();
    if (live)
      for (int t = 0; t < cnt; t++) {
        const double turns = direct_turns(k0, sx[t][0]) + direct_turns(k1, sx[t][1]) + direct_turns(k2, sx[t][2]);
        double c, s;
        sincospi(2.0 * turns, &s, &c);
        const double gr = k0 * sgr[t][0] + k1 * sgr[t][2] + k2 * sgr[t][4], gi = k0 * sgr[t][1] + k1 * sgr[t][3] + k2 * sgr[t][5];
        const double wr = sf[t][0] - two_pi * gi, wi = sf[t][1] + two_pi * gr;      // f + 2 pi i (gr + i gi)
        accr += wr * c - wi * s; acci += wr * s + wi * c;
      }
  }
  if (live) { out[2 * m] = accr; out[2 * m + 1] = acci; }
}

template <class R> struct Direct {
  typedef Plan<R> P;
  typedef Nodes<R> Nd;

  // the f_hat block of rank pid of the mesh, in memory order (what that rank's plan holds: reference
  // local_block_internal, kernel/ndft-parallel.c:416-419)
  static DirectBlock block_of(const P *p, int pid) {
    R xm[3] = {p->x_max[0], p->x_max[1], p->x_max[2]};
    return block_of_mesh(p->mesh, p->L.N, p->L.n, xm, p->L.m, p->L.c2r, p->pnfft_flags, pid);
  }
  static DirectBlock block_of_mesh(Mesh mesh, const INT *N, const INT *n, const R *xm, int m, bool c2r, unsigned flags, int pid) {
    mesh.co[0] = pid / mesh.np[1]; mesh.co[1] = pid % mesh.np[1]; mesh.rank = pid;
    Layout L;
    compute_layout<R>(L, mesh, N, n, xm, m, c2r, flags);
    DirectBlock B;
    const int ax[3] = {L.transposed ? 1 : 0, L.transposed ? 2 : 1, L.transposed ? 0 : 2};
    for (int q = 0; q < 3; q++) {
      B.axis[q] = ax[q]; B.len[q] = (int)L.local_N[ax[q]]; B.start[q] = (int)L.local_N_start[ax[q]]; B.Nax[q] = (int)L.N[ax[q]];
    }
    return B;
  }

  // device copy of a user array that may live on the host; *own tells whether it has to be freed
  static const R *on_device(const R *user, size_t count, bool *own) {
    *own = false;
    if (!user || !count) return user && is_device_ptr(user) ? user : nullptr;
    if (is_device_ptr(user)) return user;
    R *d = nullptr;
    PNB_CUDA(cudaMalloc((void **)&d, sizeof(R) * count));
    PNB_CUDA(cudaMemcpy(d, user, sizeof(R) * count, cudaMemcpyHostToDevice));
    *own = true;
    return d;
  }

  static bool too_big(const DirectBlock &B) {
    const size_t total = (size_t)B.len[0] * B.len[1] * B.len[2];
    if (2 * total * sizeof(double) > (size_t)0x7fffffff) {
      fprintf(stderr, "pnfft-b200: PNFFT_COMPUTE_DIRECT: an f_hat block of %zu coefficients is beyond the direct path's exchange buffers\n", total);
      return true;
    }
    return false;
  }

  static void trafo(P *p, Nd *nd, unsigned cf) {
    const auto t_begin = std::chrono::steady_clock::now();
    const int M = (int)nd->local_M;
    const bool cplx = !p->L.c2r;
    const int NC = cplx ? 2 : 1;
    if ((cf & 1u) && M && !nd->f) fprintf(stderr, "Error: missing memory allocation of nodes->f !!!\n");
    if ((cf & 2u) && M && !nd->grad_f) fprintf(stderr, "Error: missing memory allocation of nodes->grad_f !!!\n");
    if ((cf & 4u) && M && !nd->hessian_f) fprintf(stderr, "Error: missing memory allocation of nodes->hessian_f !!!\n");
    bool own_x = false;
    const R *dx = on_device(nd->x, (size_t)3 * M, &own_x);
    double *acc = nullptr;
    PNB_CUDA(cudaMalloc((void **)&acc, sizeof(double) * 20 * (size_t)(M ? M : 1)));
    PNB_CUDA(cudaMemset(acc, 0, sizeof(double) * 20 * (size_t)(M ? M : 1)));
    for (int pid = 0; pid < p->mesh.size; pid++) {
      const DirectBlock B = block_of(p, pid);
      const size_t total = (size_t)B.len[0] * B.len[1] * B.len[2];
      if (!total) continue;
      if (too_big(B)) break;
      const bool mine = pid == p->mesh.rank;
      R *dblk = nullptr;
      bool own_blk = false;
      if (p->mesh.size == 1) {
        dblk = const_cast<R *>(on_device((const R *)p->f_hat, 2 * total, &own_blk));
      } else {
        std::vector<R> h(2 * total);
        if (mine) PNB_CUDA(cudaMemcpy(h.data(), p->f_hat, sizeof(R) * 2 * total, cudaMemcpyDefault));
        MPI_Bcast(h.data(), (int)(sizeof(R) * 2 * total), MPI_BYTE, pid, p->comm);
        PNB_CUDA(cudaMalloc((void **)&dblk, sizeof(R) * 2 * total));
        PNB_CUDA(cudaMemcpy(dblk, h.data(), sizeof(R) * 2 * total, cudaMemcpyHostToDevice));
        own_blk = true;
      }
      if (M) {
        const unsigned nb = (unsigned)((M + 127) / 128);
        if (cplx) k_direct_trafo<R, true><<<nb, 128, 0, p->stream>>>(B, dblk, dx, M, acc);
        else k_direct_trafo<R, false><<<nb, 128, 0, p->stream>>>(B, dblk, dx, M, acc);
        PNB_CUDA(cudaGetLastError());
        p->launches++;
      }
      PNB_CUDA(cudaStreamSynchronize(p->stream));
      if (own_blk) cudaFree(dblk);
    }
    if (M) {
      // results straight into device-resident user arrays, through scratch for host arrays
      R *uf = (cf & 1u) ? nd->f : nullptr, *ug = (cf & 2u) ? nd->grad_f : nullptr, *uh = (cf & 4u) ? nd->hessian_f : nullptr;
      R *user[3] = {uf, ug, uh};
      const size_t cnt[3] = {(size_t)NC * M, (size_t)3 * NC * M, (size_t)6 * NC * M};
      R *dev[3] = {nullptr, nullptr, nullptr};
      bool own[3] = {false, false, false};
      for (int q = 0; q < 3; q++)
        if (user[q]) {
          if (is_device_ptr(user[q])) dev[q] = user[q];
          else { PNB_CUDA(cudaMalloc((void **)&dev[q], sizeof(R) * cnt[q])); own[q] = true; }
        }
      const DirectBlock B = block_of(p, p->mesh.rank);     // only the axis permutation matters here
      const unsigned nb = (unsigned)((M + 127) / 128);
      if (cplx) k_direct_trafo_store<R, true><<<nb, 128, 0, p->stream>>>(B, acc, M, dev[0], dev[1], dev[2]);
      else k_direct_trafo_store<R, false><<<nb, 128, 0, p->stream>>>(B, acc, M, dev[0], dev[1], dev[2]);
      PNB_CUDA(cudaGetLastError());
      p->launches++;
      PNB_CUDA(cudaStreamSynchronize(p->stream));
      for (int q = 0; q < 3; q++)
        if (own[q]) { PNB_CUDA(cudaMemcpy(user[q], dev[q], sizeof(R) * cnt[q], cudaMemcpyDeviceToHost)); cudaFree(dev[q]); }
    }
    cudaFree(acc);
    if (own_x) cudaFree(const_cast<R *>(dx));
    p->timer_trafo[0] += 1;
    p->timer_trafo[1] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
  }

  // f_hat was zeroed by the caller unless PNFFT_COMPUTE_ACCUMULATED (reference api/api-basic.c:355-356)
  static void adj(P *p, Nd *nd, unsigned cf) {
    const auto t_begin = std::chrono::steady_clock::now();
    const int M = (int)nd->local_M;
    const bool cplx = !p->L.c2r;
    const int NC = cplx ? 2 : 1;
    if ((cf & 1u) && M && !nd->f) fprintf(stderr, "Error: missing memory allocation of nodes->f !!!\n");
    if ((cf & 2u) && M && !nd->grad_f) fprintf(stderr, "Error: missing memory allocation of nodes->grad_f !!!\n");
    bool own_x = false, own_f = false, own_g = false;
    const R *dx = on_device(nd->x, (size_t)3 * M, &own_x);
    const R *df = (cf & 1u) ? on_device(nd->f, (size_t)NC * M, &own_f) : nullptr;
    const R *dg = (cf & 2u) ? on_device(nd->grad_f, (size_t)3 * NC * M, &own_g) : nullptr;
    for (int pid = 0; pid < p->mesh.size; pid++) {
      const DirectBlock B = block_of(p, pid);
      const size_t total = (size_t)B.len[0] * B.len[1] * B.len[2];
      if (!total) continue;
      if (too_big(B)) break;
      const bool mine = pid == p->mesh.rank;
      double *dout = nullptr;
      PNB_CUDA(cudaMalloc((void **)&dout, sizeof(double) * 2 * total));
      const unsigned nb = (unsigned)((total + 127) / 128);
      if (cplx) k_direct_adj<R, true><<<nb, 128, 0, p->stream>>>(B, dx, df, dg, M, dout);
      else k_direct_adj<R, false><<<nb, 128, 0, p->stream>>>(B, dx, df, dg, M, dout);
      PNB_CUDA(cudaGetLastError());
      p->launches++;
      PNB_CUDA(cudaStreamSynchronize(p->stream));
      std::vector<double> h(2 * total), red;
      PNB_CUDA(cudaMemcpy(h.data(), dout, sizeof(double) * 2 * total, cudaMemcpyDeviceToHost));
      cudaFree(dout);
      const double *sum = h.data();
      if (p->mesh.size > 1) {
        red.resize(2 * total);
        MPI_Reduce(h.data(), red.data(), (int)(2 * total), MPI_DOUBLE, MPI_SUM, pid, p->comm);
        sum = red.data();
      }
      if (mine && p->f_hat) {
        std::vector<R> cur(2 * total);
        PNB_CUDA(cudaMemcpy(cur.data(), p->f_hat, sizeof(R) * 2 * total, cudaMemcpyDefault));
        for (size_t i = 0; i < 2 * total; i++) cur[i] = (R)((double)cur[i] + sum[i]);
        PNB_CUDA(cudaMemcpy(p->f_hat, cur.data(), sizeof(R) * 2 * total, cudaMemcpyDefault));
      }
    }
    if (own_x) cudaFree(const_cast<R *>(dx));
    if (own_f) cudaFree(const_cast<R *>(df));
    if (own_g) cudaFree(const_cast<R *>(dg));
    p->timer_adj[0] += 1;
    p->timer_adj[1] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
  }
};

}  // namespace pnb
