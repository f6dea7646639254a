// The B / B^T matrices of PNFFT on the device: node binning, window-convolution gather (trafo) and
// scatter (adjoint).  Replaces reference kernel/ndft-parallel.c:2521-3009 (trafo_B_ad, adjoint_B_ad and
// their per-node loops) and kernel/assign.c:478-1130 (the (2m+1)^3 inner loops).
//
// Tiled kernels (default): nodes are binned by T0 x T1 x T2 tiles of the rank's oversampled-grid block.
// One CTA handles one (tile, node-chunk) work item: the tile plus its 2m halo is brought into shared
// memory with ONE TMA tensor load (gather), or accumulated in shared memory and written back with ONE
// TMA reduce-add (scatter: atomic-free inside the tile, the only contended traffic is the bulk
// reduction at L2).  Per node the three 1-d window factors are evaluated once per axis; the x axis is
// contracted in registers (psi_x taps live in registers), lanes sweep the (y,z) face of the stencil.
// Generic kernels (variant 1): run-time m, straight from global memory, atomics for the scatter.
#pragma once
#include <cuda.h>
#include <cub/cub.cuh>

#include "plan.h"

namespace pnb {

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <class R> __device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_xor(v, o);
  return v;
}

// floor(n*x) with a single rounding of the product, exactly like the reference
// (kernel/ndft-parallel.c:2171: pnfft_floor(n[t]*x[t]))
template <class R> __device__ __forceinline__ void project_node(const GridGeom<R> &g, const R *x3, R *nx, R *fl, int *cell) {
#pragma unroll
  for (int t = 0; t < 3; t++) {
    nx[t] = mul_rn(g.n[t], x3[t]);
    fl[t] = m_floor(nx[t]);
    cell[t] = (int)fl[t] - g.los[t];   // interior cell == index of tap 0 in the padded array
  }
}

// 3*(2m+1) window values (and AD-gradient weights) of one node, computed by one warp into psi_s/dpsi_s.
template <class R>
__device__ __forceinline__ void warp_window_eval(const GridGeom<R> &g, const R *nx, const R *fl, int lane,
                                                 R *psi_s, R *dpsi_s, bool want_d) {
  const int c = g.cutoff;
  if (g.kind == WIN_BSPLINE) {
    if (lane < 3) bspline_taps<R>(g.m, nx[lane] - fl[lane], g.n[lane], psi_s + lane * c, want_d ? dpsi_s + lane * c : nullptr);
  } else if (g.kind == WIN_GAUSSIAN && g.fast_gauss) {
    // reference kernel/ndft-parallel.c:1694-1714: exp(-d^2/b) * exp(2d/b)^s * exp_const[s], d = n x - (floor - m)
    for (int v = lane; v < 3 * c; v += 32) {
      const int t = v / c, s = v - t * c;
      const R d = nx[t] - (fl[t] - (R)g.m);
      const R e_sqr = m_exp(-(d * d) / g.b[t]), e_lin = m_exp((R)2 * d / g.b[t]);
      R tmp = e_sqr;
      for (int i = 0; i < s; i++) tmp *= e_lin;
      const R psi = tmp * g.exp_const[t * c + s];
      psi_s[v] = psi;
      if (want_d) dpsi_s[v] = (R)(-2.0) * g.n[t] / g.b[t] * (d - (R)s) * psi;
    }
  } else {
    for (int v = lane; v < 3 * c; v += 32) {
      const int t = v / c, s = v - t * c;
      const R y = fl[t] - nx[t] - (R)g.m + (R)s;
      R psi, dpsi = (R)0;
      window_tap<R>(g.kind, y, g.n[t], g.b[t], g.m, want_d, &psi, &dpsi);
      psi_s[v] = psi;
      if (want_d) dpsi_s[v] = dpsi;
    }
  }
}

// Same 3*(2m+1) values from the per-tap polynomials fitted at plan time (Core::fit_window_polys): tap s of
// axis t is a smooth function of frac = n x - floor(n x) on [0,1); p(u), u = 2 frac - 1, reproduces it to
// double rounding; the AD-gradient weights dpsi_s have their own polynomials, fitted to the exact derivative formulas
// (differentiating the interpolant would amplify its error by deg^2).
// Nodes sitting exactly on a grid line (frac == 0) take the exact path: compactly supported windows jump there.
template <class R>
__device__ __forceinline__ void warp_window_eval_poly(const GridGeom<R> &g, const R *poly, const R *nx, const R *fl, int lane,
                                                      R *psi_s, R *dpsi_s, bool want_d) {
  const int c = g.cutoff, nv = 3 * c;
  const R fr[3] = {nx[0] - fl[0], nx[1] - fl[1], nx[2] - fl[2]};
  if (fr[0] == (R)0 || fr[1] == (R)0 || fr[2] == (R)0) { warp_window_eval(g, nx, fl, lane, psi_s, dpsi_s, want_d); return; }
  for (int v = lane; v < nv; v += 32) {
    const int t = v / c;
    const R u = (R)2 * fr[t] - (R)1;
    const R *a = poly + v;
    R p = a[g.poly_deg * nv];
    if (want_d) {
      const R *ad = a + (g.poly_deg + 1) * nv;      // the derivative weights have their own fitted polynomials
      R dp = ad[g.poly_deg * nv];
      for (int k = g.poly_deg - 1; k >= 0; k--) { dp = dp * u + ad[k * nv]; p = p * u + a[k * nv]; }
      psi_s[v] = p;
      dpsi_s[v] = dp;
    } else {
      for (int k = g.poly_deg - 1; k >= 0; k--) p = p * u + a[k * nv];
      psi_s[v] = p;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// binning
// ------------------------------------------------------------------------------------------------
struct TileGeom {
  int T[3];        // tile extent in cells
  int nt[3];       // tiles per axis
  int ntiles;
  int chunk;       // max nodes per work item
  int family;      // kernel family the geometry was made for (Core::kernel_family)
  int sub;         // bins per tile: 1, or T[0] when nodes are also ordered by their x offset inside the tile (z-march v2)
};

template <class R>
__global__ void k_bin_nodes(GridGeom<R> g, TileGeom tg, const R *__restrict__ x, int M, int *__restrict__ tile_of,
                            int *__restrict__ idx, int *__restrict__ tile_count) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  R xs[3] = {x[3 * (size_t)j], x[3 * (size_t)j + 1], x[3 * (size_t)j + 2]}, nx[3], fl[3];
  int cell[3];
  project_node(g, xs, nx, fl, cell);
  int tile = tg.ntiles * tg.sub;  // nodes outside the rank's block (undefined behaviour in the reference) are skipped
  if (cell[0] >= 0 && cell[0] < g.lno[0] && cell[1] >= 0 && cell[1] < g.lno[1] && cell[2] >= 0 && cell[2] < g.lno[2]) {
    const int c0 = cell[0] / tg.T[0];
    tile = (c0 * tg.nt[1] + cell[1] / tg.T[1]) * tg.nt[2] + cell[2] / tg.T[2];
    if (tg.sub > 1) tile = tile * tg.sub + (cell[0] - c0 * tg.T[0]);
  }
  tile_of[j] = tile;
  idx[j] = j;
  atomicAdd(&tile_count[tile], 1);
}

// largest node count of a column tile (z-march v2 load-balance hint): bins of one column are contiguous in tile_start
static __global__ void k_max_column(const int *__restrict__ tile_start, int ncol, int bins_per_col, int *__restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  atomicMax(out, tile_start[(size_t)(c + 1) * bins_per_col] - tile_start[(size_t)c * bins_per_col]);
}

static __global__ void k_items_per_tile(TileGeom tg, const int *__restrict__ tile_count, int *__restrict__ n_items) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > tg.ntiles) return;
  n_items[t] = t < tg.ntiles ? (tile_count[t] + tg.chunk - 1) / tg.chunk : 0;
}

static __global__ void k_fill_items(TileGeom tg, const int *__restrict__ tile_count, const int *__restrict__ tile_start,
                             const int *__restrict__ item_start, int *__restrict__ items, int *__restrict__ nitems) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= tg.ntiles) return;
  const int cnt = tile_count[t], s = tile_start[t];
  int it = item_start[t];
  for (int b = 0; b < cnt; b += tg.chunk, it++) {
    items[3 * it] = t;
    items[3 * it + 1] = s + b;
    items[3 * it + 2] = s + min(cnt, b + tg.chunk);
  }
  if (t == tg.ntiles - 1) *nitems = it;
}

// ------------------------------------------------------------------------------------------------
// parity probes (integer work, bit-exact against the reference)
// ------------------------------------------------------------------------------------------------
template <class R>
__global__ void k_node_grid_index(GridGeom<R> g, const R *__restrict__ x, int M, long long *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  R xs[3] = {x[3 * (size_t)j], x[3 * (size_t)j + 1], x[3 * (size_t)j + 2]}, nx[3], fl[3];
  int cell[3];
  project_node(g, xs, nx, fl, cell);
  // u_j = floor(n x) - m - local_no_start + gcells_below (= m)   (reference :1563-1572)
  const long long u0 = cell[0], u1 = cell[1], u2 = cell[2];
  out[4 * (size_t)j] = u0; out[4 * (size_t)j + 1] = u1; out[4 * (size_t)j + 2] = u2;
  out[4 * (size_t)j + 3] = u2 + (long long)g.ngc[2] * (u1 + (long long)g.ngc[1] * u0);  // PNFFT_PLAIN_INDEX_3D, ipnfft.h:66
}

// sort key of reference kernel/ndft-parallel.c:2131-2142: row-major index of ((floor(n x - m)) mod n) in the global grid
template <class R>
__global__ void k_sort_keys(GridGeom<R> g, long long n0, long long n1, long long n2, const R *__restrict__ x, int M,
                            unsigned long long *__restrict__ keys, int *__restrict__ idx) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const long long n[3] = {n0, n1, n2};
  unsigned long long key = 0;
#pragma unroll
  for (int t = 0; t < 3; t++) {
    const R v = add_rn(mul_rn(g.n[t], x[3 * (size_t)j + t]), -(R)g.m);
    const long long help = (long long)m_floor(v);
    const long long u = ((help % n[t]) + n[t]) % n[t];
    key += (unsigned long long)u;
    if (t + 1 < 3) key *= (unsigned long long)n[t + 1];
  }
  keys[j] = key;
  idx[j] = j;
}

template <class R>
__global__ void k_window_tensor(GridGeom<R> g, const R *__restrict__ x, int M, R *__restrict__ psi, R *__restrict__ dpsi) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  R xs[3] = {x[3 * (size_t)warp], x[3 * (size_t)warp + 1], x[3 * (size_t)warp + 2]}, nx[3], fl[3];
  int cell[3];
  project_node(g, xs, nx, fl, cell);
  if (g.poly) warp_window_eval_poly(g, g.poly, nx, fl, lane, psi + (size_t)warp * 3 * g.cutoff, dpsi ? dpsi + (size_t)warp * 3 * g.cutoff : nullptr, dpsi != nullptr);
  else warp_window_eval(g, nx, fl, lane, psi + (size_t)warp * 3 * g.cutoff, dpsi ? dpsi + (size_t)warp * 3 * g.cutoff : nullptr, dpsi != nullptr);
}

// PNFFT_PRE_PSI tables, stored at the SORTED position p like the reference does (:1222-1240)
template <class R>
__global__ void k_precompute_psi(GridGeom<R> g, const R *__restrict__ x, const int *__restrict__ perm, int M,
                                 R *__restrict__ psi, R *__restrict__ dpsi) {
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (p >= M) return;
  const int j = perm[p];
  R xs[3] = {x[3 * (size_t)j], x[3 * (size_t)j + 1], x[3 * (size_t)j + 2]}, nx[3], fl[3];
  int cell[3];
  project_node(g, xs, nx, fl, cell);
  warp_window_eval(g, nx, fl, lane, psi + (size_t)p * 3 * g.cutoff, dpsi ? dpsi + (size_t)p * 3 * g.cutoff : nullptr, dpsi != nullptr);
}

// ------------------------------------------------------------------------------------------------
// node-side argument block shared by all gather / scatter kernels
// ------------------------------------------------------------------------------------------------
template <class R> struct NodeArgs {
  const R *x;
  const int *perm;       // sorted position -> node index (nullptr: identity)
  int M;
  R *f;                  // complex (2 R) or real per node; nullptr: not computed / not spread
  long long f_stride, f_off;   // element index = j*f_stride + f_off (ik-differentiation writes grad components through f)
  R *grad;               // 3 per node; nullptr: no gradient
  int accumulate;        // gather: add to the output instead of overwriting (PNFFT_COMPUTE_ACCUMULATED)
  const R *pre_psi, *pre_dpsi;  // PNFFT_PRE_PSI tables (sorted order) or nullptr
};

// ------------------------------------------------------------------------------------------------
// generic kernels: one warp per node, run-time m, padded grid in global memory
// ------------------------------------------------------------------------------------------------
template <class R, bool CPLX>
__global__ void __launch_bounds__(256) k_gather_generic(GridGeom<R> g, const R *__restrict__ grid, NodeArgs<R> na) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R *scratch = reinterpret_cast<R *>(smem_raw);
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + wib;
  if (p >= na.M) return;
  const int c = g.cutoff;
  R *psi_s = scratch + (size_t)wib * 6 * c, *dpsi_s = psi_s + 3 * c;
  const int j = na.perm ? na.perm[p] : p;
  const bool want_d = na.grad != nullptr;
  R xs[3] = {na.x[3 * (size_t)j], na.x[3 * (size_t)j + 1], na.x[3 * (size_t)j + 2]}, nx[3], fl[3];
  int cell[3];
  project_node(g, xs, nx, fl, cell);
  bool inside = true;
  for (int t = 0; t < 3; t++) inside = inside && cell[t] >= 0 && cell[t] < g.lno[t];
  if (na.pre_psi) {
    for (int v = lane; v < 3 * c; v += 32) {
      psi_s[v] = na.pre_psi[(size_t)p * 3 * c + v];
      if (want_d) dpsi_s[v] = na.pre_dpsi[(size_t)p * 3 * c + v];
    }
  } else {
    if (g.poly) warp_window_eval_poly(g, g.poly, nx, fl, lane, psi_s, dpsi_s, want_d);
    else warp_window_eval(g, nx, fl, lane, psi_s, dpsi_s, want_d);
  }
  __syncwarp();
  R af[2] = {0, 0}, a0[2] = {0, 0}, a1[2] = {0, 0}, a2[2] = {0, 0};
  if (inside) {
    const long long base = (long long)cell[0] * g.pitch0 + (long long)cell[1] * g.pitch1 + cell[2];
    for (int q = lane; q < c * c; q += 32) {
      const int l1 = q / c, l2 = q - l1 * c;
      const R w1 = psi_s[c + l1], w2 = psi_s[2 * c + l2];
      R s[2] = {0, 0}, d[2] = {0, 0};
      const long long off = base + (long long)l1 * g.pitch1 + l2;
      for (int l0 = 0; l0 < c; l0++) {
        const long long i = off + (long long)l0 * g.pitch0;
        R v0, v1 = 0;
        if (CPLX) { v0 = grid[2 * i]; v1 = grid[2 * i + 1]; } else { v0 = grid[i]; }
        const R w0 = psi_s[l0];
        s[0] += w0 * v0; s[1] += w0 * v1;
        if (want_d) { const R dw0 = dpsi_s[l0]; d[0] += dw0 * v0; d[1] += dw0 * v1; }
      }
      const R w12 = w1 * w2;
      af[0] += w12 * s[0]; af[1] += w12 * s[1];
      if (want_d) {
        const R dw1 = dpsi_s[c + l1], dw2 = dpsi_s[2 * c + l2];
        a0[0] += w12 * d[0]; a0[1] += w12 * d[1];
        const R w1d = dw1 * w2, w2d = w1 * dw2;
        a1[0] += w1d * s[0]; a1[1] += w1d * s[1];
        a2[0] += w2d * s[0]; a2[1] += w2d * s[1];
      }
    }
  }
  constexpr int NC = CPLX ? 2 : 1;
#pragma unroll
  for (int k = 0; k < NC; k++) {
    af[k] = warp_sum(af[k]);
    if (want_d) { a0[k] = warp_sum(a0[k]); a1[k] = warp_sum(a1[k]); a2[k] = warp_sum(a2[k]); }
  }
  if (lane == 0) {
    if (na.f) {
      R *o = na.f + ((size_t)j * na.f_stride + na.f_off) * NC;
      for (int k = 0; k < NC; k++) o[k] = na.accumulate ? o[k] + af[k] : af[k];
    }
    if (na.grad) {
      R *o = na.grad + (size_t)j * 3 * NC;
      for (int k = 0; k < NC; k++) {
        o[k] = na.accumulate ? o[k] + a0[k] : a0[k];
        o[NC + k] = na.accumulate ? o[NC + k] + a1[k] : a1[k];
        o[2 * NC + k] = na.accumulate ? o[2 * NC + k] + a2[k] : a2[k];
      }
    }
  }
}

template <class R, bool CPLX>
__global__ void __launch_bounds__(256) k_scatter_generic(GridGeom<R> g, R *__restrict__ grid, NodeArgs<R> na) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R *scratch = reinterpret_cast<R *>(smem_raw);
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + wib;
  if (p >= na.M) return;
  const int c = g.cutoff;
  R *psi_s = scratch + (size_t)wib * 6 * c, *dpsi_s = psi_s + 3 * c;
  const int j = na.perm ? na.perm[p] : p;
  const bool want_d = na.grad != nullptr;
  R xs[3] = {na.x[3 * (size_t)j], na.x[3 * (size_t)j + 1], na.x[3 * (size_t)j + 2]}, nx[3], fl[3];
  int cell[3];
  project_node(g, xs, nx, fl, cell);
  for (int t = 0; t < 3; t++) if (cell[t] < 0 || cell[t] >= g.lno[t]) return;
  if (na.pre_psi) {
    for (int v = lane; v < 3 * c; v += 32) {
      psi_s[v] = na.pre_psi[(size_t)p * 3 * c + v];
      if (want_d) dpsi_s[v] = na.pre_dpsi[(size_t)p * 3 * c + v];
    }
  } else {
    if (g.poly) warp_window_eval_poly(g, g.poly, nx, fl, lane, psi_s, dpsi_s, want_d);
    else warp_window_eval(g, nx, fl, lane, psi_s, dpsi_s, want_d);
  }
  __syncwarp();
  constexpr int NC = CPLX ? 2 : 1;
  R fv[2] = {0, 0}, g0[2] = {0, 0}, g1[2] = {0, 0}, g2[2] = {0, 0};
  if (na.f) for (int k = 0; k < NC; k++) fv[k] = na.f[((size_t)j * na.f_stride + na.f_off) * NC + k];
  if (na.grad)
    for (int k = 0; k < NC; k++) {
      g0[k] = na.grad[(size_t)j * 3 * NC + k];
      g1[k] = na.grad[(size_t)j * 3 * NC + NC + k];
      g2[k] = na.grad[(size_t)j * 3 * NC + 2 * NC + k];
    }
  const long long base = (long long)cell[0] * g.pitch0 + (long long)cell[1] * g.pitch1 + cell[2];
  for (int q = lane; q < c * c; q += 32) {
    const int l1 = q / c, l2 = q - l1 * c;
    const R w1 = psi_s[c + l1], w2 = psi_s[2 * c + l2];
    const R w12 = w1 * w2;
    R w1d = 0, w2d = 0;
    if (want_d) { w1d = dpsi_s[c + l1] * w2; w2d = w1 * dpsi_s[2 * c + l2]; }
    const long long off = base + (long long)l1 * g.pitch1 + l2;
    for (int l0 = 0; l0 < c; l0++) {
      const long long i = off + (long long)l0 * g.pitch0;
      const R w0 = psi_s[l0];
      R val[2];
      for (int k = 0; k < NC; k++) val[k] = (w0 * w12) * fv[k];
      if (want_d) {
        const R dw0 = dpsi_s[l0];
        for (int k = 0; k < NC; k++) val[k] += (dw0 * w12) * g0[k] + (w0 * w1d) * g1[k] + (w0 * w2d) * g2[k];
      }
      if (CPLX) { atomicAdd(&grid[2 * i], val[0]); atomicAdd(&grid[2 * i + 1], val[1]); }
      else atomicAdd(&grid[i], val[0]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tiled kernels
// ------------------------------------------------------------------------------------------------
template <int M_, int CELLB> struct TileCfg {
  static constexpr int C = 2 * M_ + 1;
  // tile extents (cells); chosen so that (T+2m)^3-ish box of 16-byte cells stays below ~200 kB
  static constexpr int T0 = (M_ <= 6) ? 8 : 6;
  static constexpr int T1 = (M_ <= 6) ? 8 : 6;
  static constexpr int T2 = (M_ <= 6) ? 16 : 8;
  static constexpr int BX = T0 + 2 * M_, BY = T1 + 2 * M_;
  // z pitch of the shared-memory box: lanes sweep the (y,z) face in flattened order q = l1*C + l2; with
  // BZ == C (mod 128/CELLB) the address of lane q is == q (mod one 128-byte wavefront) => conflict free.
  // 8- and 4-byte cells additionally need BZ*CELLB % 16 == 0 (TMA), which costs a 2-way conflict at row ends.
  static constexpr int W = 128 / CELLB;
  static constexpr int BZ0 = T2 + 2 * M_;
  static constexpr int ALIGN = (CELLB >= 16) ? 1 : 16 / CELLB;
  static constexpr int bz() {
    int z = BZ0;
    if (ALIGN == 1) { while (z % W != C % W) z++; }
    else { while (z % ALIGN != 0 || ((z % W) != ((C + 1) % W) && (z % W) != ((C + W - 1) % W))) z++; }
    return z;
  }
  static constexpr int BZ = bz();
  static constexpr int BOX_CELLS = BX * BY * BZ;
  static constexpr int BOX_BYTES = BOX_CELLS * CELLB;
  static constexpr int NCH = (C * C + 31) / 32;
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const void *src, const CUtensorMap *tm, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
               ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)) : "memory");
}

template <class R, bool CPLX> struct CellT;
template <> struct CellT<double, true> { typedef double2 type; };
template <> struct CellT<double, false> { typedef double type; };
template <> struct CellT<float, true> { typedef float2 type; };
template <> struct CellT<float, false> { typedef float type; };

__device__ __forceinline__ void fma_cell(double2 &a, double w, const double2 &v) { a.x = fma(w, v.x, a.x); a.y = fma(w, v.y, a.y); }
__device__ __forceinline__ void fma_cell(float2 &a, float w, const float2 &v) { a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); }
__device__ __forceinline__ void fma_cell(double &a, double w, const double &v) { a = fma(w, v, a); }
__device__ __forceinline__ void fma_cell(float &a, float w, const float &v) { a = fmaf(w, v, a); }
__device__ __forceinline__ void zero_cell(double2 &a) { a.x = 0; a.y = 0; }
__device__ __forceinline__ void zero_cell(float2 &a) { a.x = 0; a.y = 0; }
__device__ __forceinline__ void zero_cell(double &a) { a = 0; }
__device__ __forceinline__ void zero_cell(float &a) { a = 0; }
__device__ __forceinline__ double2 cell_sum(double2 a) { a.x = warp_sum(a.x); a.y = warp_sum(a.y); return a; }
__device__ __forceinline__ float2 cell_sum(float2 a) { a.x = warp_sum(a.x); a.y = warp_sum(a.y); return a; }
__device__ __forceinline__ double cell_sum(double a) { return warp_sum(a); }
__device__ __forceinline__ float cell_sum(float a) { return warp_sum(a); }
__device__ __forceinline__ void store_out(double *o, double2 v, int acc) { if (acc) { v.x += o[0]; v.y += o[1]; } o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ void store_out(float *o, float2 v, int acc) { if (acc) { v.x += o[0]; v.y += o[1]; } o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ void store_out(double *o, double v, int acc) { o[0] = acc ? o[0] + v : v; }
__device__ __forceinline__ void store_out(float *o, float v, int acc) { o[0] = acc ? o[0] + v : v; }
__device__ __forceinline__ double2 load_in(const double *p, double2) { return make_double2(p[0], p[1]); }
__device__ __forceinline__ float2 load_in(const float *p, float2) { return make_float2(p[0], p[1]); }
__device__ __forceinline__ double load_in(const double *p, double) { return p[0]; }
__device__ __forceinline__ float load_in(const float *p, float) { return p[0]; }
__device__ __forceinline__ double2 scale_cell(double w, double2 v) { return make_double2(w * v.x, w * v.y); }
__device__ __forceinline__ float2 scale_cell(float w, float2 v) { return make_float2(w * v.x, w * v.y); }
__device__ __forceinline__ double scale_cell(double w, double v) { return w * v; }
__device__ __forceinline__ float scale_cell(float w, float v) { return w * v; }

constexpr int kGatherWarps = 16;
constexpr int kMaxPolyCoef = 25;   // polynomial degree <= 24

template <class R, bool CPLX, int M_, bool GRAD>
__global__ void __launch_bounds__(kGatherWarps * 32, 1)
k_gather_tiled(const __grid_constant__ CUtensorMap tmap, GridGeom<R> g, TileGeom tg, const R *__restrict__ /*unused*/,
               NodeArgs<R> na, const int *__restrict__ items, const int *__restrict__ nitems) {
  typedef typename CellT<R, CPLX>::type Cell;
  typedef TileCfg<M_, (int)sizeof(Cell)> Cfg;
  constexpr int C = Cfg::C, BY = Cfg::BY, BZ = Cfg::BZ, NCH = Cfg::NCH;
  constexpr int NCOMP = CPLX ? 2 : 1;
  if ((int)blockIdx.x >= *nitems) return;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cell *box = reinterpret_cast<Cell *>(smem_raw);
  R *scratch = reinterpret_cast<R *>(smem_raw + Cfg::BOX_BYTES);
  R *poly_s = scratch + kGatherWarps * 6 * C;
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(poly_s + 2 * kMaxPolyCoef * 3 * C);

  const int tile = items[3 * blockIdx.x], begin = items[3 * blockIdx.x + 1], end = items[3 * blockIdx.x + 2];
  const int tz = tile % tg.nt[2], ty = (tile / tg.nt[2]) % tg.nt[1], tx = tile / (tg.nt[2] * tg.nt[1]);
  const int o0 = tx * Cfg::T0, o1 = ty * Cfg::T1, o2 = tz * Cfg::T2;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, (unsigned)Cfg::BOX_BYTES);
    tma_load_3d(box, &tmap, o2 * NCOMP, o1, o0, bar);
  }
  if (g.poly) for (int i = threadIdx.x; i < 2 * (g.poly_deg + 1) * 3 * C; i += kGatherWarps * 32) poly_s[i] = g.poly[i];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  R *psi_s = scratch + warp * 6 * C, *dpsi_s = psi_s + 3 * C;

  // lane -> (l1,l2) of the stencil's (y,z) face, per chunk; node independent
  int off[NCH], l1s[NCH], l2s[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ch++) {
    int q = ch * 32 + lane;
    if (q >= C * C) q = C * C - 1;
    l1s[ch] = q / C; l2s[ch] = q - l1s[ch] * C;
    off[ch] = l1s[ch] * BZ + l2s[ch];
  }

  bool waited = false;
  for (int p = begin + warp; p < end; p += kGatherWarps) {
    const int j = na.perm[p];
    R xs[3] = {na.x[3 * (size_t)j], na.x[3 * (size_t)j + 1], na.x[3 * (size_t)j + 2]}, nx[3], fl[3];
    int cell[3];
    project_node(g, xs, nx, fl, cell);
    __syncwarp();
    if (na.pre_psi) {
      for (int v = lane; v < 3 * C; v += 32) {
        psi_s[v] = na.pre_psi[(size_t)p * 3 * C + v];
        if (GRAD) dpsi_s[v] = na.pre_dpsi[(size_t)p * 3 * C + v];
      }
    } else {
      if (g.poly) warp_window_eval_poly(g, poly_s, nx, fl, lane, psi_s, dpsi_s, GRAD);
      else warp_window_eval(g, nx, fl, lane, psi_s, dpsi_s, GRAD);
    }
    __syncwarp();
    if (!waited) { mbar_wait(bar, 0); waited = true; }

    R w0[C], dw0[GRAD ? C : 1];
#pragma unroll
    for (int l0 = 0; l0 < C; l0++) { w0[l0] = psi_s[l0]; if (GRAD) dw0[l0] = dpsi_s[l0]; }

    const int base = ((cell[0] - o0) * BY + (cell[1] - o1)) * BZ + (cell[2] - o2);
    Cell af, a0, a1, a2;
    zero_cell(af); zero_cell(a0); zero_cell(a1); zero_cell(a2);
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
      const bool active = (ch * 32 + lane) < C * C;
      const Cell *ptr = box + base + off[ch];
      Cell s, d;
      zero_cell(s); zero_cell(d);
#pragma unroll
      for (int l0 = 0; l0 < C; l0++) {
        const Cell v = ptr[l0 * BY * BZ];
        fma_cell(s, w0[l0], v);
        if (GRAD) fma_cell(d, dw0[l0], v);
      }
      const R w1 = psi_s[C + l1s[ch]], w2 = psi_s[2 * C + l2s[ch]];
      const R w12 = active ? w1 * w2 : (R)0;
      fma_cell(af, w12, s);
      if (GRAD) {
        const R dw1 = dpsi_s[C + l1s[ch]], dw2 = dpsi_s[2 * C + l2s[ch]];
        fma_cell(a0, w12, d);
        fma_cell(a1, active ? dw1 * w2 : (R)0, s);
        fma_cell(a2, active ? w1 * dw2 : (R)0, s);
      }
    }
    af = cell_sum(af);
    if (GRAD) { a0 = cell_sum(a0); a1 = cell_sum(a1); a2 = cell_sum(a2); }
    if (lane == 0) {
      if (na.f) store_out(na.f + ((size_t)j * na.f_stride + na.f_off) * NCOMP, af, na.accumulate);
      if (GRAD) {
        R *o = na.grad + (size_t)j * 3 * NCOMP;
        store_out(o, a0, na.accumulate);
        store_out(o + NCOMP, a1, na.accumulate);
        store_out(o + 2 * NCOMP, a2, na.accumulate);
      }
    }
  }
  if (!waited) mbar_wait(bar, 0);   // never leave with the bulk copy in flight
}

// scatter: C warps, warp w owns the box planes X with X == w (mod C); every node touches exactly one
// owned plane per warp, so the read-modify-writes of different warps never meet.
template <int M_> struct ScatterCfg { static constexpr int NB = (M_ <= 6) ? 32 : 16; };

template <class R, bool CPLX, int M_, bool GRAD>
__global__ void __launch_bounds__((2 * M_ + 1) * 32, 1)
k_scatter_tiled(const __grid_constant__ CUtensorMap tmap, GridGeom<R> g, TileGeom tg, NodeArgs<R> na,
                const int *__restrict__ items, const int *__restrict__ nitems) {
  typedef typename CellT<R, CPLX>::type Cell;
  typedef TileCfg<M_, (int)sizeof(Cell)> Cfg;
  constexpr int C = Cfg::C, BY = Cfg::BY, BZ = Cfg::BZ, NCH = Cfg::NCH;
  constexpr int NCOMP = CPLX ? 2 : 1;
  constexpr int NT = C * 32;
  constexpr int NB = ScatterCfg<M_>::NB;
  constexpr int WPN = GRAD ? 6 * C : 3 * C;   // weights per node
  if ((int)blockIdx.x >= *nitems) return;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cell *box = reinterpret_cast<Cell *>(smem_raw);
  R *wts = reinterpret_cast<R *>(smem_raw + Cfg::BOX_BYTES);           // [NB][WPN]
  Cell *vals = reinterpret_cast<Cell *>(wts + NB * WPN);                 // [NB][4]: f, g0, g1, g2
  int *hdr = reinterpret_cast<int *>(vals + NB * 4);                     // [NB][2]: box offset of tap (0,0,0); ux
  R *poly_s = reinterpret_cast<R *>(hdr + NB * 2);

  const int tile = items[3 * blockIdx.x], begin = items[3 * blockIdx.x + 1], end = items[3 * blockIdx.x + 2];
  const int tz = tile % tg.nt[2], ty = (tile / tg.nt[2]) % tg.nt[1], tx = tile / (tg.nt[2] * tg.nt[1]);
  const int o0 = tx * Cfg::T0, o1 = ty * Cfg::T1, o2 = tz * Cfg::T2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  {
    Cell z; zero_cell(z);
    for (int i = threadIdx.x; i < Cfg::BOX_CELLS; i += NT) box[i] = z;
    if (g.poly) for (int i = threadIdx.x; i < 2 * (g.poly_deg + 1) * 3 * C; i += NT) poly_s[i] = g.poly[i];
  }
  int off[NCH], l1s[NCH], l2s[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ch++) {
    int q = ch * 32 + lane;
    if (q >= C * C) q = C * C - 1;
    l1s[ch] = q / C; l2s[ch] = q - l1s[ch] * C;
    off[ch] = l1s[ch] * BZ + l2s[ch];
  }

  for (int b0 = begin; b0 < end; b0 += NB) {
    const int nb = min(NB, end - b0);
    __syncthreads();   // previous batch fully consumed (and the zero fill is complete)
    // ---- phase A: headers and window weights of the batch ----
    if (threadIdx.x < nb) {
      const int i = threadIdx.x, j = na.perm[b0 + i];
      R xs[3] = {na.x[3 * (size_t)j], na.x[3 * (size_t)j + 1], na.x[3 * (size_t)j + 2]}, nx[3], fl[3];
      int cell[3];
      project_node(g, xs, nx, fl, cell);
      hdr[2 * i] = ((cell[0] - o0) * BY + (cell[1] - o1)) * BZ + (cell[2] - o2);
      hdr[2 * i + 1] = cell[0] - o0;
      Cell z; zero_cell(z);
      vals[4 * i] = na.f ? load_in(na.f + ((size_t)j * na.f_stride + na.f_off) * NCOMP, z) : z;
      if (GRAD) {
        const R *gp = na.grad + (size_t)j * 3 * NCOMP;
        vals[4 * i + 1] = load_in(gp, z); vals[4 * i + 2] = load_in(gp + NCOMP, z); vals[4 * i + 3] = load_in(gp + 2 * NCOMP, z);
      }
    }
    if (na.pre_psi) {
      for (int v = threadIdx.x; v < nb * 3 * C; v += NT) {
        const int i = v / (3 * C), r = v - i * 3 * C;
        wts[i * WPN + r] = na.pre_psi[(size_t)(b0 + i) * 3 * C + r];
        if (GRAD) wts[i * WPN + 3 * C + r] = na.pre_dpsi[(size_t)(b0 + i) * 3 * C + r];
      }
    } else if (g.kind == WIN_BSPLINE && !g.poly) {
      for (int v = threadIdx.x; v < nb * 3; v += NT) {
        const int i = v / 3, t = v - i * 3, j = na.perm[b0 + i];
        const R nxv = mul_rn(g.n[t], na.x[3 * (size_t)j + t]);
        const R frac = nxv - m_floor(nxv);
        bspline_taps<R>(M_, frac, g.n[t], wts + i * WPN + t * C, GRAD ? wts + i * WPN + 3 * C + t * C : nullptr);
      }
    } else {
      for (int v = threadIdx.x; v < nb * 3 * C; v += NT) {
        const int i = v / (3 * C), r = v - i * 3 * C, t = r / C, s = r - t * C, j = na.perm[b0 + i];
        const R nxv = mul_rn(g.n[t], na.x[3 * (size_t)j + t]);
        const R flv = m_floor(nxv);
        R psi, dpsi = (R)0;
        const R fr = nxv - flv;
        if (g.poly && fr != (R)0) {
          const R u = (R)2 * fr - (R)1;
          const R *a = poly_s + r;
          const R *ad = a + (g.poly_deg + 1) * 3 * C;
          psi = a[g.poly_deg * 3 * C];
          if (GRAD) dpsi = ad[g.poly_deg * 3 * C];
          for (int k = g.poly_deg - 1; k >= 0; k--) { if (GRAD) dpsi = dpsi * u + ad[k * 3 * C]; psi = psi * u + a[k * 3 * C]; }
        } else if (g.kind == WIN_GAUSSIAN && g.fast_gauss) {
          const R d = nxv - (flv - (R)M_);
          const R e_sqr = m_exp(-(d * d) / g.b[t]), e_lin = m_exp((R)2 * d / g.b[t]);
          R tmp = e_sqr;
          for (int k = 0; k < s; k++) tmp *= e_lin;
          psi = tmp * g.exp_const[t * C + s];
          dpsi = (R)(-2.0) * g.n[t] / g.b[t] * (d - (R)s) * psi;
        } else {
          window_tap<R>(g.kind, flv - nxv - (R)M_ + (R)s, g.n[t], g.b[t], M_, GRAD, &psi, &dpsi);
        }
        wts[i * WPN + r] = psi;
        if (GRAD) wts[i * WPN + 3 * C + r] = dpsi;
      }
    }
    __syncthreads();
    // ---- phase B: every warp adds its plane of every node ----
    for (int i = 0; i < nb; i++) {
      const R *w = wts + i * WPN;
      const int ux = hdr[2 * i + 1];
      int l0 = warp - ux % C;
      if (l0 < 0) l0 += C;
      Cell *plane = box + hdr[2 * i] + l0 * BY * BZ;
      const R wx = w[l0];
      Cell P = scale_cell(wx, vals[4 * i]), Q1, Q2;
      if (GRAD) {
        fma_cell(P, w[3 * C + l0], vals[4 * i + 1]);
        Q1 = scale_cell(wx, vals[4 * i + 2]);
        Q2 = scale_cell(wx, vals[4 * i + 3]);
      }
#pragma unroll
      for (int ch = 0; ch < NCH; ch++) {
        if ((ch * 32 + lane) < C * C) {
          const R w1 = w[C + l1s[ch]], w2 = w[2 * C + l2s[ch]];
          Cell v = plane[off[ch]];
          fma_cell(v, w1 * w2, P);
          if (GRAD) {
            const R dw1 = w[4 * C + l1s[ch]], dw2 = w[5 * C + l2s[ch]];
            fma_cell(v, dw1 * w2, Q1);
            fma_cell(v, w1 * dw2, Q2);
          }
          plane[off[ch]] = v;
        }
      }
    }
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_reduce_add_3d(box, &tmap, o2 * NCOMP, o1, o0);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

template <class R, bool CPLX, int M_, bool GRAD> struct TiledSmem {
  typedef typename CellT<R, CPLX>::type Cell;
  typedef TileCfg<M_, (int)sizeof(Cell)> Cfg;
  static constexpr size_t poly = (size_t)2 * kMaxPolyCoef * 3 * Cfg::C * sizeof(R);   // psi and dpsi polynomials
  static constexpr size_t gather = (size_t)Cfg::BOX_BYTES + (size_t)kGatherWarps * 6 * Cfg::C * sizeof(R) + poly + 16;
  static constexpr int NB = ScatterCfg<M_>::NB;
  static constexpr size_t scatter = (size_t)Cfg::BOX_BYTES + (size_t)NB * (GRAD ? 6 : 3) * Cfg::C * sizeof(R) +
                                    (size_t)NB * 4 * sizeof(Cell) + (size_t)NB * 2 * sizeof(int) + poly + 16;
  static_assert(gather <= 232448 && scatter <= 232448, "shared-memory budget of one CTA exceeded");
};

}  // namespace pnb
