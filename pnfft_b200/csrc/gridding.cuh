// The B / B^T matrices of PNFFT on the device: node binning, window-convolution gather (trafo) and
// scatter (adjoint).  Replaces reference kernel/ndft-parallel.c:2521-3009 (trafo_B_ad, adjoint_B_ad and
// their per-node loops) and kernel/assign.c:478-1130 (the (2m+1)^3 inner loops).
//
// Shared pieces of every kernel family: node projection (grid index, interlacing shift), window evaluation by one warp,
// binning into (column tile, z sub-chunk, x offset) bins, the integer parity probes, and the generic kernels
// (kernel_variant 1: one warp per node, run-time m, padded grid in global memory, atomics for the scatter) that serve
// every cutoff the z-marching kernels (zmarch.cuh, zmarch2.cuh) are not instantiated for.
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

// One axis of a node: nx = n x and fl = floor(n x) with a single rounding of the product, exactly like the reference
// (kernel/ndft-parallel.c:2171: pnfft_floor(n[t]*x[t])), and the interior cell fl - local_no_start.
// Second pass of an interlaced plan (reference :2732-2753): x += 0.5/n first; the cell comes from the shifted, NOT yet
// folded coordinate (it may be one past the block: that is the extra ghost cell above), then x >= 0.5 is folded to
// x - 1 with fl - n, and the window arguments fl - n x use the folded pair.
template <class R> __device__ __forceinline__ void node_axis(const GridGeom<R> &g, R x, int t, R *nx, R *fl, int *cell) {
  if (g.il_on) x = (R)((double)x + g.il[t]);
  R v = mul_rn(g.n[t], x);
  R f = m_floor(v);
  *cell = (int)f - g.los[t];   // interior cell == index of tap 0 in the padded array
  if (g.il_on && x >= (R)0.5) { x -= (R)1; f -= g.n[t]; v = mul_rn(g.n[t], x); }
  *nx = v; *fl = f;
}
template <class R> __device__ __forceinline__ void project_node(const GridGeom<R> &g, const R *x3, R *nx, R *fl, int *cell) {
#pragma unroll
  for (int t = 0; t < 3; t++) node_axis(g, x3[t], t, &nx[t], &fl[t], &cell[t]);
}
// the 0.5 of an interlaced plan rides on the x-axis factors (exact: a power of two)
template <class R> __device__ __forceinline__ void warp_scale_x(const GridGeom<R> &g, int lane, R *psi_s, R *dpsi_s, bool want_d) {
  if (g.wscale == (R)1) return;
  for (int v = lane; v < g.cutoff; v += 32) { psi_s[v] *= g.wscale; if (want_d) dpsi_s[v] *= g.wscale; }
}

// PNFFT_PRE_*_PSI: tap s of an axis from its interpolation table (reference pre_tensor_intpol, kernel/ndft-parallel.c:1586-1617,
// stencils kernel/ipnfft.h:450-495); dist = n x - floor(n x) in [0, 1)
template <class R> __device__ __forceinline__ R intpol_tap(const GridGeom<R> &g, const R *__restrict__ tab, int s, R dist) {
  const R dn = dist * (R)g.intpol_num;
  const long long k0 = (long long)m_floor(dn);
  const R d = dn - (R)k0;
  const int o1 = g.intpol_order + 1;
  const R *f = tab + (k0 * g.cutoff + s) * o1;
  switch (g.intpol_order) {
    case 0: return f[0];
    case 1: return f[0] * ((R)1 - d) + f[1] * d;
    case 2: { const R c0 = d + (R)1, c1 = d, c2 = d - (R)1; return (R)0.5 * f[0] * c1 * c2 - c0 * f[1] * c2 + (R)0.5 * c0 * c1 * f[2]; }
    default: {
      const R c0 = d + (R)1, c1 = d, c2 = d - (R)1, c3 = d - (R)2;
      return (-f[0] * c1 * c2 * c3 + (R)3 * c0 * f[1] * c2 * c3 - (R)3 * c0 * c1 * f[2] * c3 + c0 * c1 * c2 * f[3]) / (R)6;
    }
  }
}

// 3*(2m+1) window values (and AD-gradient weights) of one node, computed by one warp into psi_s/dpsi_s.
template <class R>
__device__ __forceinline__ void warp_window_eval(const GridGeom<R> &g, const R *nx, const R *fl, int lane,
                                                 R *psi_s, R *dpsi_s, bool want_d) {
  const int c = g.cutoff;
  if (g.intpol_order >= 0) {
    for (int v = lane; v < 3 * c; v += 32) {
      const int t = v / c, s = v - t * c;
      psi_s[v] = intpol_tap(g, g.intpol_tab[t], s, nx[t] - fl[t]);
      if (want_d) dpsi_s[v] = intpol_tap(g, g.intpol_tab[3 + t], s, nx[t] - fl[t]);
    }
  } else if (g.kind == WIN_BSPLINE) {
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
  if (cell[0] >= 0 && cell[0] < g.lno[0] + g.il_on && cell[1] >= 0 && cell[1] < g.lno[1] + g.il_on && cell[2] >= 0 && cell[2] < g.lno[2] + g.il_on) {
    const int c0 = cell[0] / tg.T[0];
    tile = (c0 * tg.nt[1] + cell[1] / tg.T[1]) * tg.nt[2] + cell[2] / tg.T[2];
    if (tg.sub > 1) tile = tile * tg.sub + (cell[0] - c0 * tg.T[0]);
  }
  tile_of[j] = tile;
  idx[j] = j;
  atomicAdd(&tile_count[tile], 1);
}

// 64-bit content hash of a device array (order sensitive: every word is mixed with its position before the sum), used to
// tell whether device-resident node coordinates changed since they were binned (Core::prepare_nodes)
template <class R> __global__ void k_hash_words(const R *__restrict__ v, long long n, unsigned long long *__restrict__ out) {
  unsigned long long acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long w;
    if (sizeof(R) == 8) w = (unsigned long long)__double_as_longlong((double)v[i]);
    else w = (unsigned long long)__float_as_uint((float)v[i]);
    unsigned long long z = w + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);      // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    acc += z ^ (z >> 31);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// largest node count of a column tile (z-march v2 load-balance hint): bins of one column are contiguous in tile_start
static __global__ void k_max_column(const int *__restrict__ tile_start, int ncol, int bins_per_col, int *__restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  atomicMax(out, tile_start[(size_t)(c + 1) * bins_per_col] - tile_start[(size_t)c * bins_per_col]);
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
  for (int t = 0; t < 3; t++) inside = inside && cell[t] >= 0 && cell[t] < g.lno[t] + g.il_on;
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
  warp_scale_x(g, lane, psi_s, dpsi_s, want_d);
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

// Hessian of the trafo at the nodes with analytic window derivatives (reference kernel/assign.c:881-1027 with the tables
// of kernel/ndft-parallel.c:1956-2105): one warp per node, any cutoff, padded grid in global memory.  Components in the
// reference's order xx, xy, xz, yy, yz, zz.  The windows are always evaluated on the fly (exact formulas).
template <class R, bool CPLX>
__global__ void __launch_bounds__(256) k_hessian_generic(GridGeom<R> g, const R *__restrict__ grid, NodeArgs<R> na, R *__restrict__ hess) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R *scratch = reinterpret_cast<R *>(smem_raw);
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + wib;
  if (p >= na.M) return;
  const int c = g.cutoff;
  R *psi_s = scratch + (size_t)wib * 9 * c, *dpsi_s = psi_s + 3 * c, *ddpsi_s = psi_s + 6 * c;
  const int j = na.perm ? na.perm[p] : p;
  R xs[3] = {na.x[3 * (size_t)j], na.x[3 * (size_t)j + 1], na.x[3 * (size_t)j + 2]}, nx[3], fl[3];
  int cell[3];
  project_node(g, xs, nx, fl, cell);
  bool inside = true;
  for (int t = 0; t < 3; t++) inside = inside && cell[t] >= 0 && cell[t] < g.lno[t] + g.il_on;
  warp_window_eval(g, nx, fl, lane, psi_s, dpsi_s, true);
  __syncwarp();
  for (int v = lane; v < 3 * c; v += 32) {
    const int t = v / c, s = v - t * c;
    const R y = fl[t] - nx[t] - (R)g.m + (R)s;
    if (g.intpol_order >= 0) ddpsi_s[v] = intpol_tap(g, g.intpol_tab[6 + t], s, nx[t] - fl[t]);
    else ddpsi_s[v] = window_ddtap<R>(g.kind, y, g.n[t], g.b[t], g.m, psi_s[v], dpsi_s[v]);
  }
  __syncwarp();
  if (g.wscale != (R)1) for (int v = lane; v < c; v += 32) { psi_s[v] *= g.wscale; dpsi_s[v] *= g.wscale; ddpsi_s[v] *= g.wscale; }
  __syncwarp();
  constexpr int NC = CPLX ? 2 : 1;
  R h[6][2];
  for (int q = 0; q < 6; q++) { h[q][0] = 0; h[q][1] = 0; }
  if (inside) {
    const long long base = (long long)cell[0] * g.pitch0 + (long long)cell[1] * g.pitch1 + cell[2];
    for (int q = lane; q < c * c; q += 32) {
      const int l1 = q / c, l2 = q - l1 * c;
      R s0[2] = {0, 0}, s1[2] = {0, 0}, s2[2] = {0, 0};     // sums over x with psi, dpsi, ddpsi
      const long long off = base + (long long)l1 * g.pitch1 + l2;
      for (int l0 = 0; l0 < c; l0++) {
        const long long i = off + (long long)l0 * g.pitch0;
        R v0, v1 = 0;
        if (CPLX) { v0 = grid[2 * i]; v1 = grid[2 * i + 1]; } else { v0 = grid[i]; }
        const R w = psi_s[l0], dw = dpsi_s[l0], ddw = ddpsi_s[l0];
        s0[0] += w * v0; s0[1] += w * v1; s1[0] += dw * v0; s1[1] += dw * v1; s2[0] += ddw * v0; s2[1] += ddw * v1;
      }
      const R w1 = psi_s[c + l1], d1 = dpsi_s[c + l1], dd1 = ddpsi_s[c + l1];
      const R w2 = psi_s[2 * c + l2], d2 = dpsi_s[2 * c + l2], dd2 = ddpsi_s[2 * c + l2];
      for (int k = 0; k < NC; k++) {
        h[0][k] += w1 * w2 * s2[k];      // xx
        h[1][k] += d1 * w2 * s1[k];      // xy
        h[2][k] += w1 * d2 * s1[k];      // xz
        h[3][k] += dd1 * w2 * s0[k];     // yy
        h[4][k] += d1 * d2 * s0[k];      // yz
        h[5][k] += w1 * dd2 * s0[k];     // zz
      }
    }
  }
  for (int q = 0; q < 6; q++)
    for (int k = 0; k < NC; k++) h[q][k] = warp_sum(h[q][k]);
  if (lane == 0) {
    R *o = hess + (size_t)j * 6 * NC;
    for (int q = 0; q < 6; q++)
      for (int k = 0; k < NC; k++) o[q * NC + k] = na.accumulate ? o[q * NC + k] + h[q][k] : h[q][k];
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
  for (int t = 0; t < 3; t++) if (cell[t] < 0 || cell[t] >= g.lno[t] + g.il_on) return;
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
  warp_scale_x(g, lane, psi_s, dpsi_s, want_d);
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
// mbarrier / TMA helpers and cell arithmetic shared by the z-marching kernels (zmarch.cuh, zmarch2.cuh)
// ------------------------------------------------------------------------------------------------
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

constexpr int kMaxPolyCoef = 25;   // polynomial degree <= 24

}  // namespace pnb
