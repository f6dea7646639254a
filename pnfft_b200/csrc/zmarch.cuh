// Z-marching gridding kernels: the default B / B^T path (reference kernel/assign.c:478-1130 inner loops,
// kernel/ndft-parallel.c:2703-3009 node loops).
//
// Why: a (2m+1)^3 stencil read straight from shared memory moves 16 B per complex FMA pair; at 128 B/clk/SM that
// caps an F-only gather at 25 % of the FP64 pipe (64 DFMA/clk/SM) and a read-modify-write scatter at 12.5 %.  The
// only way past that wall is to keep grid cells in REGISTERS and reuse them across nodes.
//
// How: one CTA owns a column tile of T0 x T1 cells (x, y) and marches along z.  Thread r owns the grid row
// (r0, r1) of the tile's (T0+2m) x (T1+2m) footprint and holds a sliding window of W = ZS+2m consecutive z cells
// of that row in registers.  Nodes are binned by (column tile, z sub-chunk of ZS cells); for a node with z offset
// d inside the sub-chunk the 2m+1 z taps hit window slots d .. d+2m, selected by a warp-uniform switch so every
// register index is static.  Per node and thread:
//     scatter:  v = psi_x[r0-dx] psi_y[r1-dy] f_j ;  win[d+k] += psi_z[k] v            (2m+1 complex FMAs, no atomics)
//     gather :  t = sum_k psi_z[k] win[d+k]        ;  partial = psi_x psi_y t -> reduced over the footprint rows
// Warps whose rows do not meet the node's footprint skip it (warp-uniform test).  When a sub-chunk is done the
// window advances by ZS cells: the scatter flushes the finished cells through a shared-memory staging box with ONE
// TMA reduce-add per box (cp.reduce.async.bulk.tensor: the only contended traffic, resolved in L2), the gather pulls
// the next cells in with TMA tile loads (cp.async.bulk.tensor + mbarrier), double buffered.
#pragma once
#include "gridding.cuh"

namespace pnb {

template <int M_> struct ZmCfg {
  static constexpr int C = 2 * M_ + 1;
  static constexpr int CP = 16 * ((C + 15) / 16);          // padded weight row (16-byte aligned vector loads)
  static constexpr int T0 = (M_ <= 6) ? 16 : 8;            // column tile, cells
  static constexpr int T1 = 4;
  static constexpr int ZS = (M_ <= 6) ? 8 : 4;             // z sub-chunk (window advance)
  static constexpr int ZB = 4;                             // z extent of one TMA box
  static constexpr int R0 = T0 + 2 * M_, R1 = T1 + 2 * M_; // footprint rows
  static constexpr int ROWS = R0 * R1;
  static constexpr int NT = ROWS;                          // one thread per row
  static constexpr int NWARP = NT / 32;
  static constexpr int W = ZS + 2 * M_;                    // register window (cells)
  static constexpr int ZSEG = 128 / ZS;                    // sub-chunks per work item (128 cells of z)
  static constexpr int NB = (M_ <= 6) ? 32 : 16;           // nodes per batch (weights staged in shared memory)
  static_assert(ROWS % 32 == 0, "rows must fill whole warps");
  static_assert(ZS % ZB == 0 && (2 * M_) % ZB == 0, "window advance and halo must be whole TMA boxes");
};

struct ZmGeom {
  int nc[2];       // column tiles per axis
  int nt2;         // z sub-chunks
  int nseg;        // work items per column
};

template <class R> struct WPair;
template <> struct WPair<double> { typedef double2 type; };
template <> struct WPair<float> { typedef float2 type; };

// window values of a batch of nodes -> shared memory, layout per node: psi_x[CP] psi_y[CP] psi_z[CP] (dpsi_x dpsi_y dpsi_z)
template <class R, int M_, bool GRAD, int NT>
__device__ __forceinline__ void zm_batch_weights(const GridGeom<R> &g, const NodeArgs<R> &na, const R *poly_s, int b0, int nb,
                                                 R *wts) {
  typedef ZmCfg<M_> Cfg;
  constexpr int C = Cfg::C, CP = Cfg::CP, WPN = (GRAD ? 6 : 3) * CP;
  if (na.pre_psi) {
    for (int v = threadIdx.x; v < nb * 3 * C; v += NT) {
      const int i = v / (3 * C), r = v - i * 3 * C, t = r / C, s = r - t * C;
      wts[i * WPN + t * CP + s] = na.pre_psi[(size_t)(b0 + i) * 3 * C + r];
      if (GRAD) wts[i * WPN + (3 + t) * CP + s] = na.pre_dpsi[(size_t)(b0 + i) * 3 * C + r];
    }
  } else if (g.kind == WIN_BSPLINE && !g.poly) {
    for (int v = threadIdx.x; v < nb * 3; v += NT) {
      const int i = v / 3, t = v - i * 3, j = na.perm[b0 + i];
      const R nxv = mul_rn(g.n[t], na.x[3 * (size_t)j + t]);
      bspline_taps<R>(M_, nxv - m_floor(nxv), g.n[t], wts + i * WPN + t * CP, GRAD ? wts + i * WPN + (3 + t) * CP : nullptr);
    }
  } else {
    for (int v = threadIdx.x; v < nb * 3 * C; v += NT) {
      const int i = v / (3 * C), r = v - i * 3 * C, t = r / C, s = r - t * C, j = na.perm[b0 + i];
      const R nxv = mul_rn(g.n[t], na.x[3 * (size_t)j + t]);
      const R flv = m_floor(nxv), fr = nxv - flv;
      R psi, dpsi = (R)0;
      if (g.poly && fr != (R)0) {
        const R u = (R)2 * fr - (R)1;
        const R *a = poly_s + r;
        psi = a[g.poly_deg * 3 * C];
        for (int k = g.poly_deg - 1; k >= 0; k--) { if (GRAD) dpsi = dpsi * u + psi; psi = psi * u + a[k * 3 * C]; }
        dpsi *= (R)2 * g.n[t];
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
      wts[i * WPN + t * CP + s] = psi;
      if (GRAD) wts[i * WPN + (3 + t) * CP + s] = dpsi;
    }
  }
}

// win[D+k] += wz[k] * A (+ dwz[k] * B)
template <int D, int C, bool GRAD, class R, class Cell, int W>
__device__ __forceinline__ void zm_accum(Cell (&win)[W], const R *wz, const R *dwz, const Cell &A, const Cell &B) {
  typedef typename WPair<R>::type P2;
#pragma unroll
  for (int k = 0; k + 1 < C; k += 2) {
    const P2 w = *reinterpret_cast<const P2 *>(wz + k);
    fma_cell(win[D + k], w.x, A);
    fma_cell(win[D + k + 1], w.y, A);
    if (GRAD) {
      const P2 dw = *reinterpret_cast<const P2 *>(dwz + k);
      fma_cell(win[D + k], dw.x, B);
      fma_cell(win[D + k + 1], dw.y, B);
    }
  }
  fma_cell(win[D + C - 1], wz[C - 1], A);
  if (GRAD) fma_cell(win[D + C - 1], dwz[C - 1], B);
}

template <int ZS, int C, bool GRAD, class R, class Cell, int W>
__device__ __forceinline__ void zm_accum_switch(int d, Cell (&win)[W], const R *wz, const R *dwz, const Cell &A, const Cell &B) {
  switch (d) {
    case 0: zm_accum<0, C, GRAD>(win, wz, dwz, A, B); break;
    case 1: zm_accum<1, C, GRAD>(win, wz, dwz, A, B); break;
    case 2: zm_accum<2, C, GRAD>(win, wz, dwz, A, B); break;
    case 3: zm_accum<3, C, GRAD>(win, wz, dwz, A, B); break;
    default:
      if constexpr (ZS > 4) {
        switch (d) {
          case 4: zm_accum<4, C, GRAD>(win, wz, dwz, A, B); break;
          case 5: zm_accum<5, C, GRAD>(win, wz, dwz, A, B); break;
          case 6: zm_accum<6, C, GRAD>(win, wz, dwz, A, B); break;
          default: zm_accum<7, C, GRAD>(win, wz, dwz, A, B); break;
        }
      }
      break;
  }
}

template <class R, bool CPLX, int M_, bool GRAD> struct ZmSmem {
  typedef typename CellT<R, CPLX>::type Cell;
  typedef ZmCfg<M_> Cfg;
  static constexpr int BOX_CELLS = Cfg::ROWS * Cfg::ZB;
  static constexpr size_t box_bytes = (size_t)BOX_CELLS * sizeof(Cell);
  static constexpr int WPN = (GRAD ? 6 : 3) * Cfg::CP;
  static constexpr size_t off_wts = 2 * box_bytes;                                        // two staging boxes
  static constexpr size_t off_vals = off_wts + (size_t)Cfg::NB * WPN * sizeof(R);
  static constexpr size_t off_hdr = off_vals + (size_t)Cfg::NB * 4 * sizeof(Cell);
  static constexpr size_t off_poly = off_hdr + (size_t)Cfg::NB * sizeof(int);
  static constexpr size_t off_bar = (off_poly + (size_t)kMaxPolyCoef * 3 * Cfg::C * sizeof(R) + 15) / 16 * 16;
  static constexpr size_t scatter = off_bar + 64;
  // gather: per batch node the z-contracted partials of every footprint row: t (and t' for the gradient)
  static constexpr int GNB = (M_ <= 6) ? 16 : 8;                                           // gather batch
  static constexpr int PSTRIDE = Cfg::C * Cfg::C * (GRAD ? 2 : 1);                         // partial cells per node
  static constexpr size_t off_part = off_bar + 64;
  static constexpr size_t gather = off_part + (size_t)GNB * PSTRIDE * sizeof(Cell);
  static_assert(scatter <= 232448 && gather <= 232448, "shared-memory budget of one CTA exceeded");
};

// ------------------------------------------------------------------------------------------------
// scatter (adjoint B^T)
// ------------------------------------------------------------------------------------------------
template <class R, bool CPLX, int M_, bool GRAD>
__global__ void __launch_bounds__(ZmCfg<M_>::NT, 1)
k_scatter_zm(const __grid_constant__ CUtensorMap tmap, GridGeom<R> g, ZmGeom zg, NodeArgs<R> na,
             const int *__restrict__ tile_start) {
  typedef typename CellT<R, CPLX>::type Cell;
  typedef ZmCfg<M_> Cfg;
  typedef ZmSmem<R, CPLX, M_, GRAD> Sm;
  constexpr int C = Cfg::C, CP = Cfg::CP, R1 = Cfg::R1, ZS = Cfg::ZS, ZB = Cfg::ZB, W = Cfg::W, NB = Cfg::NB, NT = Cfg::NT;
  constexpr int NCOMP = CPLX ? 2 : 1;
  constexpr int WPN = Sm::WPN;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cell *stage = reinterpret_cast<Cell *>(smem_raw);
  R *wts = reinterpret_cast<R *>(smem_raw + Sm::off_wts);
  Cell *vals = reinterpret_cast<Cell *>(smem_raw + Sm::off_vals);
  int *hdr = reinterpret_cast<int *>(smem_raw + Sm::off_hdr);
  R *poly_s = reinterpret_cast<R *>(smem_raw + Sm::off_poly);

  const int col = blockIdx.x / zg.nseg, seg = blockIdx.x - col * zg.nseg;
  const int tz0 = seg * Cfg::ZSEG, tz1 = min(zg.nt2, tz0 + Cfg::ZSEG);
  const int *ts = tile_start + (size_t)col * zg.nt2;
  if (ts[tz0] == ts[tz1]) return;                                   // no nodes in this segment
  const int cx = col / zg.nc[1], cy = col - cx * zg.nc[1];
  const int o0 = cx * Cfg::T0, o1 = cy * Cfg::T1;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r0 = tid / R1, r1 = tid - r0 * R1;
  const int wr0min = (warp * 32) / R1, wr0max = (warp * 32 + 31) / R1;

  if (g.poly) for (int i = tid; i < (g.poly_deg + 1) * 3 * C; i += NT) poly_s[i] = g.poly[i];

  Cell win[W];
#pragma unroll
  for (int i = 0; i < W; i++) zero_cell(win[i]);
  int nflush = 0, dirty = 0;   // dirty: sub-chunks since the window last received a contribution are still non-zero

  auto flush = [&](auto first_tag, int zcoord) {
    constexpr int FIRST = decltype(first_tag)::value;
    Cell *st = stage + (nflush & 1) * Sm::BOX_CELLS;
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the box issued two flushes ago was read
    __syncthreads();
#pragma unroll
    for (int q = 0; q < ZB; q++) st[tid * ZB + q] = win[FIRST + q];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      tma_reduce_add_3d(st, &tmap, zcoord * NCOMP, o1, o0);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    nflush++;
  };

  for (int tz = tz0; tz < tz1; tz++) {
    const int s = ts[tz], e = ts[tz + 1];
    const int zb = tz * ZS;
    for (int b0 = s; b0 < e; b0 += NB) {
      const int nb = min(NB, e - b0);
      __syncthreads();   // previous batch consumed
      if (tid < nb) {
        const int j = na.perm[b0 + tid];
        R xs[3] = {na.x[3 * (size_t)j], na.x[3 * (size_t)j + 1], na.x[3 * (size_t)j + 2]}, nx[3], fl[3];
        int cell[3];
        project_node(g, xs, nx, fl, cell);
        hdr[tid] = (cell[0] - o0) | ((cell[1] - o1) << 8) | ((cell[2] - zb) << 16);
        Cell z; zero_cell(z);
        vals[4 * tid] = na.f ? load_in(na.f + ((size_t)j * na.f_stride + na.f_off) * NCOMP, z) : z;
        if (GRAD) {
          const R *gp = na.grad + (size_t)j * 3 * NCOMP;
          vals[4 * tid + 1] = load_in(gp, z); vals[4 * tid + 2] = load_in(gp + NCOMP, z); vals[4 * tid + 3] = load_in(gp + 2 * NCOMP, z);
        }
      }
      zm_batch_weights<R, M_, GRAD, NT>(g, na, poly_s, b0, nb, wts);
      __syncthreads();
      for (int i = 0; i < nb; i++) {
        const int h = hdr[i];
        const int dx = h & 255, dy = (h >> 8) & 255, dz = h >> 16;
        if (dx > wr0max || dx + C - 1 < wr0min) continue;            // warp-uniform: footprint misses this warp's rows
        const int i0 = r0 - dx, i1 = r1 - dy;
        const bool in = (unsigned)i0 < (unsigned)C && (unsigned)i1 < (unsigned)C;
        const R *w = wts + i * WPN;
        const R w0 = in ? w[i0] : (R)0, w1 = in ? w[CP + i1] : (R)0;
        Cell A = scale_cell(w0 * w1, vals[4 * i]), B;
        zero_cell(B);
        if (GRAD) {
          const R dw0 = in ? w[3 * CP + i0] : (R)0, dw1 = in ? w[4 * CP + i1] : (R)0;
          fma_cell(A, dw0 * w1, vals[4 * i + 1]);
          fma_cell(A, w0 * dw1, vals[4 * i + 2]);
          B = scale_cell(w0 * w1, vals[4 * i + 3]);
        }
        zm_accum_switch<ZS, C, GRAD>(dz, win, w + 2 * CP, w + 5 * CP, A, B);
      }
    }
    if (e > s) dirty = (W + ZS - 1) / ZS;
    // the first ZS cells of the window are final: flush them, advance the window
    if (dirty > 0) {
#pragma unroll
      for (int q = 0; q < ZS / ZB; q++) {
        if (q == 0) flush(std::integral_constant<int, 0>(), zb);
        else flush(std::integral_constant<int, ZB>(), zb + ZB);
      }
      dirty--;
    }
#pragma unroll
    for (int i = 0; i < W - ZS; i++) win[i] = win[i + ZS];
#pragma unroll
    for (int i = W - ZS; i < W; i++) zero_cell(win[i]);
  }
  // tail: the 2m cells beyond the last sub-chunk
  if (dirty > 0) {
    const int zb = tz1 * ZS;
    flush(std::integral_constant<int, 0>(), zb);
    if constexpr (2 * M_ > ZB) flush(std::integral_constant<int, ZB>(), zb + ZB);
    if constexpr (2 * M_ > 2 * ZB) flush(std::integral_constant<int, 2 * ZB>(), zb + 2 * ZB);
    if constexpr (2 * M_ > 3 * ZB) flush(std::integral_constant<int, 3 * ZB>(), zb + 3 * ZB);
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging must outlive the bulk reads
}


// ------------------------------------------------------------------------------------------------
// gather (trafo B)
// ------------------------------------------------------------------------------------------------
// t = sum_k wz[k] win[D+k]  (td = sum_k dwz[k] win[D+k])
template <int D, int C, bool GRAD, class R, class Cell, int W>
__device__ __forceinline__ void zm_dot(const Cell (&win)[W], const R *wz, const R *dwz, Cell &t, Cell &td) {
  typedef typename WPair<R>::type P2;
  Cell t1, td1;
  zero_cell(t); zero_cell(td); zero_cell(t1); zero_cell(td1);
#pragma unroll
  for (int k = 0; k + 1 < C; k += 2) {
    const P2 w = *reinterpret_cast<const P2 *>(wz + k);
    fma_cell(t, w.x, win[D + k]);
    fma_cell(t1, w.y, win[D + k + 1]);
    if (GRAD) {
      const P2 dw = *reinterpret_cast<const P2 *>(dwz + k);
      fma_cell(td, dw.x, win[D + k]);
      fma_cell(td1, dw.y, win[D + k + 1]);
    }
  }
  fma_cell(t, wz[C - 1], win[D + C - 1]);
  if (GRAD) fma_cell(td, dwz[C - 1], win[D + C - 1]);
  fma_cell(t, (R)1, t1);
  if (GRAD) fma_cell(td, (R)1, td1);
}

template <int ZS, int C, bool GRAD, class R, class Cell, int W>
__device__ __forceinline__ void zm_dot_switch(int d, const Cell (&win)[W], const R *wz, const R *dwz, Cell &t, Cell &td) {
  switch (d) {
    case 0: zm_dot<0, C, GRAD>(win, wz, dwz, t, td); break;
    case 1: zm_dot<1, C, GRAD>(win, wz, dwz, t, td); break;
    case 2: zm_dot<2, C, GRAD>(win, wz, dwz, t, td); break;
    case 3: zm_dot<3, C, GRAD>(win, wz, dwz, t, td); break;
    default:
      if constexpr (ZS > 4) {
        switch (d) {
          case 4: zm_dot<4, C, GRAD>(win, wz, dwz, t, td); break;
          case 5: zm_dot<5, C, GRAD>(win, wz, dwz, t, td); break;
          case 6: zm_dot<6, C, GRAD>(win, wz, dwz, t, td); break;
          default: zm_dot<7, C, GRAD>(win, wz, dwz, t, td); break;
        }
      }
      break;
  }
}

__device__ __forceinline__ double2 shfl_down_cell(double2 v, int o, int width) {
  return make_double2(__shfl_down_sync(0xffffffffu, v.x, o, width), __shfl_down_sync(0xffffffffu, v.y, o, width));
}
__device__ __forceinline__ float2 shfl_down_cell(float2 v, int o, int width) {
  return make_float2(__shfl_down_sync(0xffffffffu, v.x, o, width), __shfl_down_sync(0xffffffffu, v.y, o, width));
}
__device__ __forceinline__ double shfl_down_cell(double v, int o, int width) { return __shfl_down_sync(0xffffffffu, v, o, width); }
__device__ __forceinline__ float shfl_down_cell(float v, int o, int width) { return __shfl_down_sync(0xffffffffu, v, o, width); }
__device__ __forceinline__ int cell_is_nan(const double2 &a) { return (a.x != a.x) | (a.y != a.y); }
__device__ __forceinline__ int cell_is_nan(const float2 &a) { return (a.x != a.x) | (a.y != a.y); }
__device__ __forceinline__ int cell_is_nan(const double &a) { return a != a; }
__device__ __forceinline__ int cell_is_nan(const float &a) { return a != a; }
__device__ __forceinline__ void add_cell(double2 &a, const double2 &b) { a.x += b.x; a.y += b.y; }
__device__ __forceinline__ void add_cell(float2 &a, const float2 &b) { a.x += b.x; a.y += b.y; }
__device__ __forceinline__ void add_cell(double &a, const double &b) { a += b; }
__device__ __forceinline__ void add_cell(float &a, const float &b) { a += b; }

template <class R, bool CPLX, int M_, bool GRAD>
__global__ void __launch_bounds__(ZmCfg<M_>::NT, 1)
k_gather_zm(const __grid_constant__ CUtensorMap tmap, GridGeom<R> g, ZmGeom zg, NodeArgs<R> na,
            const int *__restrict__ tile_start) {
  typedef typename CellT<R, CPLX>::type Cell;
  typedef ZmCfg<M_> Cfg;
  typedef ZmSmem<R, CPLX, M_, GRAD> Sm;
  constexpr int C = Cfg::C, CP = Cfg::CP, R1 = Cfg::R1, ZS = Cfg::ZS, ZB = Cfg::ZB, W = Cfg::W, NT = Cfg::NT;
  constexpr int NCOMP = CPLX ? 2 : 1;
  constexpr int WPN = Sm::WPN, GNB = Sm::GNB;
  constexpr int PSTRIDE = Sm::PSTRIDE;
  constexpr int LPN = 16;                              // lanes per node in the reduction phase
  constexpr unsigned BOX_BYTES = (unsigned)Sm::box_bytes;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cell *stage = reinterpret_cast<Cell *>(smem_raw);
  R *wts = reinterpret_cast<R *>(smem_raw + Sm::off_wts);
  int *hdr = reinterpret_cast<int *>(smem_raw + Sm::off_hdr);
  R *poly_s = reinterpret_cast<R *>(smem_raw + Sm::off_poly);
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw + Sm::off_bar);
  Cell *part = reinterpret_cast<Cell *>(smem_raw + Sm::off_part);

  const int col = blockIdx.x / zg.nseg, seg = blockIdx.x - col * zg.nseg;
  const int tz0 = seg * Cfg::ZSEG, tz1 = min(zg.nt2, tz0 + Cfg::ZSEG);
  const int *ts = tile_start + (size_t)col * zg.nt2;
  if (ts[tz0] == ts[tz1]) return;
  const int cx = col / zg.nc[1], cy = col - cx * zg.nc[1];
  const int o0 = cx * Cfg::T0, o1 = cy * Cfg::T1;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r0 = tid / R1, r1 = tid - r0 * R1;
  const int wr0min = (warp * 32) / R1, wr0max = (warp * 32 + 31) / R1;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (g.poly) for (int i = tid; i < (g.poly_deg + 1) * 3 * C; i += NT) poly_s[i] = g.poly[i];
  __syncthreads();

  unsigned phase[2] = {0u, 0u};
  auto issue = [&](int buf, int zcoord) {           // thread 0: TMA box [R0][R1][ZB] at z = zcoord -> staging buffer
    mbar_expect_tx(&bar[buf], BOX_BYTES);
    tma_load_3d(stage + buf * Sm::BOX_CELLS, &tmap, zcoord * NCOMP, o1, o0, &bar[buf]);
  };
  Cell win[W];
  // all threads: wait for the box, copy my row's ZB cells into the window.  Returns a predicate that depends on the
  // loaded values: the caller feeds it to __syncthreads_or() so that every thread's shared-memory reads have
  // RETURNED before the barrier releases thread 0 to re-arm the staging buffer with the next TMA load.  (A plain
  // bar.sync only orders the issue of the loads; with the load/store unit backed up the async-proxy write of the
  // next box can overtake generic-proxy reads that are still queued -- observed as cells of box b+2 in window b.)
  auto take = [&](auto first_tag, int buf) -> int {
    constexpr int FIRST = decltype(first_tag)::value;
    mbar_wait(&bar[buf], phase[buf]);
    phase[buf] ^= 1u;
    const Cell *st = stage + buf * Sm::BOX_CELLS + tid * ZB;
    int nan = 0;
#pragma unroll
    for (int q = 0; q < ZB; q++) { win[FIRST + q] = st[q]; nan |= cell_is_nan(win[FIRST + q]); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    return nan;
  };

  // ---- prologue: window = cells [zb0, zb0 + W) ----
  {
    const int zb0 = tz0 * ZS;
    static_assert(W % ZB == 0, "window must be whole boxes");
    constexpr int NBOX = W / ZB;
    if (tid == 0) { issue(0, zb0); if (NBOX > 1) issue(1, zb0 + ZB); }
#pragma unroll
    for (int b = 0; b < NBOX; b++) {
      // static slot index through a small dispatch (b is a compile-time constant after unrolling)
      int pr = 0;
      if (b == 0) pr = take(std::integral_constant<int, 0>(), 0);
      else if (b == 1) pr = take(std::integral_constant<int, ZB>(), 1);
      else if (b == 2) pr = take(std::integral_constant<int, 2 * ZB < W ? 2 * ZB : 0>(), 0);
      else if (b == 3) pr = take(std::integral_constant<int, 3 * ZB < W ? 3 * ZB : 0>(), 1);
      else if (b == 4) pr = take(std::integral_constant<int, 4 * ZB < W ? 4 * ZB : 0>(), 0);
      else if (b == 5) pr = take(std::integral_constant<int, 5 * ZB < W ? 5 * ZB : 0>(), 1);
      (void)__syncthreads_or(pr);                    // everyone has read (and received) buffer (b & 1)
      if (tid == 0 && b + 2 < NBOX) issue(b & 1, zb0 + (b + 2) * ZB);
    }
  }

  for (int tz = tz0; tz < tz1; tz++) {
    const int s = ts[tz], e = ts[tz + 1];
    const int zb = tz * ZS;
    const bool more = tz + 1 < tz1;
    // prefetch the cells the next advance needs: [zb + W, zb + W + ZS)
    if (tid == 0 && more) { issue(0, zb + W); if (ZS / ZB > 1) issue(1, zb + W + ZB); }

    for (int b0 = s; b0 < e; b0 += GNB) {
      const int nb = min(GNB, e - b0);
      __syncthreads();   // previous batch fully reduced
      if (tid < nb) {
        const int j = na.perm[b0 + tid];
        R xs[3] = {na.x[3 * (size_t)j], na.x[3 * (size_t)j + 1], na.x[3 * (size_t)j + 2]}, nx[3], fl[3];
        int cell[3];
        project_node(g, xs, nx, fl, cell);
        hdr[tid] = (cell[0] - o0) | ((cell[1] - o1) << 8) | ((cell[2] - zb) << 16);
      }
      zm_batch_weights<R, M_, GRAD, NT>(g, na, poly_s, b0, nb, wts);
      __syncthreads();
      // ---- phase B: z contraction from the register window, one partial per footprint row ----
      for (int i = 0; i < nb; i++) {
        const int h = hdr[i];
        const int dx = h & 255, dy = (h >> 8) & 255, dz = h >> 16;
        if (dx > wr0max || dx + C - 1 < wr0min) continue;
        const R *w = wts + i * WPN;
        Cell t, td;
        zm_dot_switch<ZS, C, GRAD>(dz, win, w + 2 * CP, w + 5 * CP, t, td);
        const int i0 = r0 - dx, i1 = r1 - dy;
        if ((unsigned)i0 < (unsigned)C && (unsigned)i1 < (unsigned)C) {
          Cell *p = part + i * PSTRIDE + i0 * C + i1;
          p[0] = t;
          if (GRAD) p[C * C] = td;
        }
      }
      __syncthreads();
      // ---- phase C: weighted reduction over the (2m+1)^2 rows, LPN lanes per node ----
      if (warp * (32 / LPN) < nb) {                    // warp-uniform: this warp owns at least one node of the batch
        const int grp = tid / LPN, sub = tid - grp * LPN;
        const bool act = grp < nb;
        const R *w = wts + (act ? grp : 0) * WPN;
        const Cell *p = part + (act ? grp : 0) * PSTRIDE;
        Cell af, a0, a1, a2;
        zero_cell(af); zero_cell(a0); zero_cell(a1); zero_cell(a2);
        for (int q = act ? sub : C * C; q < C * C; q += LPN) {
          const int i0 = q / C, i1 = q - i0 * C;
          const R w0 = w[i0], w1 = w[CP + i1];
          const Cell t = p[q];
          fma_cell(af, w0 * w1, t);
          if (GRAD) {
            const R dw0 = w[3 * CP + i0], dw1 = w[4 * CP + i1];
            fma_cell(a0, dw0 * w1, t);
            fma_cell(a1, w0 * dw1, t);
            fma_cell(a2, w0 * w1, p[C * C + q]);
          }
        }
#pragma unroll
        for (int o = LPN / 2; o > 0; o >>= 1) {
          add_cell(af, shfl_down_cell(af, o, LPN));
          if (GRAD) { add_cell(a0, shfl_down_cell(a0, o, LPN)); add_cell(a1, shfl_down_cell(a1, o, LPN)); add_cell(a2, shfl_down_cell(a2, o, LPN)); }
        }
        if (act && sub == 0) {
          const int j = na.perm[b0 + grp];
          if (na.f) store_out(na.f + ((size_t)j * na.f_stride + na.f_off) * NCOMP, af, na.accumulate);
          if (GRAD) {
            R *o = na.grad + (size_t)j * 3 * NCOMP;
            store_out(o, a0, na.accumulate);
            store_out(o + NCOMP, a1, na.accumulate);
            store_out(o + 2 * NCOMP, a2, na.accumulate);
          }
        }
      }
    }
    // ---- advance the window by ZS cells ----
    if (more) {
#pragma unroll
      for (int i = 0; i < W - ZS; i++) win[i] = win[i + ZS];
      int pr = take(std::integral_constant<int, W - ZS>(), 0);
      if constexpr (ZS / ZB > 1) pr |= take(std::integral_constant<int, W - ZS + ZB>(), 1);
      (void)__syncthreads_or(pr);   // staging buffers free for the next prefetch
    }
  }
}

}  // namespace pnb
