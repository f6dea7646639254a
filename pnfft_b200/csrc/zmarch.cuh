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
  static constexpr int T0 = (M_ <= 6) ? 16 : 8;            // column tile, cells
  static constexpr int T1 = 4;
  static constexpr int ZS = (M_ <= 6) ? 8 : 4;             // z sub-chunk (window advance)
  static constexpr int ZB = 4;                             // z extent of one TMA box
  static constexpr int R0 = T0 + 2 * M_, R1 = T1 + 2 * M_; // footprint rows
  static constexpr int ROWS = R0 * R1;                     // one consumer thread per row
  static constexpr int NWARP = ROWS / 32;                  // consumer warps
  static constexpr int NT = ROWS + 32;                     // + one producer warp (streams the node table)
  static constexpr int W = ZS + 2 * M_;                    // register window (cells)
  static constexpr int ZSEG = 128 / ZS;                    // sub-chunks per work item (128 cells of z)
  static_assert(ROWS % 32 == 0, "rows must fill whole warps");
  static_assert(ZS % ZB == 0 && (2 * M_) % ZB == 0, "window advance and halo must be whole TMA boxes");
};

struct ZmGeom {
  int nc[2];       // column tiles per axis
  int nt2;         // z sub-chunks
  int nseg;        // work items per column
  int sub;         // bins per (column, sub-chunk) in tile_start (1, or the x-offset bins of the tensor-core gather's sort key)
};

template <class R> struct WPair;
template <> struct WPair<double> { typedef double2 type; };
template <> struct WPair<float> { typedef float2 type; };

// Per-call node table (sorted order): everything a gridding kernel needs about a node, evaluated ONCE per node
// and axis by k_node_table and streamed into the gridding kernels with 1-d bulk copies.  Row layout (units of R):
//   psi_x[CP] psi_y[CP] psi_z[CP] (dpsi_x[CP] dpsi_y[CP] dpsi_z[CP])  vals[VS]
// slot C of psi_x / psi_y / psi_z holds the node's cell offset inside its tile (dx, dy, dz) as an int bit pattern;
// vals = f (and grad_f) of the node for the adjoint.
template <class R, int M_, bool GRAD> struct ZmTab {
  static constexpr int C = 2 * M_ + 1;
  static constexpr int CP = (sizeof(R) == 8) ? (C + 1) : ((C + 1 + 3) / 4 * 4);
  static constexpr int NROW = GRAD ? 6 : 3;
  static constexpr int WPN = NROW * CP;
  static constexpr int VS = GRAD ? 8 : (sizeof(R) == 8 ? 2 : 4);
  static constexpr int ROWLEN = WPN + VS;
  static constexpr int ROWBYTES = ROWLEN * (int)sizeof(R);
  static_assert(ROWBYTES % 16 == 0, "bulk copies need 16-byte granules");
  static_assert((CP * sizeof(R)) % (2 * sizeof(R)) == 0, "weight rows must allow paired loads");
};

__device__ __forceinline__ int as_int_bits(double v) { return __double2loint(v); }
__device__ __forceinline__ int as_int_bits(float v) { return __float_as_int(v); }
__device__ __forceinline__ double int_bits_as(int i, double) { return __hiloint2double(0, i); }
__device__ __forceinline__ float int_bits_as(int i, float) { return __int_as_float(i); }

// one thread per (node, axis): 2m+1 window values (and derivatives) of that axis
template <class R, int M_, bool GRAD>
__global__ void __launch_bounds__(192)
k_node_table(GridGeom<R> g, NodeArgs<R> na, int ncomp, int with_vals, R *__restrict__ tab) {
  typedef ZmCfg<M_> Cfg;
  typedef ZmTab<R, M_, GRAD> Tab;
  constexpr int C = Cfg::C, CP = Tab::CP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R *poly_s = reinterpret_cast<R *>(smem_raw);
  if (g.poly) {
    for (int i = threadIdx.x; i < 2 * (g.poly_deg + 1) * 3 * C; i += blockDim.x) poly_s[i] = g.poly[i];
    __syncthreads();
  }
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int p = (int)(gid / 3), t = (int)(gid - 3LL * p);
  if (p >= na.M) return;
  const int j = na.perm[p];
  R nxv, flv;
  int cell;
  node_axis(g, na.x[3 * (size_t)j + t], t, &nxv, &flv, &cell);
  const R fr = nxv - flv;
  R *row = tab + (size_t)p * Tab::ROWLEN;
  R *rp = row + t * CP, *rd = row + (3 + t) * CP;
  if (na.pre_psi) {
    for (int s = 0; s < C; s++) {
      rp[s] = na.pre_psi[(size_t)p * 3 * C + t * C + s];
      if (GRAD) rd[s] = na.pre_dpsi[(size_t)p * 3 * C + t * C + s];
    }
  } else if (g.intpol_order >= 0) {
    for (int s = 0; s < C; s++) {
      rp[s] = intpol_tap(g, g.intpol_tab[t], s, fr);
      if (GRAD) rd[s] = intpol_tap(g, g.intpol_tab[3 + t], s, fr);
    }
  } else if (g.poly && fr != (R)0) {
    // per-tap polynomials in u = 2 frac - 1 (Core::fit_window_polys); all taps advance together (Horner)
    R psi[C], dpsi[GRAD ? C : 1];
    const R u = (R)2 * fr - (R)1;
    const R *a = poly_s + t * C;
    const int nv = 3 * C;
    const R *ad = a + (g.poly_deg + 1) * nv;
#pragma unroll
    for (int s = 0; s < C; s++) { psi[s] = a[g.poly_deg * nv + s]; if (GRAD) dpsi[s] = ad[g.poly_deg * nv + s]; }
    for (int k = g.poly_deg - 1; k >= 0; k--) {
#pragma unroll
      for (int s = 0; s < C; s++) {
        if (GRAD) dpsi[s] = dpsi[s] * u + ad[k * nv + s];
        psi[s] = psi[s] * u + a[k * nv + s];
      }
    }
#pragma unroll
    for (int s = 0; s < C; s++) { rp[s] = psi[s]; if (GRAD) rd[s] = dpsi[s]; }
  } else if (g.kind == WIN_BSPLINE) {
    bspline_taps<R>(M_, fr, g.n[t], rp, GRAD ? rd : nullptr);
  } else if (g.kind == WIN_GAUSSIAN && g.fast_gauss) {
    const R d = nxv - (flv - (R)M_);
    const R e_sqr = m_exp(-(d * d) / g.b[t]), e_lin = m_exp((R)2 * d / g.b[t]);
    R tmp = e_sqr;
    for (int s = 0; s < C; s++) {
      const R v = tmp * g.exp_const[t * C + s];
      rp[s] = v;
      if (GRAD) rd[s] = (R)(-2.0) * g.n[t] / g.b[t] * (d - (R)s) * v;
      tmp *= e_lin;
    }
  } else {
#pragma unroll 1
    for (int s = 0; s < C; s++) {
      R a = (R)0, b = (R)0;
      window_tap<R>(g.kind, flv - nxv - (R)M_ + (R)s, g.n[t], g.b[t], M_, GRAD, &a, &b);
      rp[s] = a;
      if (GRAD) rd[s] = b;
    }
  }
  if (t == 0 && g.wscale != (R)1)     // the 0.5 of an interlaced plan rides on the x-axis factors
    for (int s = 0; s < C; s++) { rp[s] *= g.wscale; if (GRAD) rd[s] *= g.wscale; }
  // cell offset inside the (T0, T1, ZS) tile
  const int T = t == 0 ? Cfg::T0 : (t == 1 ? Cfg::T1 : Cfg::ZS);
  row[t * CP + C] = int_bits_as(cell - (cell / T) * T, (R)0);
  if (t == 0 && with_vals) {
    R *v = row + Tab::WPN;
    if (na.f) for (int c = 0; c < ncomp; c++) v[c] = na.f[((size_t)j * na.f_stride + na.f_off) * ncomp + c];
    else for (int c = 0; c < ncomp; c++) v[c] = (R)0;
    if (GRAD) for (int c = 0; c < 3 * ncomp; c++) v[ncomp + c] = na.grad[(size_t)j * 3 * ncomp + c];
  }
}

// ---- mbarrier / bulk-copy helpers of the node-table ring ----
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int N> __device__ __forceinline__ void consumer_sync() {   // named barrier over the N consumer threads
  asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory");
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
  typedef ZmTab<R, M_, GRAD> Tab;
  static constexpr int BOX_CELLS = Cfg::ROWS * Cfg::ZB;
  static constexpr size_t box_bytes = (size_t)BOX_CELLS * sizeof(Cell);
  static constexpr int S = 4;                                   // ring stages
  static constexpr int SGB = 16;                                // scatter: nodes per batch
  static constexpr int GGB = (M_ <= 6) ? (GRAD ? 8 : 16) : (GRAD ? 4 : 8);   // gather: nodes per batch
  static constexpr size_t off_ring = 2 * box_bytes;             // after the two staging boxes
  static constexpr size_t sc_off_bar = off_ring + (size_t)S * SGB * Tab::ROWBYTES;
  static constexpr size_t scatter = sc_off_bar + 128;
  static constexpr int PSTRIDE = Cfg::C * Cfg::C * (GRAD ? 2 : 1);     // partial cells per node
  static constexpr size_t ga_off_bar = off_ring + (size_t)S * GGB * Tab::ROWBYTES;
  static constexpr size_t ga_off_part = (ga_off_bar + 128 + 15) / 16 * 16;
  static constexpr size_t gather = ga_off_part + (size_t)2 * GGB * PSTRIDE * sizeof(Cell);
  static_assert(scatter <= 232448 && gather <= 232448, "shared-memory budget of one CTA exceeded");
};

// output side of the gather
template <class R> struct GatherOut {
  const int *perm;
  R *f;
  long long f_stride, f_off;
  R *grad;
  int accumulate;
};

// ------------------------------------------------------------------------------------------------
// scatter (adjoint B^T)
// ------------------------------------------------------------------------------------------------
template <class R, bool CPLX, int M_, bool GRAD>
__global__ void __launch_bounds__(ZmCfg<M_>::NT, 1)
k_scatter_zm(const __grid_constant__ CUtensorMap tmap, ZmGeom zg, const R *__restrict__ tab, const int *__restrict__ tile_start) {
  typedef typename CellT<R, CPLX>::type Cell;
  typedef ZmCfg<M_> Cfg;
  typedef ZmTab<R, M_, GRAD> Tab;
  typedef ZmSmem<R, CPLX, M_, GRAD> Sm;
  constexpr int C = Cfg::C, CP = Tab::CP, R1 = Cfg::R1, ZS = Cfg::ZS, ZB = Cfg::ZB, W = Cfg::W, ROWS = Cfg::ROWS;
  constexpr int NCOMP = CPLX ? 2 : 1, S = Sm::S, GB = Sm::SGB, ROWLEN = Tab::ROWLEN;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cell *stage = reinterpret_cast<Cell *>(smem_raw);
  R *ring = reinterpret_cast<R *>(smem_raw + Sm::off_ring);
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + Sm::sc_off_bar);
  unsigned long long *empty = full + S;

  const int col = blockIdx.x / zg.nseg, seg = blockIdx.x - col * zg.nseg;
  const int tz0 = seg * Cfg::ZSEG, tz1 = min(zg.nt2, tz0 + Cfg::ZSEG);
  const int *ts = tile_start + (size_t)col * zg.nt2 * zg.sub;
  if (ts[(size_t)tz0 * zg.sub] == ts[(size_t)tz1 * zg.sub]) return;                                   // no nodes in this segment
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], Cfg::NWARP); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == Cfg::NWARP) {
    // ---- producer warp: stream the node-table rows of every batch into the ring ----
    if (lane == 0) {
      int kb = 0;
      for (int tz = tz0; tz < tz1; tz++) {
        const int s = ts[(size_t)tz * zg.sub], e = ts[(size_t)(tz + 1) * zg.sub];
        for (int b0 = s; b0 < e; b0 += GB, kb++) {
          const int st = kb % S;
          mbar_wait(&empty[st], ((unsigned)(kb / S) & 1u) ^ 1u);
          const unsigned bytes = (unsigned)(min(GB, e - b0) * Tab::ROWBYTES);
          mbar_expect_tx(&full[st], bytes);
          bulk_load_1d(ring + (size_t)st * GB * ROWLEN, tab + (size_t)b0 * ROWLEN, bytes, &full[st]);
        }
      }
    }
    return;
  }

  const int cx = col / zg.nc[1], cy = col - cx * zg.nc[1];
  const int o0 = cx * Cfg::T0, o1 = cy * Cfg::T1;
  const int r0 = tid / R1, r1 = tid - r0 * R1;
  const int wr0min = (warp * 32) / R1, wr0max = (warp * 32 + 31) / R1;

  Cell win[W];
#pragma unroll
  for (int i = 0; i < W; i++) zero_cell(win[i]);
  int nflush = 0, dirty = 0;   // dirty: flushes still needed until the window is all zero again

  auto flush = [&](auto first_tag, int zcoord) {
    constexpr int FIRST = decltype(first_tag)::value;
    Cell *st = stage + (nflush & 1) * Sm::BOX_CELLS;
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the box issued two flushes ago was read
    consumer_sync<ROWS>();
#pragma unroll
    for (int q = 0; q < ZB; q++) st[tid * ZB + q] = win[FIRST + q];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    consumer_sync<ROWS>();
    if (tid == 0) {
      tma_reduce_add_3d(st, &tmap, zcoord * NCOMP, o1, o0);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    nflush++;
  };

  int kb = 0;
  for (int tz = tz0; tz < tz1; tz++) {
    const int s = ts[(size_t)tz * zg.sub], e = ts[(size_t)(tz + 1) * zg.sub];
    const int zb = tz * ZS;
    for (int b0 = s; b0 < e; b0 += GB, kb++) {
      const int nb = min(GB, e - b0), st = kb % S;
      mbar_wait(&full[st], (unsigned)(kb / S) & 1u);
      const R *wb = ring + (size_t)st * GB * ROWLEN;
      for (int i = 0; i < nb; i++) {
        const R *w = wb + i * ROWLEN;
        const int dx = as_int_bits(w[C]);
        if (dx > wr0max || dx + C - 1 < wr0min) continue;            // warp-uniform: footprint misses this warp's rows
        const int dy = as_int_bits(w[CP + C]), dz = as_int_bits(w[2 * CP + C]);
        const int i0 = r0 - dx, i1 = r1 - dy;
        const bool in = (unsigned)i0 < (unsigned)C && (unsigned)i1 < (unsigned)C;
        const R w0 = in ? w[i0] : (R)0, w1 = in ? w[CP + i1] : (R)0;
        const R *v = w + Tab::WPN;
        Cell z; zero_cell(z);
        Cell A = scale_cell(w0 * w1, load_in(v, z)), B;
        zero_cell(B);
        if (GRAD) {
          const R dw0 = in ? w[3 * CP + i0] : (R)0, dw1 = in ? w[4 * CP + i1] : (R)0;
          fma_cell(A, dw0 * w1, load_in(v + NCOMP, z));
          fma_cell(A, w0 * dw1, load_in(v + 2 * NCOMP, z));
          B = scale_cell(w0 * w1, load_in(v + 3 * NCOMP, z));
        }
        zm_accum_switch<ZS, C, GRAD>(dz, win, w + 2 * CP, w + 5 * CP, A, B);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
    if (e > s) dirty = (W + ZS - 1) / ZS;
    // the first ZS cells of the window are final: flush them, advance the window
    if (dirty > 0) {
      flush(std::integral_constant<int, 0>(), zb);
      if constexpr (ZS / ZB > 1) flush(std::integral_constant<int, ZB>(), zb + ZB);
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


// named-barrier OR-reduction over the N consumer threads (bar.red needs every thread's predicate, i.e. the values it
// depends on have arrived in registers: see the note at take() below)
template <int N> __device__ __forceinline__ int consumer_sync_or(int pred) {
  int res;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.s32 p, %1, 0;\n"
      "bar.red.or.pred q, 1, %2, p;\n"
      "selp.s32 %0, 1, 0, q;\n"
      "}\n" : "=r"(res) : "r"(pred), "n"(N) : "memory");
  return res;
}

template <class R, bool CPLX, int M_, bool GRAD>
__global__ void __launch_bounds__(ZmCfg<M_>::NT, 1)
k_gather_zm(const __grid_constant__ CUtensorMap tmap, ZmGeom zg, const R *__restrict__ tab, const int *__restrict__ tile_start,
            GatherOut<R> out) {
  typedef typename CellT<R, CPLX>::type Cell;
  typedef ZmCfg<M_> Cfg;
  typedef ZmTab<R, M_, GRAD> Tab;
  typedef ZmSmem<R, CPLX, M_, GRAD> Sm;
  constexpr int C = Cfg::C, CP = Tab::CP, R1 = Cfg::R1, ZS = Cfg::ZS, ZB = Cfg::ZB, W = Cfg::W, ROWS = Cfg::ROWS;
  constexpr int NCOMP = CPLX ? 2 : 1, S = Sm::S, GB = Sm::GGB, ROWLEN = Tab::ROWLEN, PSTRIDE = Sm::PSTRIDE;
  constexpr int LPN = 16;                              // lanes per node in the reduction phase
  constexpr int NCW = GB / 2;                          // warps that reduce (two nodes each): the first and last NCW/2,
  constexpr int NWARP = Cfg::NWARP;                    // whose rows meet the fewest footprints
  static_assert(NCW <= NWARP, "not enough warps for the reduction phase");
  constexpr unsigned BOX_BYTES = (unsigned)Sm::box_bytes;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cell *stage = reinterpret_cast<Cell *>(smem_raw);
  R *ring = reinterpret_cast<R *>(smem_raw + Sm::off_ring);
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + Sm::ga_off_bar);
  unsigned long long *empty = full + S;
  unsigned long long *tbar = empty + S;
  Cell *part = reinterpret_cast<Cell *>(smem_raw + Sm::ga_off_part);

  const int col = blockIdx.x / zg.nseg, seg = blockIdx.x - col * zg.nseg;
  const int tz0 = seg * Cfg::ZSEG, tz1 = min(zg.nt2, tz0 + Cfg::ZSEG);
  const int *ts = tile_start + (size_t)col * zg.nt2 * zg.sub;
  if (ts[(size_t)tz0 * zg.sub] == ts[(size_t)tz1 * zg.sub]) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], NWARP); }
    mbar_init(&tbar[0], 1);
    mbar_init(&tbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NWARP) {
    if (lane == 0) {
      int kb = 0;
      for (int tz = tz0; tz < tz1; tz++) {
        const int s = ts[(size_t)tz * zg.sub], e = ts[(size_t)(tz + 1) * zg.sub];
        for (int b0 = s; b0 < e; b0 += GB, kb++) {
          const int st = kb % S;
          mbar_wait(&empty[st], ((unsigned)(kb / S) & 1u) ^ 1u);
          const unsigned bytes = (unsigned)(min(GB, e - b0) * Tab::ROWBYTES);
          mbar_expect_tx(&full[st], bytes);
          bulk_load_1d(ring + (size_t)st * GB * ROWLEN, tab + (size_t)b0 * ROWLEN, bytes, &full[st]);
        }
      }
    }
    return;
  }

  const int cx = col / zg.nc[1], cy = col - cx * zg.nc[1];
  const int o0 = cx * Cfg::T0, o1 = cy * Cfg::T1;
  const int r0 = tid / R1, r1 = tid - r0 * R1;
  const int wr0min = (warp * 32) / R1, wr0max = (warp * 32 + 31) / R1;
  const int cw = warp < NCW / 2 ? warp : (warp >= NWARP - NCW / 2 ? warp - (NWARP - NCW) : -1);

  unsigned phase[2] = {0u, 0u};
  auto issue = [&](int buf, int zcoord) {           // thread 0: TMA box [R0][R1][ZB] at z = zcoord -> staging buffer
    mbar_expect_tx(&tbar[buf], BOX_BYTES);
    tma_load_3d(stage + buf * Sm::BOX_CELLS, &tmap, zcoord * NCOMP, o1, o0, &tbar[buf]);
  };
  Cell win[W];
  // all threads: wait for the box, copy my row's ZB cells into the window.  Returns a predicate that depends on the
  // loaded values: the caller feeds it to consumer_sync_or() so that every thread's shared-memory reads have
  // RETURNED before the barrier releases thread 0 to re-arm the staging buffer with the next TMA load.  (A plain
  // bar.sync only orders the issue of the loads; with the load/store unit backed up the async-proxy write of the
  // next box can overtake generic-proxy reads that are still queued -- observed as cells of box b+2 in window b.)
  auto take = [&](auto first_tag, int buf) -> int {
    constexpr int FIRST = decltype(first_tag)::value;
    mbar_wait(&tbar[buf], phase[buf]);
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
      int pr = 0;
      if (b == 0) pr = take(std::integral_constant<int, 0>(), 0);
      else if (b == 1) pr = take(std::integral_constant<int, ZB>(), 1);
      else if (b == 2) pr = take(std::integral_constant<int, 2 * ZB < W ? 2 * ZB : 0>(), 0);
      else if (b == 3) pr = take(std::integral_constant<int, 3 * ZB < W ? 3 * ZB : 0>(), 1);
      else if (b == 4) pr = take(std::integral_constant<int, 4 * ZB < W ? 4 * ZB : 0>(), 0);
      else if (b == 5) pr = take(std::integral_constant<int, 5 * ZB < W ? 5 * ZB : 0>(), 1);
      (void)consumer_sync_or<ROWS>(pr);              // everyone has read (and received) buffer (b & 1)
      if (tid == 0 && b + 2 < NBOX) issue(b & 1, zb0 + (b + 2) * ZB);
    }
  }

  // weighted reduction of batch (st_, b0_, nb_) over the (2m+1)^2 footprint rows, LPN lanes per node
  auto reduce_batch = [&](int st_, int b0_, int nb_, int pb_) {
    if (cw < 0 || cw * 2 >= nb_) return;             // warp-uniform
    const int grp = cw * 2 + (lane >> 4), sub = lane & (LPN - 1);
    const bool act = grp < nb_;
    const R *w = ring + ((size_t)st_ * GB + (act ? grp : 0)) * ROWLEN;
    const Cell *p = part + (size_t)(pb_ * GB + (act ? grp : 0)) * PSTRIDE;
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
      const int j = out.perm[b0_ + grp];
      if (out.f) store_out(out.f + ((size_t)j * out.f_stride + out.f_off) * NCOMP, af, out.accumulate);
      if (GRAD) {
        R *o = out.grad + (size_t)j * 3 * NCOMP;
        store_out(o, a0, out.accumulate);
        store_out(o + NCOMP, a1, out.accumulate);
        store_out(o + 2 * NCOMP, a2, out.accumulate);
      }
    }
  };

  int kb = 0, prev_st = -1, prev_b0 = 0, prev_nb = 0;
  for (int tz = tz0; tz < tz1; tz++) {
    const int s = ts[(size_t)tz * zg.sub], e = ts[(size_t)(tz + 1) * zg.sub];
    const int zb = tz * ZS;
    const bool more = tz + 1 < tz1;
    // prefetch the cells the next advance needs: [zb + W, zb + W + ZS)
    if (tid == 0 && more) { issue(0, zb + W); if (ZS / ZB > 1) issue(1, zb + W + ZB); }

    for (int b0 = s; b0 < e; b0 += GB, kb++) {
      const int nb = min(GB, e - b0), st = kb % S, pb = kb & 1;
      mbar_wait(&full[st], (unsigned)(kb / S) & 1u);
      const R *wb = ring + (size_t)st * GB * ROWLEN;
      // ---- z contraction of batch kb from the register window: one partial per footprint row ----
      for (int i = 0; i < nb; i++) {
        const R *w = wb + i * ROWLEN;
        const int dx = as_int_bits(w[C]);
        if (dx > wr0max || dx + C - 1 < wr0min) continue;
        const int dy = as_int_bits(w[CP + C]), dz = as_int_bits(w[2 * CP + C]);
        Cell t, td;
        zm_dot_switch<ZS, C, GRAD>(dz, win, w + 2 * CP, w + 5 * CP, t, td);
        const int i0 = r0 - dx, i1 = r1 - dy;
        if ((unsigned)i0 < (unsigned)C && (unsigned)i1 < (unsigned)C) {
          Cell *p = part + (size_t)(pb * GB + i) * PSTRIDE + i0 * C + i1;
          p[0] = t;
          if (GRAD) p[C * C] = td;
        }
      }
      // ---- reduction of batch kb-1 (its partials became visible at the barrier that ended the last iteration) ----
      if (prev_st >= 0) {
        reduce_batch(prev_st, prev_b0, prev_nb, pb ^ 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[prev_st]);       // the ring stage of batch kb-1 is free
      }
      consumer_sync<ROWS>();
      prev_st = st; prev_b0 = b0; prev_nb = nb;
    }
    // ---- advance the window by ZS cells ----
    if (more) {
#pragma unroll
      for (int i = 0; i < W - ZS; i++) win[i] = win[i + ZS];
      int pr = take(std::integral_constant<int, W - ZS>(), 0);
      if constexpr (ZS / ZB > 1) pr |= take(std::integral_constant<int, W - ZS + ZB>(), 1);
      (void)consumer_sync_or<ROWS>(pr);   // staging buffers free for the next prefetch
    }
  }
  if (prev_st >= 0) reduce_batch(prev_st, prev_b0, prev_nb, (kb - 1) & 1);
}

}  // namespace pnb
