// Warp-autonomous z-marching gridding kernels (v2, the default B / B^T path for m = 4 and 6).
// Reference loops replaced: kernel/assign.c:478-1130 (the (2m+1)^3 inner loops), kernel/ndft-parallel.c:2703-3009.
//
// What v1 (zmarch.cuh) taught (profiles/r1_zmarch_v1_ncu_full.md): with the grid cells in registers the FP64 pipe is
// the right bound, but v1 spent 45 % of its stall samples in CTA-wide barriers (row warps see very different node
// counts) and only 19-24 % of its issued instructions were DFMA (every warp walked every node and tested whether the
// footprint met its rows).  v2 removes both:
//   * no CTA-wide barrier after start-up: every consumer warp owns the grid rows (x = 2w, 2w+1; all 16 y rows) of the
//     column tile's footprint, keeps their sliding z window in registers, and loads / flushes ITS rows with its own TMA
//     boxes [2][16][ZB] and its own mbarrier.  Warps drift apart by up to the depth of the node ring.
//   * nodes are sorted by (column tile, z sub-chunk, dx): the rows of warp w meet exactly the nodes with
//     dx in [2w-2m, 2w+1], a contiguous range of every chunk, found from a 17-entry prefix table in the chunk header.
//     No per-node skip test, no predicated weight loads (the x / y weight rows are stored zero-padded).
//   * the gather leaves its per-row partial sums in mbarrier-guarded shared-memory stages; whichever warps are idle
//     (the edge rows see few nodes) grab nodes of the previous chunk from an atomic ticket and do their x-y reduction.
//   * the node loops are software pipelined by hand (header / x / y weights of the next node are fetched while the
//     z taps of the current one run), because one warp per CTA meets every node and issues in order.
// A producer warp streams the node table (k_node_table2: window factors once per node and axis, written coalesced)
// into a shared-memory ring with 1-d bulk copies and builds the chunk headers.
#pragma once
#include <climits>

#include "zmarch.cuh"

namespace pnb {

#ifndef ZM2_PARK_NS
#define ZM2_PARK_NS 500u
#endif
#ifndef ZM2_RPT
#define ZM2_RPT 1
#endif
#ifndef ZM2_DZ_STATIC
#define ZM2_DZ_STATIC 0     // measured on B200 (C3): 19 % fewer DFMA, but gather F +8 %, scatter F +17 %, gather F+grad -2 %
#endif
#ifndef ZM2_GCHAINS
#define ZM2_GCHAINS 1       // accumulation chains per sum of the gradient gather
#endif
#ifndef ZM2_DEPTH
#define ZM2_DEPTH 4         // z weight quads in flight
#endif
template <int M_> struct Zm2Cfg {
  static constexpr int C = 2 * M_ + 1;
  static constexpr int R1 = 16;                  // footprint rows along y: half a warp
  static constexpr int T1 = R1 - 2 * M_;         // column tile, cells (m=6: 4, m=4: 8)
  static constexpr int RPT = ZM2_RPT;            // x rows per thread (register blocking)
#ifdef ZM2_T0
  static constexpr int T0 = ZM2_T0;
#else
  static constexpr int T0 = RPT == 2 ? 12 : 10;  // R0 = 24 (m=6): 6 consumer warps of 4 x rows / R0 = 22: 11 warps of 2 rows
#endif
  static constexpr int SUB = 16;                 // x-offset bins per tile in the sort key (>= T0)
  static constexpr int ZS = 4;                   // z sub-chunk == window advance
  static constexpr int ZB = 4;                   // z extent of one TMA box
  static constexpr int XW = 2 * RPT;             // x rows per warp: RPT per thread, two lane halves
  static constexpr int R0 = T0 + 2 * M_;
  static constexpr int NCW = R0 / XW;            // consumer warps
  static constexpr int W = ZS + 2 * M_;          // register window (cells) per row
  static constexpr int NFL = (W + ZS - 1) / ZS;  // flushes until a touched window is all zero again
  static constexpr int XLEAD = XW - 1;           // zero padding in front of the x weights
  static constexpr int YLEAD = T1 - 1;           // ... and of the y weights
  // z taps: true = the table stores psi_z unshifted and the node loops dispatch on the node's z offset inside its
  // sub-chunk (ZS statically indexed copies of the tap loop, 2m+1 taps each); false = psi_z stored shifted by that
  // offset and zero padded to W (one branch-free loop over all W window cells, (W - 2m - 1) / W of its FMAs on zeros)
  static constexpr bool DZS = ZM2_DZ_STATIC != 0;
  static_assert(T1 >= 1 && R0 % XW == 0 && W % ZB == 0 && ZS % ZB == 0 && T0 <= SUB, "unsupported cutoff");
};

struct Zm2Geom {
  int nc[2];       // column tiles per axis
  int nt2;         // z sub-chunks per column
  int zseg;        // sub-chunks per work item
  int nseg;        // work items per column
  int col0;        // first column of this launch (the node table may be built and consumed in column batches)
  int target;      // nodes per work item a heavy column is cut into (zm_segment)
  int fill;        // pieces every column is cut into regardless (few columns: fills the GPU)
};

// Work items of a column: up to nseg pieces along z with (nearly) equal NODE counts, cut at sub-chunk boundaries found in
// the column's bin prefix; a column is split into as many pieces as it has multiples of `target` nodes (at least `fill`), so light
// columns stay whole (every piece pays a window prologue) and the items of a heavy column weigh about `target` each.  Equal-length
// pieces leave half of a clustered column (a Gaussian blob along z) in one item, and the heaviest item bounds the launch
// once the node set is spread over several GPUs (C4 on 8 GPUs).  Items beyond the column's piece count are empty.
__host__ __device__ __forceinline__ void zm_segment(const int *__restrict__ bs, int nt2, int sub, int seg, int nseg, int target, int fill, int &tz0, int &tz1) {
  if (nseg <= 1) { tz0 = 0; tz1 = nt2; return; }
  if (target <= 0) {            // pieces of equal length
    const int zseg = (nt2 + nseg - 1) / nseg;
    tz0 = seg * zseg < nt2 ? seg * zseg : nt2; tz1 = tz0 + zseg < nt2 ? tz0 + zseg : nt2;
    return;
  }
  const int s0 = bs[0];
  const long long total = bs[(size_t)nt2 * sub] - s0;
  long long want = (total + target - 1) / (target > 0 ? target : 1);
  if (want < (fill > 1 ? fill : 1)) want = fill > 1 ? fill : 1;
  const int pieces = (int)(want < nseg ? want : nseg);
  auto cut = [&](int k) -> int {      // smallest sub-chunk t with at least total * k / pieces nodes in front of it
    if (k <= 0) return 0;
    if (k >= pieces) return nt2;
    const long long goal = total * k / pieces;
    int lo = 0, hi = nt2;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((long long)(bs[(size_t)mid * sub] - s0) >= goal) hi = mid; else lo = mid + 1;
    }
    return lo;
  };
  if (seg >= pieces) { tz0 = 0; tz1 = 0; return; }
  tz0 = cut(seg);
  tz1 = cut(seg + 1);
}

// Node-table row (units of R).  hdr = 8 ints {-dx*sizeof(R), -dy*sizeof(R), dz, dx, node index j, 0, 0, 0};
// X = [0 x XLEAD, psi_x[0..C), 0 x XLEAD...], Y = [0 x (T1-1), psi_y[0..C), 0 x (T1-1)...], Z = psi_z[0..C) zero padded to ZP
// (Cfg::DZS: the node loops dispatch on the header's dz) or [0 x dz, psi_z[0..C), 0...] of length W, already aligned with
// the register window of the node's sub-chunk (no dispatch, but W instead of 2m+1 taps per node and row);
// the same three rows of dpsi when GRAD; vals = f (and grad_f) of the node for the adjoint.
template <class R, class Cfg_, bool GRAD, bool VALS, bool CPLX> struct ZmRowOf {
  typedef Cfg_ Cfg;
  static constexpr int AL = 16 / (int)sizeof(R);
  static constexpr int up(int v) { return (v + AL - 1) / AL * AL; }
  static constexpr int C = Cfg::C;
  static constexpr int HDR = 32 / (int)sizeof(R);
  static constexpr int XP = up(C + 2 * Cfg::XLEAD), YP = up(C + 2 * Cfg::YLEAD), ZP = up(Cfg::W);
  static constexpr int oX = HDR, oY = oX + XP, oZ = oY + YP;
  static constexpr int oDX = oZ + ZP, oDY = oDX + XP, oDZ = oDY + YP;
  static constexpr int oV = GRAD ? oDZ + ZP : oZ + ZP;
  static constexpr int NV = VALS ? up((CPLX ? 2 : 1) * (GRAD ? 4 : 1)) : 0;
  static constexpr int ROWLEN = oV + NV;
  static constexpr int ROWBYTES = ROWLEN * (int)sizeof(R);
  static_assert(ROWBYTES % 16 == 0, "bulk copies need 16-byte granules");
};
template <class R, int M_, bool GRAD, bool VALS, bool CPLX> struct Zm2Row : ZmRowOf<R, Zm2Cfg<M_>, GRAD, VALS, CPLX> {};

constexpr int kZm2HdrBytes = 128;    // chunk header: {tz, count, first sorted index, 0, start[0..T0]}
#ifndef ZM2_TABNODES
#define ZM2_TABNODES 64
#endif
constexpr int kZm2TabNodes = ZM2_TABNODES;     // nodes per block of the table kernel

// ------------------------------------------------------------------------------------------------
// node table: one thread per (node, axis); rows assembled in shared memory, written out coalesced
// ------------------------------------------------------------------------------------------------
template <class R, int M_, bool GRAD, bool VALS, bool CPLX, class Cfg_ = Zm2Cfg<M_>>
// (four blocks per SM is what shared memory allows for v3's rows; without the bound the register-resident B-spline /
// fast-Gaussian branches raise the kernel to 96 registers and cost every window a block of occupancy)
__global__ void __launch_bounds__(3 * kZm2TabNodes, 4)
k_node_table2(GridGeom<R> g, NodeArgs<R> na, R *__restrict__ tab, int first) {
  typedef Cfg_ Cfg;
  typedef ZmRowOf<R, Cfg_, GRAD, VALS, CPLX> Row;
  constexpr int C = Cfg::C, NCOMP = CPLX ? 2 : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R *rows = reinterpret_cast<R *>(smem_raw);
  R *poly_s = rows + (size_t)kZm2TabNodes * Row::ROWLEN;
  if (g.poly) for (int i = threadIdx.x; i < (GRAD ? 2 : 1) * (g.poly_deg + 1) * 3 * C; i += blockDim.x) poly_s[i] = g.poly[i];
  // the rows are zero padded: clear them once with 16-byte stores, the threads then write the taps only
  for (int i = threadIdx.x; i < kZm2TabNodes * (Row::ROWBYTES / 16); i += blockDim.x) reinterpret_cast<uint4 *>(rows)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const int p0 = first + blockIdx.x * kZm2TabNodes;      // rows [first, na.M) of the sorted order; tab is indexed absolutely
  const int nn = min(kZm2TabNodes, na.M - p0);
  const int ln = threadIdx.x / 3, t = threadIdx.x - 3 * ln;
  if (ln < nn) {
    const int p = p0 + ln;
    const int j = na.perm[p];
    R nxv, flv;
    int cell;
    node_axis(g, na.x[3 * (size_t)j + t], t, &nxv, &flv, &cell);
    const R fr = nxv - flv;
    R psi[C], dpsi[GRAD ? C : 1];
    unsigned slow = 0;       // taps to redo with the library-call formulas
    if (na.pre_psi) {
#pragma unroll
      for (int s = 0; s < C; s++) {
        psi[s] = na.pre_psi[(size_t)p * 3 * C + t * C + s];
        if (GRAD) dpsi[s] = na.pre_dpsi[(size_t)p * 3 * C + t * C + s];
      }
    } else if (g.intpol_order >= 0) {
      // (statically indexed like every other branch: psi / dpsi must stay in registers)
      const R *tp0 = t == 0 ? g.intpol_tab[0] : (t == 1 ? g.intpol_tab[1] : g.intpol_tab[2]);
      const R *tp1 = t == 0 ? g.intpol_tab[3] : (t == 1 ? g.intpol_tab[4] : g.intpol_tab[5]);
#pragma unroll
      for (int s = 0; s < C; s++) {
        psi[s] = intpol_tap(g, tp0, s, fr);
        if (GRAD) dpsi[s] = intpol_tap(g, tp1, s, fr);
      }
    } else if (g.poly && fr != (R)0 && (!GRAD || g.poly_deg <= 16)) {
      // per-tap polynomials in u = 2 frac - 1 (Core::fit_window_polys); all taps advance together (Horner).  With the
      // gradient the exact formulas are cheaper than two degree > 16 Horner chains per tap (measured).
      const R u = (R)2 * fr - (R)1;
      const int dg = GRAD ? g.poly_deg : g.poly_deg_psi;
      const R *a = poly_s + t * C;
      const int nv = 3 * C;
      const R *ad = a + (g.poly_deg + 1) * nv;      // the derivative weights have their own fitted polynomials
#pragma unroll
      for (int s = 0; s < C; s++) { psi[s] = a[dg * nv + s]; if (GRAD) dpsi[s] = ad[dg * nv + s]; }
      for (int k = dg - 1; k >= 0; k--) {
#pragma unroll
        for (int s = 0; s < C; s++) {
          if (GRAD) dpsi[s] = dpsi[s] * u + ad[k * nv + s];
          psi[s] = psi[s] * u + a[k * nv + s];
        }
      }
    } else if (sizeof(R) == 8 && g.kind == WIN_KAISER_BESSEL) {
      // one exponential and one division per tap (window.h: kb_tap_fast); the rare small-argument taps are redone with
      // the library-call formulas below, straight into the shared-memory row
#pragma unroll
      for (int s = 0; s < C; s++) {
        double a = 0, b = 0;
        if (!kb_tap_fast((double)(flv - nxv - (R)M_ + (R)s), (double)g.n[t], (double)g.b[t], M_, GRAD, &a, &b)) slow |= 1u << s;
        psi[s] = (R)a;
        if (GRAD) dpsi[s] = (R)b;
      }
    } else if (g.kind == WIN_BSPLINE) {
      bspline_taps_fixed<R, M_, GRAD>(fr, g.n[t], psi, dpsi);      // de Boor triangle in registers
    } else if (g.kind == WIN_GAUSSIAN && g.fast_gauss) {
      // PNFFT_FAST_GAUSSIAN (reference ndft-parallel.c:1790-1822): two exponentials per axis, the taps by recurrence
      const R d = nxv - (flv - (R)M_);
      const R e_sqr = m_exp(-(d * d) / g.b[t]), e_lin = m_exp((R)2 * d / g.b[t]);
      R tmp = e_sqr;
#pragma unroll
      for (int s = 0; s < C; s++) {
        const R v = tmp * g.exp_const[t * C + s];
        psi[s] = v;
        if (GRAD) dpsi[s] = (R)(-2.0) * g.n[t] / g.b[t] * (d - (R)s) * v;
        tmp *= e_lin;
      }
    } else if (sizeof(R) == 8 && g.kind == WIN_GAUSSIAN && !g.fast_gauss) {
      // Gaussian taps in double (window.h: window_tap) with the branch-free exponential of the Kaiser-Bessel path and one
      // division per thread instead of two per tap: psi = exp(-y^2 / b) / sqrt(pi b), dpsi = 2 n / b * y * psi
      const double ib = 1.0 / (double)g.b[t];
      const double c0 = 1.0 / sqrt(3.14159265358979323846 * (double)g.b[t]);
      const double c1 = 2.0 * (double)g.n[t] * ib;
      const double y0 = (double)(flv - nxv) - (double)M_;
#pragma unroll
      for (int s = 0; s < C; s++) {
        const double y = y0 + (double)s;
        const double a = exp_mid(-(y * y) * ib) * c0;
        psi[s] = (R)a;
        if (GRAD) dpsi[s] = (R)(c1 * y * a);
      }
    } else {
      // exact formulas (nodes on a grid line, windows without a polynomial fit): through a local scratch row
      R tp[C], td[C];
      if (g.kind == WIN_BSPLINE) {
        bspline_taps<R>(M_, fr, g.n[t], tp, GRAD ? td : nullptr);
      } else if (g.kind == WIN_GAUSSIAN && g.fast_gauss) {
        const R d = nxv - (flv - (R)M_);
        const R e_sqr = m_exp(-(d * d) / g.b[t]), e_lin = m_exp((R)2 * d / g.b[t]);
        R tmp = e_sqr;
        for (int s = 0; s < C; s++) {
          const R v = tmp * g.exp_const[t * C + s];
          tp[s] = v;
          td[s] = (R)(-2.0) * g.n[t] / g.b[t] * (d - (R)s) * v;
          tmp *= e_lin;
        }
      } else {
#pragma unroll 1
        for (int s = 0; s < C; s++) {
          R a = (R)0, b = (R)0;
          window_tap<R>(g.kind, flv - nxv - (R)M_ + (R)s, g.n[t], g.b[t], M_, GRAD, &a, &b);
          tp[s] = a; td[s] = b;
        }
      }
#pragma unroll
      for (int s = 0; s < C; s++) { psi[s] = tp[s]; if (GRAD) dpsi[s] = td[s]; }
    }
    R *row = rows + (size_t)ln * Row::ROWLEN;
    const int off = t == 0 ? Row::oX : (t == 1 ? Row::oY : Row::oZ);
    const int T = t == 0 ? Cfg::T0 : (t == 1 ? Cfg::T1 : Cfg::ZS);
    const int d = cell - (cell / T) * T;
    if (t == 0 && g.wscale != (R)1) {     // the 0.5 of an interlaced plan rides on the x-axis factors (exact)
#pragma unroll
      for (int s = 0; s < C; s++) { psi[s] *= g.wscale; if (GRAD) dpsi[s] *= g.wscale; }
    }
    const int dzc = min(max(d, 0), Cfg::ZS - 1);
    const int lead = t == 0 ? Cfg::XLEAD : (t == 1 ? Cfg::YLEAD : (Cfg::DZS ? 0 : dzc));
#pragma unroll
    for (int s = 0; s < C; s++) { row[off + lead + s] = psi[s]; if (GRAD) row[off + (Row::oDX - Row::oX) + lead + s] = dpsi[s]; }
    while (slow) {
      const int s = __ffs(slow) - 1;
      slow &= slow - 1;
      R a = (R)0, b = (R)0;
      window_tap<R>(WIN_KAISER_BESSEL, flv - nxv - (R)M_ + (R)s, g.n[t], g.b[t], M_, GRAD, &a, &b);
      if (t == 0) { a *= g.wscale; b *= g.wscale; }
      row[off + lead + s] = a;
      if (GRAD) row[off + (Row::oDX - Row::oX) + lead + s] = b;
    }
    // cell offset inside the (T0, T1, ZS) tile
    int *h = reinterpret_cast<int *>(row);
    if (t == 0) { h[0] = -d * (int)sizeof(R); h[3] = d; h[4] = j; }
    else if (t == 1) { h[1] = -d * (int)sizeof(R); h[5] = 0; }
    else { h[2] = dzc; h[6] = 0; h[7] = 0; }
    if (VALS && t == 0) {
      R *v = row + Row::oV;
      for (int c = 0; c < Row::NV; c++) v[c] = (R)0;
      if (na.f) for (int c = 0; c < NCOMP; c++) v[c] = na.f[((size_t)j * na.f_stride + na.f_off) * NCOMP + c];
      if (GRAD) for (int c = 0; c < 3 * NCOMP; c++) v[NCOMP + c] = na.grad[(size_t)j * 3 * NCOMP + c];
    }
  }
  __syncthreads();
  const uint4 *src = reinterpret_cast<const uint4 *>(rows);
  uint4 *dst = reinterpret_cast<uint4 *>(tab + (size_t)p0 * Row::ROWLEN);
  const int n16 = nn * (Row::ROWBYTES / 16);
  for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up
// ------------------------------------------------------------------------------------------------
template <class R, bool CPLX, int M_, bool GRAD> struct Zm2Smem {
  typedef typename CellT<R, CPLX>::type Cell;
  typedef Zm2Cfg<M_> Cfg;
  typedef Zm2Row<R, M_, GRAD, true, CPLX> RowS;
  typedef Zm2Row<R, M_, GRAD, false, CPLX> RowG;
  static constexpr int WARP_ROWS = Cfg::XW * 16;                           // grid rows one warp owns
  static constexpr int WARP_BOX = WARP_ROWS * Cfg::ZS * (int)sizeof(Cell); // bytes one warp stages per window advance
  // scatter: two staging buffers per warp, ring of S chunks of GB nodes
  static constexpr int SS = 4, SGB = 32;
  static constexpr int s_stage = kZm2HdrBytes + SGB * RowS::ROWBYTES;
  static constexpr int s_off_ring = Cfg::NCW * 2 * WARP_BOX;
  static constexpr int s_off_bar = s_off_ring + SS * s_stage;
  static constexpr int scatter = s_off_bar + 2 * SS * 8;
  // gather: one staging buffer per warp, ring, P stages of per-row partial sums
#ifndef ZM2_GP
#define ZM2_GP 3
#endif
#ifndef ZM2_GGB
#define ZM2_GGB 8
#endif
  static constexpr int GS = 4, GP = ZM2_GP;
  static constexpr int GGB = GRAD ? ZM2_GGB : 2 * ZM2_GGB;
  static constexpr int PN = Cfg::C * 16 * (GRAD ? 2 : 1);                 // partial cells per node
  static constexpr int g_stage = kZm2HdrBytes + GGB * RowG::ROWBYTES;
  static constexpr int g_off_ring = Cfg::NCW * WARP_BOX;
  static constexpr int g_off_part = g_off_ring + GS * g_stage;
  static constexpr int g_off_bar = g_off_part + GP * GGB * PN * (int)sizeof(Cell);
  static constexpr int gather = g_off_bar + (2 * GS + 2 * GP + Cfg::NCW) * 8;
  static_assert(s_stage % 16 == 0 && g_stage % 16 == 0 && s_off_ring % 128 == 0 && g_off_ring % 128 == 0 && WARP_BOX % 512 == 0, "alignment");
  static_assert(scatter <= 232448 && gather <= 232448, "shared-memory budget of one CTA exceeded");
};

// mbarrier wait that lets the hardware park the warp (suspend-time hint) instead of spinning on the issue port
__device__ __forceinline__ void mbar_wait_park(unsigned long long *bar, unsigned phase, unsigned ns = ZM2_PARK_NS) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITP_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONEP_%=;\n"
      "bra WAITP_%=;\n"
      "DONEP_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(phase), "r"(ns) : "memory");
}

// z taps of one node from / into the register window; the z offset picks one of ZS statically indexed variants
// through a two- (three-) level branch tree (cheaper than the jump table a switch compiles to)
template <int D, int C, bool GRAD, class R, class Cell, int W, int ZP>
__device__ __forceinline__ void zm2_acc(Cell (&win)[W], const R (&wz)[ZP], const R (&dwz)[ZP], const Cell &A, const Cell &B) {
#pragma unroll
  for (int k = 0; k < C; k++) {
    fma_cell(win[D + k], wz[k], A);
    if (GRAD) fma_cell(win[D + k], dwz[k], B);
  }
}
template <int D, int C, bool GRAD, class R, class Cell, int W, int ZP>
__device__ __forceinline__ void zm2_dot(const Cell (&win)[W], const R (&wz)[ZP], const R (&dwz)[ZP], Cell &t, Cell &td) {
  Cell t1, td1;    // two accumulation chains per sum
  zero_cell(t); zero_cell(td); zero_cell(t1); zero_cell(td1);
#pragma unroll
  for (int k = 0; k < C; k++) {
    if (k & 1) { fma_cell(t1, wz[k], win[D + k]); if (GRAD) fma_cell(td1, dwz[k], win[D + k]); }
    else { fma_cell(t, wz[k], win[D + k]); if (GRAD) fma_cell(td, dwz[k], win[D + k]); }
  }
  add_cell(t, t1);
  if (GRAD) add_cell(td, td1);
}
#define ZM2_TREE(d, CALL)                                         \
  do {                                                            \
    if constexpr (ZS == 4) {                                      \
      if ((d) & 2) { if ((d) & 1) { CALL(3); } else { CALL(2); } } \
      else { if ((d) & 1) { CALL(1); } else { CALL(0); } }        \
    } else {                                                      \
      static_assert(ZS == 8, "z sub-chunk must be 4 or 8");       \
      if ((d) & 4) {                                              \
        if ((d) & 2) { if ((d) & 1) { CALL(7); } else { CALL(6); } } \
        else { if ((d) & 1) { CALL(5); } else { CALL(4); } }      \
      } else {                                                    \
        if ((d) & 2) { if ((d) & 1) { CALL(3); } else { CALL(2); } } \
        else { if ((d) & 1) { CALL(1); } else { CALL(0); } }      \
      }                                                           \
    }                                                             \
  } while (0)

__device__ __forceinline__ int warp_any_volatile(int pred) {
  int res;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.s32 p, %1, 0;\n"
      "vote.sync.any.pred q, p, 0xffffffff;\n"
      "selp.s32 %0, 1, 0, q;\n"
      "}\n" : "=r"(res) : "r"(pred) : "memory");
  return res;
}

// producer warp: chunk headers + node-table rows of one work item into the ring; NEND end markers at the end
template <int T0, int SUB, int S, int GB, int ROWBYTES, int STAGE, int NEND>
__device__ __forceinline__ void zm2_produce(unsigned char *ring, unsigned long long *full, unsigned long long *empty,
                                            const unsigned char *tab, const int *__restrict__ bs, int tz0, int tz1, int lane) {
  int kb = 0;
  int gs_next = (lane <= T0) ? bs[(size_t)tz0 * SUB + lane] : 0;
  for (int tz = tz0; tz < tz1; tz++) {
    const int gs = gs_next;
    if (tz + 1 < tz1) gs_next = (lane <= T0) ? bs[(size_t)(tz + 1) * SUB + lane] : 0;
    const int s = __shfl_sync(0xffffffffu, gs, 0), e = __shfl_sync(0xffffffffu, gs, T0);
    for (int c0 = s; c0 < e; c0 += GB, kb++) {
      const int st = kb % S;
      mbar_wait_park(&empty[st], (((unsigned)(kb / S)) & 1u) ^ 1u);
      const int cnt = min(GB, e - c0);
      unsigned char *sp = ring + (size_t)st * STAGE;
      int *h = reinterpret_cast<int *>(sp);
      if (lane <= T0) h[4 + lane] = min(max(gs - c0, 0), cnt);
      if (lane == 0) { h[0] = tz; h[1] = cnt; h[2] = c0; h[3] = 0; }
      __syncwarp();
      if (lane == 0) {
        const unsigned bytes = (unsigned)(cnt * ROWBYTES);
        mbar_expect_tx(&full[st], bytes);
        bulk_load_1d(sp + kZm2HdrBytes, tab + (size_t)c0 * ROWBYTES, bytes, &full[st]);
      }
    }
  }
  for (int k = 0; k < NEND; k++, kb++) {
    const int st = kb % S;
    mbar_wait_park(&empty[st], (((unsigned)(kb / S)) & 1u) ^ 1u);
    if (lane == 0) {
      int *h = reinterpret_cast<int *>(ring + (size_t)st * STAGE);
      h[0] = INT_MAX; h[1] = 0;
      mbar_arrive(&full[st]);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// staging layout: one TMA box is [XW][16][ZB] cells, z innermost; with 64- or 32-byte box rows the tensor map uses the
// matching shared-memory swizzle so that the 16-byte chunks a quarter warp touches fall into different banks
// ------------------------------------------------------------------------------------------------
template <int IB> __device__ __forceinline__ int zm2_swz(int byte_off) {
  constexpr int mask = IB == 64 ? 3 : (IB == 32 ? 1 : 0);
  return byte_off ^ (((byte_off >> 7) & mask) << 4);
}
template <class Cell, int ZB> __device__ __forceinline__ Cell *zm2_stg(void *buf, int rho, int k) {
  constexpr int IB = ZB * (int)sizeof(Cell);
  return reinterpret_cast<Cell *>(reinterpret_cast<unsigned char *>(buf) + zm2_swz<IB>(rho * IB + k * (int)sizeof(Cell)));
}
template <class Cell, int ZB> constexpr int zm2_swizzle_mode() {   // 0 none, 1 = 32 B, 2 = 64 B
  return ZB * (int)sizeof(Cell) == 64 ? 2 : (ZB * (int)sizeof(Cell) == 32 ? 1 : 0);
}

// The z weights of a node stream through a small rotating set of 16-byte register "quads" (2 doubles / 4 floats):
// slot j % D holds quad j; as soon as a quad is consumed its slot is refilled with quad j + D of the same node or, past
// the end of the row, with the first quads of the NEXT node, so the shared-memory latency never drains the pipeline.
template <class R> struct ZQuad;
template <> struct ZQuad<double> { typedef double2 type; static constexpr int PER = 2; };
template <> struct ZQuad<float> { typedef float4 type; static constexpr int PER = 4; };
__device__ __forceinline__ double zq_get(const double2 &q, int e) { return e == 0 ? q.x : q.y; }
__device__ __forceinline__ float zq_get(const float4 &q, int e) { return e == 0 ? q.x : (e == 1 ? q.y : (e == 2 ? q.z : q.w)); }
// 16-byte shared-memory load the compiler may not move past its neighbours of the same kind (keeps the weight refills
// where the source puts them: interleaved with the taps, not bunched at the end of the node loop)
#ifndef ZM2_ASM_LDS
#define ZM2_ASM_LDS 0
#endif
__device__ __forceinline__ double2 lds_pinned(const double2 *p) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ float4 lds_pinned(const float4 *p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
template <class Q> __device__ __forceinline__ Q zm2_ldq(const void *p) {
  if (ZM2_ASM_LDS) return lds_pinned(reinterpret_cast<const Q *>(p));
  return *reinterpret_cast<const Q *>(p);
}
template <int NQT, bool GRAD> constexpr int zm2_depth() {
  constexpr int dmax = ZM2_DEPTH;   // (two quads in flight measured 2 % slower for the gradient gather, splitting its sums into two chains 3 % slower)
  return (dmax >= 4 && NQT % 4 == 0) ? 4 : ((dmax >= 3 && NQT % 3 == 0) ? 3 : ((NQT % 2 == 0) ? 2 : 1));
}

// ------------------------------------------------------------------------------------------------
// scatter (adjoint B^T)
// ------------------------------------------------------------------------------------------------
template <class R, bool CPLX, int M_, bool GRAD>
__global__ void __launch_bounds__((Zm2Cfg<M_>::NCW + 1) * 32, 1)
k_scatter_zm2(const __grid_constant__ CUtensorMap tmap, Zm2Geom zg, const R *__restrict__ tab, const int *__restrict__ bin_start) {
  typedef typename CellT<R, CPLX>::type Cell;
  typedef Zm2Cfg<M_> Cfg;
  typedef Zm2Smem<R, CPLX, M_, GRAD> Sm;
  typedef typename Sm::RowS Row;
  constexpr int C = Cfg::C, T0 = Cfg::T0, T1 = Cfg::T1, ZS = Cfg::ZS, ZB = Cfg::ZB, W = Cfg::W, NCW = Cfg::NCW, XW = Cfg::XW;
  constexpr int NCOMP = CPLX ? 2 : 1, S = Sm::SS, GB = Sm::SGB, ROWBYTES = Row::ROWBYTES, STAGE = Sm::s_stage, ZP = Row::ZP;
  constexpr int SZ = (int)sizeof(R), WROWS = Sm::WARP_ROWS, BOXB = WROWS * ZB * (int)sizeof(Cell);

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *ring = smem_raw + Sm::s_off_ring;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + Sm::s_off_bar);
  unsigned long long *empty = full + S;

  const int colr = blockIdx.x / zg.nseg, seg = blockIdx.x - colr * zg.nseg, col = zg.col0 + colr;
  const int *bs = bin_start + (size_t)col * zg.nt2 * Cfg::SUB;
  int tz0, tz1;
  zm_segment(bs, zg.nt2, Cfg::SUB, seg, zg.nseg, zg.target, zg.fill, tz0, tz1);
  if (bs[(size_t)tz0 * Cfg::SUB] == bs[(size_t)tz1 * Cfg::SUB]) return;            // no nodes in this work item
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NCW) {
    zm2_produce<T0, Cfg::SUB, S, GB, ROWBYTES, STAGE, 1>(ring, full, empty, reinterpret_cast<const unsigned char *>(tab), bs, tz0, tz1, lane);
    return;
  }

  const int cx = col / zg.nc[1], cy = col - cx * zg.nc[1];
  const int o0 = cx * T0 + XW * warp, o1 = cy * T1;
  const int hlf = lane >> 4, r1 = lane & 15;
  constexpr int RPT = Cfg::RPT;
  const int rbase = XW * warp + RPT * hlf;               // the first of my RPT x rows
  const int rho0 = (RPT * hlf) * 16 + r1;                // its row inside the warp's TMA box (the next one is 16 further)
  const int aX = (Row::oX + Cfg::XLEAD + rbase) * SZ, aY = (Row::oY + T1 - 1 + r1) * SZ;
  constexpr int dOff = (Row::oDX - Row::oX) * SZ;
  const int dxlo = max(0, XW * warp - (C - 1)), dxhi1 = min(T0 - 1, XW * warp + XW - 1) + 1;
  unsigned char *mystg = smem_raw + (size_t)warp * 2 * Sm::WARP_BOX;

  Cell win[RPT][W];
#pragma unroll
  for (int e = 0; e < RPT; e++)
#pragma unroll
    for (int i = 0; i < W; i++) zero_cell(win[e][i]);
  int cur = tz0, dirty = 0, nfl = 0;

  // the first ZS cells of the windows are final: reduce-add them into the grid, advance the windows
  auto flush_advance = [&]() {
    unsigned char *sb = mystg + (nfl & 1) * Sm::WARP_BOX;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the box issued two flushes ago was read
    __syncwarp();
#pragma unroll
    for (int q = 0; q < ZS; q++)
#pragma unroll
      for (int e = 0; e < RPT; e++) *zm2_stg<Cell, ZB>(sb + (q / ZB) * BOXB, rho0 + 16 * e, q % ZB) = win[e][q];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int b = 0; b < ZS / ZB; b++) tma_reduce_add_3d(sb + b * BOXB, &tmap, (cur * ZS + b * ZB) * NCOMP, o1, o0);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    nfl++;
#pragma unroll
    for (int e = 0; e < RPT; e++) {
#pragma unroll
      for (int i = 0; i < W - ZS; i++) win[e][i] = win[e][i + ZS];
#pragma unroll
      for (int i = W - ZS; i < W; i++) zero_cell(win[e][i]);
    }
    cur++;
  };

  // operands of one node that depend on this thread's rows: x and y weights (and derivatives), node values
  struct Ops { R w0[RPT], w1, dw0[RPT], dw1; Cell f, g0, g1, g2; };
  auto fetch = [&](const unsigned char *row, const int4 &hd, Ops &o) {
#pragma unroll
    for (int e = 0; e < RPT; e++) o.w0[e] = *reinterpret_cast<const R *>(row + aX + e * SZ + hd.x);
    o.w1 = *reinterpret_cast<const R *>(row + aY + hd.y);
    const R *v = reinterpret_cast<const R *>(row) + Row::oV;
    Cell z; zero_cell(z);
    o.f = load_in(v, z);
    if (GRAD) {
#pragma unroll
      for (int e = 0; e < RPT; e++) o.dw0[e] = *reinterpret_cast<const R *>(row + aX + e * SZ + dOff + hd.x);
      o.dw1 = *reinterpret_cast<const R *>(row + aY + dOff + hd.y);
      o.g0 = load_in(v + NCOMP, z); o.g1 = load_in(v + 2 * NCOMP, z); o.g2 = load_in(v + 3 * NCOMP, z);
    }
  };

  for (int kb = 0;; kb++) {
    const int st = kb % S;
    mbar_wait_park(&full[st], ((unsigned)(kb / S)) & 1u);
    const unsigned char *sp = ring + (size_t)st * STAGE;
    const int *h = reinterpret_cast<const int *>(sp);
    const int tz = h[0];
    if (tz == INT_MAX) break;
    const int lo = h[4 + dxlo], hi = h[4 + dxhi1];
    if (hi > lo) {
      while (cur < tz) {
        if (dirty > 0) { flush_advance(); dirty--; }
        else cur = tz;
      }
      const unsigned char *row = sp + kZm2HdrBytes + (size_t)lo * ROWBYTES;
      const unsigned char *last = sp + kZm2HdrBytes + (size_t)(hi - 1) * ROWBYTES;
      // software pipeline: header two nodes ahead, thread-dependent operands one node ahead, z weights streamed
      typedef typename ZQuad<R>::type Quad;
      constexpr int PER = ZQuad<R>::PER, NQT = ZP / PER, D = zm2_depth<NQT, true>();   // two quads in flight: the windows leave no room for more
      int4 hd = *reinterpret_cast<const int4 *>(row);
      const unsigned char *row1 = row + ROWBYTES < last ? row + ROWBYTES : last;
      int4 hn = *reinterpret_cast<const int4 *>(row1);
      Ops op;
      fetch(row, hd, op);
      Quad Q[D], DQ[D];
#pragma unroll
      for (int j = 0; j < D; j++) {
        Q[j] = *reinterpret_cast<const Quad *>(row + (Row::oZ + j * PER) * SZ);
        if (GRAD) DQ[j] = *reinterpret_cast<const Quad *>(row + (Row::oDZ + j * PER) * SZ);
      }
      for (int i = lo; i < hi; i++) {
        // per-row amplitudes: A_e = w0_e (w1 f + dw1 g1) + dw0_e (w1 g0), B_e = w0_e (w1 g2)
        Cell u = scale_cell(op.w1, op.f), A[RPT], B[RPT];
        if (GRAD) {
          fma_cell(u, op.dw1, op.g1);
          const Cell v = scale_cell(op.w1, op.g0), sg = scale_cell(op.w1, op.g2);
#pragma unroll
          for (int e = 0; e < RPT; e++) { A[e] = scale_cell(op.w0[e], u); fma_cell(A[e], op.dw0[e], v); B[e] = scale_cell(op.w0[e], sg); }
        } else {
#pragma unroll
          for (int e = 0; e < RPT; e++) { A[e] = scale_cell(op.w0[e], u); zero_cell(B[e]); }
        }
        // prefetch
        const unsigned char *row2 = row1 + ROWBYTES < last ? row1 + ROWBYTES : last;
        const int4 hn2 = *reinterpret_cast<const int4 *>(row2);
        fetch(row1, hn, op);
        auto ztaps = [&](auto dz_tag) {
          constexpr int DZ = decltype(dz_tag)::value;     // first window cell of the node's taps
          constexpr int NT = Cfg::DZS ? C : W;
#pragma unroll
          for (int j = 0; j < NQT; j++) {
            const Quad w = Q[j % D];
            Quad dw = w;
            if (GRAD) dw = DQ[j % D];
            const unsigned char *src = (j + D < NQT) ? row : row1;
            const int jq = (j + D < NQT) ? j + D : j + D - NQT;
            Q[j % D] = *reinterpret_cast<const Quad *>(src + (Row::oZ + jq * PER) * SZ);
            if (GRAD) DQ[j % D] = *reinterpret_cast<const Quad *>(src + (Row::oDZ + jq * PER) * SZ);
#pragma unroll
            for (int e = 0; e < PER; e++) {
              const int k = j * PER + e;
              if (k < NT) {
#pragma unroll
                for (int r = 0; r < RPT; r++) {
                  fma_cell(win[r][DZ + k], zq_get(w, e), A[r]);
                  if (GRAD) fma_cell(win[r][DZ + k], zq_get(dw, e), B[r]);
                }
              }
            }
          }
        };
        if constexpr (Cfg::DZS) {
          const int dz = hd.z;
#define ZM2_CALL(d_) ztaps(std::integral_constant<int, (d_)>())
          ZM2_TREE(dz, ZM2_CALL);
#undef ZM2_CALL
        } else {
          ztaps(std::integral_constant<int, 0>());
        }
        hd = hn; hn = hn2; row = row1; row1 = row2;
      }
      dirty = Cfg::NFL;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }
  while (dirty > 0) { flush_advance(); dirty--; }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging must outlive the bulk reads
}

// ------------------------------------------------------------------------------------------------
// gather (trafo B)
// ------------------------------------------------------------------------------------------------
// sum NV values over the 32 lanes with NV-1 + log2(32/NV) exchanges; the lane whose upper bits spell k ends up with
// the total of value k (returned in v[0], k in *which); all lanes sharing those upper bits hold the same total
template <int NV, int WIDTH, class R> __device__ __forceinline__ void lanes_reduce(R (&v)[NV], int lane, int *which) {
  int w = 0;
  int bit = WIDTH / 2;
#pragma unroll
  for (int n = NV; n > 1; n >>= 1, bit >>= 1) {
    const bool hi = (lane & bit) != 0;
#pragma unroll
    for (int k = 0; k < n / 2; k++) {
      const R send = hi ? v[k] : v[k + n / 2];
      const R keep = hi ? v[k + n / 2] : v[k];
      v[k] = keep + shfl_xor(send, bit);
    }
    w += hi ? n / 2 : 0;
  }
  for (; bit > 0; bit >>= 1) v[0] += shfl_xor(v[0], bit);
  *which = w;
}

// x-y contraction of the per-row partial sums of NPP nodes by one warp, LPN = 32 / NPP lanes per node.  Lane l of a
// group walks the node's (2m+1) x 16 partials LPN at a time: entry e = LPN it + l is x row e / 16 (the same for the
// whole group) and y row e % 16, so a lane meets 16 / LPN y weights only and applies them once at the end.  The
// LPN-lane reduction then leaves one output value per lane group slot (lanes_reduce).
template <int NPP, class R, bool CPLX, int M_, bool GRAD, class Row, class Cell>
__device__ __forceinline__ void zm2_reduce_quad(const unsigned char *row, const Cell *pp, bool live, int lane, const GatherOut<R> &out) {
  typedef Zm2Cfg<M_> Cfg;
  constexpr int C = Cfg::C, T1 = Cfg::T1, NCOMP = CPLX ? 2 : 1, SZ = (int)sizeof(R);
  constexpr int LPN = 32 / NPP;          // lanes per node: 8 (four nodes per pass) or 16 (two nodes per pass)
  constexpr int NY = 16 / LPN;           // y weights a lane meets
  static_assert(LPN == 8 || LPN == 16, "two or four nodes per pass");
  const int l = lane & (LPN - 1);
  const int *hd = reinterpret_cast<const int *>(row);
  const int ny = hd[1], j = hd[4];
  const R *rr = reinterpret_cast<const R *>(row);
  R w1[NY], dw1[NY];
#pragma unroll
  for (int b = 0; b < NY; b++) {
    w1[b] = *reinterpret_cast<const R *>(row + (Row::oY + T1 - 1 + LPN * b + l) * SZ + ny);
    dw1[b] = GRAD ? *reinterpret_cast<const R *>(row + (Row::oDY + T1 - 1 + LPN * b + l) * SZ + ny) : (R)0;
  }
  Cell s[NY], sd[NY], u[NY];
#pragma unroll
  for (int b = 0; b < NY; b++) { zero_cell(s[b]); zero_cell(sd[b]); zero_cell(u[b]); }
#pragma unroll
  for (int it = 0; it < NY * C; it++) {
    const int i0 = it / NY, b = it % NY;
    const R w0 = rr[Row::oX + Cfg::XLEAD + i0];
    const Cell t = pp[LPN * it + l];
    fma_cell(s[b], w0, t);
    if (GRAD) {
      const R dw0 = rr[Row::oDX + Cfg::XLEAD + i0];
      const Cell td = pp[C * 16 + LPN * it + l];
      fma_cell(sd[b], dw0, t);
      fma_cell(u[b], w0, td);
    }
  }
  constexpr int NVAL = NCOMP * (GRAD ? 4 : 1);
  R v[NVAL];
  Cell af = scale_cell(w1[0], s[0]), a0, a1, a2;
  zero_cell(a0); zero_cell(a1); zero_cell(a2);
  if (GRAD) { a0 = scale_cell(w1[0], sd[0]); a1 = scale_cell(dw1[0], s[0]); a2 = scale_cell(w1[0], u[0]); }
#pragma unroll
  for (int b = 1; b < NY; b++) {
    fma_cell(af, w1[b], s[b]);
    if (GRAD) { fma_cell(a0, w1[b], sd[b]); fma_cell(a1, dw1[b], s[b]); fma_cell(a2, w1[b], u[b]); }
  }
  if constexpr (CPLX) {
    v[0] = af.x; v[1] = af.y;
    if constexpr (GRAD) { v[2] = a0.x; v[3] = a0.y; v[4] = a1.x; v[5] = a1.y; v[6] = a2.x; v[7] = a2.y; }
  } else {
    v[0] = af;
    if constexpr (GRAD) { v[1] = a0; v[2] = a1; v[3] = a2; }
  }
  int which;
  lanes_reduce<NVAL, LPN>(v, lane, &which);
  if (live && (l & (LPN / NVAL - 1)) == 0) {
    R *o = nullptr;
    if (which < NCOMP) { if (out.f) o = out.f + ((size_t)j * out.f_stride + out.f_off) * NCOMP + which; }
    else if (GRAD) o = out.grad + (size_t)j * 3 * NCOMP + (which - NCOMP);
    if (o) *o = out.accumulate ? *o + v[0] : v[0];
  }
}

#ifdef ZM2_TIMING
__device__ long long *g_zm2_timing = nullptr;   // [warp][6]: wait full, wait pempty, window advance, node loop, arrive+help, loop top
#endif
template <class R, bool CPLX, int M_, bool GRAD>
__global__ void __launch_bounds__((Zm2Cfg<M_>::NCW + 1) * 32, 1)
k_gather_zm2(const __grid_constant__ CUtensorMap tmap, Zm2Geom zg, const R *__restrict__ tab, const int *__restrict__ bin_start,
             GatherOut<R> out) {
  typedef typename CellT<R, CPLX>::type Cell;
  typedef Zm2Cfg<M_> Cfg;
  typedef Zm2Smem<R, CPLX, M_, GRAD> Sm;
  typedef typename Sm::RowG Row;
  constexpr int C = Cfg::C, T0 = Cfg::T0, T1 = Cfg::T1, ZS = Cfg::ZS, ZB = Cfg::ZB, W = Cfg::W, NCW = Cfg::NCW, XW = Cfg::XW;
  constexpr int NCOMP = CPLX ? 2 : 1, S = Sm::GS, P = Sm::GP, GB = Sm::GGB, ROWBYTES = Row::ROWBYTES, STAGE = Sm::g_stage, PN = Sm::PN;
  constexpr int ZP = Row::ZP, WROWS = Sm::WARP_ROWS;
  constexpr unsigned BOXB = WROWS * ZB * (unsigned)sizeof(Cell);
  static_assert(S >= 3 && P >= 2, "ring / partial stages too shallow for the deferred reduction");

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *ring = smem_raw + Sm::g_off_ring;
  Cell *part = reinterpret_cast<Cell *>(smem_raw + Sm::g_off_part);
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + Sm::g_off_bar);
  unsigned long long *empty = full + S;
  unsigned long long *pfull = empty + S;
  unsigned long long *pempty = pfull + P;
  unsigned long long *wbar = pempty + P;

  const int colr = blockIdx.x / zg.nseg, seg = blockIdx.x - colr * zg.nseg, col = zg.col0 + colr;
  const int *bs = bin_start + (size_t)col * zg.nt2 * Cfg::SUB;
  int tz0, tz1;
  zm_segment(bs, zg.nt2, Cfg::SUB, seg, zg.nseg, zg.target, zg.fill, tz0, tz1);
  if (bs[(size_t)tz0 * Cfg::SUB] == bs[(size_t)tz1 * Cfg::SUB]) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], NCW); }
    for (int i = 0; i < P; i++) { mbar_init(&pfull[i], NCW); mbar_init(&pempty[i], NCW); }
    for (int i = 0; i < NCW; i++) mbar_init(&wbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NCW) {
    zm2_produce<T0, Cfg::SUB, S, GB, ROWBYTES, STAGE, 1>(ring, full, empty, reinterpret_cast<const unsigned char *>(tab), bs, tz0, tz1, lane);
    return;
  }

  // ---- consumer warps ----
  const int cx = col / zg.nc[1], cy = col - cx * zg.nc[1];
  const int o0 = cx * T0 + XW * warp, o1 = cy * T1;
  const int hlf = lane >> 4, r1 = lane & 15;
  constexpr int RPT = Cfg::RPT;
  const int rbase = XW * warp + RPT * hlf;
  const int rho0 = (RPT * hlf) * 16 + r1;
  const int dxlo = max(0, XW * warp - (C - 1)), dxhi1 = min(T0 - 1, XW * warp + XW - 1) + 1;
  unsigned char *mystg = smem_raw + (size_t)warp * Sm::WARP_BOX;
  unsigned long long *mybar = &wbar[warp];
  unsigned wph = 0;

  Cell win[RPT][W];
  int cur = INT_MIN / 2;       // sub-chunk whose cells [cur*ZS, cur*ZS + W) are in the windows
  bool pending = false;        // a load of the cells [cur*ZS + W, +ZS) is in flight

  auto issue = [&](int zc, int nbox) {       // lane 0: nbox TMA boxes [XW][16][ZB] starting at cell zc
    if (lane == 0) {
      mbar_expect_tx(mybar, BOXB * (unsigned)nbox);
      for (int b = 0; b < nbox; b++) tma_load_3d(mystg + b * BOXB, &tmap, (zc + b * ZB) * NCOMP, o1, o0, mybar);
    }
  };
  // wait for the staged boxes and copy my rows' cells into the windows.  The vote consumes the loaded values, so every
  // lane's shared-memory reads have RETURNED before lane 0 may re-arm the staging buffer with the next TMA load.
  auto take = [&](auto first_tag, auto nbox_tag) {
    constexpr int FIRST = decltype(first_tag)::value, NBOX = decltype(nbox_tag)::value;
    mbar_wait(mybar, wph);
    wph ^= 1u;
    int nan = 0;
#pragma unroll
    for (int b = 0; b < NBOX; b++)
#pragma unroll
      for (int q = 0; q < ZB; q++)
#pragma unroll
        for (int e = 0; e < RPT; e++) {
          win[e][FIRST + b * ZB + q] = *zm2_stg<Cell, ZB>(mystg + b * BOXB, rho0 + 16 * e, q);
          nan |= cell_is_nan(win[e][FIRST + b * ZB + q]);
        }
    // No cross-proxy fence here: the staged cells were only READ by this warp, and the vote below consumes the loaded
    // values, so every read has returned before lane 0 can issue the TMA load that overwrites the buffer (the same
    // read -> release -> TMA-overwrite order every mbarrier pipeline relies on).  The fence cost 4 % of the gradient
    // gather (measured: 40.3 -> 38.5 ms); -DZM2_TAKE_FENCE restores it.
#if defined(ZM2_TAKE_FENCE)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
    (void)warp_any_volatile(nan);
  };
  auto drop_pending = [&]() {
    if (pending) { mbar_wait(mybar, wph); wph ^= 1u; pending = false; }
  };
  auto reload = [&](int tz) {      // whole windows at sub-chunk tz, then prefetch the next advance
    drop_pending();
    const int z0 = tz * ZS;
    constexpr int PER = ZS / ZB;   // boxes per staging buffer
    constexpr int NB = W / ZB;     // boxes of the whole window
    static_assert(NB <= 6 * PER, "window reload written for at most six rounds");
#define ZM2_ROUND(r)                                                                                                   \
    if constexpr ((r) * PER < NB) {                                                                                    \
      constexpr int nb_ = (NB - (r) * PER) < PER ? (NB - (r) * PER) : PER;                                             \
      issue(z0 + (r) * ZS, nb_);                                                                                       \
      take(std::integral_constant<int, ((r) * PER < NB) ? (r) * ZS : 0>(), std::integral_constant<int, nb_>());        \
    }
    ZM2_ROUND(0) ZM2_ROUND(1) ZM2_ROUND(2) ZM2_ROUND(3) ZM2_ROUND(4) ZM2_ROUND(5)
#undef ZM2_ROUND
    cur = tz;
    issue(z0 + W, PER);
    pending = true;
  };
  auto advance1 = [&]() {
#pragma unroll
    for (int i = 0; i < W - ZS; i++)
#pragma unroll
      for (int e = 0; e < RPT; e++) win[e][i] = win[e][i + ZS];
    take(std::integral_constant<int, W - ZS>(), std::integral_constant<int, ZS / ZB>());
    cur++;
    issue(cur * ZS + W, ZS / ZB);
  };

  // the warps whose rows meet every node of the tile are the critical path: they leave the reduction to the others
  const bool helper = (dxhi1 - dxlo) < T0;
  // deferred x-y reduction of chunk kr: wait until every warp has left its partial sums, then take nodes from the
  // chunk's ticket counter (in its ring header) until none are left; releases the partial stage and the ring stage
  auto help = [&](int kr) {
    const int st = kr % S, ps = kr % P;
    unsigned char *sp = ring + (size_t)st * STAGE;
    int *h = reinterpret_cast<int *>(sp);
    // (the critical warps neither reduce nor wait: their arrivals below only count them out of the two stages)
    if (helper) {
      mbar_wait_park(&pfull[ps], ((unsigned)(kr / P)) & 1u, 50u);   // the reduction latency is on the critical path
      const int cnt = h[1];
      for (;;) {
        int i = 0;
        constexpr int NPP = 4;     // nodes per warp pass (2: half a warp each, measured no faster)
        if (lane == 0) i = atomicAdd(&h[3], NPP);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= cnt) break;
        const int mine = i + lane / (32 / NPP);
        const int ic = mine < cnt ? mine : cnt - 1;
        zm2_reduce_quad<NPP, R, CPLX, M_, GRAD, Row>(sp + kZm2HdrBytes + (size_t)ic * ROWBYTES, part + (size_t)(ps * GB + ic) * PN, mine < cnt, lane, out);
      }
    }
    __syncwarp();
    if (lane == 0) { mbar_arrive(&pempty[ps]); mbar_arrive(&empty[st]); }
  };

  const int aP = (rbase * 16 + r1);
  int kb = 0;
#ifdef ZM2_TIMING
  long long tq[6] = {0, 0, 0, 0, 0, 0}, tc = clock64(), tit = 0, itsum = 0, itn = 0, itmin = 1LL << 40;
#define ZM2_T(i) do { const long long now_ = clock64(); tq[i] += now_ - tc; tc = now_; } while (0)
#else
#define ZM2_T(i) do { } while (0)
#endif
  for (;; kb++) {
    const int st = kb % S, ps = kb % P;
    ZM2_T(5);
    mbar_wait_park(&full[st], ((unsigned)(kb / S)) & 1u);
    ZM2_T(0);
    const unsigned char *sp = ring + (size_t)st * STAGE;
    const int *h = reinterpret_cast<const int *>(sp);
    const int tz = h[0];
    if (tz == INT_MAX) break;
    const int lo = h[4 + dxlo], hi = h[4 + dxhi1];
    mbar_wait_park(&pempty[ps], (((unsigned)(kb / P)) & 1u) ^ 1u);
    ZM2_T(1);
    if (hi > lo) {
      if (tz != cur) {
        if (pending && tz == cur + 1) advance1();
        else if (pending && tz == cur + 2) { advance1(); advance1(); }
        else reload(tz);
      }
      ZM2_T(2);
      const unsigned char *row = sp + kZm2HdrBytes + (size_t)lo * ROWBYTES;
      const unsigned char *last = sp + kZm2HdrBytes + (size_t)(hi - 1) * ROWBYTES;
      Cell *pb = part + (size_t)(ps * GB + lo) * PN + aP;
      typedef typename ZQuad<R>::type Quad;
      constexpr int SZ = (int)sizeof(R), PER = ZQuad<R>::PER, NQT = ZP / PER, D = zm2_depth<NQT, GRAD>();
      int4 hd = *reinterpret_cast<const int4 *>(row);
      Quad Q[D], DQ[D];
#pragma unroll
      for (int j = 0; j < D; j++) {
        Q[j] = *reinterpret_cast<const Quad *>(row + (Row::oZ + j * PER) * SZ);
        if (GRAD) DQ[j] = *reinterpret_cast<const Quad *>(row + (Row::oDZ + j * PER) * SZ);
      }
      for (int i = lo; i < hi; i++, pb += PN) {
#ifdef ZM2_TIMING
        { const long long now_ = clock64(); if (i > lo) { const long long d_ = now_ - tit; itsum += d_; itn++; if (d_ < itmin) itmin = d_; } tit = now_; }
#endif
        const unsigned char *row1 = row + ROWBYTES < last ? row + ROWBYTES : last;
        const int4 hn = *reinterpret_cast<const int4 *>(row1);       // prefetch the next header
        Cell t[RPT], td[RPT], t1[RPT], td1[RPT];       // F only: two chains per sum; with the gradient the sums are chains enough
#pragma unroll
        for (int r = 0; r < RPT; r++) { zero_cell(t[r]); zero_cell(td[r]); zero_cell(t1[r]); zero_cell(td1[r]); }
        auto ztaps = [&](auto dz_tag) {
          constexpr int DZ = decltype(dz_tag)::value;     // first window cell of the node's taps
          constexpr int NT = Cfg::DZS ? C : W;
#pragma unroll
          for (int j = 0; j < NQT; j++) {
            const Quad w = Q[j % D];
            Quad dw = w;
            if (GRAD) dw = DQ[j % D];
            const unsigned char *src = (j + D < NQT) ? row : row1;
            const int jq = (j + D < NQT) ? j + D : j + D - NQT;
            Q[j % D] = zm2_ldq<Quad>(src + (Row::oZ + jq * PER) * SZ);
            if (GRAD) DQ[j % D] = zm2_ldq<Quad>(src + (Row::oDZ + jq * PER) * SZ);
#pragma unroll
            for (int e = 0; e < PER; e++) {
              const int k = j * PER + e;
              if (k < NT) {
#pragma unroll
                for (int r = 0; r < RPT; r++) {
                  if (GRAD && ZM2_GCHAINS == 2) {
                    if (k & 1) { fma_cell(t1[r], zq_get(w, e), win[r][DZ + k]); fma_cell(td1[r], zq_get(dw, e), win[r][DZ + k]); }
                    else { fma_cell(t[r], zq_get(w, e), win[r][DZ + k]); fma_cell(td[r], zq_get(dw, e), win[r][DZ + k]); }
                  }
                  else if (GRAD) { fma_cell(t[r], zq_get(w, e), win[r][DZ + k]); fma_cell(td[r], zq_get(dw, e), win[r][DZ + k]); }
                  else if (k & 1) fma_cell(t1[r], zq_get(w, e), win[r][DZ + k]);
                  else fma_cell(t[r], zq_get(w, e), win[r][DZ + k]);
                }
              }
            }
          }
        };
        if constexpr (Cfg::DZS) {
          const int dz = hd.z;
#define ZM2_CALL(d_) ztaps(std::integral_constant<int, (d_)>())
          ZM2_TREE(dz, ZM2_CALL);
#undef ZM2_CALL
        } else {
          ztaps(std::integral_constant<int, 0>());
        }
        const int i0 = rbase - hd.w;
        Cell *p = pb - hd.w * 16;
#pragma unroll
        for (int r = 0; r < RPT; r++) {
          if (!GRAD || ZM2_GCHAINS == 2) add_cell(t[r], t1[r]);
          if (GRAD && ZM2_GCHAINS == 2) add_cell(td[r], td1[r]);
          if ((unsigned)(i0 + r) < (unsigned)C) { p[16 * r] = t[r]; if (GRAD) p[C * 16 + 16 * r] = td[r]; }
        }
        hd = hn; row = row1;
      }
    }
    ZM2_T(3);
    __syncwarp();
    if (lane == 0) mbar_arrive(&pfull[ps]);
    if (kb >= 1) help(kb - 1);
    ZM2_T(4);
  }
  if (kb >= 1) help(kb - 1);
#ifdef ZM2_TIMING
  if (lane == 0 && g_zm2_timing) {
    for (int i = 0; i < 6; i++) atomicAdd((unsigned long long *)&g_zm2_timing[warp * 6 + i], (unsigned long long)tq[i]);
    atomicAdd((unsigned long long *)&g_zm2_timing[72 + warp * 2], (unsigned long long)itsum);
    atomicAdd((unsigned long long *)&g_zm2_timing[72 + warp * 2 + 1], (unsigned long long)itn);
    atomicMin((long long *)&g_zm2_timing[96 + warp], itmin);
  }
#endif
  drop_pending();    // no TMA write may still be in flight when the CTA retires
}

}  // namespace pnb
