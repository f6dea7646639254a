// Tensor-core gather for the cutoffs whose window does not fit the register file (v4: m = 5, 7, 8 in double; the
// clustered BASELINE config 4 runs Gaussian / B-spline windows with m = 8).
// Reference loops replaced: kernel/assign.c:667-1130 (assign_f / assign_f_and_grad_f), kernel/ndft-parallel.c:2703-2888.
//
// v3 (zmarch3.cuh) keeps every grid row of a column tile in the registers of the warp that owns it; at m = 8 a tile's
// footprint is (T + 16)^2 rows x 20 z cells and no longer fits beside the MMA fragments.  v4 keeps the window in SHARED
// memory instead: a ring of >= KS + 1 z sub-chunks of the footprint [R0][R1][4] cells, each brought in by ONE TMA box, and a
// warp reads its B fragments from there (one conflict-free LDS.64 per DMMA pair; +4 % issue time in the cost model of
// profiles/r2_dmma_overlap_ubench.log).  With the window shared, rows have no owner:
//   * whole node batches (8 nodes, consecutive in the chunk's dx order) go to whichever warp asks next (one ticket counter
//     per CTA); the warp contracts z for exactly the x rows its batch touches and ALL y rows on the tensor cores
//     (T[node, column] = sum_k psi_z[node, k] * window[k, column], M = 8 nodes, N = 8 columns, K = 4 cells per DMMA), runs
//     the x-y contraction on its C fragments and writes the finished nodes out: no cross-warp reduction, no partial
//     stages, and clustered node sets balance themselves inside a CTA;
//   * warps may be at different z sub-chunks at the same time: a window chunk is re-used once every consumer warp has
//     walked past it (one mbarrier per slot, NW arrivals), and every warp derives the load sequence (and with it the wait
//     parity of every slot) from the bin table itself, exactly as the service warp that issues the loads does;
//   * node-table rows (ZmRowOf over Zm4Cfg, built by k_node_table2) are read straight from global memory: a batch keeps a
//     warp busy for >= 10^4 cycles, the 20 row loads in front of it and the two x weights fetched one row ahead cost nothing.
// The sort key is v2 / v3's (column tile, z sub-chunk, dx) on the column tile of the v1 kernels of the same cutoff (8 x 4
// cells at m = 8), so the v1 scatter of the adjoint runs on the same bins (ZmGeom::sub).
#pragma once
#include "zmarch3.cuh"

namespace pnb {

template <int M_> struct Zm4Ok { static constexpr bool value = (M_ == 5 || M_ == 7 || M_ == 8); };

template <int M_> struct Zm4Cfg {
  static constexpr int C = 2 * M_ + 1;
  static constexpr int T0 = 8, T1 = 4, ZS = 4;   // column tile and z sub-chunk (cells)
  static constexpr int SUB = 8;                  // x-offset bins per tile in the sort key
  static constexpr int KS = (ZS + 2 * M_ + 3) / 4;   // window chunks = k-steps per row
  static constexpr int W = (ZS + 2 * M_ + 7) / 8 * 8;   // z weights per node, zero padded (pre-shifted by the node's dz): the
                                                 // gather reads 4 KS of them, the scatter's circular accumulator window all W
  static constexpr int NFL = (ZS + 2 * M_ + ZS - 1) / ZS;   // flushes until a touched accumulator window is all zero again
  static constexpr int R0 = T0 + 2 * M_;         // footprint rows along x
  static constexpr int R1C = (T1 + 2 * M_ + 3) / 4 * 4, R1R = (T1 + 2 * M_ + 7) / 8 * 8;   // y rows in whole n-blocks (c2c / c2r)
  static constexpr int XLEAD = T0 - 1;
  static constexpr int YLEAD = (R1R - C > T1 - 1) ? R1R - C : T1 - 1;   // a lane reads the y weight of every row of its n-blocks
  static constexpr bool DZS = false;
  static constexpr int NW = 15;                  // consumer warps (+ 1 service warp = 512 threads, 128 registers)
};

// The shared-window gather on v3's geometry (m = 6: column tile 10 x 4, 16 x-offset bins, v3's node-table rows, so that it
// pairs with v3's scatter on one binning and one table).  v3's rows carry ONE leading zero in front of the x weights, not
// T0 - 1: the gather tests the tap range instead of reading padding (Zm4Smem::XPRED).
template <int M_> struct Zm4on2Cfg : Zm2Cfg<M_> {
  typedef Zm2Cfg<M_> B;
  static constexpr int KS = (B::ZS + 2 * M_ + 3) / 4;
  static constexpr int R1C = (B::T1 + 2 * M_ + 3) / 4 * 4, R1R = (B::T1 + 2 * M_ + 7) / 8 * 8;
  static constexpr int NW = 19;                  // consumer warps (+ 1 service warp = 640 threads, 102 registers)
  static_assert(R1R - B::C <= B::YLEAD, "y weights of every footprint row must lie inside the row's y section");
};

template <bool CPLX, int M_, class Cfg_ = Zm4Cfg<M_>> struct Zm4Smem {
  typedef Cfg_ Cfg;
  static constexpr bool XPRED = Cfg::XLEAD < Cfg::T0 - 1;
  static constexpr int NCOMP = CPLX ? 2 : 1, CELLB = 8 * NCOMP;
  static constexpr int R1 = CPLX ? Cfg::R1C : Cfg::R1R;
  static constexpr int NYB = R1 * NCOMP / 8;     // n-blocks per x row
  static constexpr int SLOTB = Cfg::R0 * R1 * Cfg::ZS * CELLB;
  // window ring: the chunks of one window plus as many more as shared memory holds (at most 16).  The consumer warps may
  // be spread over NSLOT - KS + 1 window positions at a time; with few nodes per sub-chunk (uniform node sets: 2-3 batches
  // per position) that spread is what keeps them busy
  static constexpr int NSLOT_FIT = (232448 - 1024) / SLOTB;
  static constexpr int NSLOT = NSLOT_FIT > 16 ? 16 : (NSLOT_FIT < Cfg::KS + 1 ? Cfg::KS + 1 : NSLOT_FIT);
  static constexpr int off_bar = NSLOT * SLOTB;
  static constexpr int gather = off_bar + (2 * NSLOT + 2) * 8;
  static_assert(SLOTB % 128 == 0, "alignment");
  static_assert(gather <= 232448, "shared-memory budget of one CTA exceeded");
  // scatter: a CTA owns one y half of the footprint (NYB0 n-blocks of whole rows, the second half may be shorter), one x
  // row per consumer warp; two staging boxes [1][RH][ZS] per warp, ring of SS stages {header, SGB rows, SGB value rows}
  static constexpr int NYB0 = (NYB + 1) / 2, NYB1 = NYB - NYB0;
  static constexpr int RH = NYB0 * 8 / NCOMP;    // y rows of a half's TMA box
  static constexpr int NCW = Cfg::R0;            // consumer warps
  static constexpr int WARP_BOX = RH * Cfg::ZS * CELLB;
  static constexpr int SS = 3, SGB = 16;
  template <bool GRAD, bool RG> struct Scat {
    typedef ZmRowOf<double, Cfg, RG, false, CPLX> Row;
    static constexpr int NVAL = NCOMP * (GRAD ? 4 : 1), NVP = (NVAL + 1) / 2 * 2;
    static constexpr int off_vals = kZm2HdrBytes + SGB * Row::ROWBYTES;
    static constexpr int stage = off_vals + SGB * NVP * 8;
    static constexpr int off_ring = (NCW * 2 * WARP_BOX + 127) / 128 * 128;
    static constexpr int off_bar = off_ring + SS * stage;
    static constexpr int bytes = off_bar + 2 * SS * 8;
    static_assert(stage % 16 == 0 && off_vals % 16 == 0 && WARP_BOX % 128 == 0, "alignment");
    static_assert(bytes <= 232448, "shared-memory budget of one CTA exceeded");
  };
};

__device__ __forceinline__ double ldg_f64(const unsigned char *p) { return __ldg(reinterpret_cast<const double *>(p)); }

// ------------------------------------------------------------------------------------------------
// gather (trafo B)
// ------------------------------------------------------------------------------------------------
template <bool CPLX, int M_, bool GRAD, bool RG, class Cfg_ = Zm4Cfg<M_>>
__global__ void __launch_bounds__((Cfg_::NW + 1) * 32, 1)
k_gather_mma4(const __grid_constant__ CUtensorMap tmap, Zm2Geom zg, const double *__restrict__ tab, const int *__restrict__ bin_start,
              GatherOut<double> out) {
  static_assert(RG || !GRAD, "a gradient kernel needs rows with derivative sections");
  typedef Cfg_ Cfg;
  typedef Zm4Smem<CPLX, M_, Cfg_> Sm;
  typedef ZmRowOf<double, Cfg, RG, false, CPLX> Row;
  constexpr int C = Cfg::C, ZS = Cfg::ZS, KS = Cfg::KS, NW = Cfg::NW, NSLOT = Sm::NSLOT, SUB = Cfg::SUB;
  constexpr int NCOMP = Sm::NCOMP, CELLB = Sm::CELLB, R1 = Sm::R1, NYB = Sm::NYB, SLOTB = Sm::SLOTB;
  constexpr int NVAL = NCOMP * (GRAD ? 4 : 1), ROWBYTES = Row::ROWBYTES;
  constexpr int NWY = CPLX ? NYB : 2 * NYB;        // y weights a lane needs: one per C-fragment row it holds
  constexpr int dOff = RG ? (Row::oDX - Row::oX) * 8 : 0;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *win = smem_raw;
  unsigned long long *wfull = reinterpret_cast<unsigned long long *>(smem_raw + Sm::off_bar);
  unsigned long long *wempty = wfull + NSLOT;
  int *ticket = reinterpret_cast<int *>(wempty + NSLOT);

  const int colr = blockIdx.x / zg.nseg, seg = blockIdx.x - colr * zg.nseg, col = zg.col0 + colr;
  const int *bs = bin_start + (size_t)col * zg.nt2 * SUB;
  int tz0, tz1;
  zm_segment(bs, zg.nt2, SUB, seg, zg.nseg, zg.target, zg.fill, tz0, tz1);
  if (bs[(size_t)tz0 * SUB] == bs[(size_t)tz1 * SUB]) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cx = col / zg.nc[1], cy = col - cx * zg.nc[1];
  const int o0 = cx * Cfg::T0, o1 = cy * Cfg::T1;

  if (tid == 0) {
    for (int i = 0; i < NSLOT; i++) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], NW); }
    *ticket = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NW) {
    // ---- service warp: the window chunks of every populated sub-chunk, in order ----
    if (lane == 0) {
      int hi = INT_MIN;             // chunks below hi have been requested
      unsigned loads[NSLOT];        // loads issued into each slot so far
#pragma unroll
      for (int i = 0; i < NSLOT; i++) loads[i] = 0;
      int prev = bs[(size_t)tz0 * SUB];
      for (int tz = tz0; tz < tz1; tz++) {
        const int nxt = bs[(size_t)(tz + 1) * SUB];
        if (nxt == prev) continue;
        prev = nxt;
        for (int c = max(tz, hi); c < tz + KS; c++) {
          const int sl = c % NSLOT;
#pragma unroll
          for (int i = 0; i < NSLOT; i++)
            if (i == sl) {
              if (loads[i] > 0) mbar_wait_park(&wempty[i], (loads[i] - 1) & 1u);
              loads[i]++;
            }
          mbar_expect_tx(&wfull[sl], (unsigned)SLOTB);
          tma_load_3d(win + (size_t)sl * SLOTB, &tmap, c * ZS * NCOMP, o1, o0, &wfull[sl]);
        }
        hi = tz + KS;
      }
    }
    return;
  }

  // ---- consumer warps ----
  const int g = lane >> 2, t = lane & 3;
  // my element of an n-block inside a window slot: [x row][y row][z][comp]; 256 contiguous bytes per n-block and warp
  const int frag_off = CPLX ? ((g >> 1) * (ZS * CELLB) + t * CELLB + (g & 1) * 8) : (g * (ZS * CELLB) + t * CELLB);
  const int aYl = (Row::oY + Cfg::YLEAD + (CPLX ? t : 2 * t)) * 8;
  const int aXl = (Row::oX + Cfg::XLEAD) * 8;
  const unsigned char *tabb = reinterpret_cast<const unsigned char *>(tab);

  // walk state, identical in every warp: position, first sorted node and batch count of the sub-chunk at the position,
  // batches in front of it, last populated sub-chunk at or before the position, and the load sequence replayed so far
  int pos = tz0, s_pos = bs[(size_t)tz0 * SUB], e_pos = bs[(size_t)(tz0 + 1) * SUB];
  int base = 0, last_ne = INT_MIN / 2, hi = INT_MIN;
  unsigned parbits = 0;            // bit sl: parity of the number of loads into slot sl
  for (;;) {
    int T = 0;
    if (lane == 0) T = atomicAdd(ticket, 1);
    T = __shfl_sync(0xffffffffu, T, 0);
    // advance to the sub-chunk that holds batch T
    bool done = false;
    for (;;) {
      const int nb = (e_pos - s_pos + 7) >> 3;
      if (nb > 0 && last_ne != pos) {            // first visit of a populated sub-chunk: replay its loads
        for (int c = max(pos, hi); c < pos + KS; c++) parbits ^= 1u << (c % NSLOT);
        hi = pos + KS;
        last_ne = pos;
      }
      if (T < base + nb) break;
      // leave pos: its chunk is part of no later window.  The release must land in the mbarrier phase of THIS occupant of
      // the slot: a warp that skips sub-chunks whose batches others took could otherwise arrive before the chunk was
      // even loaded (its load waits for the slowest warp to release the previous occupant) and complete the wrong phase
      if (pos - last_ne < KS) {
        const int sl = pos % NSLOT;
        mbar_wait_park(&wfull[sl], ((parbits >> sl) & 1u) ^ 1u);
        if (lane == 0) mbar_arrive(&wempty[sl]);
      }
      base += nb;
      pos++;
      if (pos >= tz1) { done = true; break; }
      s_pos = e_pos;
      e_pos = bs[(size_t)(pos + 1) * SUB];
    }
    if (done) break;
    const int tz = pos;
    const unsigned char *slot[KS];
#pragma unroll
    for (int q = 0; q < KS; q++) {
      const int sl = (tz + q) % NSLOT;
      mbar_wait_park(&wfull[sl], ((parbits >> sl) & 1u) ^ 1u);
      slot[q] = win + (size_t)sl * SLOTB + frag_off;
    }
    const int b0 = s_pos + (T - base) * 8;
    const int i = min(b0 + g, e_pos - 1);
    const bool live = b0 + g < e_pos;
    const unsigned char *row = tabb + (size_t)i * ROWBYTES;
    const int4 hd = __ldg(reinterpret_cast<const int4 *>(row));       // {-dx*8, -dy*8, dz, dx}
    const int j = __ldg(reinterpret_cast<const int *>(row) + 4);
    double az[KS], adz[GRAD ? KS : 1];
#pragma unroll
    for (int q = 0; q < KS; q++) {
      az[q] = ldg_f64(row + (Row::oZ + 4 * q + t) * 8);
      if (GRAD) adz[q] = ldg_f64(row + (Row::oZ + 4 * q + t) * 8 + dOff);
    }
    double wy[NWY], dwy[GRAD ? NWY : 1];
#pragma unroll
    for (int q = 0; q < NWY; q++) {
      const int yo = CPLX ? 4 * q : (8 * (q >> 1) + (q & 1));
      wy[q] = ldg_f64(row + aYl + yo * 8 + hd.y);
      if (GRAD) dwy[q] = ldg_f64(row + aYl + yo * 8 + dOff + hd.y);
    }
    // x rows the batch touches: the nodes are in dx order
    const int xlo = __shfl_sync(0xffffffffu, hd.w, 0), xhi = __shfl_sync(0xffffffffu, hd.w, 28) + C;
    const unsigned char *xw = row + aXl + hd.x;
    double v[NVAL];
#pragma unroll
    for (int q = 0; q < NVAL; q++) v[q] = 0;
    // x weight of row X for my node: padding in the row (v4's own rows), or the tap range tested (v3's rows)
    auto ld_wx = [&](int X, double &w, double &dw) {
      if (!Sm::XPRED || (unsigned)(X - hd.w) < (unsigned)C) {
        w = ldg_f64(xw + X * 8);
        if (GRAD) dw = ldg_f64(xw + X * 8 + dOff);
      } else { w = 0; dw = 0; }
    };
    double wx_n = 0, dwx_n = 0;
    ld_wx(xlo, wx_n, dwx_n);
    for (int X = xlo; X < xhi; X++) {
      const double wx = wx_n, dwx = dwx_n;
      if (X + 1 < xhi) ld_wx(X + 1, wx_n, dwx_n);
      const int xoff = X * (R1 * ZS * CELLB);
      double a[NCOMP], bb[NCOMP], c[NCOMP];
#pragma unroll
      for (int q = 0; q < NCOMP; q++) { a[q] = 0; bb[q] = 0; c[q] = 0; }
#pragma unroll
      for (int nb = 0; nb < NYB; nb++) {
        double cp[2] = {0, 0}, cd[2] = {0, 0};
#pragma unroll
        for (int q = 0; q < KS; q++) {
          const double B = *reinterpret_cast<const double *>(slot[q] + xoff + nb * 256);
          dmma884(cp, az[q], B);
          if (GRAD) dmma884(cd, adz[q], B);
        }
        if constexpr (CPLX) {
          a[0] = fma(wy[nb], cp[0], a[0]); a[1] = fma(wy[nb], cp[1], a[1]);
          if (GRAD) {
            bb[0] = fma(dwy[nb], cp[0], bb[0]); bb[1] = fma(dwy[nb], cp[1], bb[1]);
            c[0] = fma(wy[nb], cd[0], c[0]); c[1] = fma(wy[nb], cd[1], c[1]);
          }
        } else {
          a[0] = fma(wy[2 * nb], cp[0], a[0]); a[0] = fma(wy[2 * nb + 1], cp[1], a[0]);
          if (GRAD) {
            bb[0] = fma(dwy[2 * nb], cp[0], bb[0]); bb[0] = fma(dwy[2 * nb + 1], cp[1], bb[0]);
            c[0] = fma(wy[2 * nb], cd[0], c[0]); c[0] = fma(wy[2 * nb + 1], cd[1], c[0]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < NCOMP; q++) {
        v[q] = fma(wx, a[q], v[q]);
        if (GRAD) {
          v[NCOMP + q] = fma(dwx, a[q], v[NCOMP + q]);
          v[2 * NCOMP + q] = fma(wx, bb[q], v[2 * NCOMP + q]);
          v[3 * NCOMP + q] = fma(wx, c[q], v[3 * NCOMP + q]);
        }
      }
    }
    // sum over the lane quad (the quad's lanes hold different y rows); every value of node g ends in exactly one lane
    if constexpr (NVAL >= 4) {
      constexpr int H = NVAL / 2, Q = NVAL / 4;
      const bool hi2 = (t & 2) != 0, hi1 = (t & 1) != 0;
      double k2[H];
#pragma unroll
      for (int q = 0; q < H; q++) {
        const double send = hi2 ? v[q] : v[q + H], keep = hi2 ? v[q + H] : v[q];
        k2[q] = keep + shfl_xor(send, 2);
      }
      double k1[Q];
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const double send = hi1 ? k2[q] : k2[q + Q], keep = hi1 ? k2[q + Q] : k2[q];
        k1[q] = keep + shfl_xor(send, 1);
      }
      if (live) {
#pragma unroll
        for (int q = 0; q < Q; q++) {
          const int vi = t * Q + q;          // value index: [f, grad x, grad y, grad z] x NCOMP
          double *o = nullptr;
          if (vi < NCOMP) { if (out.f) o = out.f + ((size_t)j * out.f_stride + out.f_off) * NCOMP + vi; }
          else if (out.grad) o = out.grad + (size_t)j * 3 * NCOMP + (vi - NCOMP);
          if (o) *o = out.accumulate ? *o + k1[q] : k1[q];
        }
      }
    } else if constexpr (NVAL == 2) {
      const bool hi2 = (t & 2) != 0;
      double k = (hi2 ? v[1] : v[0]) + shfl_xor(hi2 ? v[0] : v[1], 2);
      k += shfl_xor(k, 1);
      if (live && (t & 1) == 0 && out.f) {
        double *o = out.f + ((size_t)j * out.f_stride + out.f_off) * NCOMP + (t >> 1);
        *o = out.accumulate ? *o + k : k;
      }
    } else {
      double k = v[0] + shfl_xor(v[0], 2);
      k += shfl_xor(k, 1);
      if (live && t == 0 && out.f) {
        double *o = out.f + ((size_t)j * out.f_stride + out.f_off) * NCOMP;
        *o = out.accumulate ? *o + k : k;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// scatter (adjoint B^T): v3's tensor-core scatter (zmarch3.cuh: k_scatter_mma) on the v4 geometry
// ------------------------------------------------------------------------------------------------
// window[k, column] += sum_node psi_z[node, k] * amp[node, column], M = 8 z cells, N = 8 columns, K = 4 nodes per DMMA.  The
// accumulator window of a row does not fit one warp's registers at m = 8 together with its neighbours', so a CTA owns one
// y HALF of the column tile's footprint (blockIdx parity) and each of its R0 consumer warps ONE x row of it: NYBH n-blocks
// x W / 8 cell blocks of C fragments.  The window is circular in its W cell slots (W = 24 at m = 7, 8: not a power of two);
// everything else is v3's: the service warp streams rows and sorted node values into a ring, a warp meets the nodes whose x
// support covers its row (a contiguous range of the chunk's dx-sorted nodes), finished chunks leave through a staging box
// and ONE TMA reduce-add per warp and advance.
template <bool CPLX, int M_, bool GRAD, bool RG, int NYBH, class Sm, class Sc>
__device__ __forceinline__ void zm4_scatter_rows(const CUtensorMap &tmap, unsigned char *smem_raw, int warp, int lane, int col, int yh,
                                                 const Zm2Geom &zg, int tz0) {
  typedef Zm4Cfg<M_> Cfg;
  typedef typename Sc::Row Row;
  constexpr int C = Cfg::C, T0 = Cfg::T0, ZS = Cfg::ZS, W = Cfg::W, NZB = W / 8;
  constexpr int NCOMP = Sm::NCOMP, S = Sm::SS, ROWBYTES = Row::ROWBYTES, STAGE = Sc::stage, NVP = Sc::NVP;
  constexpr bool PF = !GRAD;             // operands of the next batch built while this batch's MMAs run (register budget)
  unsigned char *ring = smem_raw + Sc::off_ring;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + Sc::off_bar);
  unsigned long long *empty = full + S;

  const int g = lane >> 2, t = lane & 3;
  const int cx = col / zg.nc[1], cy = col - cx * zg.nc[1];
  const int o0 = cx * T0 + warp, o1 = cy * Cfg::T1 + yh * Sm::RH;
  const int dxlo = max(0, warp - (C - 1)), dxhi1 = min(T0 - 1, warp) + 1;
  unsigned char *mystg = smem_raw + (size_t)warp * 2 * Sm::WARP_BOX;
  // B fragment: column 8 nb + g is component g & 1 of row 4 nb + (g >> 1) (complex) or row 8 nb + g (real) of my half
  const int aX = (Row::oX + Cfg::XLEAD + warp) * 8;
  const int aY = (Row::oY + Cfg::YLEAD + yh * Sm::RH + (CPLX ? (g >> 1) : g)) * 8;
  const int aV = CPLX ? (g & 1) * 8 : 0;
  constexpr int dOff = RG ? (Row::oDX - Row::oX) * 8 : 0;
  // C fragment -> staging box [RH][4]: cell z = g & 3 of row 4 nb + t (complex), rows 8 nb + 2t, 2t + 1 (real)
  const int stg_off = CPLX ? (t * 64 + (g & 3) * 16) : (2 * t * 32 + (g & 3) * 8);

  double acc[NYBH][NZB][2];
#pragma unroll
  for (int nb = 0; nb < NYBH; nb++)
#pragma unroll
    for (int zb = 0; zb < NZB; zb++) { acc[nb][zb][0] = 0; acc[nb][zb][1] = 0; }
  int cur = tz0, dirty = 0, nfl = 0;

  auto flush_advance = [&]() {
    unsigned char *sb = mystg + (nfl & 1) * Sm::WARP_BOX;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the box issued two flushes ago was read
    __syncwarp();
    const int slot0 = (cur * ZS) % W;
    const bool mine = (g >> 2) == ((slot0 >> 2) & 1);
    const int zsel = slot0 >> 3;
    if (NYBH * 8 / NCOMP < Sm::RH) {       // the shorter half: rows of the box my n-blocks do not cover stay zero
      for (int i = lane; i < Sm::WARP_BOX / 16; i += 32) reinterpret_cast<uint4 *>(sb)[i] = make_uint4(0u, 0u, 0u, 0u);
      __syncwarp();
    }
#pragma unroll
    for (int zb = 0; zb < NZB; zb++)
      if (zb == zsel) {
#pragma unroll
        for (int nb = 0; nb < NYBH; nb++) {
          if (mine) {
            if constexpr (CPLX) {
              *reinterpret_cast<double2 *>(sb + nb * 256 + stg_off) = make_double2(acc[nb][zb][0], acc[nb][zb][1]);
            } else {
              *reinterpret_cast<double *>(sb + nb * 256 + stg_off) = acc[nb][zb][0];
              *reinterpret_cast<double *>(sb + nb * 256 + stg_off + 32) = acc[nb][zb][1];
            }
            acc[nb][zb][0] = 0; acc[nb][zb][1] = 0;
          }
        }
      }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      tma_reduce_add_3d(sb, &tmap, cur * ZS * NCOMP, o1, o0);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    nfl++;
    cur++;
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
      // z weight offsets of my cell slots for this window position
      int offA[NZB];
#pragma unroll
      for (int zb = 0; zb < NZB; zb++) offA[zb] = (Row::oZ + (((8 * zb + g - cur * ZS) % W + W) % W)) * 8;
      constexpr int NZG = GRAD ? NZB : 1, NYG = GRAD ? NYBH : 1;
      auto prep = [&](int b0, double (&az)[NZB], double (&adz)[NZG], double (&bA)[NYBH], double (&bB)[NYG]) {
        const bool valid = b0 + t < hi;
        const int i = min(b0 + t, hi - 1);
        const unsigned char *row = sp + kZm2HdrBytes + (size_t)i * ROWBYTES;
        const int4 hd = *reinterpret_cast<const int4 *>(row);       // {-dx*8, -dy*8, dz, dx}
#pragma unroll
        for (int zb = 0; zb < NZB; zb++) {
          az[zb] = *reinterpret_cast<const double *>(row + offA[zb]);
          if (GRAD) adz[zb] = *reinterpret_cast<const double *>(row + offA[zb] + dOff);
        }
        const unsigned char *vrow = sp + Sc::off_vals + (size_t)i * NVP * 8 + aV;
        double f = *reinterpret_cast<const double *>(vrow);
        if (!valid) f = 0;
        double g0 = 0, g1 = 0, g2 = 0;
        if (GRAD) {
          g0 = *reinterpret_cast<const double *>(vrow + NCOMP * 8);
          g1 = *reinterpret_cast<const double *>(vrow + 2 * NCOMP * 8);
          g2 = *reinterpret_cast<const double *>(vrow + 3 * NCOMP * 8);
          if (!valid) { g0 = 0; g1 = 0; g2 = 0; }
        }
        // amplitudes: A = psi_x (psi_y f + dpsi_y g1) + dpsi_x psi_y g0, B = psi_x psi_y g2 (paired with dpsi_z)
        const double w0 = *reinterpret_cast<const double *>(row + aX + hd.x);
        double ax = w0 * f, bx = 0, cxw = 0;
        if (GRAD) {
          const double dw0 = *reinterpret_cast<const double *>(row + aX + dOff + hd.x);
          ax = fma(dw0, g0, ax);
          bx = w0 * g1;
          cxw = w0 * g2;
        }
#pragma unroll
        for (int jy = 0; jy < NYBH; jy++) {
          const int yo = CPLX ? 4 * jy : 8 * jy;
          const double w1 = *reinterpret_cast<const double *>(row + aY + yo * 8 + hd.y);
          bA[jy] = w1 * ax;
          if (GRAD) {
            const double dw1 = *reinterpret_cast<const double *>(row + aY + yo * 8 + dOff + hd.y);
            bA[jy] = fma(dw1, bx, bA[jy]);
            bB[jy] = w1 * cxw;
          }
        }
      };
      double az[NZB], adz[NZG], bA[NYBH], bB[NYG];
      if constexpr (PF) {
        double az_n[NZB], adz_n[NZG], bA_n[NYBH], bB_n[NYG];
        prep(lo, az, adz, bA, bB);
        for (int b0 = lo; b0 < hi; b0 += 4) {
          prep(b0 + 4, az_n, adz_n, bA_n, bB_n);
#pragma unroll
          for (int nb = 0; nb < NYBH; nb++)
#pragma unroll
            for (int zb = 0; zb < NZB; zb++) dmma884(acc[nb][zb], az[zb], bA[nb]);
#pragma unroll
          for (int zb = 0; zb < NZB; zb++) az[zb] = az_n[zb];
#pragma unroll
          for (int nb = 0; nb < NYBH; nb++) bA[nb] = bA_n[nb];
        }
      } else {
        for (int b0 = lo; b0 < hi; b0 += 4) {
          prep(b0, az, adz, bA, bB);
#pragma unroll
          for (int nb = 0; nb < NYBH; nb++) {
#pragma unroll
            for (int zb = 0; zb < NZB; zb++) dmma884(acc[nb][zb], az[zb], bA[nb]);
            if (GRAD) {
#pragma unroll
              for (int zb = 0; zb < NZB; zb++) dmma884(acc[nb][zb], adz[zb], bB[nb]);
            }
          }
        }
      }
      dirty = Cfg::NFL;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }
  while (dirty > 0) { flush_advance(); dirty--; }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging must outlive the bulk reads
}

template <bool CPLX, int M_, bool GRAD, bool RG>
__global__ void __launch_bounds__((Zm4Cfg<M_>::R0 + 1) * 32, 1)
k_scatter_mma4(const __grid_constant__ CUtensorMap tmap, Zm2Geom zg, const double *__restrict__ tab, const double *__restrict__ vals,
               const int *__restrict__ bin_start) {
  typedef Zm4Cfg<M_> Cfg;
  typedef Zm4Smem<CPLX, M_> Sm;
  typedef typename Sm::template Scat<GRAD, RG> Sc;
  typedef typename Sc::Row Row;
  constexpr int T0 = Cfg::T0, NCW = Sm::NCW, S = Sm::SS, GB = Sm::SGB, ROWBYTES = Row::ROWBYTES, STAGE = Sc::stage, NVP = Sc::NVP;
  static_assert(GB % 4 == 0 && GB <= 32 && 4 + T0 + 1 <= kZm2HdrBytes / 4 && (Row::oDX * 8) % 16 == 0, "ring stages hold whole node batches; header layout");

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *ring = smem_raw + Sc::off_ring;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + Sc::off_bar);
  unsigned long long *empty = full + S;

  const int yh = blockIdx.x & 1, item = blockIdx.x >> 1;
  const int colr = item / zg.nseg, seg = item - colr * zg.nseg, col = zg.col0 + colr;
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
    // service warp: rows and sorted node values of GB nodes per ring stage, with the chunk's dx prefix table
    const unsigned char *tabb = reinterpret_cast<const unsigned char *>(tab);
    int kb = 0;
    int gs_next = (lane <= T0) ? bs[(size_t)tz0 * Cfg::SUB + lane] : 0;
    for (int tz = tz0; tz < tz1; tz++) {
      const int gs = gs_next;
      if (tz + 1 < tz1) gs_next = (lane <= T0) ? bs[(size_t)(tz + 1) * Cfg::SUB + lane] : 0;
      const int s0 = __shfl_sync(0xffffffffu, gs, 0), e = __shfl_sync(0xffffffffu, gs, T0);
      for (int c0 = s0; c0 < e; c0 += GB, kb++) {
        const int st = kb % S;
        mbar_wait_park(&empty[st], (((unsigned)(kb / S)) & 1u) ^ 1u);
        const int cnt = min(GB, e - c0);
        unsigned char *sp = ring + (size_t)st * STAGE;
        int *h = reinterpret_cast<int *>(sp);
        if (lane <= T0) h[4 + lane] = min(max(gs - c0, 0), cnt);
        if (lane == 0) { h[0] = tz; h[1] = cnt; h[2] = c0; h[3] = 0; }
        __syncwarp();
        // an F-only scatter on rows that carry derivative sections reads the first (psi) half of every row only
        constexpr bool HALF = RG && !GRAD;
        constexpr unsigned PSIB = (unsigned)Row::oDX * 8u;
        const unsigned vb = (unsigned)(cnt * NVP * 8);
        if (lane == 0) {
          const unsigned rb = HALF ? (unsigned)cnt * PSIB : (unsigned)(cnt * ROWBYTES);
          mbar_expect_tx(&full[st], rb + vb);
          if (!HALF) bulk_load_1d(sp + kZm2HdrBytes, tabb + (size_t)c0 * ROWBYTES, rb, &full[st]);
          bulk_load_1d(sp + Sc::off_vals, vals + (size_t)c0 * NVP, vb, &full[st]);
        }
        if (HALF) {
          __syncwarp();      // the expected byte count is posted before any copy can complete
          if (lane < cnt)
            bulk_load_1d(sp + kZm2HdrBytes + (size_t)lane * ROWBYTES, tabb + (size_t)(c0 + lane) * ROWBYTES, PSIB, &full[st]);
        }
      }
    }
    const int st = kb % S;
    mbar_wait_park(&empty[st], (((unsigned)(kb / S)) & 1u) ^ 1u);
    if (lane == 0) {
      int *h = reinterpret_cast<int *>(ring + (size_t)st * STAGE);
      h[0] = INT_MAX; h[1] = 0;
      mbar_arrive(&full[st]);
    }
    return;
  }
  if (yh == 0) zm4_scatter_rows<CPLX, M_, GRAD, RG, Sm::NYB0, Sm, Sc>(tmap, smem_raw, warp, lane, col, 0, zg, tz0);
  else if constexpr (Sm::NYB1 > 0) zm4_scatter_rows<CPLX, M_, GRAD, RG, Sm::NYB1, Sm, Sc>(tmap, smem_raw, warp, lane, col, 1, zg, tz0);
}

}  // namespace pnb
