// Z-marching gridding kernels on the FP64 tensor-core path (v3: DMMA m8n8k4), the default B / B^T path in double for
// the cutoffs whose register window is a whole number of k-steps (m = 6; see Zm3Ok).
// Reference loops replaced: kernel/assign.c:478-1130 (the (2m+1)^3 inner loops), kernel/ndft-parallel.c:2703-3009.
//
// What v2 (zmarch2.cuh) taught (profiles/r1b_ncu_full.md): with the grid cells in registers the FP64 pipe is the right
// bound, but a DFMA node loop is issue- and latency-bound (one in-order warp per row pair meets every node; 44-48 % of the
// pipe at best, 59 % in an idealised model with every warp busy).  Measured on B200 (tools/ubench_dmma.cu,
// profiles/r2_dmma_ubench.md): DMMA.8x8x4 runs at the full rate of the FP64 pipe (64 FMA/clk/SM, the same peak as DFMA,
// same pipe) with ONE instruction per 256 FMAs, and a single warp per scheduler with two independent accumulator tiles
// already reaches 94 % of it.  The contraction over the marching axis is a small GEMM per warp and node batch:
//   gather :  T[node, row]  = sum_k psi_z[node, k] * window[k, row]          (M = 8 nodes, N = 8 columns, K = 4 cells per DMMA)
//   scatter:  window[k, row] += sum_node psi_z[node, k] * amp[node, row]     (M = 8 cells, N = 8 columns, K = 4 nodes per DMMA)
// with row = one (x, y) grid line of the warp's footprint (a complex cell is two columns) and the z window of every row
// held in registers in the fragment layout of the instruction, so the grid still moves HBM -> shared memory (TMA box) ->
// registers once per CTA and z sub-chunk.  The window is circular in its k-step slots (gather) / cell slots (scatter):
// an advance overwrites (flushes) the slot of the chunk that left and the weight fragments are read with the rotated
// offset, so no cell ever moves between registers or lanes.
// The x-y contraction of the gather stays on the DFMA path but shrinks to 8 rows per lane and node (the C fragment of a
// batch is one node per lane quad); the per-warp partial outputs (8 values per node instead of 32 x 2 partial sums) meet
// in a small shared-memory stage that one reducer warp sums in a fixed order (bit-reproducible) and writes out.
// Everything around the node loops is v2's: sort key (column tile, z sub-chunk, dx), node table rows (Zm2Row), chunk
// headers with the dx prefix table, the producer warp and its bulk-copy ring, per-warp TMA boxes and mbarriers.
#pragma once
#include "zmarch2.cuh"

namespace pnb {

template <int M_> struct Zm3Ok { static constexpr bool value = (M_ == 6); };

#ifdef ZM3_DBG_NOEPI      // timing experiment only (wrong results): the MMAs stay, the x-y contraction goes
#define ZM3_ASM asm volatile
#else
#define ZM3_ASM asm
#endif
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  ZM3_ASM("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// non-blocking phase test (the service warp polls two pipelines)
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, unsigned phase) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  return __any_sync(0xffffffffu, ok != 0);
}

// RG: the node-table rows carry the derivative sections (a table built for a gradient transform also serves F-only
// gathers and scatters of the same node set: Core keeps the rows while the binning they belong to stays valid).  The rows
// never hold node values; the scatter streams f / grad_f as a second, small section of every ring stage (k_pack_vals).
template <bool CPLX, int M_, bool GRAD, bool RG> struct Zm3Smem {
  static_assert(RG || !GRAD, "a gradient kernel needs rows with derivative sections");
  typedef Zm2Cfg<M_> Cfg;
  typedef Zm2Row<double, M_, RG, false, CPLX> Row;
  static constexpr int NCOMP = CPLX ? 2 : 1;
  static constexpr int CELLB = 8 * NCOMP;
  static constexpr int WARP_BOX = Cfg::XW * 16 * Cfg::ZS * CELLB;          // bytes one warp stages per window advance
  static constexpr int NVAL = NCOMP * (GRAD ? 4 : 1);                      // values per node (f, grad_f)
  static constexpr int NVP = (NVAL + 1) / 2 * 2;                           // ... padded to 16-byte granules
#ifndef ZM3_GP
#define ZM3_GP 3
#endif
#ifndef ZM3_GS
#define ZM3_GS 4
#endif
#ifndef ZM3_GGB
#define ZM3_GGB 32
#endif
  // gather: one staging box per warp, ring of S stages of GB nodes, P stages of per-warp partial outputs
  static constexpr int GS = ZM3_GS, GP = ZM3_GP, GGB = ZM3_GGB;
  static constexpr int g_stage = kZm2HdrBytes + GGB * Row::ROWBYTES;
  static constexpr int g_off_ring = Cfg::NCW * WARP_BOX;
  static constexpr int g_off_part = g_off_ring + GS * g_stage;
  static constexpr int g_part_stage = Cfg::NCW * GGB * NVAL * 8;
  static constexpr int g_off_bar = g_off_part + GP * g_part_stage;
  static constexpr int gather = g_off_bar + (2 * GS + 2 * GP + Cfg::NCW) * 8;
  // scatter: two staging boxes per warp, ring of stages {header, GB rows, GB value rows}
  static constexpr int SS = 4, SGB = 32;
  static constexpr int s_off_vals = kZm2HdrBytes + SGB * Row::ROWBYTES;
  static constexpr int s_stage = s_off_vals + SGB * NVP * 8;
  static constexpr int s_off_ring = Cfg::NCW * 2 * WARP_BOX;
  static constexpr int s_off_bar = s_off_ring + SS * s_stage;
  static constexpr int scatter = s_off_bar + 2 * SS * 8;
  static_assert(g_stage % 16 == 0 && s_stage % 16 == 0 && s_off_vals % 16 == 0 && g_off_ring % 128 == 0 && s_off_ring % 128 == 0 && WARP_BOX % 256 == 0, "alignment");
  static_assert(scatter <= 232448 && gather <= 232448, "shared-memory budget of one CTA exceeded");
};

// node values in sorted order for the scatter: vals[p] = {f, grad_f[0..3)} of node perm[p], NVP reals per node
template <bool CPLX, bool GRAD> __global__ void __launch_bounds__(256) k_pack_vals(NodeArgs<double> na, double *__restrict__ vals) {
  constexpr int NCOMP = CPLX ? 2 : 1, NVAL = NCOMP * (GRAD ? 4 : 1), NVP = (NVAL + 1) / 2 * 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)na.M * NVP) return;
  const int p = (int)(i / NVP), c = (int)(i - (long long)p * NVP);
  const int j = na.perm ? na.perm[p] : p;
  double v = 0;
  if (c < NCOMP) { if (na.f) v = na.f[((size_t)j * na.f_stride + na.f_off) * NCOMP + c]; }
  else if (GRAD && c < NVAL) v = na.grad[(size_t)j * 3 * NCOMP + (c - NCOMP)];
  vals[i] = v;
}

// ------------------------------------------------------------------------------------------------
// gather (trafo B)
// ------------------------------------------------------------------------------------------------
// Lane (g, t) = (lane >> 2, lane & 3) of a consumer warp.  B fragment of n-block nb, k-step slot s: window cell
// z = 4 s' + t (s' = the chunk in slot s) of column 8 nb + g; complex grids: column c is component c & 1 of row c >> 1,
// real grids: column c is row c.  Row r of the warp is x row r / 16, y row r % 16 of its [XW][16] footprint slice.
// A fragment of slot s: z weight 4 q + t of node g of the batch, q = (s - cur) mod KS the chunk's place in the window.
// C fragment: node g, columns 2t, 2t+1 of the n-block = one complex partial sum (two real ones) per lane.
template <bool CPLX, int M_, bool GRAD, bool RG>
__global__ void __launch_bounds__((Zm2Cfg<M_>::NCW + 1) * 32, 1)
k_gather_mma(const __grid_constant__ CUtensorMap tmap, Zm2Geom zg, const double *__restrict__ tab, const int *__restrict__ bin_start,
             GatherOut<double> out) {
  typedef Zm2Cfg<M_> Cfg;
  typedef Zm3Smem<CPLX, M_, GRAD, RG> Sm;
  typedef typename Sm::Row Row;
  constexpr int C = Cfg::C, T0 = Cfg::T0, T1 = Cfg::T1, ZS = Cfg::ZS, W = Cfg::W, NCW = Cfg::NCW, XW = Cfg::XW;
  constexpr int NCOMP = Sm::NCOMP, S = Sm::GS, P = Sm::GP, GB = Sm::GGB, ROWBYTES = Row::ROWBYTES, STAGE = Sm::g_stage, NVAL = Sm::NVAL;
  constexpr int KS = W / 4;                       // k-steps = chunks in the window
  constexpr int NYB = 16 * NCOMP / 8;             // n-blocks per x row
  constexpr int NB = XW * NYB;                    // n-blocks per warp
  constexpr unsigned BOXB = Sm::WARP_BOX;
  static_assert(W % 4 == 0 && ZS == 4 && Cfg::ZB == 4 && Cfg::RPT == 1, "window must be whole k-steps of one chunk each");
  static_assert(GB % 8 == 0, "ring stages hold whole node batches");

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *ring = smem_raw + Sm::g_off_ring;
  double *part = reinterpret_cast<double *>(smem_raw + Sm::g_off_part);
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
    for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], NCW + 1); }
    for (int i = 0; i < P; i++) { mbar_init(&pfull[i], NCW); mbar_init(&pempty[i], 1); }
    for (int i = 0; i < NCW; i++) mbar_init(&wbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NCW) {
    // ---- service warp: feeds the node ring (zm2_produce as a resumable state machine) and, between two stages, sums the
    // partial outputs of the stages the consumers have finished: per node the warps that met it, in warp order ----
    constexpr int NPP = 32 / NVAL;                 // nodes per reduction pass
    const int sub = lane / NVAL, v = lane - sub * NVAL;
    const unsigned char *tabb = reinterpret_cast<const unsigned char *>(tab);
    int ptz = tz0 - 1, c0 = 0, e = 0, kbp = 0, kbr = 0, nreal = 0, gs = 0;
    int gs_next = (lane <= T0) ? bs[(size_t)tz0 * Cfg::SUB + lane] : 0;
    bool at_end = false, pdone = false;
    for (;;) {
      bool did = false;
      if (!pdone) {
        while (!at_end && c0 >= e) {               // next z sub-chunk with nodes
          ptz++;
          if (ptz >= tz1) { at_end = true; break; }
          gs = gs_next;
          if (ptz + 1 < tz1) gs_next = (lane <= T0) ? bs[(size_t)(ptz + 1) * Cfg::SUB + lane] : 0;
          c0 = __shfl_sync(0xffffffffu, gs, 0);
          e = __shfl_sync(0xffffffffu, gs, T0);
        }
        const int st = kbp % S;
        if (mbar_test(&empty[st], (((unsigned)(kbp / S)) & 1u) ^ 1u)) {
          unsigned char *sp = ring + (size_t)st * STAGE;
          int *h = reinterpret_cast<int *>(sp);
          if (at_end) {
            if (lane == 0) { h[0] = INT_MAX; h[1] = 0; mbar_arrive(&full[st]); }
            pdone = true;
          } else {
            const int cnt = min(GB, e - c0);
            if (lane <= T0) h[4 + lane] = min(max(gs - c0, 0), cnt);
            if (lane == 0) { h[0] = ptz; h[1] = cnt; h[2] = c0; h[3] = 0; }
            __syncwarp();
            if (lane == 0) {
              const unsigned bytes = (unsigned)(cnt * ROWBYTES);
              mbar_expect_tx(&full[st], bytes);
              bulk_load_1d(sp + kZm2HdrBytes, tabb + (size_t)c0 * ROWBYTES, bytes, &full[st]);
            }
            c0 += GB;
            nreal++;
          }
          kbp++;
          did = true;
          __syncwarp();
        }
      }
      if (kbr < nreal) {
        const int st = kbr % S, ps = kbr % P;
        if (mbar_test(&pfull[ps], ((unsigned)(kbr / P)) & 1u)) {
          mbar_wait(&full[st], ((unsigned)(kbr / S)) & 1u);     // passed long ago: orders my reads behind the bulk copy
          const unsigned char *sp = ring + (size_t)st * STAGE;
          const int cnt = reinterpret_cast<const int *>(sp)[1];
          const double *pp = part + (size_t)ps * (Sm::g_part_stage / 8);
          for (int i0 = 0; i0 < cnt; i0 += NPP) {
            const int i = i0 + sub;
            if (i < cnt) {
              const int *hd = reinterpret_cast<const int *>(sp + kZm2HdrBytes + (size_t)i * ROWBYTES);
              const int dx = hd[3], j = hd[4];
              const int wlo = dx / XW, whi = min(NCW - 1, (dx + C - 1) / XW);
              double sum = 0;
              for (int w = wlo; w <= whi; w++) sum += pp[((size_t)w * GB + i) * NVAL + v];
              double *o = nullptr;
              if (v < NCOMP) { if (out.f) o = out.f + ((size_t)j * out.f_stride + out.f_off) * NCOMP + v; }
              else if (GRAD) o = out.grad + (size_t)j * 3 * NCOMP + (v - NCOMP);
              if (o) *o = out.accumulate ? *o + sum : sum;
            }
          }
          __syncwarp();
          if (lane == 0) { mbar_arrive(&pempty[ps]); mbar_arrive(&empty[st]); }
          kbr++;
          did = true;
        }
      }
      if (pdone && kbr == nreal) break;
      if (!did) __nanosleep(64);
    }
    return;
  }

  // ---- consumer warps ----
  const int g = lane >> 2, t = lane & 3;
  const int cx = col / zg.nc[1], cy = col - cx * zg.nc[1];
  const int o0 = cx * T0 + XW * warp, o1 = cy * T1;
  const int dxlo = max(0, XW * warp - (C - 1)), dxhi1 = min(T0 - 1, XW * warp + XW - 1) + 1;
  unsigned char *mystg = smem_raw + (size_t)warp * Sm::WARP_BOX;
  unsigned long long *mybar = &wbar[warp];
  unsigned wph = 0;
  // my element of a staged box [XW][16][4] (z innermost): the 32 lanes of an n-block read 256 contiguous bytes
  const int stg_off = CPLX ? ((g >> 1) * 64 + (g & 1) * 8 + t * 16) : (g * 32 + t * 8);
  // weights of the rows my C fragments hold: y rows 4 jy + t (complex) or 8 jy + 2 t + e (real)
  const int aX = (Row::oX + Cfg::XLEAD + XW * warp) * 8;
  const int aY = (Row::oY + T1 - 1 + (CPLX ? t : 2 * t)) * 8;
  constexpr int dOff = RG ? (Row::oDX - Row::oX) * 8 : 0;

  double win[NB][KS];
  int cur = INT_MIN / 2;       // sub-chunk whose cells [cur*ZS, cur*ZS + W) are in the window
  int rot = 0;                 // cur mod KS: the slot of chunk cur
  bool pending = false;        // a load of chunk cur + KS is in flight

  auto issue = [&](int chunk) {
    if (lane == 0) {
      mbar_expect_tx(mybar, BOXB);
      tma_load_3d(mystg, &tmap, chunk * ZS * NCOMP, o1, o0, mybar);
    }
  };
  // wait for the staged box and copy my cells into slot `slot` (see zmarch2.cuh take(): the vote consumes the loaded
  // values, so every lane's reads have returned before lane 0 re-arms the staging buffer)
  auto take = [&](int slot) {
    mbar_wait(mybar, wph);
    wph ^= 1u;
    int nan = 0;
#pragma unroll
    for (int s = 0; s < KS; s++)
      if (s == slot) {
#pragma unroll
        for (int nb = 0; nb < NB; nb++) {
          win[nb][s] = *reinterpret_cast<const double *>(mystg + nb * 256 + stg_off);
          nan |= (win[nb][s] != win[nb][s]);
        }
      }
    (void)warp_any_volatile(nan);
  };
  auto set_cur = [&](int c) { cur = c; rot = ((c % KS) + KS) % KS; };
  auto drop_pending = [&]() {
    if (pending) { mbar_wait(mybar, wph); wph ^= 1u; pending = false; }
  };
  auto reload = [&](int tz) {
    drop_pending();
    set_cur(tz);
    for (int i = 0; i < KS; i++) {
      issue(tz + i);
      take((rot + i) % KS);
    }
    issue(tz + KS);
    pending = true;
  };
  auto advance1 = [&]() {       // chunk cur leaves, chunk cur + KS (in flight) takes its slot
    take(rot);
    set_cur(cur + 1);
    issue(cur + KS);
  };

#ifdef ZM2_TIMING
  long long tq[6] = {0, 0, 0, 0, 0, 0}, tc = clock64(), tit = 0, itsum = 0, itn = 0, itmin = 1LL << 40;
#endif
  for (int kb = 0;; kb++) {
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
      double *pw = part + (size_t)ps * (Sm::g_part_stage / 8) + (size_t)warp * GB * NVAL;
      // A fragments (z weights of node g of the batch) are fetched one batch ahead
      auto load_a = [&](int b0, const unsigned char *&row, int4 &hd, double (&az)[KS], double (&adz)[KS]) {
        const int i = min(b0 + g, hi - 1);
        row = sp + kZm2HdrBytes + (size_t)i * ROWBYTES;
        hd = *reinterpret_cast<const int4 *>(row);       // {-dx*8, -dy*8, dz, dx}
#pragma unroll
        for (int s = 0; s < KS; s++) {
          const int q = (s - rot + KS) % KS;
          az[s] = *reinterpret_cast<const double *>(row + (Row::oZ + 4 * q + t) * 8);
          if (GRAD) adz[s] = *reinterpret_cast<const double *>(row + (Row::oZ + 4 * q + t) * 8 + dOff);
        }
      };
      const unsigned char *row, *row_n;
      int4 hd, hd_n;
      double az[KS], adz[KS], az_n[KS], adz_n[KS];
      load_a(lo, row, hd, az, adz);
      for (int b0 = lo; b0 < hi; b0 += 8) {
#ifdef ZM2_TIMING
        { const long long now_ = clock64(); if (b0 > lo) { const long long d_ = now_ - tit; itsum += d_; itn++; if (d_ < itmin) itmin = d_; } tit = now_; }
#endif
        load_a(b0 + 8, row_n, hd_n, az_n, adz_n);
        double wx[XW], dwx[XW], wy[4], dwy[4];
#pragma unroll
        for (int e = 0; e < XW; e++) {
          wx[e] = *reinterpret_cast<const double *>(row + aX + e * 8 + hd.x);
          if (GRAD) dwx[e] = *reinterpret_cast<const double *>(row + aX + e * 8 + dOff + hd.x);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int yo = CPLX ? 4 * q : (8 * (q >> 1) + (q & 1));
          wy[q] = *reinterpret_cast<const double *>(row + aY + yo * 8 + hd.y);
          if (GRAD) dwy[q] = *reinterpret_cast<const double *>(row + aY + yo * 8 + dOff + hd.y);
        }
        double v[NVAL];
#pragma unroll
        for (int q = 0; q < NVAL; q++) v[q] = 0;
#pragma unroll
        for (int x = 0; x < XW; x++) {
          double a[NCOMP], b[NCOMP], c[NCOMP];
#pragma unroll
          for (int q = 0; q < NCOMP; q++) { a[q] = 0; b[q] = 0; c[q] = 0; }
#pragma unroll
          for (int jy = 0; jy < NYB; jy++) {
            const int nb = x * NYB + jy;
            double cp[2] = {0, 0}, cd[2] = {0, 0};
#pragma unroll
            for (int s = 0; s < KS; s++) {
              dmma884(cp, az[s], win[nb][s]);
              if (GRAD) dmma884(cd, adz[s], win[nb][s]);
            }
#ifdef ZM3_DBG_NOEPI
            a[0] += cp[0];
            if (false)
#endif
            if constexpr (CPLX) {
              a[0] = fma(wy[jy], cp[0], a[0]); a[1] = fma(wy[jy], cp[1], a[1]);
              if (GRAD) {
                b[0] = fma(dwy[jy], cp[0], b[0]); b[1] = fma(dwy[jy], cp[1], b[1]);
                c[0] = fma(wy[jy], cd[0], c[0]); c[1] = fma(wy[jy], cd[1], c[1]);
              }
            } else {
              a[0] = fma(wy[2 * jy], cp[0], a[0]); a[0] = fma(wy[2 * jy + 1], cp[1], a[0]);
              if (GRAD) {
                b[0] = fma(dwy[2 * jy], cp[0], b[0]); b[0] = fma(dwy[2 * jy + 1], cp[1], b[0]);
                c[0] = fma(wy[2 * jy], cd[0], c[0]); c[0] = fma(wy[2 * jy + 1], cd[1], c[0]);
              }
            }
          }
#pragma unroll
          for (int q = 0; q < NCOMP; q++) {
            v[q] = fma(wx[x], a[q], v[q]);
            if (GRAD) {
              v[NCOMP + q] = fma(dwx[x], a[q], v[NCOMP + q]);
              v[2 * NCOMP + q] = fma(wx[x], b[q], v[2 * NCOMP + q]);
              v[3 * NCOMP + q] = fma(wx[x], c[q], v[3 * NCOMP + q]);
            }
          }
        }
        // sum over the lane quad (the four t of node g); with NVAL >= 4 every exchange also halves what a lane keeps,
        // lane t ends up with values [t * NVAL/4, (t+1) * NVAL/4)
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
          if (b0 + g < hi) {
            double *o = pw + (size_t)(b0 + g) * NVAL + t * Q;
            if constexpr (Q == 2) *reinterpret_cast<double2 *>(o) = make_double2(k1[0], k1[1]);
            else o[0] = k1[0];
          }
        } else if constexpr (NVAL == 2) {
          const bool hi2 = (t & 2) != 0;
          double k = (hi2 ? v[1] : v[0]) + shfl_xor(hi2 ? v[0] : v[1], 2);
          k += shfl_xor(k, 1);
          if (b0 + g < hi && (t & 1) == 0) pw[(size_t)(b0 + g) * NVAL + (t >> 1)] = k;
        } else {
          double k = v[0] + shfl_xor(v[0], 2);
          k += shfl_xor(k, 1);
          if (b0 + g < hi && t == 0) pw[(size_t)(b0 + g) * NVAL] = k;
        }
        row = row_n; hd = hd_n;
#pragma unroll
        for (int s = 0; s < KS; s++) { az[s] = az_n[s]; if (GRAD) adz[s] = adz_n[s]; }
      }
    }
    ZM2_T(3);
    __syncwarp();
    if (lane == 0) { mbar_arrive(&pfull[ps]); mbar_arrive(&empty[st]); }
    ZM2_T(4);
  }
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

// ------------------------------------------------------------------------------------------------
// scatter (adjoint B^T)
// ------------------------------------------------------------------------------------------------
// The window is the ACCUMULATOR: C fragment of n-block nb, cell block zb: cell slot 8 zb + g, columns 2t, 2t+1 (one
// complex cell or two real ones of neighbouring rows).  Slot sigma holds grid cell z with z mod W == sigma (W = 16), so
// the chunk that leaves at an advance is the four slots (4 cur) mod 16 .. +3: half the lanes of one cell block store them
// to the staging box (TMA reduce-add) and clear them.  A fragment: z weight ((8 zb + g) - 4 cur) mod 16 of node t of the
// batch; B fragment: amplitude of node t in column 8 nb + g = psi_x psi_y f (+ gradient terms), built per lane.
template <bool CPLX, int M_, bool GRAD, bool RG>
__global__ void __launch_bounds__((Zm2Cfg<M_>::NCW + 1) * 32, 1)
k_scatter_mma(const __grid_constant__ CUtensorMap tmap, Zm2Geom zg, const double *__restrict__ tab, const double *__restrict__ vals,
              const int *__restrict__ bin_start) {
  typedef Zm2Cfg<M_> Cfg;
  typedef Zm3Smem<CPLX, M_, GRAD, RG> Sm;
  typedef typename Sm::Row Row;
  constexpr int C = Cfg::C, T0 = Cfg::T0, T1 = Cfg::T1, ZS = Cfg::ZS, W = Cfg::W, NCW = Cfg::NCW, XW = Cfg::XW;
  constexpr int NCOMP = Sm::NCOMP, S = Sm::SS, GB = Sm::SGB, ROWBYTES = Row::ROWBYTES, STAGE = Sm::s_stage, NVP = Sm::NVP;
  constexpr int NYB = 16 * NCOMP / 8, NB = XW * NYB, NZB = W / 8;
  static_assert(W == 16 && ZS == 4 && Cfg::ZB == 4 && Cfg::RPT == 1, "accumulator window = two cell blocks of eight slots");
  static_assert(GB % 4 == 0, "ring stages hold whole node batches");

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
    // producer: zm2_produce with a second bulk copy per stage, the node values
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
        if (lane == 0) {
          const unsigned rb = (unsigned)(cnt * ROWBYTES), vb = (unsigned)(cnt * NVP * 8);
          mbar_expect_tx(&full[st], rb + vb);
          bulk_load_1d(sp + kZm2HdrBytes, tabb + (size_t)c0 * ROWBYTES, rb, &full[st]);
          bulk_load_1d(sp + Sm::s_off_vals, vals + (size_t)c0 * NVP, vb, &full[st]);
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

  const int g = lane >> 2, t = lane & 3;
  const int cx = col / zg.nc[1], cy = col - cx * zg.nc[1];
  const int o0 = cx * T0 + XW * warp, o1 = cy * T1;
  const int dxlo = max(0, XW * warp - (C - 1)), dxhi1 = min(T0 - 1, XW * warp + XW - 1) + 1;
  unsigned char *mystg = smem_raw + (size_t)warp * 2 * Sm::WARP_BOX;
  // B fragment: column 8 nb + g is component g & 1 of row 4 nb + (g >> 1) (complex) or row 8 nb + g (real)
  const int aX = (Row::oX + Cfg::XLEAD + XW * warp) * 8;
  const int aY = (Row::oY + T1 - 1 + (CPLX ? (g >> 1) : g)) * 8;
  const int aV = CPLX ? (g & 1) * 8 : 0;          // my component of the node's values (second section of the stage)
  constexpr int dOff = RG ? (Row::oDX - Row::oX) * 8 : 0;
  // C fragment -> staging box [XW][16][4]: cell z = g & 3 of row 4 nb + t (complex), rows 8 nb + 2t, 2t + 1 (real)
  const int stg_off = CPLX ? (t * 64 + (g & 3) * 16) : (2 * t * 32 + (g & 3) * 8);

  double acc[NB][NZB][2];
#pragma unroll
  for (int nb = 0; nb < NB; nb++)
#pragma unroll
    for (int zb = 0; zb < NZB; zb++) { acc[nb][zb][0] = 0; acc[nb][zb][1] = 0; }
  int cur = tz0, dirty = 0, nfl = 0;

  // chunk cur is final: reduce-add its four cell slots into the grid and clear them
  auto flush_advance = [&]() {
    unsigned char *sb = mystg + (nfl & 1) * Sm::WARP_BOX;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the box issued two flushes ago was read
    __syncwarp();
    const int slot0 = (cur * ZS) & (W - 1);
    const bool mine = (g >> 2) == ((slot0 >> 2) & 1);
    const int zsel = slot0 >> 3;
#pragma unroll
    for (int zb = 0; zb < NZB; zb++)
      if (zb == zsel) {
#pragma unroll
        for (int nb = 0; nb < NB; nb++) {
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
      // z weight offsets of my two cell slots for this window position
      int offA[NZB];
#pragma unroll
      for (int zb = 0; zb < NZB; zb++) offA[zb] = (Row::oZ + ((8 * zb + g - cur * ZS) & (W - 1))) * 8;
      // operands of one batch: z weight fragments and the amplitude of node t in each of my columns
      auto prep = [&](int b0, double (&az)[NZB], double (&adz)[NZB], double (&bA)[NB], double (&bB)[NB]) {
        const bool valid = b0 + t < hi;
        const int i = min(b0 + t, hi - 1);
        const unsigned char *row = sp + kZm2HdrBytes + (size_t)i * ROWBYTES;
        const int4 hd = *reinterpret_cast<const int4 *>(row);       // {-dx*8, -dy*8, dz, dx}
#pragma unroll
        for (int zb = 0; zb < NZB; zb++) {
          az[zb] = *reinterpret_cast<const double *>(row + offA[zb]);
          if (GRAD) adz[zb] = *reinterpret_cast<const double *>(row + offA[zb] + dOff);
        }
        const unsigned char *vrow = sp + Sm::s_off_vals + (size_t)i * NVP * 8 + aV;
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
        double ax[XW], bx[XW], cxw[XW];
#pragma unroll
        for (int e = 0; e < XW; e++) {
          const double w0 = *reinterpret_cast<const double *>(row + aX + e * 8 + hd.x);
          ax[e] = w0 * f;                    // times psi_y
          if (GRAD) {
            const double dw0 = *reinterpret_cast<const double *>(row + aX + e * 8 + dOff + hd.x);
            ax[e] = fma(dw0, g0, ax[e]);
            bx[e] = w0 * g1;                 // times dpsi_y
            cxw[e] = w0 * g2;                // times psi_y, paired with dpsi_z
          }
        }
#pragma unroll
        for (int jy = 0; jy < NYB; jy++) {
          const int yo = CPLX ? 4 * jy : 8 * jy;
          const double w1 = *reinterpret_cast<const double *>(row + aY + yo * 8 + hd.y);
          double dw1 = 0;
          if (GRAD) dw1 = *reinterpret_cast<const double *>(row + aY + yo * 8 + dOff + hd.y);
#pragma unroll
          for (int x = 0; x < XW; x++) {
            const int nb = x * NYB + jy;
            bA[nb] = w1 * ax[x];
            if (GRAD) { bA[nb] = fma(dw1, bx[x], bA[nb]); bB[nb] = w1 * cxw[x]; }
          }
        }
      };
      double az[NZB], adz[NZB], bA[NB], bB[NB], az_n[NZB], adz_n[NZB], bA_n[NB], bB_n[NB];
      prep(lo, az, adz, bA, bB);
      for (int b0 = lo; b0 < hi; b0 += 4) {
        prep(b0 + 4, az_n, adz_n, bA_n, bB_n);        // the next batch's operands build while this batch's MMAs run
#pragma unroll
        for (int nb = 0; nb < NB; nb++) {
#pragma unroll
          for (int zb = 0; zb < NZB; zb++) dmma884(acc[nb][zb], az[zb], bA[nb]);
          if (GRAD) {
#pragma unroll
            for (int zb = 0; zb < NZB; zb++) dmma884(acc[nb][zb], adz[zb], bB[nb]);
          }
        }
#pragma unroll
        for (int zb = 0; zb < NZB; zb++) { az[zb] = az_n[zb]; if (GRAD) adz[zb] = adz_n[zb]; }
#pragma unroll
        for (int nb = 0; nb < NB; nb++) { bA[nb] = bA_n[nb]; if (GRAD) bB[nb] = bB_n[nb]; }
      }
      dirty = Cfg::NFL;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }
  while (dirty > 0) { flush_advance(); dirty--; }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging must outlive the bulk reads
}

}  // namespace pnb
