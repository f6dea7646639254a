// Plan / node-set life cycle and the trafo / adj launch sequences (host side).
// Mirrors the reference's driver layer: api/api-basic.c:170-378 (trafo, adj and their ik variants),
// kernel/ndft-parallel.c:869-1069 (init_internal), :734-775 (node_borders), :788-822 (local sizes),
// api/api-guru.c:195-220 (fft_output_size) -- with every numeric stage a CUDA kernel on one stream.
#pragma once
#include <cmath>
#include <cstring>

#include "fftown.cuh"
#include "zmarch4.cuh"
#include "gridops.cuh"

namespace pnb {

void select_device_for_rank();

// flag values (ABI, include/pnfft.h)
enum : unsigned {
  F_PRE_PHI_HAT = 1u << 0, F_FAST_GAUSSIAN = 1u << 1, F_MALLOC_F_HAT = 1u << 6, F_INTERLACED = 1u << 8,
  F_TRANSPOSED_F_HAT = 1u << 11, F_DIFF_IK = 1u << 12, F_WIN_GAUSSIAN = 1u << 13, F_WIN_BSPLINE = 1u << 14,
  F_WIN_SINC_POWER = 1u << 15, F_WIN_BESSEL_I0 = 1u << 16, F_USE_FK_GAUSSIAN_T = 1u << 17, F_SORT_NODES = 1u << 18
};
enum : unsigned { N_MALLOC_X = 1u << 0, N_MALLOC_F = 1u << 1, N_MALLOC_GRAD_F = 1u << 2, N_MALLOC_HESSIAN_F = 1u << 3 };
enum : unsigned { P_PRE_FULL = 1u << 0, P_PRE_PSI = 1u << 1, P_PRE_GRAD_PSI = 1u << 2 };
enum : unsigned {
  C_F = 1u << 0, C_GRAD_F = 1u << 1, C_HESSIAN_F = 1u << 2, C_DIRECT = 1u << 3, C_ACCUMULATED = 1u << 4,
  C_OMIT_DECONV = 1u << 5, C_OMIT_FFT = 1u << 6, C_OMIT_CONV = 1u << 7
};
// job lists of the peer-memory data plane (p2p.cuh)
enum { LIST_FWD0 = 0, LIST_FWD1, LIST_FWD2, LIST_BWD2, LIST_BWD1, LIST_BWD0, LIST_FILL1, LIST_FILL0, LIST_RED0, LIST_RED1, LIST_COUNT };
enum { T_ITER = 0, T_WHOLE, T_LOOP_B, T_SORT_NODES, T_GCELLS, T_MATRIX_B, T_MATRIX_F, T_MATRIX_D, T_SHIFT_IN, T_SHIFT_OUT };

inline bool is_device_ptr(const void *p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

inline int window_kind(unsigned flags) {
  // precedence as in reference kernel/ndft-parallel.c:1650-1675
  if (flags & F_WIN_GAUSSIAN) return WIN_GAUSSIAN;
  if (flags & F_WIN_BSPLINE) return WIN_BSPLINE;
  if (flags & F_WIN_SINC_POWER) return WIN_SINC_POWER;
  if (flags & F_WIN_BESSEL_I0) return WIN_BESSEL_I0;
  return WIN_KAISER_BESSEL;
}

// window kind of the Fourier coefficients (D matrix, pnfft_phi_hat / pnfft_inv_phi_hat): PNFFT_WINDOW_GAUSSIAN_T keeps the
// Gaussian psi and takes the coefficients of its truncation (reference kernel/matrix_D.c:195-217)
inline int window_hat_kind(unsigned flags) {
  const int kind = window_kind(flags);
  return (kind == WIN_GAUSSIAN && (flags & F_USE_FK_GAUSSIAN_T)) ? WIN_GAUSSIAN_T : kind;
}

inline bool get_mesh(MPI_Comm comm, Mesh &M) {
  int nd = 0, dims[3] = {1, 1, 1}, periods[3], coords[3] = {0, 0, 0};
  if (MPI_Comm_rank(comm, &M.rank) != MPI_SUCCESS) return false;
  MPI_Comm_size(comm, &M.size);
  if (MPI_Cartdim_get(comm, &nd) != MPI_SUCCESS) return false;
  if (nd == 0) {
    if (M.size != 1) { fprintf(stderr, "pnfft-b200: communicator is not Cartesian; use pnfft_create_procmesh_2d\n"); return false; }
    M.np[0] = M.np[1] = 1; M.co[0] = M.co[1] = 0;
    return true;
  }
  MPI_Cart_get(comm, nd, dims, periods, coords);
  if (nd == 3 && dims[2] != 1) { fprintf(stderr, "pnfft-b200: 3-d process meshes are not supported (use a 2-d pencil mesh)\n"); return false; }
  M.np[0] = dims[0]; M.np[1] = nd > 1 ? dims[1] : 1;
  M.co[0] = coords[0]; M.co[1] = nd > 1 ? coords[1] : 0;
  return true;
}

template <class R>
inline void compute_layout(Layout &L, const Mesh &M, const INT *N, const INT *n, const R *x_max, int m, bool c2r, unsigned flags) {
  L.m = m; L.cutoff = 2 * m + 1; L.c2r = c2r;
  for (int t = 0; t < 3; t++) {
    L.N[t] = N[t]; L.n[t] = n[t];
    // reference api/api-guru.c:216-219: no = min(n, 2*(lrint(floor(n*x_max)) + m + 2))
    const INT c = (INT)lrint((double)m_floor((R)n[t] * x_max[t]));
    const INT cand = 2 * (c + m + 2);
    L.no[t] = n[t] < cand ? n[t] : cand;
    L.o_off[t] = n[t] / 2 - L.no[t] / 2;
    L.gcb[t] = m;
    L.gca[t] = L.cutoff - m - 1 + ((flags & F_INTERLACED) ? 1 : 0);
  }
  L.Nc2 = c2r ? N[2] / 2 + 1 : N[2];
  L.transposed = (flags & F_TRANSPOSED_F_HAT) != 0;
  const INT ext[3] = {N[0], N[1], L.Nc2};
  if (!L.transposed) {
    for (int t = 0; t < 2; t++) block_1d(ext[t], M.np[t], M.co[t], &L.local_N[t], &L.local_N_start[t]);
    L.local_N[2] = ext[2]; L.local_N_start[2] = 0;
  } else {
    // k1 over mesh dim 0, k2 over mesh dim 1, k0 whole (PFFT_TRANSPOSED_IN of a 2-d mesh, reference doc/intro.tex:39-44)
    block_1d(ext[1], M.np[0], M.co[0], &L.local_N[1], &L.local_N_start[1]);
    block_1d(ext[2], M.np[1], M.co[1], &L.local_N[2], &L.local_N_start[2]);
    L.local_N[0] = ext[0]; L.local_N_start[0] = 0;
  }
  for (int t = 0; t < 2; t++) block_1d(L.no[t], M.np[t], M.co[t], &L.local_no[t], &L.local_no_start[t]);
  L.local_no[2] = L.no[2]; L.local_no_start[2] = 0;
  for (int t = 0; t < 3; t++) {
    L.local_N_start[t] -= N[t] / 2;
    L.local_no_start[t] -= L.no[t] / 2;
    L.ngc[t] = L.local_no[t] + L.gcb[t] + L.gca[t];
  }
  L.pitch2 = (L.ngc[2] + 3) / 4 * 4;
}

template <class R>
inline void node_borders(const Layout &L, const R *x_max, R *lo, R *up) {
  for (int t = 0; t < 3; t++) {
    lo[t] = (R)L.local_no_start[t] / (R)L.n[t];
    up[t] = (R)(L.local_no_start[t] + L.local_no[t]) / (R)L.n[t];
    if (lo[t] < -x_max[t]) lo[t] = -x_max[t];
    if (lo[t] > x_max[t]) lo[t] = x_max[t];
    if (up[t] < -x_max[t]) up[t] = -x_max[t];
    if (up[t] > x_max[t]) up[t] = x_max[t];
  }
}

// ---------------------------------------------------------------------------------------------
// TMA descriptor of the padded grid
// ---------------------------------------------------------------------------------------------
typedef CUresult (*TmapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    PNB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) { fprintf(stderr, "pnfft-b200: cuTensorMapEncodeTiled unavailable\n"); abort(); }
    fn = (TmapEncodeFn)p;
  }
  return fn;
}

template <class R>
inline CUtensorMap make_grid_tmap(void *grid, const Layout &L, int ncomp, int bx, int by, int bz, int swizzle = 0) {
  CUtensorMap tm;
  const cuuint64_t gdim[3] = {(cuuint64_t)L.pitch2 * ncomp, (cuuint64_t)L.ngc[1], (cuuint64_t)L.ngc[0]};
  const cuuint64_t gstr[2] = {(cuuint64_t)L.pitch2 * ncomp * sizeof(R), (cuuint64_t)L.ngc[1] * L.pitch2 * ncomp * sizeof(R)};
  const cuuint32_t box[3] = {(cuuint32_t)(bz * ncomp), (cuuint32_t)by, (cuuint32_t)bx};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = sizeof(R) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : (swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUresult r = tmap_encoder()(&tm, dt, 3, grid, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "pnfft-b200: cuTensorMapEncodeTiled failed (%d)\n", (int)r); abort(); }
  return tm;
}

}  // namespace pnb
#include "direct.cuh"
namespace pnb {

// ---------------------------------------------------------------------------------------------
template <class R> struct Core {
  typedef Plan<R> P;
  typedef Nodes<R> Nd;
  typedef typename Vec2<R>::type C;

  static size_t cell_bytes(const P *p) { return p->L.c2r ? sizeof(R) : sizeof(C); }
  static INT local_N_total(const P *p) { return p->L.local_N[0] * p->L.local_N[1] * p->L.local_N[2]; }

  static GridGeom<R> geom(const P *p) {
    GridGeom<R> g;
    g.m = p->L.m; g.cutoff = p->L.cutoff; g.kind = p->kind;
    g.fast_gauss = (p->pnfft_flags & F_FAST_GAUSSIAN) ? 1 : 0;
    for (int t = 0; t < 3; t++) {
      g.n[t] = (R)p->L.n[t]; g.b[t] = p->b[t];
      g.los[t] = (int)p->L.local_no_start[t]; g.lno[t] = (int)p->L.local_no[t]; g.ngc[t] = (int)p->L.ngc[t];
    }
    g.pitch1 = p->L.pitch2;
    g.pitch0 = (long long)p->L.ngc[1] * p->L.pitch2;
    g.exp_const = p->d_exp_const;
    g.poly = (p->use_poly && p->poly_deg >= 0 && p->intpol_order < 0) ? p->d_poly : nullptr;
    g.intpol_order = p->intpol_order; g.intpol_num = p->intpol_num;
    for (int q = 0; q < 9; q++) g.intpol_tab[q] = p->d_intpol[q];
    g.poly_deg = p->poly_deg;
    g.poly_deg_psi = p->poly_deg_psi;
    const bool il = (p->pnfft_flags & F_INTERLACED) != 0;
    g.il_on = (il && p->il_pass == 1) ? 1 : 0;
    for (int t = 0; t < 3; t++) g.il[t] = 0.5 / (double)p->L.n[t];
    g.wscale = il ? (R)0.5 : (R)1;
    return g;
  }

  // (re)compute everything that depends on the window shape b: 1/phi_hat tables, fast-Gaussian constants
  // (reference PNX(init_precompute_window) kernel/ndft-parallel.c:1072-1140, matrix_D.c:453-482)
  static void upload_window_tables(P *p) {
    const Layout &L = p->L;
    for (int t = 0; t < 3; t++) {
      const INT len = L.local_N[t];
      std::vector<R> h((size_t)(len > 0 ? len : 1));
      for (INT i = 0; i < len; i++)
        h[(size_t)i] = phi_hat_any<R>(window_hat_kind(p->pnfft_flags), (long)(L.local_N_start[t] + i), (long)L.n[t], p->b[t], L.m, true);
      if (!p->d_invphi[t]) PNB_CUDA(cudaMalloc(&p->d_invphi[t], sizeof(R) * h.size()));
      PNB_CUDA(cudaMemcpy(p->d_invphi[t], h.data(), sizeof(R) * h.size(), cudaMemcpyHostToDevice));
    }
    if (p->pnfft_flags & F_FAST_GAUSSIAN) {
      const int c = L.cutoff;
      std::vector<R> h((size_t)3 * c);
      for (int t = 0; t < 3; t++)
        for (int s = 0; s < c; s++)
          h[(size_t)(t * c + s)] = m_exp(-(R)(s * s) / p->b[t]) / m_sqrt(m_pi<R>() * p->b[t]);
      if (!p->d_exp_const) PNB_CUDA(cudaMalloc(&p->d_exp_const, sizeof(R) * h.size()));
      PNB_CUDA(cudaMemcpy(p->d_exp_const, h.data(), sizeof(R) * h.size(), cudaMemcpyHostToDevice));
    }
    fit_window_polys(p);
    build_intpol_tables(p);
  }

  // PNFFT_PRE_{CONST,LIN,QUAD,CUB}_PSI tables (reference init_intpol_table_psi kernel/ndft-parallel.c:321-353, sizes
  // :1088-1122): entry [k][c][i] = psi((m + (k + i) / num - c) / n), i = -order/2 .. (order+1)/2, for psi, dpsi, ddpsi
  static void build_intpol_tables(P *p) {
    if (p->intpol_order < 0) return;
    const Layout &L = p->L;
    const int c = L.cutoff, m = L.m, o = p->intpol_order;
    p->intpol_num = (int)std::ceil((2.0 * 15.0 + 1.0) / c) * 2048;
    const size_t len = (size_t)p->intpol_num * c * (o + 1);
    std::vector<R> h(len);
    for (int q = 0; q < 3; q++)
      for (int t = 0; t < 3; t++) {
        const R n = (R)L.n[t], b = p->b[t];
        size_t ind = 0;
        for (int k = 0; k < p->intpol_num; k++)
          for (int cc = 0; cc < c; cc++)
            for (int i = -o / 2; i <= (o + 1) / 2; i++, ind++) {
              const R x = ((R)m + (R)(k + i) / (R)p->intpol_num - (R)cc) / n;      // the reference's argument, its arithmetic
              const R z = n * x;
              R psi = 0, d = 0;
              if (p->kind == WIN_BSPLINE) {
                psi = bspline<R>(2 * m, z + (R)m);
                d = n * (bspline<R>(2 * m - 1, z + (R)m) - bspline<R>(2 * m - 1, z + (R)m - (R)1));
              } else {
                window_tap<R>(p->kind, -z, n, b, m, true, &psi, &d);
              }
              h[ind] = q == 0 ? psi : (q == 1 ? d : window_ddtap<R>(p->kind, -z, n, b, m, psi, d));
            }
        R *&dst = p->d_intpol[3 * q + t];
        if (!dst) PNB_CUDA(cudaMalloc((void **)&dst, sizeof(R) * len));
        PNB_CUDA(cudaMemcpy(dst, h.data(), sizeof(R) * len, cudaMemcpyHostToDevice));
      }
  }


  // Per-tap polynomial form of the window (the kernels' on-the-fly evaluation): tap s of axis t as a function
  // of frac = n x - floor(n x) in [0,1) is analytic for every supported window, so a Chebyshev interpolant of
  // modest degree reproduces it to rounding.  Fitted here in double from the exact formulas of window.h, the
  // degree is raised until the next Chebyshev coefficients are below 5e-16 (double) / 1e-9 (float) of the window
  // maximum; if degree 24 is not enough the kernels keep the exact evaluation.
  static void fit_window_polys(P *p) {
    // Double precision keeps the exact formulas: measured on B200 (C3) the node-table kernel is bound by row assembly
    // and its HBM write, not by the window evaluation (exact 4.2 / 8.8 ms against 4.4 / 11.8 ms with degree 13 / 21
    // polynomials), and the exact path follows the reference's arithmetic more closely.  Float uses the polynomials.
    if (sizeof(R) == 8 && !getenv("PNFFT_B200_POLY_DOUBLE")) { p->poly_deg = p->poly_deg_psi = -1; return; }
    const Layout &L = p->L;
    const int c = L.cutoff, nv = 3 * c, m = L.m;
    const int NP = 48;
    const double pi = 3.14159265358979323846;
    // Chebyshev coefficients of psi (set 0) and of the AD-gradient weight dpsi (set 1), each tap and axis
    std::vector<double> ck((size_t)2 * nv * NP);
    double vmax[2] = {0, 0};
    for (int v = 0; v < nv; v++) {
      const int t = v / c, s = v - t * c;
      double f[2][NP];
      for (int j = 0; j < NP; j++) {
        const double u = cos(pi * (j + 0.5) / NP), frac = 0.5 * (u + 1.0);
        double psi = 0, d = 0;
        window_tap<double>(p->kind, (double)s - (double)m - frac, (double)L.n[t], (double)p->b[t], m, true, &psi, &d);
        f[0][j] = psi; f[1][j] = d;
        vmax[0] = std::max(vmax[0], fabs(psi));
        vmax[1] = std::max(vmax[1], fabs(d));
      }
      for (int q = 0; q < 2; q++)
        for (int k = 0; k < NP; k++) {
          long double acc = 0;
          for (int j = 0; j < NP; j++) acc += (long double)f[q][j] * cosl((long double)pi * k * (j + 0.5L) / NP);
          ck[((size_t)q * nv + v) * NP + k] = (double)(acc * 2.0L / NP * (k == 0 ? 0.5L : 1.0L));
        }
    }
    // The coefficients come from double evaluations of the window, so they bottom out at a noise floor of ~1e-16 of
    // the maximum; a tail-sum criterion below that floor can never be met (it silently disabled the polynomials in
    // double).  Take the first degree after which three consecutive coefficients of both sets are below tol.
    const double rtol = sizeof(R) == 8 ? 5e-16 : 1e-9;
    int deg = -1;
    for (int D = 2; D <= kMaxPolyCoef - 1 && deg < 0; D++) {
      bool ok = true;
      for (int q = 0; q < 2 && ok; q++)
        for (int v = 0; v < nv && ok; v++)
          for (int k = D + 1; k <= D + 3; k++)
            if (fabs(ck[((size_t)q * nv + v) * NP + k]) > rtol * vmax[q]) { ok = false; break; }
      if (ok) deg = D;
    }
    p->poly_deg = deg;
    // psi alone usually needs a lower degree than the derivative weights (Kaiser-Bessel m=6: 13 against 21)
    p->poly_deg_psi = deg;
    for (int D = 2; D < deg; D++) {
      bool ok = true;
      for (int v = 0; v < nv && ok; v++)
        for (int k = D + 1; k <= D + 3; k++)
          if (fabs(ck[(size_t)v * NP + k]) > rtol * vmax[0]) { ok = false; break; }
      if (ok) { p->poly_deg_psi = D; break; }
    }
    if (deg < 0) return;
    // Chebyshev -> monomial coefficients in u; device layout [set][k][v]
    std::vector<R> h((size_t)2 * (deg + 1) * nv);
    std::vector<long double> Tkm1((size_t)deg + 1), Tk((size_t)deg + 1), Tn((size_t)deg + 1), a((size_t)deg + 1);
    for (int q = 0; q < 2; q++)
      for (int v = 0; v < nv; v++) {
        const double *cv = &ck[((size_t)q * nv + v) * NP];
        std::fill(a.begin(), a.end(), 0.0L);
        std::fill(Tkm1.begin(), Tkm1.end(), 0.0L);
        std::fill(Tk.begin(), Tk.end(), 0.0L);
        Tkm1[0] = 1.0L;                       // T_0
        if (deg >= 1) Tk[1] = 1.0L;           // T_1
        a[0] += cv[0] * Tkm1[0];
        if (deg >= 1) a[1] += cv[1];
        for (int k = 2; k <= deg; k++) {
          std::fill(Tn.begin(), Tn.end(), 0.0L);
          for (int i = 0; i < deg; i++) Tn[(size_t)i + 1] += 2.0L * Tk[(size_t)i];
          for (int i = 0; i <= deg; i++) Tn[(size_t)i] -= Tkm1[(size_t)i];
          for (int i = 0; i <= deg; i++) a[(size_t)i] += (long double)cv[k] * Tn[(size_t)i];
          Tkm1 = Tk; Tk = Tn;
        }
        for (int k = 0; k <= deg; k++) h[((size_t)q * (deg + 1) + k) * nv + v] = (R)a[(size_t)k];
      }
    if (!p->d_poly) PNB_CUDA(cudaMalloc((void **)&p->d_poly, sizeof(R) * (size_t)2 * kMaxPolyCoef * nv));
    PNB_CUDA(cudaMemcpy(p->d_poly, h.data(), sizeof(R) * h.size(), cudaMemcpyHostToDevice));
  }

  static P *init(const INT *N, const INT *n, const R *x_max, int m, unsigned pnfft_flags, unsigned pfft_flags,
                 MPI_Comm comm_cart, bool c2r) {
    Mesh mesh;
    if (!get_mesh(comm_cart, mesh)) return nullptr;
    for (int t = 0; t < 3; t++) {
      if (n[t] % 2) { fprintf(stderr, "pnfft-b200: odd FFT sizes n are not supported\n"); return nullptr; }
      if (n[t] < N[t]) { fprintf(stderr, "pnfft-b200: n < N\n"); return nullptr; }
    }
    if (m < 1 || m > kMaxM) { fprintf(stderr, "pnfft-b200: window cutoff m must be in [1,%d]\n", kMaxM); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      fprintf(stderr, "pnfft-b200: no CUDA device -- this library has no CPU path\n");
      return nullptr;
    }
    select_device_for_rank();
    if (mesh.size > 1) (void)world_nccl();   // collective: create the NCCL communicator outside of any ncclGroup

    P *p = new P();
    p->mesh = mesh;
    MPI_Comm_dup(comm_cart, &p->comm);
    p->pnfft_flags = pnfft_flags; p->pfft_flags = pfft_flags;
    p->kind = window_kind(pnfft_flags);
    {
      // interpolation order, with the reference's precedence (kernel/ndft-parallel.c:997-1006) AND its flag promotion
      // (api/api-guru.c:152-155 tests the plan flags against precompute-namespace constants: PNFFT_PRE_LIN_PSI (1 << 3)
      // turns PNFFT_PRE_CONST_PSI (1 << 2) on, so "linear" interpolates with order 0 there; a drop-in does the same)
      unsigned fl = pnfft_flags;
      if (fl & (1u << 3)) fl |= 1u << 2;
      p->intpol_order = (fl & (1u << 2)) ? 0 : ((fl & (1u << 3)) ? 1 : ((fl & (1u << 4)) ? 2 : ((fl & (1u << 5)) ? 3 : -1)));
    }
    compute_layout<R>(p->L, mesh, N, n, x_max, m, c2r, pnfft_flags);
    Layout &L = p->L;
    for (int t = 0; t < 3; t++) {
      p->x_max[t] = x_max[t];
      p->sigma[t] = (R)n[t] / (R)N[t];
      p->b[t] = window_shape<R>(p->kind, m, p->sigma[t]);
    }
    // every rank evaluates the condition for EVERY block of the split axes, so that all ranks refuse together (a rank
    // that alone returned NULL would leave the others waiting in the first exchange)
    for (int t = 0; t < 2; t++)
      if (mesh.np[t] > 1) {
        const INT need = std::max(L.gcb[t], L.gca[t]);
        for (int c = 0; c < mesh.np[t]; c++) {
          INT len, start;
          block_1d(L.no[t], mesh.np[t], c, &len, &start);
          if (len < need) {
            if (mesh.rank == 0)
              fprintf(stderr, "pnfft-b200: grid block %d of axis %d (%td cells) is narrower than the halo (%td): use fewer ranks along it\n",
                      c, t, len, need);
            MPI_Comm_free(&p->comm);
            delete p; return nullptr;
          }
        }
      }
    // the plan's stream carries F and the ghost-cell exchange, whose peer-memory stages other ranks wait for: it gets the
    // highest priority, the node-side stream (binning, window table: 10^5 small blocks that would otherwise occupy every
    // SM while F's kernels queue behind them) the lowest.  PNFFT_B200_STREAM_PRIO=0 creates them all alike.
    int prio_lo = 0, prio_hi = 0;
    static const bool use_prio = env_flag("PNFFT_B200_STREAM_PRIO", true);
    if (use_prio) PNB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    PNB_CUDA(cudaStreamCreateWithPriority(&p->stream, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < 16; i++) PNB_CUDA(cudaEventCreate(&p->ev[i]));
    PNB_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    PNB_CUDA(cudaStreamCreateWithPriority(&p->node_stream, cudaStreamNonBlocking, prio_lo));
    for (int i = 0; i < 3; i++) PNB_CUDA(cudaEventCreateWithFlags(&p->ev_copy[i], cudaEventDisableTiming));
    memset(p->timer_trafo, 0, sizeof p->timer_trafo);
    memset(p->timer_adj, 0, sizeof p->timer_adj);
    memset(p->stage_ms, 0, sizeof p->stage_ms);

    const size_t nloc = (size_t)local_N_total(p);
    if (pnfft_flags & F_MALLOC_F_HAT) {
      PNB_CUDA(cudaHostAlloc((void **)&p->f_hat, sizeof(C) * (nloc ? nloc : 1), cudaHostAllocDefault));
      p->owns_f_hat = true;
    }
    PNB_CUDA(cudaMalloc((void **)&p->d_f_hat, sizeof(C) * (nloc ? nloc : 1)));
    PNB_CUDA(cudaMalloc((void **)&p->d_g1, sizeof(C) * (nloc ? nloc : 1)));
    if (pnfft_flags & F_DIFF_IK) PNB_CUDA(cudaMalloc((void **)&p->d_g1_buffer, sizeof(C) * (nloc ? nloc : 1)));
    p->grid_bytes = (size_t)L.ngc[0] * L.ngc[1] * L.pitch2 * cell_bytes(p);
    PNB_CUDA(cudaMalloc(&p->d_grid, p->grid_bytes ? p->grid_bytes : 16));
    PNB_CUDA(cudaMemsetAsync(p->d_grid, 0, p->grid_bytes, p->stream));
    p->pipe = build_pipe(L, mesh);
    p->work_bytes = sizeof(C) * (size_t)p->pipe.buf_elems;
    // halo exchange buffers also live in the work buffers
    if (mesh.size > 1) {
      const size_t h0 = (size_t)(L.gcb[0] + L.gca[0]) * L.ngc[1] * L.pitch2 * cell_bytes(p) * 2;
      const size_t h1 = (size_t)(L.gcb[1] + L.gca[1]) * L.ngc[0] * L.pitch2 * cell_bytes(p) * 2;
      p->work_bytes = std::max(p->work_bytes, std::max(h0, h1));
    }
    for (int i = 0; i < 3; i++) PNB_CUDA(cudaMalloc(&p->d_work[i], p->work_bytes));
    upload_window_tables(p);
    make_fft_plans(p);
    PNB_CUDA(cudaStreamSynchronize(p->stream));
    setup_peer(p, N, n, x_max, m, c2r);
    return p;
  }

  // -------------------------------------------------------------------------------------------
  // peer-memory data plane (p2p.cuh): IPC mapping of every rank's buffers, job lists of all exchange stages
  // -------------------------------------------------------------------------------------------
  static void setup_peer(P *p, const INT *N, const INT *n, const R *x_max, int m, bool c2r) {
    PeerPlane &pp = p->peer;
    const Mesh &M = p->mesh;
    pp.nranks = M.size; pp.rank = M.rank;
    static const bool want = env_flag("PNFFT_B200_P2P", true);
    if (M.size <= 1 || !want) return;
    const int Pn = M.size;
    if (Pn > 32) return;
    int *my_flags = nullptr;
    PNB_CUDA(cudaMalloc((void **)&my_flags, sizeof(int) * 64));
    PNB_CUDA(cudaMemset(my_flags, 0, sizeof(int) * 64));
    void *mine[6] = {p->d_work[0], p->d_work[1], p->d_work[2], p->d_grid, (void *)p->d_g1, (void *)my_flags};
    std::vector<cudaIpcMemHandle_t> hs((size_t)6 * Pn);
    int ok = 1;
    for (int k = 0; k < 6; k++)
      if (cudaIpcGetMemHandle(&hs[(size_t)6 * M.rank + k], mine[k]) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    for (int q = 0; q < Pn; q++) MPI_Bcast(&hs[(size_t)6 * q], (int)(6 * sizeof(cudaIpcMemHandle_t)), MPI_BYTE, q, p->comm);
    for (int k = 0; k < 3; k++) pp.work[k].assign((size_t)Pn, nullptr);
    pp.grid.assign((size_t)Pn, nullptr); pp.g1.assign((size_t)Pn, nullptr); pp.flags.assign((size_t)Pn, nullptr);
    for (int q = 0; q < Pn && ok; q++) {
      void *ptr[6];
      for (int k = 0; k < 6; k++) {
        if (q == M.rank) { ptr[k] = mine[k]; continue; }
        ptr[k] = nullptr;
        if (cudaIpcOpenMemHandle(&ptr[k], hs[(size_t)6 * q + k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      }
      for (int k = 0; k < 3; k++) pp.work[k][(size_t)q] = ptr[k];
      pp.grid[(size_t)q] = ptr[3]; pp.g1[(size_t)q] = ptr[4]; pp.flags[(size_t)q] = (int *)ptr[5];
    }
    // ---- job lists: every rank's layout and pipe, maps composed rank to rank ----
    std::vector<JobList> lists((size_t)LIST_COUNT);
    for (auto &l : lists) { l.n = 0; l.nblocks = 0; }
    if (ok) ok = build_peer_jobs(p, N, n, x_max, m, c2r, lists) ? 1 : 0;
    int all_ok = 0;
    MPI_Allreduce(&ok, &all_ok, 1, MPI_INT, MPI_MIN, p->comm);
    if (!all_ok) {
      if (M.rank == 0) fprintf(stderr, "pnfft-b200: peer-memory exchange unavailable (CUDA IPC or map composition); using NCCL send/recv\n");
      close_peer(p);
      cudaFree(my_flags);
      return;
    }
    for (int k = 0; k < LIST_COUNT; k++) { finish_jobs(lists[(size_t)k]); pp.nblocks[k] = lists[(size_t)k].nblocks; }
    PNB_CUDA(cudaMalloc((void **)&pp.d_jobs, sizeof(JobList) * LIST_COUNT));
    PNB_CUDA(cudaMemcpy(pp.d_jobs, lists.data(), sizeof(JobList) * LIST_COUNT, cudaMemcpyHostToDevice));
    PNB_CUDA(cudaMalloc((void **)&pp.d_flag_ptrs, sizeof(int *) * Pn));
    PNB_CUDA(cudaMemcpy(pp.d_flag_ptrs, pp.flags.data(), sizeof(int *) * Pn, cudaMemcpyHostToDevice));
    PNB_CUDA(cudaMalloc((void **)&pp.d_error, sizeof(int)));
    PNB_CUDA(cudaMemset(pp.d_error, 0, sizeof(int)));
    PNB_CUDA(cudaHostAlloc((void **)&pp.h_error, sizeof(int), cudaHostAllocDefault));
    *pp.h_error = 0;
    pp.on = true;
    MPI_Barrier(p->comm);       // every rank's flags are zeroed and mapped before the first epoch is released
  }
  static void close_peer(P *p) {
    PeerPlane &pp = p->peer;
    for (int q = 0; q < (int)pp.grid.size(); q++) {
      if (q == pp.rank) continue;
      for (int k = 0; k < 3; k++) if (pp.work[k][(size_t)q]) cudaIpcCloseMemHandle(pp.work[k][(size_t)q]);
      if (pp.grid[(size_t)q]) cudaIpcCloseMemHandle(pp.grid[(size_t)q]);
      if (pp.g1[(size_t)q]) cudaIpcCloseMemHandle(pp.g1[(size_t)q]);
      if (pp.flags[(size_t)q]) cudaIpcCloseMemHandle(pp.flags[(size_t)q]);
    }
    for (int k = 0; k < 3; k++) pp.work[k].clear();
    pp.grid.clear(); pp.g1.clear(); pp.flags.clear();
    pp.on = false;
  }

  static bool build_peer_jobs(P *p, const INT *N, const INT *n, const R *x_max, int m, bool c2r, std::vector<JobList> &lists) {
    const PeerPlane &pp = p->peer;
    const Mesh &Mm = p->mesh;
    const int Pn = Mm.size, me = Mm.rank;
    std::vector<Layout> Ls((size_t)Pn);
    std::vector<PipeGeom> pipes((size_t)Pn);
    for (int q = 0; q < Pn; q++) {
      Mesh mq = Mm;
      mq.rank = q; mq.co[0] = q / Mm.np[1]; mq.co[1] = q % Mm.np[1];
      compute_layout<R>(Ls[(size_t)q], mq, N, n, x_max, m, c2r, p->pnfft_flags);
      pipes[(size_t)q] = build_pipe(Ls[(size_t)q], mq);
    }
    auto entry_for = [&](const Stage &S, int peer) -> const Transfer * {
      for (const auto &T : S.tr) if (T.peer == peer) return &T;
      return nullptr;
    };
    // work buffer that holds the destination-side / source-side array of each pipeline stage (Core::fft_forward / _backward)
    const int dst_buf[3] = {1, 0, 1};      // L1 in W1, L3 in W0, L4 in W1
    auto src_ptr = [&](int s, int q) -> void * { return s == 0 ? pp.g1[(size_t)q] : pp.work[dst_buf[s - 1]][(size_t)q]; };
    const PipeGeom &G = p->pipe;
    for (int s = 0; s < 3; s++) {
      JobList &LF = lists[(size_t)(LIST_FWD0 + s)], &LB = lists[(size_t)(LIST_BWD0 - s)];
      for (const auto &T : G.st[s].tr) {
        const int q = T.peer;
        if (q == me && !T.self_maps.empty()) {
          for (const auto &bm : T.self_maps) {
            if (!push_job(LF, bm, pp.work[dst_buf[s]][(size_t)me], true, src_ptr(s, me), T.send_sign, false)) return false;
            if (!push_job(LB, bm, src_ptr(s, me), false, pp.work[dst_buf[s]][(size_t)me], T.send_sign, false)) return false;
          }
          continue;
        }
        const Transfer *Tq = entry_for(pipes[(size_t)q].st[s], me);
        if (!Tq) return false;
        // forward: my source array -> q's destination array  (my send map o q's receive maps)
        if (T.send_elems > 0) {
          if (T.send_maps.size() != 1) return false;
          for (const auto &rm : Tq->recv_maps) {
            BoxMap c;
            if (!compose_self_map(T.send_maps[0], rm, &c)) return false;
            if (!push_job(LF, c, pp.work[dst_buf[s]][(size_t)q], true, src_ptr(s, me), T.send_sign, false)) return false;
          }
        }
        // backward: my destination-side array -> q's source-side array  (q's send map o my receive maps)
        if (T.recv_elems > 0) {
          if (Tq->send_maps.size() != 1) return false;
          for (const auto &rm : T.recv_maps) {
            BoxMap c;
            if (!compose_self_map(Tq->send_maps[0], rm, &c)) return false;
            if (!push_job(LB, c, src_ptr(s, q), false, pp.work[dst_buf[s]][(size_t)me], Tq->send_sign, false)) return false;
          }
        }
      }
    }
    // ---- ghost cells along the split axes: fill = push my border rows into the neighbours' halos (axis 1, then axis 0,
    //      which carries the axis-1 halos => corners); reduce = pull the neighbours' halos and add (axis 0, then axis 1) ----
    const Layout &L = p->L;
    for (int axis = 0; axis < 2; axis++) {
      if (Mm.np[axis] <= 1) continue;
      JobList &LFi = lists[(size_t)(axis == 1 ? LIST_FILL1 : LIST_FILL0)], &LR = lists[(size_t)(axis == 0 ? LIST_RED0 : LIST_RED1)];
      int cu[2] = {Mm.co[0], Mm.co[1]}, cd[2] = {Mm.co[0], Mm.co[1]};
      cu[axis] += 1; cd[axis] -= 1;
      const int up = Mm.rank_of(cu[0], cu[1]), down = Mm.rank_of(cd[0], cd[1]);
      const Layout &Lu = Ls[(size_t)up], &Ld = Ls[(size_t)down];
      const long long gcb = L.gcb[axis], gca = L.gca[axis], lno = L.local_no[axis];
      long long lo[3], ext[3];
      for (int t = 0; t < 3; t++) {
        const bool wide = t > axis;            // axes filled before this one span the padded extent (Core::halo_T)
        lo[t] = wide ? 0 : L.gcb[t];
        ext[t] = wide ? L.ngc[t] : L.local_no[t];
      }
      auto region = [&](const Layout &LL, long long start, long long width, long long *off, long long *str, long long *dims) {
        long long s3[3] = {lo[0], lo[1], lo[2]}, d3[3] = {ext[0], ext[1], ext[2]};
        s3[axis] = start; d3[axis] = width;
        str[0] = (long long)LL.ngc[1] * LL.pitch2; str[1] = LL.pitch2; str[2] = 1;
        *off = s3[0] * str[0] + s3[1] * str[1] + s3[2];
        for (int t = 0; t < 3; t++) dims[t] = d3[t];
      };
      auto add_job = [&](JobList &JL, void *dst, const Layout &Ldst, long long dstart, const void *src, const Layout &Lsrc, long long sstart,
                         long long width, bool add) -> bool {
        if (JL.n >= kMaxJobs) return false;
        BoxJob &J = JL.j[JL.n++];
        long long dd[3];
        region(Ldst, dstart, width, &J.d_off, J.d_str, J.dims);
        region(Lsrc, sstart, width, &J.s_off, J.s_str, dd);
        J.dst = dst; J.src = src; J.parity = 0; J.sign = 0; J.add = add ? 1 : 0; J.first_block = 0; J.bx = J.by = 1;
        return true;
      };
      void *gme = pp.grid[(size_t)me];
      // fill: my top gcb interior rows -> up's below halo; my bottom gca interior rows -> down's above halo
      if (!add_job(LFi, pp.grid[(size_t)up], Lu, 0, gme, L, gcb + lno - gcb, gcb, false)) return false;
      if (!add_job(LFi, pp.grid[(size_t)down], Ld, Ld.gcb[axis] + Ld.local_no[axis], gme, L, gcb, gca, false)) return false;
      // reduce: down's above halo -> my bottom interior rows; up's below halo -> my top interior rows
      if (!add_job(LR, gme, L, gcb, pp.grid[(size_t)down], Ld, Ld.gcb[axis] + Ld.local_no[axis], gca, true)) return false;
      if (!add_job(LR, gme, L, gcb + lno - gcb, pp.grid[(size_t)up], Lu, 0, gcb, true)) return false;
    }
    return true;
  }

  static void peer_barrier(P *p) {
    PeerPlane &pp = p->peer;
    pp.epoch++;
    k_peer_barrier<<<1, 32, 0, p->stream>>>(pp.d_flag_ptrs, pp.flags[(size_t)pp.rank], pp.rank, pp.nranks, pp.epoch, pp.d_error);
    p->launches++;
  }
  template <class T> static void peer_jobs(P *p, int list) {
    PeerPlane &pp = p->peer;
    if (pp.nblocks[list] <= 0) return;
    k_box_jobs<T><<<(unsigned)pp.nblocks[list], 256, 0, p->stream>>>(pp.d_jobs + list);
    p->launches++;
  }
  // after a call: did a barrier give up waiting for a peer?
  static void peer_check(P *p) {
    PeerPlane &pp = p->peer;
    if (!pp.on) return;
    PNB_CUDA(cudaMemcpyAsync(pp.h_error, pp.d_error, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    PNB_CUDA(cudaStreamSynchronize(p->stream));
    if (*pp.h_error) { fprintf(stderr, "pnfft-b200: rank %d: a peer did not reach exchange barrier %d (epoch now %d)\n", pp.rank, *pp.h_error, pp.epoch); abort(); }
  }

  static void make_fft_plans(P *p) {
    const Layout &L = p->L;
    const PipeGeom &G = p->pipe;
    // complex passes of power-of-two length can run on the library's own kernels (fftown.cuh): PNFFT_B200_OWN_FFT=1.
    // Measured on B200 (n = 512^3, profiles/r2_fft_own.md): cuFFT's kernels are faster than this first version, so cuFFT
    // stays the default.
    const bool force_cufft = !(getenv("PNFFT_B200_OWN_FFT") && atoi(getenv("PNFFT_B200_OWN_FFT")) != 0);    // read per plan
    for (int t = 0; t < 3; t++) {
      p->own_fft[t] = !force_cufft && fft_own_ok(L.n[t]) && !(t == 2 && L.c2r);
      if (p->own_fft[t]) {
        for (int u = 0; u < t; u++) if (p->own_fft[u] && L.n[u] == L.n[t]) p->d_tw[t] = p->d_tw[u];
        if (!p->d_tw[t]) { C *tw = nullptr; fft_make_twiddles<C>((int)L.n[t], &tw); p->d_tw[t] = tw; }
      }
    }
    if (!p->own_fft[0]) p->fft_x = make_plan_1d(L.n[0], G.S1, 1, G.S1, FftType<R>::c2c, 0, p->stream);
    if (!p->own_fft[1]) p->fft_y = make_plan_1d(L.n[1], G.S3, 1, G.S3, FftType<R>::c2c, 0, p->stream);
    const long long nb = (long long)L.local_no[0] * L.local_no[1];
    if (p->own_fft[2]) {
    } else if (!L.c2r) {
      p->fft_z_fwd = make_plan_1d(L.n[2], 1, L.n[2], nb, FftType<R>::c2c, 0, p->stream);
      p->fft_z_bwd = p->fft_z_fwd;
    } else {
      p->fft_z_fwd = make_plan_1d(L.n[2], 1, 0, nb, FftType<R>::c2r, L.n[2], p->stream);
      p->fft_z_bwd = make_plan_1d(L.n[2], 1, 0, nb, FftType<R>::r2c, L.n[2], p->stream);
    }
  }

  static void finalize(P *p, unsigned flags) {
    if (!p) return;
    cudaStreamSynchronize(p->stream);
    if (p->peer.on) {          // nobody unmaps or frees while a peer may still touch the buffers
      int *my_flags = p->peer.flags[(size_t)p->peer.rank];
      MPI_Barrier(p->comm);
      close_peer(p);
      MPI_Barrier(p->comm);
      cudaFree(my_flags); cudaFree(p->peer.d_jobs); cudaFree(p->peer.d_flag_ptrs); cudaFree(p->peer.d_error); cudaFreeHost(p->peer.h_error);
    }
    if (p->fft_x) cufftDestroy(p->fft_x);
    if (p->fft_y) cufftDestroy(p->fft_y);
    if (p->fft_z_fwd) cufftDestroy(p->fft_z_fwd);
    if (p->fft_z_bwd && p->fft_z_bwd != p->fft_z_fwd) cufftDestroy(p->fft_z_bwd);
    if (p->owns_f_hat && (flags & F_MALLOC_F_HAT) && p->f_hat) cudaFreeHost(p->f_hat);
    cudaFree(p->d_f_hat); cudaFree(p->d_g1); cudaFree(p->d_g1_buffer); cudaFree(p->d_grid);
    cudaFree(p->d_work[0]); cudaFree(p->d_work[1]); cudaFree(p->d_work[2]); cudaFree(p->d_exp_const); cudaFree(p->d_sort_tmp); cudaFree(p->d_poly);
    for (int t = 0; t < 3; t++) cudaFree(p->d_invphi[t]);
    for (int q = 0; q < 9; q++) cudaFree(p->d_intpol[q]);
    for (int t = 0; t < 3; t++) {
      bool shared = false;
      for (int u = 0; u < t; u++) if (p->d_tw[u] == p->d_tw[t]) shared = true;
      if (!shared) cudaFree(p->d_tw[t]);
    }
    for (int i = 0; i < 16; i++) cudaEventDestroy(p->ev[i]);
    cudaStreamDestroy(p->stream);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    if (p->node_stream) cudaStreamDestroy(p->node_stream);
    for (int i = 0; i < 3; i++) if (p->ev_copy[i]) cudaEventDestroy(p->ev_copy[i]);
    MPI_Comm_free(&p->comm);
    delete p;
  }

  // -------------------------------------------------------------------------------------------
  // F : d_g1 -> padded grid interior   /   F^H : padded grid interior -> d_g1
  // -------------------------------------------------------------------------------------------
  static void fft_forward(P *p) {
    const Layout &L = p->L;
    const PipeGeom &G = p->pipe;
    cudaStream_t st = p->stream;
    C *W0 = (C *)p->d_work[0], *W1 = (C *)p->d_work[1], *PK = (C *)p->d_work[2];
    const bool peer = p->peer.on;
    auto fft_x = [&]() {
      if (p->own_fft[0]) { fft_strided_launch<C>(W1, (int)L.n[0], G.S1, G.S1, -1, (const C *)p->d_tw[0], st); p->launches++; }
      else if (p->fft_x) { FftType<R>::exec_c2c(p->fft_x, W1, CUFFT_FORWARD); p->lib_launches++; }
    };
    auto fft_y = [&]() {
      if (p->own_fft[1]) { fft_strided_launch<C>(W0, (int)L.n[1], G.S3, G.S3, -1, (const C *)p->d_tw[1], st); p->launches++; }
      else if (p->fft_y) { FftType<R>::exec_c2c(p->fft_y, W0, CUFFT_FORWARD); p->lib_launches++; }
    };
    if (peer) {
      // Peer-memory re-distributions with FOUR flag barriers per transform instead of six: a stage's "every push has
      // landed" barrier also tells that every rank is done with the stage's source buffers, which are the next stage's
      // destinations, provided each rank zeroes what the next stage leaves untouched BEFORE it arrives.
      //   W1 (L1) and W0 (L3) are idle when the transform starts: zeroed before the opening barrier;
      //   W1 (L4) is the source of my stage-1 pushes: zeroed right after them, before the stage-1 barrier.
      peer_zero_forward(p, 0, W1);
      peer_zero_forward(p, 1, W0);
      peer_barrier(p);                       // every rank is inside this transform, its L1 / L3 buffers zeroed and idle
      peer_jobs<C>(p, LIST_FWD0 + 0);
      peer_barrier(p);                       // L1 complete everywhere
      fft_x();
      peer_jobs<C>(p, LIST_FWD0 + 1);
      peer_zero_forward(p, 2, W1);
      peer_barrier(p);                       // L3 complete everywhere; every W1 zeroed and idle
      fft_y();
      peer_jobs<C>(p, LIST_FWD0 + 2);
      peer_barrier(p);                       // L4 complete everywhere
    } else {
      run_stage_forward<C>(G.st[0], p->mesh, p->d_g1, W0, W1, PK, st, &p->launches);       // L1 in W1
      fft_x();
      run_stage_forward<C>(G.st[1], p->mesh, W1, W1, W0, PK, st, &p->launches);            // L3 in W0
      fft_y();
      run_stage_forward<C>(G.st[2], p->mesh, W0, W0, W1, PK, st, &p->launches);            // L4 in W1
    }
    const long long lno0 = L.local_no[0], lno1 = L.local_no[1];
    if (p->own_fft[2]) {
      // z pass fused with the crop / embed into the padded grid
      fft_z_grid_launch<C>(W1, (C *)p->d_grid, (int)L.n[2], grid_map(p), -1, (const C *)p->d_tw[2], st);
      p->launches++;
    } else if (!L.c2r) {
      if (p->fft_z_fwd) { FftType<R>::exec_c2c(p->fft_z_fwd, W1, CUFFT_FORWARD); p->lib_launches++; }
      // crop l2 in [o_off2, o_off2+no2) and embed into the padded grid
      BoxMap bm = dense_map(lno0, lno1, L.no[2], L.ngc[1], L.pitch2, L.gcb[0], L.gcb[1], L.gcb[2], L.o_off[2]);
      bm.c_str[0] = lno1 * L.n[2]; bm.c_str[1] = L.n[2];
      box_copy<C>(st, (C *)p->d_grid, W1, bm, BOX_C2A, false, &p->launches);
    } else {
      if (p->fft_z_fwd) { FftType<R>::exec_c2r(p->fft_z_fwd, W1, (R *)W0); p->lib_launches++; }
      BoxMap bm = dense_map(lno0, lno1, L.no[2], L.ngc[1], L.pitch2, L.gcb[0], L.gcb[1], L.gcb[2], L.o_off[2]);
      bm.c_str[0] = lno1 * L.n[2]; bm.c_str[1] = L.n[2];
      box_copy<R>(st, (R *)p->d_grid, (R *)W0, bm, BOX_C2A, false, &p->launches);
    }
  }

  static FftGridMap grid_map(const P *p) {
    const Layout &L = p->L;
    FftGridMap gm;
    gm.rows = (long long)L.local_no[0] * L.local_no[1]; gm.rows1 = L.local_no[1];
    gm.ngc1 = L.ngc[1]; gm.pitch2 = L.pitch2;
    gm.gcb0 = (int)L.gcb[0]; gm.gcb1 = (int)L.gcb[1]; gm.gcb2 = (int)L.gcb[2]; gm.o_off = (int)L.o_off[2]; gm.no = (int)L.no[2];
    return gm;
  }

  static void fft_backward(P *p) {
    const Layout &L = p->L;
    const PipeGeom &G = p->pipe;
    cudaStream_t st = p->stream;
    C *W0 = (C *)p->d_work[0], *W1 = (C *)p->d_work[1], *PK = (C *)p->d_work[2];
    const long long lno0 = L.local_no[0], lno1 = L.local_no[1];
    const bool pruned2 = L.no[2] < L.n[2];
    if (p->own_fft[2]) {
      fft_z_grid_launch<C>(W1, (C *)p->d_grid, (int)L.n[2], grid_map(p), 1, (const C *)p->d_tw[2], st);
      p->launches++;
    } else if (!L.c2r) {
      if (pruned2) PNB_CUDA(cudaMemsetAsync(W1, 0, sizeof(C) * (size_t)G.L4_elems, st));
      BoxMap bm = dense_map(lno0, lno1, L.no[2], L.ngc[1], L.pitch2, L.gcb[0], L.gcb[1], L.gcb[2], L.o_off[2]);
      bm.c_str[0] = lno1 * L.n[2]; bm.c_str[1] = L.n[2];
      box_copy<C>(st, (C *)p->d_grid, W1, bm, BOX_A2C, false, &p->launches);
      if (p->fft_z_bwd) { FftType<R>::exec_c2c(p->fft_z_bwd, W1, CUFFT_INVERSE); p->lib_launches++; }
    } else {
      if (pruned2) PNB_CUDA(cudaMemsetAsync(W0, 0, sizeof(R) * (size_t)(lno0 * lno1 * L.n[2]), st));
      BoxMap bm = dense_map(lno0, lno1, L.no[2], L.ngc[1], L.pitch2, L.gcb[0], L.gcb[1], L.gcb[2], L.o_off[2]);
      bm.c_str[0] = lno1 * L.n[2]; bm.c_str[1] = L.n[2];
      box_copy<R>(st, (R *)p->d_grid, (R *)W0, bm, BOX_A2C, false, &p->launches);
      if (p->fft_z_bwd) { FftType<R>::exec_r2c(p->fft_z_bwd, (R *)W0, W1); p->lib_launches++; }
    }
    const bool peer = p->peer.on;
    auto ifft_y = [&]() {
      if (p->own_fft[1]) { fft_strided_launch<C>(W0, (int)L.n[1], G.S3, G.S3, 1, (const C *)p->d_tw[1], st); p->launches++; }
      else if (p->fft_y) { FftType<R>::exec_c2c(p->fft_y, W0, CUFFT_INVERSE); p->lib_launches++; }
    };
    auto ifft_x = [&]() {
      if (p->own_fft[0]) { fft_strided_launch<C>(W1, (int)L.n[0], G.S1, G.S1, 1, (const C *)p->d_tw[0], st); p->launches++; }
      else if (p->fft_x) { FftType<R>::exec_c2c(p->fft_x, W1, CUFFT_INVERSE); p->lib_launches++; }
    };
    if (peer) {
      // four barriers, as in fft_forward: W0 (L3) is idle at the start; W1 (L1) is the source of my stage-2 pushes and is
      // zeroed right behind them; g1 is idle during the whole transform
      if (L.no[1] < L.n[1]) PNB_CUDA(cudaMemsetAsync(W0, 0, sizeof(C) * (size_t)G.L3_elems, st));
      peer_barrier(p);
      peer_jobs<C>(p, LIST_BWD0 - 2);        // L4 in W1 -> L3 in W0
      if (L.no[0] < L.n[0]) PNB_CUDA(cudaMemsetAsync(W1, 0, sizeof(C) * (size_t)G.L1_elems, st));
      peer_barrier(p);
      ifft_y();
      peer_jobs<C>(p, LIST_BWD0 - 1);        // L3 in W0 -> L1 in W1
      peer_barrier(p);
      ifft_x();
      peer_jobs<C>(p, LIST_BWD0 - 0);        // L1 in W1 -> g1
      peer_barrier(p);
    } else {
      // L4 in W1 -> L3 in W0
      if (L.no[1] < L.n[1]) stage_backward_zero(p, G.st[2], W1, W0, W0, G.L3_elems);
      else run_stage_backward<C>(G.st[2], p->mesh, W1, W0, W0, PK, st, &p->launches);
      ifft_y();
      // L3 in W0 -> L1 in W1
      if (L.no[0] < L.n[0]) stage_backward_zero(p, G.st[1], W0, W1, W1, G.L1_elems);
      else run_stage_backward<C>(G.st[1], p->mesh, W0, W1, W1, PK, st, &p->launches);
      ifft_x();
      // L1 in W1 -> g1
      run_stage_backward<C>(G.st[0], p->mesh, W1, W0, p->d_g1, PK, st, &p->launches);
    }
  }

  // what a forward stage leaves untouched in its destination array must be zero before any peer pushes into it
  static void peer_zero_forward(P *p, int s, C *dest) {
    const Stage &S = p->pipe.st[s];
    if (S.zero_all) PNB_CUDA(cudaMemsetAsync(dest, 0, sizeof(C) * (size_t)S.dst_elems, p->stream));
    else if (S.zero_len > 0) PNB_CUDA(cudaMemsetAsync(dest + S.zero_off, 0, sizeof(C) * (size_t)S.zero_len, p->stream));
  }

  // backward stage whose source-side array has rows outside the pruned output range: they must be zero
  static void stage_backward_zero(P *p, const Stage &S, C *arr, C *w, C *out, long long out_elems) {
    cudaStream_t st = p->stream;
    for (const auto &T : S.tr)
      for (const auto &bm0 : T.recv_maps) { BoxMap bm = bm0; bm.c_off += T.recv_off; box_copy<C>(st, arr, w, bm, BOX_A2C, false, &p->launches); }
    exchange_chunks<C>(S, p->mesh, w, arr, false, st);
    PNB_CUDA(cudaMemsetAsync(out, 0, sizeof(C) * (size_t)out_elems, st));
    for (const auto &T : S.tr)
      for (const auto &bm0 : T.send_maps) { BoxMap bm = bm0; bm.c_off += T.send_off; box_copy<C>(st, out, arr, bm, BOX_C2A, T.send_sign, &p->launches); }
  }

  // -------------------------------------------------------------------------------------------
  // ghost cells
  // -------------------------------------------------------------------------------------------
  template <class T> static void halo_axis_local(P *p, int axis, bool reduce, const int *lo, const int *hi) {
    const Layout &L = p->L;
    HaloGeom hg;
    hg.pitch1 = L.pitch2; hg.pitch0 = (long long)L.ngc[1] * L.pitch2;
    for (int t = 0; t < 3; t++) { hg.lo[t] = lo[t]; hg.hi[t] = hi[t]; }
    hg.axis = axis; hg.gcb = (int)L.gcb[axis]; hg.gca = (int)L.gca[axis]; hg.lno = (int)L.local_no[axis];
    long long total = 1;
    for (int t = 0; t < 3; t++) total *= (t == axis) ? (reduce ? 1 : hg.gcb + hg.gca) : (hi[t] - lo[t]);
    if (total <= 0) return;
    const int bs = 256;
    const int nb = (int)std::min<long long>((total + bs - 1) / bs, 148 * 32);
    if (reduce) k_halo_reduce<T><<<nb, bs, 0, p->stream>>>((T *)p->d_grid, hg);
    else k_halo_fill<T><<<nb, bs, 0, p->stream>>>((T *)p->d_grid, hg);
    p->launches++;
  }

  // split axis: m-wide slabs travel to / from the two mesh neighbours over NCCL
  template <class T> static void halo_axis_nccl(P *p, int axis, bool reduce, const int *lo, const int *hi) {
    const Layout &L = p->L;
    const Mesh &M = p->mesh;
    cudaStream_t st = p->stream;
    const int gcb = (int)L.gcb[axis], gca = (int)L.gca[axis], lno = (int)L.local_no[axis];
    int cu[2] = {M.co[0], M.co[1]}, cd[2] = {M.co[0], M.co[1]};
    cu[axis] += 1; cd[axis] -= 1;
    const int up = M.rank_of(cu[0], cu[1]), down = M.rank_of(cd[0], cd[1]);
    // box extents in the other dims
    long long ext[3];
    for (int t = 0; t < 3; t++) ext[t] = hi[t] - lo[t];
    auto slab = [&](int start, int width, long long coff) {
      long long d[3] = {ext[0], ext[1], ext[2]}, s[3] = {lo[0], lo[1], lo[2]};
      d[axis] = width; s[axis] = start;
      return dense_map(d[0], d[1], d[2], L.ngc[1], L.pitch2, s[0], s[1], s[2], coff);
    };
    long long per = 1;
    for (int t = 0; t < 3; t++) if (t != axis) per *= ext[t];
    T *sb = (T *)p->d_work[0], *rb = (T *)p->d_work[1];
    T *grid = (T *)p->d_grid;
    // fill : my top gcb interior rows -> up's below halo ; my bottom gca interior rows -> down's above halo
    // reduce: my above halo (gca) -> up's bottom interior rows ; my below halo (gcb) -> down's top interior rows
    const long long n_up = per * (reduce ? gca : gcb), n_dn = per * (reduce ? gcb : gca);
    if (!reduce) {
      box_copy<T>(st, grid, sb, slab(gcb + lno - gcb, gcb, 0), BOX_A2C, false, &p->launches);
      box_copy<T>(st, grid, sb, slab(gcb, gca, n_up), BOX_A2C, false, &p->launches);
    } else {
      box_copy<T>(st, grid, sb, slab(gcb + lno, gca, 0), BOX_A2C, false, &p->launches);
      box_copy<T>(st, grid, sb, slab(0, gcb, n_up), BOX_A2C, false, &p->launches);
    }
    // what arrives: from down: its "to up" message (same size as my n_up); from up: its "to down" message
    PNB_NCCL(nccl_api().GroupStart());
    PNB_NCCL(nccl_api().Send(sb, (size_t)n_up * sizeof(T), ncclChar, up, world_nccl(), st));
    PNB_NCCL(nccl_api().Send(sb + n_up, (size_t)n_dn * sizeof(T), ncclChar, down, world_nccl(), st));
    PNB_NCCL(nccl_api().Recv(rb, (size_t)n_up * sizeof(T), ncclChar, down, world_nccl(), st));
    PNB_NCCL(nccl_api().Recv(rb + n_up, (size_t)n_dn * sizeof(T), ncclChar, up, world_nccl(), st));
    PNB_NCCL(nccl_api().GroupEnd());
    if (!reduce) {
      box_copy<T>(st, grid, rb, slab(0, gcb, 0), BOX_C2A, false, &p->launches);               // from down: below halo
      box_copy<T>(st, grid, rb, slab(gcb + lno, gca, n_up), BOX_C2A, false, &p->launches);    // from up: above halo
    } else {
      box_copy<T>(st, grid, rb, slab(gcb, gca, 0), BOX_C2A_ADD, false, &p->launches);               // down's above halo -> my bottom rows
      box_copy<T>(st, grid, rb, slab(gcb + lno - gcb, gcb, n_up), BOX_C2A_ADD, false, &p->launches);  // up's below halo -> my top rows
    }
  }

  template <class T> static void halo_T(P *p, bool reduce) {
    const Layout &L = p->L;
    const Mesh &M = p->mesh;
    const bool peer = p->peer.on;
    bool pushed = false, any_peer = false;     // peer stores in flight that the next step (any axis) must see
    // fill order z, y, x (later axes carry the earlier halos => corners); reduce runs backwards
    for (int k = 0; k < 3; k++) {
      const int axis = reduce ? k : 2 - k;
      int lo[3], hi[3];
      for (int t = 0; t < 3; t++) {
        // axes processed "before" this one (in fill order: larger index) span the padded extent
        const bool wide = (t > axis);
        lo[t] = wide ? 0 : (int)L.gcb[t];
        hi[t] = wide ? (int)L.ngc[t] : (int)(L.gcb[t] + L.local_no[t]);
      }
      const bool split = axis < 2 && M.np[axis] > 1;
      if (split && peer) {
        // the neighbours must have finished whatever produced the rows this step reads or overwrites
        peer_barrier(p);
        if (!reduce) { peer_jobs<T>(p, axis == 1 ? LIST_FILL1 : LIST_FILL0); pushed = true; }
        else peer_jobs<T>(p, axis == 0 ? LIST_RED0 : LIST_RED1);
        any_peer = true;
      } else {
        if (pushed) { peer_barrier(p); pushed = false; }
        if (split) halo_axis_nccl<T>(p, axis, reduce, lo, hi);
        else halo_axis_local<T>(p, axis, reduce, lo, hi);
      }
    }
    // fill: the last pushes must have landed before the gather; reduce: nobody may overwrite a halo (next call) that a
    // neighbour is still reading
    if (any_peer) peer_barrier(p);
  }
  static void halo(P *p, bool reduce) {
    if (p->L.c2r) halo_T<R>(p, reduce); else halo_T<C>(p, reduce);
  }

  // -------------------------------------------------------------------------------------------
  // nodes
  // -------------------------------------------------------------------------------------------
  static Nd *init_nodes(INT local_M, unsigned malloc_flags) {
    Nd *nd = new Nd();
    nd->local_M = local_M;
    nd->malloc_flags = malloc_flags;
    const size_t M = (size_t)(local_M > 0 ? local_M : 1);
    // sizes as in reference kernel/ndft-parallel.c:1484-1522 (f: 2 R per node, grad_f: 6 R per node)
    if (malloc_flags & N_MALLOC_X) PNB_CUDA(cudaHostAlloc((void **)&nd->x, sizeof(R) * 3 * M, cudaHostAllocDefault));
    if (malloc_flags & N_MALLOC_F) PNB_CUDA(cudaHostAlloc((void **)&nd->f, sizeof(R) * 2 * M, cudaHostAllocDefault));
    if (malloc_flags & N_MALLOC_GRAD_F) PNB_CUDA(cudaHostAlloc((void **)&nd->grad_f, sizeof(R) * 6 * M, cudaHostAllocDefault));
    if (malloc_flags & N_MALLOC_HESSIAN_F) PNB_CUDA(cudaHostAlloc((void **)&nd->hessian_f, sizeof(R) * 12 * M, cudaHostAllocDefault));
    return nd;
  }

  static void free_nodes(Nd *nd, unsigned flags) {
    if (!nd) return;
    if ((flags & N_MALLOC_X) && nd->x) { if (is_device_ptr(nd->x)) cudaFree(nd->x); else cudaFreeHost(nd->x); }
    if ((flags & N_MALLOC_F) && nd->f) { if (is_device_ptr(nd->f)) cudaFree(nd->f); else cudaFreeHost(nd->f); }
    if ((flags & N_MALLOC_GRAD_F) && nd->grad_f) { if (is_device_ptr(nd->grad_f)) cudaFree(nd->grad_f); else cudaFreeHost(nd->grad_f); }
    if ((flags & N_MALLOC_HESSIAN_F) && nd->hessian_f) cudaFreeHost(nd->hessian_f);
    cudaFree(nd->d_x); cudaFree(nd->d_x_alt); cudaFree(nd->d_f); cudaFree(nd->d_grad_f); cudaFree(nd->d_hess); cudaFree(nd->d_wtab); cudaFree(nd->d_vals);
    for (int k = 0; k < 2; k++) {
      BinState<R> &b = k ? nd->il : *static_cast<BinState<R> *>(nd);
      cudaFree(b.d_tile); cudaFree(b.d_tile_sorted); cudaFree(b.d_perm); cudaFree(b.d_idx);
      cudaFree(b.d_tile_count); cudaFree(b.d_tile_start); cudaFree(b.d_pre_psi); cudaFree(b.d_pre_dpsi); cudaFree(b.d_rows);
      if (b.h_maxcol) { cudaFreeHost(b.h_maxcol); cudaFree(b.d_maxcol); }
    }
    if (nd->h_hash) { cudaFreeHost(nd->h_hash); cudaFree(nd->d_hash); }
    delete nd;
  }

  static void ensure(R **buf, size_t *cap, size_t need) {
    if (*cap >= need && *buf) return;
    if (*buf) cudaFree(*buf);
    PNB_CUDA(cudaMalloc((void **)buf, sizeof(R) * (need ? need : 1)));
    *cap = need;
  }

  // device view of a user array: the pointer itself if it is device memory, else the (uploaded) mirror
  static R *dev_in(P *p, const R *user, R **mirror, size_t *cap, size_t count, bool upload) {
    if (!user) return nullptr;
    if (is_device_ptr(user)) return const_cast<R *>(user);
    ensure(mirror, cap, count);
    if (upload && count) PNB_CUDA(cudaMemcpyAsync(*mirror, user, sizeof(R) * count, cudaMemcpyHostToDevice, p->stream));
    return *mirror;
  }

  // kernel families: 2 = warp-autonomous z-marching register kernels (default for m = 4, 6), 0 = z-marching v1
  // (CTA-synchronous; m = 8, and variant 8 forces it), 1 = generic global-memory kernels (any m; variant 1 forces them)
  static int kernel_family(const P *p) {
    const int m = p->L.m;
    const bool fits = (m == 4 || m == 6 || m == 8);
    const int kv = p->kernel_variant & 9;
    // family 3 (double, m = 5, 7, 8): gather on the FP64 tensor cores with the window in shared memory (zmarch4.cuh); the
    // scatter is v1's at m = 8 (same column tile, bins shared through ZmGeom::sub) and the generic kernel at m = 5, 7.
    // PNFFT_B200_NO_MMA4=1 or any kernel variant bit switches it off.
    static const bool off4 = getenv("PNFFT_B200_NO_MMA4") && atoi(getenv("PNFFT_B200_NO_MMA4")) != 0;
    if (sizeof(R) == 8 && (m == 5 || m == 7 || m == 8) && !off4 && !(p->kernel_variant & 13)) return 3;
    if (kv == 1 || !fits) return 1;
    if (kv == 8 || m == 8) return 0;
    return 2;
  }
  // family 2 in double runs its node loops on the FP64 tensor cores (zmarch3.cuh) unless variant bit 4 or
  // PNFFT_B200_NO_MMA=1 asks for the DFMA loops of zmarch2.cuh (kept for single precision and for comparison)
  template <int M_> static bool use_mma(const P *p) {
    static const bool off = getenv("PNFFT_B200_NO_MMA") && atoi(getenv("PNFFT_B200_NO_MMA")) != 0;
    return sizeof(R) == 8 && Zm3Ok<M_>::value && !off && !(p->kernel_variant & 4);
  }
  static TileGeom tile_geom(const P *p, bool *tiled_ok) {
    TileGeom tg;
    const int m = p->L.m;
    const int fam = kernel_family(p);
    tg.sub = 1;
    if (fam == 2) { tg.T[0] = Zm2Cfg<6>::T0; tg.T[1] = 16 - 2 * m; tg.T[2] = Zm2Cfg<6>::ZS; tg.sub = Zm2Cfg<6>::SUB; }   // == Zm2Cfg<m>::T0, T1, ZS
    else if (fam == 3) { tg.T[0] = Zm4Cfg<8>::T0; tg.T[1] = Zm4Cfg<8>::T1; tg.T[2] = Zm4Cfg<8>::ZS; tg.sub = Zm4Cfg<8>::SUB; }
    else if (fam == 0) { tg.T[0] = (m <= 6) ? 16 : 8; tg.T[1] = 4; tg.T[2] = (m <= 6) ? 8 : 4; }   // == ZmCfg<m>::T0, T1, ZS
    else { tg.T[0] = 8; tg.T[1] = 8; tg.T[2] = 16; }    // generic kernels: the bins only order the nodes for locality
    // a shifted node of an interlaced plan may sit one cell past the block (the extra ghost cell above)
    const int extra = (p->pnfft_flags & F_INTERLACED) ? 1 : 0;
    for (int t = 0; t < 3; t++) tg.nt[t] = (int)((p->L.local_no[t] + extra + tg.T[t] - 1) / tg.T[t]);
    for (int t = 0; t < 3; t++) if (tg.nt[t] < 1) tg.nt[t] = 1;
    tg.ntiles = tg.nt[0] * tg.nt[1] * tg.nt[2];
    if (tiled_ok) *tiled_ok = fam != 1;
    tg.family = fam;
    return tg;
  }

  static void bin_nodes(P *p, Nd *nd, const R *d_x) {
    const int M = (int)nd->local_M;
    const TileGeom tg = tile_geom(p, nullptr);
    cudaStream_t st = p->stream;
    if (nd->cap_nodes < (size_t)M || !nd->d_tile) {
      cudaFree(nd->d_tile); cudaFree(nd->d_tile_sorted); cudaFree(nd->d_perm); cudaFree(nd->d_idx);
      const size_t c = (size_t)(M > 0 ? M : 1);
      PNB_CUDA(cudaMalloc((void **)&nd->d_tile, sizeof(int) * c));
      PNB_CUDA(cudaMalloc((void **)&nd->d_tile_sorted, sizeof(int) * c));
      PNB_CUDA(cudaMalloc((void **)&nd->d_perm, sizeof(int) * c));
      PNB_CUDA(cudaMalloc((void **)&nd->d_idx, sizeof(int) * c));
      nd->cap_nodes = c;
    }
    const size_t nt1 = (size_t)tg.ntiles * tg.sub + 2;
    if (nd->cap_tiles < nt1) {
      cudaFree(nd->d_tile_count); cudaFree(nd->d_tile_start);
      PNB_CUDA(cudaMalloc((void **)&nd->d_tile_count, sizeof(int) * nt1));
      PNB_CUDA(cudaMalloc((void **)&nd->d_tile_start, sizeof(int) * nt1));
      nd->cap_tiles = nt1;
    }
    PNB_CUDA(cudaMemsetAsync(nd->d_tile_count, 0, sizeof(int) * nt1, st));
    if (M == 0) return;
    const GridGeom<R> g = geom(p);
    k_bin_nodes<R><<<(M + 255) / 256, 256, 0, st>>>(g, tg, d_x, M, nd->d_tile, nd->d_idx, nd->d_tile_count);
    int bits = 1;
    while ((1LL << bits) <= (long long)tg.ntiles * tg.sub) bits++;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, nd->d_tile, nd->d_tile_sorted, nd->d_idx, nd->d_perm, M, 0, bits, st);
    size_t tmp2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, nd->d_tile_count, nd->d_tile_start, (int)nt1, st);
    tmp = std::max(tmp, tmp2);
    if (p->sort_tmp_bytes < tmp) {
      cudaFree(p->d_sort_tmp);
      PNB_CUDA(cudaMalloc(&p->d_sort_tmp, tmp));
      p->sort_tmp_bytes = tmp;
    }
    cub::DeviceRadixSort::SortPairs(p->d_sort_tmp, tmp, nd->d_tile, nd->d_tile_sorted, nd->d_idx, nd->d_perm, M, 0, bits, st);
    cub::DeviceScan::ExclusiveSum(p->d_sort_tmp, tmp, nd->d_tile_count, nd->d_tile_start, (int)nt1, st);
    p->launches += 1;       // k_bin_nodes
    p->lib_launches += 2;   // cub radix sort + scan
    if (kernel_family(p) == 2 || kernel_family(p) == 3) {
      // load-balance hint for the NEXT gridding launches: lands in pinned host memory by an async copy that nobody waits
      // for (the host reads whatever the last finished binning left there: the old value or the new one, never a zero)
      if (!nd->h_maxcol) {
        PNB_CUDA(cudaHostAlloc((void **)&nd->h_maxcol, sizeof(int), cudaHostAllocDefault)); *nd->h_maxcol = 0;
        PNB_CUDA(cudaMalloc((void **)&nd->d_maxcol, sizeof(int)));
      }
      PNB_CUDA(cudaMemsetAsync(nd->d_maxcol, 0, sizeof(int), st));
      const int ncol = tg.nt[0] * tg.nt[1];
      k_max_column<<<(ncol + 255) / 256, 256, 0, st>>>(nd->d_tile_start, ncol, tg.nt[2] * tg.sub, nd->d_maxcol);
      PNB_CUDA(cudaMemcpyAsync(nd->h_maxcol, nd->d_maxcol, sizeof(int), cudaMemcpyDeviceToHost, st));
      p->launches++;
    }
  }

  template <bool CPLX, int M_, bool GRAD>
  static void launch_zm(P *p, Nd *nd, const NodeArgs<R> &na, bool scatter) {
    typedef ZmCfg<M_> Cfg;
    typedef ZmTab<R, M_, GRAD> Tab;
    typedef ZmSmem<R, CPLX, M_, GRAD> Sm;
    const TileGeom tg = tile_geom(p, nullptr);
    const GridGeom<R> g = geom(p);
    const CUtensorMap tm = make_grid_tmap<R>(p->d_grid, p->L, CPLX ? 2 : 1, Cfg::R0, Cfg::R1, Cfg::ZB);
    ZmGeom zg;
    zg.nc[0] = tg.nt[0]; zg.nc[1] = tg.nt[1]; zg.nt2 = tg.nt[2];
    zg.nseg = (tg.nt[2] + Cfg::ZSEG - 1) / Cfg::ZSEG;
    zg.sub = tg.sub;
    const unsigned nblk = (unsigned)(tg.nt[0] * tg.nt[1] * zg.nseg);
    // 1. node table: window factors evaluated once per node and axis (+ f / grad_f for the adjoint), sorted order
    const size_t need = (size_t)na.M * Tab::ROWLEN + 64;
    ensure(&nd->d_wtab, &nd->cap_wtab, need);
    {
      auto kt = k_node_table<R, M_, GRAD>;
      const long long nthr = 3LL * na.M;
      const size_t psm = g.poly ? sizeof(R) * (size_t)2 * (g.poly_deg + 1) * 3 * Cfg::C : 0;
      kt<<<(unsigned)((nthr + 191) / 192), 192, psm, p->stream>>>(g, na, CPLX ? 2 : 1, scatter ? 1 : 0, nd->d_wtab);
      p->launches++;
    }
    // 2. gridding
    if (!scatter) {
      GatherOut<R> out;
      out.perm = na.perm; out.f = na.f; out.f_stride = na.f_stride; out.f_off = na.f_off; out.grad = na.grad; out.accumulate = na.accumulate;
      auto kern = k_gather_zm<R, CPLX, M_, GRAD>;
      PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sm::gather));
      kern<<<nblk, Cfg::NT, Sm::gather, p->stream>>>(tm, zg, nd->d_wtab, nd->d_tile_start, out);
    } else {
      auto kern = k_scatter_zm<R, CPLX, M_, GRAD>;
      PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sm::scatter));
      kern<<<nblk, Cfg::NT, Sm::scatter, p->stream>>>(tm, zg, nd->d_wtab, nd->d_tile_start);
    }
    PNB_CUDA(cudaGetLastError());
    p->launches++;
  }

  // z-march v2 (zmarch2.cuh): coalesced node table, then the warp-autonomous gather / scatter
  template <bool CPLX, int M_, bool GRAD>
  static void launch_zm2(P *p, Nd *nd, const NodeArgs<R> &na, bool scatter) {
    if constexpr (M_ == 4 || M_ == 6) {
      typedef Zm2Cfg<M_> Cfg;
      typedef Zm2Smem<R, CPLX, M_, GRAD> Sm;
      const TileGeom tg = tile_geom(p, nullptr);
      const GridGeom<R> g = geom(p);
      typedef typename CellT<R, CPLX>::type Cell;
      const CUtensorMap tm = make_grid_tmap<R>(p->d_grid, p->L, CPLX ? 2 : 1, Cfg::XW, 16, Cfg::ZB, zm2_swizzle_mode<Cell, Cfg::ZB>());
      Zm2Geom zg = zm2_work_items(p, tg, nd, na.M);
      const int ncol = tg.nt[0] * tg.nt[1];
      const size_t psm = g.poly ? sizeof(R) * (size_t)2 * (g.poly_deg + 1) * 3 * Cfg::C : 0;
      typedef typename Sm::RowG RowG;
      typedef typename Sm::RowS RowS;
      const size_t rowlen = scatter ? RowS::ROWLEN : RowG::ROWLEN;
      // The node table (rowlen * M reals) is built and consumed in column batches when it would not fit the budget
      // (PNFFT_B200_TABLE_GB, default 24 GB): 2^27 nodes on one GPU need 58 - 116 GB of rows otherwise.
      static const double cap_gb = getenv("PNFFT_B200_TABLE_GB") ? atof(getenv("PNFFT_B200_TABLE_GB")) : 24.0;
      const double need_gb = (double)na.M * rowlen * sizeof(R) / 1073741824.0;
      int nbatch = 1;
      if (need_gb > cap_gb && !(p->b_phase != 3)) nbatch = std::min(ncol, (int)std::ceil(2.0 * need_gb / cap_gb));
      std::vector<int> cb((size_t)nbatch + 1), nb((size_t)nbatch + 1);
      for (int k = 0; k <= nbatch; k++) cb[(size_t)k] = (int)((long long)ncol * k / nbatch);
      nb[0] = 0; nb[(size_t)nbatch] = na.M;
      if (nbatch > 1) {          // first sorted node of every batch: the bins of a column are contiguous in tile_start
        const size_t per_col = (size_t)tg.nt[2] * Cfg::SUB;
        for (int k = 1; k < nbatch; k++)
          PNB_CUDA(cudaMemcpyAsync(&nb[(size_t)k], nd->d_tile_start + (size_t)cb[(size_t)k] * per_col, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
        PNB_CUDA(cudaStreamSynchronize(p->stream));
      }
      size_t max_rows = 0;
      for (int k = 0; k < nbatch; k++) max_rows = std::max(max_rows, (size_t)(nb[(size_t)k + 1] - nb[(size_t)k]));
      ensure(&nd->d_wtab, &nd->cap_wtab, max_rows * rowlen + 64);
      for (int k = 0; k < nbatch; k++) {
        const int first = nb[(size_t)k], last = nb[(size_t)k + 1];
        if (last <= first) continue;
        NodeArgs<R> nb_args = na;
        nb_args.M = last;
        R *tab = nd->d_wtab - (size_t)first * rowlen;        // rows are addressed by their absolute sorted position
        zg.col0 = cb[(size_t)k];
        const unsigned nblk = (unsigned)((cb[(size_t)k + 1] - cb[(size_t)k]) * zg.nseg);
        const unsigned ntb = (unsigned)((last - first + kZm2TabNodes - 1) / kZm2TabNodes);
        if (!scatter) {
          if (p->b_phase & 1) {
            auto kt = k_node_table2<R, M_, GRAD, false, CPLX>;
            const size_t tsm = (size_t)kZm2TabNodes * RowG::ROWBYTES + psm;
            PNB_CUDA(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
            kt<<<ntb, 3 * kZm2TabNodes, tsm, p->stream>>>(g, nb_args, tab, first);
            p->launches++;
          }
          if (p->b_phase & 2) {
            GatherOut<R> out;
            out.perm = na.perm; out.f = na.f; out.f_stride = na.f_stride; out.f_off = na.f_off; out.grad = na.grad; out.accumulate = na.accumulate;
            auto kern = k_gather_zm2<R, CPLX, M_, GRAD>;
            PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sm::gather));
            kern<<<nblk, (Cfg::NCW + 1) * 32, Sm::gather, p->stream>>>(tm, zg, tab, nd->d_tile_start, out);
            p->launches++;
          }
        } else {
          auto kt = k_node_table2<R, M_, GRAD, true, CPLX>;
          const size_t tsm = (size_t)kZm2TabNodes * RowS::ROWBYTES + psm;
          PNB_CUDA(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
          kt<<<ntb, 3 * kZm2TabNodes, tsm, p->stream>>>(g, nb_args, tab, first);
          auto kern = k_scatter_zm2<R, CPLX, M_, GRAD>;
          PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sm::scatter));
          kern<<<nblk, (Cfg::NCW + 1) * 32, Sm::scatter, p->stream>>>(tm, zg, tab, nd->d_tile_start);
          p->launches += 2;
        }
      }
      PNB_CUDA(cudaGetLastError());
    }
  }

  // work items of the family-2 geometry: whole columns when there are enough of them to fill the GPU, else split along z;
  // clustered node sets (hint from the previous binning) get smaller items
  static Zm2Geom zm2_work_items(const P *p, const TileGeom &tg, const Nd *nd, int M) {
    Zm2Geom zg;
    zg.nc[0] = tg.nt[0]; zg.nc[1] = tg.nt[1]; zg.nt2 = tg.nt[2]; zg.col0 = 0;
    const int ncol = tg.nt[0] * tg.nt[1];
    // a column is cut into items of about `target` nodes (zm_segment): a quarter of an SM's share of the node set
    zg.target = std::max(256, M / (148 * 4));
    const int target_nodes = zg.target;
    int nseg = 1;
    while (ncol * nseg < 8 * 148 && nseg * 2 <= tg.nt[2]) nseg *= 2;
    zg.fill = nseg;                          // whole columns per item when there are enough of them to fill the GPU
    static const int nseg_env = getenv("PNFFT_B200_NSEG") ? atoi(getenv("PNFFT_B200_NSEG")) : 0;
    int nseg_min = nseg_env;
    if (!nseg_env && nd->h_maxcol) {
      // clustered node sets put most nodes into few columns (hint: the largest column of the previous binning)
      const long long maxcol = *(volatile int *)nd->h_maxcol;
      nseg_min = 1;
      while (nseg_min < 32 && (long long)nseg_min * target_nodes < maxcol) nseg_min *= 2;
    }
    while (nseg < nseg_min && nseg * 2 <= tg.nt[2]) nseg *= 2;
    if (nseg_env) zg.fill = nseg;            // a forced count splits every column
    // Pieces of equal node count only where the heaviest item bounds the launch: a rank of a multi-GPU plan that holds a
    // part of a cluster in few columns (C4 on 2x4 GPUs: 28.4 -> 22.4 ms per step).  On one GPU (C4: 16384 columns) pieces of
    // equal LENGTH measured 9 % faster (light columns in several concurrent items make up for the few batches per window
    // position), and uniform node sets have nothing to balance.  PNFFT_B200_SEG_BALANCE=0 / 1 forces the choice.
    static const int bal_env = getenv("PNFFT_B200_SEG_BALANCE") ? atoi(getenv("PNFFT_B200_SEG_BALANCE")) : -1;
    const bool balance = bal_env >= 0 ? bal_env != 0 : (p->mesh.size > 1 && nseg_min > 1 && !nseg_env);
    if (!balance) zg.target = 0;
    zg.zseg = (tg.nt[2] + nseg - 1) / nseg;
    zg.nseg = nseg;
    return zg;
  }

  // z-march v3 (zmarch3.cuh, double only): node-table rows WITHOUT node values, kept across calls while the binning they
  // were made from stays valid (device-resident x with an unchanged content hash, x declared static, PNFFT_PRE_PSI); a
  // table with derivative sections serves F-only calls too.  PNFFT_B200_ROW_CACHE=0 rebuilds the rows in every call.
  template <bool CPLX, int M_, bool GRAD, bool RG>
  static void launch_zm3_kernels(P *p, Nd *nd, const NodeArgs<R> &na, bool scatter, const Zm2Geom &zg, unsigned nblk, const R *tab) {
    if constexpr (sizeof(R) == 8 && Zm3Ok<M_>::value) {
      typedef Zm2Cfg<M_> Cfg;
      typedef Zm3Smem<CPLX, M_, GRAD, RG> Sm;
      const CUtensorMap tm = make_grid_tmap<R>(p->d_grid, p->L, CPLX ? 2 : 1, Cfg::XW, 16, Cfg::ZB, 0);   // boxes are read 256 contiguous bytes per n-block: no swizzle
      // PNFFT_B200_GATHER4=1: the shared-window gather of zmarch4.cuh on v3's bins and rows instead of k_gather_mma
      static const bool g4 = env_flag("PNFFT_B200_GATHER4", false);
      if (!scatter && g4) {
        typedef Zm4on2Cfg<M_> C4;
        typedef Zm4Smem<CPLX, M_, C4> Sm4;
        const CUtensorMap tm4 = make_grid_tmap<R>(p->d_grid, p->L, CPLX ? 2 : 1, C4::R0, Sm4::R1, C4::ZS, 0);
        GatherOut<R> out;
        out.perm = na.perm; out.f = na.f; out.f_stride = na.f_stride; out.f_off = na.f_off; out.grad = na.grad; out.accumulate = na.accumulate;
        auto kern = k_gather_mma4<CPLX, M_, GRAD, RG, C4>;
        PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sm4::gather));
        kern<<<nblk, (C4::NW + 1) * 32, Sm4::gather, p->stream>>>(tm4, zg, tab, nd->d_tile_start, out);
      } else if (!scatter) {
        GatherOut<R> out;
        out.perm = na.perm; out.f = na.f; out.f_stride = na.f_stride; out.f_off = na.f_off; out.grad = na.grad; out.accumulate = na.accumulate;
        auto kern = k_gather_mma<CPLX, M_, GRAD, RG>;
        PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sm::gather));
        kern<<<nblk, (Cfg::NCW + 1) * 32, Sm::gather, p->stream>>>(tm, zg, tab, nd->d_tile_start, out);
      } else {
        auto kern = k_scatter_mma<CPLX, M_, GRAD, RG>;
        PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sm::scatter));
        kern<<<nblk, (Cfg::NCW + 1) * 32, Sm::scatter, p->stream>>>(tm, zg, tab, nd->d_vals, nd->d_tile_start);
      }
      p->launches++;
    }
  }
  template <bool CPLX, int M_, bool RG>
  static void launch_zm3_table(P *p, const GridGeom<R> &g, const NodeArgs<R> &na_upto, R *tab, int first) {
    if constexpr (sizeof(R) == 8 && Zm3Ok<M_>::value) {
      typedef Zm2Row<R, M_, RG, false, CPLX> Row;
      const size_t psm = g.poly ? sizeof(R) * (size_t)2 * (g.poly_deg + 1) * 3 * Zm2Cfg<M_>::C : 0;
      const unsigned ntb = (unsigned)((na_upto.M - first + kZm2TabNodes - 1) / kZm2TabNodes);
      auto kt = k_node_table2<R, M_, RG, false, CPLX>;
      const size_t tsm = (size_t)kZm2TabNodes * Row::ROWBYTES + psm;
      PNB_CUDA(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
      kt<<<ntb, 3 * kZm2TabNodes, tsm, p->stream>>>(g, na_upto, tab, first);
      p->launches++;
    }
  }
  template <bool CPLX, int M_, bool GRAD>
  static void launch_zm3(P *p, Nd *nd, const NodeArgs<R> &na, bool scatter) {
    if constexpr (sizeof(R) == 8 && Zm3Ok<M_>::value) {
      typedef Zm2Cfg<M_> Cfg;
      const TileGeom tg = tile_geom(p, nullptr);
      const GridGeom<R> g = geom(p);
      Zm2Geom zg = zm2_work_items(p, tg, nd, na.M);
      const int ncol = tg.nt[0] * tg.nt[1];
      if (scatter && (p->b_phase & 2)) {       // node values in sorted order
        constexpr int NVP = Zm3Smem<CPLX, M_, GRAD, GRAD>::NVP;
        ensure(&nd->d_vals, &nd->cap_vals, (size_t)na.M * NVP + 64);
        const long long n = (long long)na.M * NVP;
        k_pack_vals<CPLX, GRAD><<<(unsigned)((n + 255) / 256), 256, 0, p->stream>>>(na, nd->d_vals);
        p->launches++;
      }
      const size_t len_g = Zm2Row<R, M_, true, false, CPLX>::ROWLEN, len_f = Zm2Row<R, M_, false, false, CPLX>::ROWLEN;
      static const double cap_gb = getenv("PNFFT_B200_TABLE_GB") ? atof(getenv("PNFFT_B200_TABLE_GB")) : 24.0;
      const char *rc = getenv("PNFFT_B200_ROW_CACHE");
      const bool cache_on = !(rc && atoi(rc) == 0);
      const double need_gb = (double)na.M * (GRAD ? len_g : len_f) * sizeof(R) / 1073741824.0;
      if (need_gb <= cap_gb || p->b_phase != 3) {
        // one table for the whole node set, in the binning's own buffer
        const bool reuse = cache_on && nd->binned && nd->rows_plan == (const void *)p && nd->rows_gen == p->win_gen && nd->rows_flavor >= (GRAD ? 1 : 0) && nd->d_rows;
        int flavor = reuse ? nd->rows_flavor : (GRAD ? 1 : 0);
        if (!reuse && (p->b_phase & 1)) {
          ensure(&nd->d_rows, &nd->cap_rows, (size_t)na.M * (flavor ? len_g : len_f) + 64);
          if (flavor) launch_zm3_table<CPLX, M_, true>(p, g, na, nd->d_rows, 0);
          else launch_zm3_table<CPLX, M_, false>(p, g, na, nd->d_rows, 0);
          nd->rows_flavor = flavor; nd->rows_plan = p; nd->rows_gen = p->win_gen;
        } else if (!reuse) {
          flavor = nd->rows_flavor;      // the table phase of this call ran earlier (side stream)
        }
        if (p->b_phase & 2) {
          const unsigned nblk = (unsigned)(ncol * zg.nseg);
          if (flavor) launch_zm3_kernels<CPLX, M_, GRAD, true>(p, nd, na, scatter, zg, nblk, nd->d_rows);
          else if constexpr (!GRAD) launch_zm3_kernels<CPLX, M_, false, false>(p, nd, na, scatter, zg, nblk, nd->d_rows);
        }
      } else {
        // the table would not fit the budget: built and consumed in column batches, nothing kept
        const size_t rowlen = GRAD ? len_g : len_f;
        const int nbatch = std::min(ncol, (int)std::ceil(2.0 * need_gb / cap_gb));
        std::vector<int> cb((size_t)nbatch + 1), nb((size_t)nbatch + 1);
        for (int k = 0; k <= nbatch; k++) cb[(size_t)k] = (int)((long long)ncol * k / nbatch);
        nb[0] = 0; nb[(size_t)nbatch] = na.M;
        const size_t per_col = (size_t)tg.nt[2] * Cfg::SUB;
        for (int k = 1; k < nbatch; k++)
          PNB_CUDA(cudaMemcpyAsync(&nb[(size_t)k], nd->d_tile_start + (size_t)cb[(size_t)k] * per_col, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
        PNB_CUDA(cudaStreamSynchronize(p->stream));
        size_t max_rows = 0;
        for (int k = 0; k < nbatch; k++) max_rows = std::max(max_rows, (size_t)(nb[(size_t)k + 1] - nb[(size_t)k]));
        ensure(&nd->d_wtab, &nd->cap_wtab, max_rows * rowlen + 64);
        nd->rows_flavor = -1;
        for (int k = 0; k < nbatch; k++) {
          const int first = nb[(size_t)k], last = nb[(size_t)k + 1];
          if (last <= first) continue;
          NodeArgs<R> nb_args = na;
          nb_args.M = last;
          R *tab = nd->d_wtab - (size_t)first * rowlen;        // rows are addressed by their absolute sorted position
          zg.col0 = cb[(size_t)k];
          const unsigned nblk = (unsigned)((cb[(size_t)k + 1] - cb[(size_t)k]) * zg.nseg);
          launch_zm3_table<CPLX, M_, GRAD>(p, g, nb_args, tab, first);
          launch_zm3_kernels<CPLX, M_, GRAD, GRAD>(p, nd, na, scatter, zg, nblk, tab);
        }
      }
      PNB_CUDA(cudaGetLastError());
    }
  }

  // z-march v4 (zmarch4.cuh, double only): gather and scatter of family 3.  Rows without node values in nd->d_rows, cached
  // like v3's; the scatter streams f / grad_f from the sorted value array of k_pack_vals.
  template <bool CPLX, int M_, bool GRAD>
  static int zm4_rows(P *p, Nd *nd, const NodeArgs<R> &na) {      // returns the flavor of the rows at hand
    int flavor = -1;
    if constexpr (sizeof(R) == 8 && Zm4Ok<M_>::value) {
      typedef Zm4Cfg<M_> Cfg;
      const GridGeom<R> g = geom(p);
      const size_t len_g = ZmRowOf<R, Cfg, true, false, CPLX>::ROWLEN, len_f = ZmRowOf<R, Cfg, false, false, CPLX>::ROWLEN;
      const char *rc = getenv("PNFFT_B200_ROW_CACHE");
      const bool cache_on = !(rc && atoi(rc) == 0);
      const bool reuse = cache_on && nd->binned && nd->rows_plan == (const void *)p && nd->rows_gen == p->win_gen && nd->rows_flavor >= (GRAD ? 1 : 0) && nd->d_rows;
      flavor = reuse ? nd->rows_flavor : (GRAD ? 1 : 0);
      if (!reuse && (p->b_phase & 1)) {
        ensure(&nd->d_rows, &nd->cap_rows, (size_t)na.M * (flavor ? len_g : len_f) + 64);
        const size_t psm = g.poly ? sizeof(R) * (size_t)2 * (g.poly_deg + 1) * 3 * Cfg::C : 0;
        const unsigned ntb = (unsigned)((na.M + kZm2TabNodes - 1) / kZm2TabNodes);
        auto table = [&](auto kt, size_t rowbytes) {
          const size_t tsm = (size_t)kZm2TabNodes * rowbytes + psm;
          PNB_CUDA(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
          kt<<<ntb, 3 * kZm2TabNodes, tsm, p->stream>>>(g, na, nd->d_rows, 0);
          p->launches++;
        };
        if (flavor) table(k_node_table2<R, M_, true, false, CPLX, Cfg>, ZmRowOf<R, Cfg, true, false, CPLX>::ROWBYTES);
        else table(k_node_table2<R, M_, false, false, CPLX, Cfg>, ZmRowOf<R, Cfg, false, false, CPLX>::ROWBYTES);
        nd->rows_flavor = flavor; nd->rows_plan = p; nd->rows_gen = p->win_gen;
      } else if (!reuse) {
        flavor = nd->rows_flavor;      // the table phase of this call ran earlier (side stream)
      }
    }
    return flavor;
  }
  template <bool CPLX, int M_, bool GRAD>
  static void launch_zm4(P *p, Nd *nd, const NodeArgs<R> &na, bool scatter) {
    if constexpr (sizeof(R) == 8 && Zm4Ok<M_>::value) {
      typedef Zm4Cfg<M_> Cfg;
      typedef Zm4Smem<CPLX, M_> Sm;
      const TileGeom tg = tile_geom(p, nullptr);
      const Zm2Geom zg = zm2_work_items(p, tg, nd, na.M);
      const int ncol = tg.nt[0] * tg.nt[1];
      if (scatter && (p->b_phase & 2)) {       // node values in sorted order
        constexpr int NVP = Sm::template Scat<GRAD, GRAD>::NVP;
        ensure(&nd->d_vals, &nd->cap_vals, (size_t)na.M * NVP + 64);
        const long long n = (long long)na.M * NVP;
        k_pack_vals<CPLX, GRAD><<<(unsigned)((n + 255) / 256), 256, 0, p->stream>>>(na, nd->d_vals);
        p->launches++;
      }
      // kernels of one launch: all columns, or one column batch with its own piece of the table
      auto run = [&](const Zm2Geom &zgk, unsigned nblk, const R *tab, int flavor) {
        if (!scatter) {
          const CUtensorMap tm = make_grid_tmap<R>(p->d_grid, p->L, CPLX ? 2 : 1, Cfg::R0, Sm::R1, Cfg::ZS, 0);
          GatherOut<R> out;
          out.perm = na.perm; out.f = na.f; out.f_stride = na.f_stride; out.f_off = na.f_off; out.grad = na.grad; out.accumulate = na.accumulate;
          auto go = [&](auto kern) {
            PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sm::gather));
            kern<<<nblk, (Cfg::NW + 1) * 32, Sm::gather, p->stream>>>(tm, zgk, tab, nd->d_tile_start, out);
            p->launches++;
          };
          if (flavor) go(k_gather_mma4<CPLX, M_, GRAD, true>);
          else if constexpr (!GRAD) go(k_gather_mma4<CPLX, M_, false, false>);
        } else {
          const CUtensorMap tm = make_grid_tmap<R>(p->d_grid, p->L, CPLX ? 2 : 1, 1, Sm::RH, Cfg::ZS, 0);
          auto go = [&](auto kern, int smem) {
            PNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            kern<<<2 * nblk, (Sm::NCW + 1) * 32, (size_t)smem, p->stream>>>(tm, zgk, tab, nd->d_vals, nd->d_tile_start);
            p->launches++;
          };
          if (flavor) go(k_scatter_mma4<CPLX, M_, GRAD, true>, Sm::template Scat<GRAD, true>::bytes);
          else if constexpr (!GRAD) go(k_scatter_mma4<CPLX, M_, false, false>, Sm::template Scat<false, false>::bytes);
        }
      };
      const size_t rowlen = ZmRowOf<R, Cfg, GRAD, false, CPLX>::ROWLEN;
      static const double cap_gb = getenv("PNFFT_B200_TABLE_GB") ? atof(getenv("PNFFT_B200_TABLE_GB")) : 24.0;
      const double need_gb = (double)na.M * rowlen * sizeof(R) / 1073741824.0;
      const bool cached = nd->binned && nd->rows_plan == (const void *)p && nd->rows_gen == p->win_gen && nd->rows_flavor >= (GRAD ? 1 : 0) && nd->d_rows;
      if (need_gb <= cap_gb || p->b_phase != 3 || cached) {
        const int flavor = zm4_rows<CPLX, M_, GRAD>(p, nd, na);
        if (p->b_phase & 2) run(zg, (unsigned)(ncol * zg.nseg), nd->d_rows, flavor);
      } else {
        // the table would not fit the budget: built and consumed in column batches, nothing kept (as launch_zm3)
        const GridGeom<R> g = geom(p);
        const int nbatch = std::min(ncol, (int)std::ceil(2.0 * need_gb / cap_gb));
        std::vector<int> cb((size_t)nbatch + 1), nb((size_t)nbatch + 1);
        for (int k = 0; k <= nbatch; k++) cb[(size_t)k] = (int)((long long)ncol * k / nbatch);
        nb[0] = 0; nb[(size_t)nbatch] = na.M;
        const size_t per_col = (size_t)tg.nt[2] * Cfg::SUB;
        for (int k = 1; k < nbatch; k++)
          PNB_CUDA(cudaMemcpyAsync(&nb[(size_t)k], nd->d_tile_start + (size_t)cb[(size_t)k] * per_col, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
        PNB_CUDA(cudaStreamSynchronize(p->stream));
        size_t max_rows = 0;
        for (int k = 0; k < nbatch; k++) max_rows = std::max(max_rows, (size_t)(nb[(size_t)k + 1] - nb[(size_t)k]));
        ensure(&nd->d_wtab, &nd->cap_wtab, max_rows * rowlen + 64);
        const size_t psm = g.poly ? sizeof(R) * (size_t)2 * (g.poly_deg + 1) * 3 * Cfg::C : 0;
        for (int k = 0; k < nbatch; k++) {
          const int first = nb[(size_t)k], last = nb[(size_t)k + 1];
          if (last <= first) continue;
          NodeArgs<R> nb_args = na;
          nb_args.M = last;
          R *tab = nd->d_wtab - (size_t)first * rowlen;        // rows are addressed by their absolute sorted position
          Zm2Geom zgk = zg;
          zgk.col0 = cb[(size_t)k];
          auto kt = k_node_table2<R, M_, GRAD, false, CPLX, Cfg>;
          const size_t tsm = (size_t)kZm2TabNodes * ZmRowOf<R, Cfg, GRAD, false, CPLX>::ROWBYTES + psm;
          PNB_CUDA(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
          kt<<<(unsigned)((last - first + kZm2TabNodes - 1) / kZm2TabNodes), 3 * kZm2TabNodes, tsm, p->stream>>>(g, nb_args, tab, first);
          p->launches++;
          run(zgk, (unsigned)((cb[(size_t)k + 1] - cb[(size_t)k]) * zg.nseg), tab, GRAD ? 1 : 0);
        }
      }
      PNB_CUDA(cudaGetLastError());
    }
  }
  static void launch_generic(P *p, const NodeArgs<R> &na, bool scatter, bool cplx) {
    const GridGeom<R> g = geom(p);
    const int wpb = 8;
    const size_t smem = (size_t)wpb * 6 * g.cutoff * sizeof(R);
    const int nblk = (na.M + wpb - 1) / wpb;
    if (cplx) {
      if (!scatter) k_gather_generic<R, true><<<nblk, wpb * 32, smem, p->stream>>>(g, (const R *)p->d_grid, na);
      else k_scatter_generic<R, true><<<nblk, wpb * 32, smem, p->stream>>>(g, (R *)p->d_grid, na);
    } else {
      if (!scatter) k_gather_generic<R, false><<<nblk, wpb * 32, smem, p->stream>>>(g, (const R *)p->d_grid, na);
      else k_scatter_generic<R, false><<<nblk, wpb * 32, smem, p->stream>>>(g, (R *)p->d_grid, na);
    }
    PNB_CUDA(cudaGetLastError());
    p->launches++;
  }
  // PNFFT_B200_MMA4_SCATTER=0: the adjoint of family 3 on the round-1 kernels (v1 z-march at m = 8, generic at m = 5, 7)
  template <bool CPLX, int M_, bool GRAD>
  static void launch_fam3(P *p, Nd *nd, const NodeArgs<R> &na, bool scatter) {
    static const bool sc4 = env_flag("PNFFT_B200_MMA4_SCATTER", true);
    if (!scatter || sc4) launch_zm4<CPLX, M_, GRAD>(p, nd, na, scatter);
    else if constexpr (M_ == 8) { if (p->b_phase & 2) launch_zm<CPLX, M_, GRAD>(p, nd, na, true); }
    else { if (p->b_phase & 2) launch_generic(p, na, true, CPLX); }
  }

  template <bool CPLX, int M_, bool GRAD>
  static void launch_zmarch(P *p, Nd *nd, const NodeArgs<R> &na, bool scatter) {
    if (kernel_family(p) == 3) { launch_fam3<CPLX, M_, GRAD>(p, nd, na, scatter); return; }
    if (kernel_family(p) == 2) {
      if (use_mma<M_>(p)) launch_zm3<CPLX, M_, GRAD>(p, nd, na, scatter);
      else launch_zm2<CPLX, M_, GRAD>(p, nd, na, scatter);
    } else launch_zm<CPLX, M_, GRAD>(p, nd, na, scatter);
  }

  template <bool CPLX> static void launch_B(P *p, Nd *nd, const NodeArgs<R> &na, bool scatter) {
    bool tiled = false;
    tile_geom(p, &tiled);
    const bool grad = na.grad != nullptr;
    if (na.M == 0) return;
    if (tiled) {
      switch (p->L.m) {
        case 4: if (grad) launch_zmarch<CPLX, 4, true>(p, nd, na, scatter); else launch_zmarch<CPLX, 4, false>(p, nd, na, scatter); return;
        case 6: if (grad) launch_zmarch<CPLX, 6, true>(p, nd, na, scatter); else launch_zmarch<CPLX, 6, false>(p, nd, na, scatter); return;
        case 8: if (grad) launch_zmarch<CPLX, 8, true>(p, nd, na, scatter); else launch_zmarch<CPLX, 8, false>(p, nd, na, scatter); return;
        case 5: if (kernel_family(p) == 3) { if (grad) launch_fam3<CPLX, 5, true>(p, nd, na, scatter); else launch_fam3<CPLX, 5, false>(p, nd, na, scatter); return; } break;
        case 7: if (kernel_family(p) == 3) { if (grad) launch_fam3<CPLX, 7, true>(p, nd, na, scatter); else launch_fam3<CPLX, 7, false>(p, nd, na, scatter); return; } break;
        default: break;
      }
    }
    launch_generic(p, na, scatter, CPLX);
  }
  static void launch_B_any(P *p, Nd *nd, const NodeArgs<R> &na, bool scatter) {
    if (p->L.c2r) launch_B<false>(p, nd, na, scatter); else launch_B<true>(p, nd, na, scatter);
  }

  // -------------------------------------------------------------------------------------------
  // node coordinates on the device, binning, and when both may be reused
  //   * pnfft_precompute_psi pins the node set (the reference's own contract: tables belong to the x they were made from)
  //   * nodes->x_static (extension pnfft_b200_nodes_x_static, or PNFFT_B200_X_STATIC=1) promises that x does not change
  //     between calls: the upload and the binning of the first call are reused until pnfft_set_x
  //   * device-resident x: a 64-bit content hash (one streaming pass, ~0.1 ms for 2^24 nodes) tells whether the bins of
  //     the previous call still belong to these coordinates (PNFFT_B200_X_HASH=0 switches it off)
  // -------------------------------------------------------------------------------------------
  static bool env_flag(const char *name, bool dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) != 0 : dflt;
  }
  // `after`: stream the coordinates were just written on (an upload), nullptr if they have been resident all along
  static void hash_device_x_queue(P *p, Nd *nd, const R *dx, cudaStream_t after = nullptr) {
    const size_t n = 3 * (size_t)nd->local_M;
    if (!nd->h_hash) {
      PNB_CUDA(cudaHostAlloc((void **)&nd->h_hash, sizeof(unsigned long long), cudaHostAllocDefault));
      PNB_CUDA(cudaMalloc((void **)&nd->d_hash, sizeof(unsigned long long)));
    }
    if (after && after != p->copy_stream) {
      PNB_CUDA(cudaEventRecord(p->ev_copy[2], after));
      PNB_CUDA(cudaStreamWaitEvent(p->copy_stream, p->ev_copy[2], 0));
    }
    PNB_CUDA(cudaMemsetAsync(nd->d_hash, 0, sizeof(unsigned long long), p->copy_stream));
    if (n) k_hash_words<R><<<148 * 8, 256, 0, p->copy_stream>>>(dx, (long long)n, nd->d_hash);
    PNB_CUDA(cudaMemcpyAsync(nd->h_hash, nd->d_hash, sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->copy_stream));
    p->launches++;
  }
  static unsigned long long hash_device_x_result(P *p, Nd *nd) {
    PNB_CUDA(cudaStreamSynchronize(p->copy_stream));
    return *nd->h_hash | 1ull;      // never 0 (0 = "no hash")
  }
  static unsigned long long hash_device_x(P *p, Nd *nd, const R *dx, cudaStream_t after = nullptr) {
    hash_device_x_queue(p, nd, dx, after);
    return hash_device_x_result(p, nd);
  }
  // dx_known: the coordinates are on the device already (second interlacing pass)
  // adj_spec: the caller is pnfft_adj and can redo its work (see Nodes::d_x_alt)
  static const R *prepare_nodes(P *p, Nd *nd, int ev_after_copy, int ev_after_bin, const R *dx_known = nullptr, int adj_spec = -1) {
    const size_t M = (size_t)nd->local_M;
    static const bool env_static = env_flag("PNFFT_B200_X_STATIC", false);
    static const bool use_hash = env_flag("PNFFT_B200_X_HASH", true);
    static const bool hash_up = env_flag("PNFFT_B200_X_HASH_UPLOADS", true);
    const bool is_static = nd->x_static || env_static;
    const bool pinned = (nd->precompute_flags & P_PRE_PSI) != 0;
    const int fam = kernel_family(p);
    const R *dx;
    unsigned long long h = 0;
    static const bool spec_env = env_flag("PNFFT_B200_ADJ_SPECULATE", true);
    const bool host_x = nd->x && !is_device_ptr(nd->x);
    bool hashed_upload = false;
    if (dx_known) { dx = dx_known; h = nd->known_hash; nd->known_hash = 0; }
    else if (adj_spec == 1 && spec_env && host_x && use_hash && hash_up && !pinned && !is_static && M && nd->adj_same_x_last &&
             nd->binned && nd->bin_plan == (const void *)p && nd->bin_family == fam && nd->d_x && nd->d_x_bound == nd->d_x && nd->bin_hash != 0) {
      // the coordinates were the trafo's last time: start on the bins at hand, check behind the transform's back
      ensure(&nd->d_x_alt, &nd->cap_x_alt, 3 * M);
      PNB_CUDA(cudaEventRecord(p->ev_copy[0], p->stream));             // behind f / grad_f on the bus
      PNB_CUDA(cudaStreamWaitEvent(p->copy_stream, p->ev_copy[0], 0));
      PNB_CUDA(cudaMemcpyAsync(nd->d_x_alt, nd->x, sizeof(R) * 3 * M, cudaMemcpyHostToDevice, p->copy_stream));
      hash_device_x_queue(p, nd, nd->d_x_alt);
      nd->spec_pending = true;
      dx = nd->d_x;
      h = nd->bin_hash;
    }
    else if ((pinned || is_static) && nd->binned && nd->d_x_bound && nd->bin_plan == (const void *)p) dx = nd->d_x_bound;
    else if (nd->x && is_device_ptr(nd->x)) {
      dx = nd->x;
      if (use_hash && !pinned && M) h = hash_device_x(p, nd, dx);
    } else if (is_static && nd->x_uploaded && nd->d_x) dx = nd->d_x;
    else if (p->x_via_copy_stream && nd->x && M) {
      // host-resident coordinates travel on the copy stream, right behind f_hat on the bus, while D and F run on the
      // plan's stream (trafo recorded ev_copy[0] after the f_hat upload)
      ensure(&nd->d_x, &nd->cap_x, 3 * M);
      PNB_CUDA(cudaStreamWaitEvent(p->copy_stream, p->ev_copy[0], 0));
      PNB_CUDA(cudaMemcpyAsync(nd->d_x, nd->x, sizeof(R) * 3 * M, cudaMemcpyHostToDevice, p->copy_stream));
      PNB_CUDA(cudaEventRecord(p->ev_copy[1], p->copy_stream));
      PNB_CUDA(cudaStreamWaitEvent(p->stream, p->ev_copy[1], 0));
      dx = nd->d_x;
      nd->x_uploaded = true;
      // the uploaded coordinates may be the ones the bins and the window table were made from (trafo then adj of one
      // step): one streaming pass over them tells, and saves the binning and the table
      if (use_hash && hash_up && !pinned) { h = hash_device_x(p, nd, dx, p->copy_stream); hashed_upload = true; }
    } else {
      dx = dev_in(p, nd->x, &nd->d_x, &nd->cap_x, 3 * M, true);
      nd->x_uploaded = true;
      if (use_hash && hash_up && !pinned && M && host_x) { h = hash_device_x(p, nd, dx, p->stream); hashed_upload = true; }
    }
    if (ev_after_copy >= 0) PNB_CUDA(cudaEventRecord(p->ev[ev_after_copy], p->stream));
    bool valid = nd->binned && nd->bin_plan == (const void *)p && nd->bin_family == fam && nd->d_x_bound == dx;
    if (valid && !pinned && !is_static) valid = h != 0 && nd->bin_hash == h;     // device x: same content as last time?
    if (adj_spec >= 0 && hashed_upload) nd->adj_same_x_last = valid;
    static const bool dbg = env_flag("PNFFT_B200_DEBUG_BIN", false);
    if (dbg) fprintf(stderr, "prepare_nodes: valid=%d binned=%d plan=%d fam=%d/%d bound=%p dx=%p h=%llx bin_hash=%llx static=%d pinned=%d\n", (int)valid, (int)nd->binned,
                     (int)(nd->bin_plan == (const void *)p), nd->bin_family, fam, (const void *)nd->d_x_bound, (const void *)dx, h, nd->bin_hash, (int)is_static, (int)pinned);
    if (!valid) {
      nd->rows_flavor = -1;          // the cached node-table rows belong to the old bins
      bin_nodes(p, nd, dx);
      nd->bin_plan = p; nd->bin_family = fam; nd->d_x_bound = dx; nd->bin_hash = h;
      nd->binned = pinned || is_static || h != 0;
      if (pinned && nd->d_pre_psi) fill_pre_psi(p, nd, dx);    // the tables are stored in sorted order: redo them with the bins
    }
    if (ev_after_bin >= 0) PNB_CUDA(cudaEventRecord(p->ev[ev_after_bin], p->stream));
    return dx;
  }

  static void fill_pre_psi(P *p, Nd *nd, const R *dx) {
    const size_t M = (size_t)nd->local_M;
    if (!M) return;
    const GridGeom<R> g = geom(p);
    k_precompute_psi<R><<<(unsigned)((M * 32 + 255) / 256), 256, 0, p->stream>>>(g, dx, nd->d_perm, (int)M, nd->d_pre_psi, nd->d_pre_dpsi);
    p->launches++;
  }

  static void precompute_psi(P *p, Nd *nd, unsigned pre_flags) {
    if (!p || !nd) return;
    for (int k = 0; k < 2; k++) {
      BinState<R> &b = k ? nd->il : *static_cast<BinState<R> *>(nd);
      cudaFree(b.d_pre_psi); cudaFree(b.d_pre_dpsi);
      b.d_pre_psi = b.d_pre_dpsi = nullptr;
      b.binned = false; b.d_x_bound = nullptr;
    }
    nd->precompute_flags = 0;
    nd->x_uploaded = false;
    if (!(pre_flags & P_PRE_PSI)) return;
    if (pre_flags & P_PRE_FULL) {
      fprintf(stderr, "pnfft-b200: PNFFT_PRE_FULL is not supported; using the tensor-product tables (PNFFT_PRE_PSI)\n");
      pre_flags &= ~P_PRE_FULL;
    }
    const size_t M = (size_t)nd->local_M;
    const int c = p->L.cutoff;
    // the reference never fills pre_dpsi (PNFFT_DIFF_AD == 0 makes its test always false, SURVEY 8a defect 2);
    // here PRE_GRAD_PSI is honoured whenever the gradient is taken analytically
    const bool want_d = (pre_flags & P_PRE_GRAD_PSI) && !(p->pnfft_flags & F_DIFF_IK);
    const int npass = (p->pnfft_flags & F_INTERLACED) ? 2 : 1;    // reference pre_psi / pre_psi_il (ndft-parallel.c:1184-1240)
    const R *dx = nullptr;
    for (int pass = 0; pass < npass; pass++) {
      p->il_pass = pass;
      if (pass) nd->swap_il();
      dx = prepare_nodes(p, nd, -1, -1, dx);
      nd->binned = true;
      PNB_CUDA(cudaMalloc((void **)&nd->d_pre_psi, sizeof(R) * 3 * c * (M ? M : 1)));
      if (want_d) PNB_CUDA(cudaMalloc((void **)&nd->d_pre_dpsi, sizeof(R) * 3 * c * (M ? M : 1)));
      fill_pre_psi(p, nd, dx);
      if (pass) nd->swap_il();
    }
    p->il_pass = 0;
    nd->precompute_flags = pre_flags;
    PNB_CUDA(cudaStreamSynchronize(p->stream));
  }

  // -------------------------------------------------------------------------------------------
  // D / D^H and ik scaling launches
  // -------------------------------------------------------------------------------------------
  // the local f_hat block in memory order: (k0, k1, k2), or (k1, k2, k0) for PNFFT_TRANSPOSED_F_HAT
  static void fhat_axes(const P *p, int ax[3]) {
    if (p->L.transposed) { ax[0] = 1; ax[1] = 2; ax[2] = 0; } else { ax[0] = 0; ax[1] = 1; ax[2] = 2; }
  }
  static FhatGeom fhat_geom(const P *p, int il_sign) {
    FhatGeom fg;
    int ax[3];
    fhat_axes(p, ax);
    for (int k = 0; k < 3; k++) {
      fg.l[k] = (int)p->L.local_N[ax[k]]; fg.s[k] = (int)p->L.local_N_start[ax[k]]; fg.n[k] = (double)p->L.n[ax[k]];
    }
    fg.il_sign = il_sign;
    return fg;
  }
  static void run_deconv(P *p, const C *src_f_hat, C *f_hat_acc, bool adjoint) {
    // second pass of an interlaced plan: exp(+ pi i sum k/n) after D, exp(- pi i sum k/n) before D^H (reference matrix_D.c:247-262)
    const bool il = (p->pnfft_flags & F_INTERLACED) && p->il_pass == 1;
    const FhatGeom fg = fhat_geom(p, il ? (adjoint ? -1 : +1) : 0);
    if (fg.l[0] <= 0 || fg.l[1] <= 0 || fg.l[2] <= 0) return;
    int ax[3];
    fhat_axes(p, ax);
    const int bs = fg.l[2] >= 128 ? 128 : 32;
    const R *cA = p->d_invphi[ax[0]], *cB = p->d_invphi[ax[1]], *cC = p->d_invphi[ax[2]];
    if (!adjoint) k_deconv_fwd<R, C><<<grid3(fg.l[0], fg.l[1], fg.l[2], bs), bs, 0, p->stream>>>(src_f_hat, p->d_g1, cA, cB, cC, fg);
    else k_deconv_adj<R, C><<<grid3(fg.l[0], fg.l[1], fg.l[2], bs), bs, 0, p->stream>>>(f_hat_acc, p->d_g1, cA, cB, cC, fg);
    p->launches++;
  }
  static void run_ik(P *p, const C *in, C *out, int mode, int dim) {
    const FhatGeom fg = fhat_geom(p, 0);
    if (fg.l[0] <= 0 || fg.l[1] <= 0 || fg.l[2] <= 0) return;
    int ax[3], axis = 0;
    fhat_axes(p, ax);
    for (int k = 0; k < 3; k++) if (ax[k] == dim) axis = k;
    const int bs = fg.l[2] >= 128 ? 128 : 32;
    k_ik_scale<R, C><<<grid3(fg.l[0], fg.l[1], fg.l[2], bs), bs, 0, p->stream>>>(in, out, mode, axis, fg);
    p->launches++;
  }

  static void run_ik2(P *p, const C *in, C *out, int comp) {
    static const int t1[6] = {0, 0, 0, 1, 1, 2}, t2[6] = {0, 1, 2, 1, 2, 2};
    const FhatGeom fg = fhat_geom(p, 0);
    if (fg.l[0] <= 0 || fg.l[1] <= 0 || fg.l[2] <= 0) return;
    int ax[3], a = 0, b = 0;
    fhat_axes(p, ax);
    for (int k = 0; k < 3; k++) { if (ax[k] == t1[comp]) a = k; if (ax[k] == t2[comp]) b = k; }
    const int bs = fg.l[2] >= 128 ? 128 : 32;
    k_ik_scale2<R, C><<<grid3(fg.l[0], fg.l[1], fg.l[2], bs), bs, 0, p->stream>>>(in, out, a, b, fg);
    p->launches++;
  }
  // Hessian with analytic window derivatives: one generic pass over the padded grid the gather just read
  static void launch_hessian(P *p, const NodeArgs<R> &na, R *dh) {
    if (na.M == 0 || !dh) return;
    const GridGeom<R> g = geom(p);
    const int wpb = 8;
    const size_t smem = (size_t)wpb * 9 * g.cutoff * sizeof(R);
    const int nblk = (na.M + wpb - 1) / wpb;
    if (p->L.c2r) k_hessian_generic<R, false><<<nblk, wpb * 32, smem, p->stream>>>(g, (const R *)p->d_grid, na, dh);
    else k_hessian_generic<R, true><<<nblk, wpb * 32, smem, p->stream>>>(g, (const R *)p->d_grid, na, dh);
    PNB_CUDA(cudaGetLastError());
    p->launches++;
  }

  static void rec(P *p, int i) { PNB_CUDA(cudaEventRecord(p->ev[i], p->stream)); }
  static double ms(P *p, int a, int b) { float t = 0; cudaEventElapsedTime(&t, p->ev[a], p->ev[b]); return (double)t; }

  // -------------------------------------------------------------------------------------------
  // trafo  (reference api/api-basic.c:170-244)
  // -------------------------------------------------------------------------------------------
  static void trafo(P *p, Nd *nd, unsigned cf) {
    if (!p) return;
    if (!nd && !(cf & C_OMIT_CONV)) return;
    if (cf & C_DIRECT) {       // the slow NDFT (reference api/api-basic.c:224-231): csrc/direct.cuh
      if (!nd) return;
      if (!p->f_hat) { fprintf(stderr, "pnfft-b200: f_hat is not set\n"); return; }
      Direct<R>::trafo(p, nd, cf);
      return;
    }
    cudaStream_t st = p->stream;
    const Layout &L = p->L;
    const size_t nloc = (size_t)local_N_total(p);
    const int NC = L.c2r ? 1 : 2;
    const bool ik = (p->pnfft_flags & F_DIFF_IK) != 0;
    const int npass = (p->pnfft_flags & F_INTERLACED) ? 2 : 1;   // reference api-basic.c:233-240
    const bool want_h = (cf & C_HESSIAN_F) && nd && nd->hessian_f;      // reference api-basic.c:148-166 (ik), assign.c:881 (AD)
    if ((cf & C_HESSIAN_F) && nd && !nd->hessian_f && !p->warned_hessian) {
      fprintf(stderr, "pnfft-b200: PNFFT_COMPUTE_HESSIAN_F without hessian_f (PNFFT_MALLOC_HESSIAN_F / pnfft_set_hessian_f)\n");
      p->warned_hessian = true;
    }
    rec(p, 0);
    const size_t M = nd ? (size_t)nd->local_M : 0;
    static const bool no_prefetch = getenv("PNFFT_B200_NO_PREFETCH") && atoi(getenv("PNFFT_B200_NO_PREFETCH")) != 0;
    // host-resident f_hat AND host-resident coordinates on one GPU: the coordinates go over the bus FIRST, so that binning
    // and the window table (node stream) run while f_hat is still travelling and D and F wait for it; the other order
    // leaves the table (6 ms at C3) behind both uploads.  PNFFT_B200_X_FIRST=0 restores f_hat first.
    static const bool xf_env = env_flag("PNFFT_B200_X_FIRST", true);
    const bool x_first = xf_env && !no_prefetch && !(cf & (C_OMIT_DECONV | C_OMIT_CONV)) && p->f_hat && !is_device_ptr(p->f_hat) && nd && nd->x &&
                         !is_device_ptr(nd->x) && M > 0 && p->mesh.size == 1 && !(p->pnfft_flags & (F_DIFF_IK | F_INTERLACED)) &&
                         kernel_family(p) == 2 && (cf & (C_F | C_GRAD_F)) && !(nd->precompute_flags & P_PRE_PSI) &&
                         (double)M * 108.0 * sizeof(R) <= 24.0 * 1073741824.0;   // the two-phase launch builds the whole window table at once
    // ---- f_hat on the device ----
    const C *fh = nullptr;
    if (!(cf & C_OMIT_DECONV)) {
      if (!p->f_hat) { fprintf(stderr, "pnfft-b200: f_hat is not set\n"); return; }
      if (is_device_ptr(p->f_hat)) fh = p->f_hat;
      else { if (!x_first) PNB_CUDA(cudaMemcpyAsync(p->d_f_hat, p->f_hat, sizeof(C) * nloc, cudaMemcpyHostToDevice, st)); fh = p->d_f_hat; }
    }
    if (!x_first) rec(p, 1);
    // the x upload of this call may overtake D and F (prepare_nodes); PNFFT_B200_NO_PREFETCH=1 keeps the single queue
    PNB_CUDA(cudaEventRecord(p->ev_copy[0], st));
    p->x_via_copy_stream = !no_prefetch;

    R *df = nullptr, *dg = nullptr, *dh = nullptr;
    const R *dx = nullptr;
    const bool conv = !(cf & C_OMIT_CONV);
    const bool acc = (cf & C_ACCUMULATED) != 0;
    p->side_nodes = false;
    p->x_first = x_first;
    for (int pass = 0; pass < npass; pass++) {
      p->il_pass = pass;
      if (pass) nd->swap_il();
      const bool acc_pass = acc || pass > 0;       // the second pass adds to the first (reference assign.c:689-691)
      // ---- D ----
      if (!x_first) {
        if (fh) run_deconv(p, fh, nullptr, false);
        rec(p, 2);
      }
      // node-side preparation is shared by all B passes of this call
      auto node_setup = [&]() {
        dx = prepare_nodes(p, nd, 4, 5, dx);
        if (pass == 0) {
          if (cf & C_F) df = dev_in(p, nd->f, &nd->d_f, &nd->cap_f, (size_t)NC * M, acc);
          if (cf & C_GRAD_F) dg = dev_in(p, nd->grad_f, &nd->d_grad_f, &nd->cap_grad, (size_t)3 * NC * M, acc);
          if (want_h) dh = dev_in(p, nd->hessian_f, &nd->d_hess, &nd->cap_hess, (size_t)6 * NC * M, acc);
        }
      };
      auto base_args = [&]() {
        NodeArgs<R> na;
        na.x = dx; na.perm = nd->d_perm; na.M = (int)M; na.f = nullptr; na.f_stride = 1; na.f_off = 0; na.grad = nullptr;
        na.accumulate = acc_pass ? 1 : 0;
        const bool use_pre = (nd->precompute_flags & P_PRE_PSI) && nd->d_pre_psi;
        na.pre_psi = use_pre ? nd->d_pre_psi : nullptr;
        na.pre_dpsi = use_pre ? nd->d_pre_dpsi : nullptr;
        return na;
      };

      if (!ik) {
        // z-march v2: the node side of the call (x upload, binning, node table) needs nothing from the grid, so it runs on
        // its own stream while D, F and the halo exchange run here; the gather waits for both.
        // On by default for multi-rank plans, where F and the halo exchange are latency bound (measured on 1x2 B200:
        // step 36.57 -> 36.35 ms); on one GPU the two sides only compete for HBM (65.4 vs 65.5 ms) and the stage timers
        // stay cleaner without it.  PNFFT_B200_SIDE_STREAM=0 / 1 forces it off / on.
        static const int side_env = getenv("PNFFT_B200_SIDE_STREAM") ? atoi(getenv("PNFFT_B200_SIDE_STREAM")) : -1;
        const bool want_side = x_first || (side_env >= 0 ? side_env != 0 : p->mesh.size > 1);
        p->side_nodes = want_side && npass == 1 && conv && M > 0 && kernel_family(p) == 2 && (cf & (C_F | C_GRAD_F));
        NodeArgs<R> na_side;
        // the grid side is queued FIRST unless the coordinates go over the bus first: prepare_nodes blocks the host until
        // an upload of host-resident coordinates has been hashed, and F must not wait in the host's queue behind that
        const bool grid_first = p->side_nodes && !x_first;
        if (grid_first) {
          if (!(cf & C_OMIT_FFT)) fft_forward(p);
          rec(p, 3);
          halo(p, false);
          rec(p, 6);
        }
        if (p->side_nodes) {
          p->stream = p->node_stream;
          PNB_CUDA(cudaStreamWaitEvent(p->node_stream, p->ev_copy[0], 0));   // the node stream starts where this call started
          rec(p, 11);
          node_setup();                 // records ev[4] (x on the device) and ev[5] (binned) on the node stream
          rec(p, 9);
          na_side = base_args();
          na_side.f = df; na_side.grad = dg;
          if (na_side.grad && na_side.pre_psi && !na_side.pre_dpsi) na_side.pre_psi = nullptr;
          p->b_phase = 1;
          if (df || dg) launch_B_any(p, nd, na_side, false);
          p->b_phase = 3;
          rec(p, 10);
          p->stream = st;
        }
        if (x_first) {      // the coordinates are on the device by now (prepare_nodes waited for their hash)
          PNB_CUDA(cudaMemcpyAsync(p->d_f_hat, p->f_hat, sizeof(C) * nloc, cudaMemcpyHostToDevice, st));
          rec(p, 1);
          run_deconv(p, fh, nullptr, false);
          rec(p, 2);
        }
        if (!grid_first) {
          if (!(cf & C_OMIT_FFT)) fft_forward(p);
          rec(p, 3);
        }
        if (p->side_nodes) {
          if (!grid_first) {
            halo(p, false);
            rec(p, 6);
          }
          PNB_CUDA(cudaStreamWaitEvent(st, p->ev[10], 0));
          rec(p, 12);                    // grid side and node side have met: the gather starts here
          p->b_phase = 2;
          if (df || dg) launch_B_any(p, nd, na_side, false);
          p->b_phase = 3;
          if (dh) launch_hessian(p, na_side, dh);
          rec(p, 7);
        } else if (conv) {
          node_setup();
          halo(p, false);
          rec(p, 6);
          NodeArgs<R> na = base_args();
          na.f = df; na.grad = dg;
          if (na.grad && na.pre_psi && !na.pre_dpsi) na.pre_psi = nullptr;   // no dpsi table: evaluate on the fly
          if (df || dg) launch_B_any(p, nd, na, false);
          if (dh) launch_hessian(p, na, dh);
          rec(p, 7);
        } else { rec(p, 4); rec(p, 5); rec(p, 6); rec(p, 7); }
      } else {
        // ik differentiation: 1 (f) + 3 (grad) passes of F and B (reference api-basic.c:100-167)
        if (((cf & C_GRAD_F) || want_h) && !(cf & C_OMIT_DECONV))
          PNB_CUDA(cudaMemcpyAsync(p->d_g1_buffer, p->d_g1, sizeof(C) * nloc, cudaMemcpyDeviceToDevice, st));
        rec(p, 3);
        if (conv) node_setup(); else { rec(p, 4); rec(p, 5); }
        if (cf & C_F) {
          if (!(cf & C_OMIT_FFT)) fft_forward(p);
          if (conv) { halo(p, false); NodeArgs<R> na = base_args(); na.f = df; launch_B_any(p, nd, na, false); }
        }
        if (cf & C_GRAD_F)
          for (int dim = 0; dim < 3; dim++) {
            if (!(cf & C_OMIT_DECONV)) run_ik(p, p->d_g1_buffer, p->d_g1, 0, dim);
            if (!(cf & C_OMIT_FFT)) fft_forward(p);
            if (conv) { halo(p, false); NodeArgs<R> na = base_args(); na.f = dg; na.f_stride = 3; na.f_off = dim; launch_B_any(p, nd, na, false); }
          }
        if (want_h)
          for (int comp = 0; comp < 6; comp++) {
            if (!(cf & C_OMIT_DECONV)) run_ik2(p, p->d_g1_buffer, p->d_g1, comp);
            if (!(cf & C_OMIT_FFT)) fft_forward(p);
            if (conv) { halo(p, false); NodeArgs<R> na = base_args(); na.f = dh; na.f_stride = 6; na.f_off = comp; launch_B_any(p, nd, na, false); }
          }
        rec(p, 6); rec(p, 7);
      }
      if (pass) nd->swap_il();
    }
    p->il_pass = 0;
    // ---- results back to the host where the user arrays live ----
    if (conv) {
      if (df && !is_device_ptr(nd->f)) PNB_CUDA(cudaMemcpyAsync(nd->f, df, sizeof(R) * NC * M, cudaMemcpyDeviceToHost, st));
      if (dg && !is_device_ptr(nd->grad_f)) PNB_CUDA(cudaMemcpyAsync(nd->grad_f, dg, sizeof(R) * 3 * NC * M, cudaMemcpyDeviceToHost, st));
      if (dh && !is_device_ptr(nd->hessian_f)) PNB_CUDA(cudaMemcpyAsync(nd->hessian_f, dh, sizeof(R) * 6 * NC * M, cudaMemcpyDeviceToHost, st));
    }
    rec(p, 8);
    p->x_via_copy_stream = false;
    PNB_CUDA(cudaStreamSynchronize(st));
    peer_check(p);
    finish_timers(p, false, ik);
  }

  // -------------------------------------------------------------------------------------------
  // adjoint  (reference api/api-basic.c:249-378)
  // -------------------------------------------------------------------------------------------
  static void adj(P *p, Nd *nd, unsigned cf) {
    if (!p) return;
    if (!nd && !(cf & C_OMIT_CONV)) return;
    if (cf & C_DIRECT) {       // the slow adjoint NDFT (reference api/api-basic.c:355-365): csrc/direct.cuh
      if (!nd) return;
      if (!p->f_hat) { fprintf(stderr, "pnfft-b200: f_hat is not set\n"); return; }
      if (!(cf & C_ACCUMULATED)) {
        const size_t bytes = sizeof(C) * (size_t)local_N_total(p);
        if (is_device_ptr(p->f_hat)) PNB_CUDA(cudaMemset(p->f_hat, 0, bytes)); else memset(p->f_hat, 0, bytes);
      }
      Direct<R>::adj(p, nd, cf);
      return;
    }
    cudaStream_t st = p->stream;
    const Layout &L = p->L;
    const size_t nloc = (size_t)local_N_total(p);
    const int NC = L.c2r ? 1 : 2;
    const bool ik = (p->pnfft_flags & F_DIFF_IK) != 0;
    const bool conv = !(cf & C_OMIT_CONV);
    const bool acc = (cf & C_ACCUMULATED) != 0;
    const size_t M = nd ? (size_t)nd->local_M : 0;
    const int npass = (p->pnfft_flags & F_INTERLACED) ? 2 : 1;   // reference api-basic.c:367-374
    if ((cf & C_HESSIAN_F) && !p->warned_hessian) {
      fprintf(stderr, "pnfft-b200: pnfft_adj has no Hessian part (as in the reference); PNFFT_COMPUTE_HESSIAN_F is ignored\n");
      p->warned_hessian = true;
    }
    rec(p, 0);
    R *df = nullptr, *dg = nullptr;
    const R *dx = nullptr;
    if (conv) {
      if (cf & C_F) df = dev_in(p, nd->f, &nd->d_f, &nd->cap_f, (size_t)NC * M, true);
      if (cf & C_GRAD_F) dg = dev_in(p, nd->grad_f, &nd->d_grad_f, &nd->cap_grad, (size_t)3 * NC * M, true);
    }
    // ---- f_hat accumulator on the device: zeroed (PNX(zero_f_hat), api-basic.c:355) or the caller's values ----
    C *fh = nullptr;
    bool fh_dev = false;
    if (p->f_hat) {
      fh_dev = is_device_ptr(p->f_hat);
      fh = fh_dev ? p->f_hat : p->d_f_hat;
      if (!acc) PNB_CUDA(cudaMemsetAsync(fh, 0, sizeof(C) * nloc, st));
      else if (!fh_dev) PNB_CUDA(cudaMemcpyAsync(fh, p->f_hat, sizeof(C) * nloc, cudaMemcpyHostToDevice, st));
    }
    // host-resident coordinates that were the trafo's last time: run on the bins at hand while they are uploaded and hashed
    // behind the transform (prepare_nodes); a different hash means new coordinates, and the adjoint is redone on them
    const bool may_spec = conv && npass == 1 && !acc && p->mesh.size == 1;
    for (int attempt = 0; attempt < 2; attempt++) {
      if (attempt && fh) PNB_CUDA(cudaMemsetAsync(fh, 0, sizeof(C) * nloc, st));
      for (int pass = 0; pass < npass; pass++) {
        p->il_pass = pass;
        if (pass) nd->swap_il();
        if (conv) dx = prepare_nodes(p, nd, 1, 2, dx, may_spec && attempt == 0 ? 1 : 0); else { rec(p, 1); rec(p, 2); }
        auto base_args = [&]() {
          NodeArgs<R> na;
          na.x = dx; na.perm = nd->d_perm; na.M = (int)M; na.f = nullptr; na.f_stride = 1; na.f_off = 0; na.grad = nullptr;
          na.accumulate = 0;
          const bool use_pre = (nd->precompute_flags & P_PRE_PSI) && nd->d_pre_psi;
          na.pre_psi = use_pre ? nd->d_pre_psi : nullptr;
          na.pre_dpsi = use_pre ? nd->d_pre_dpsi : nullptr;
          return na;
        };
        auto spread = [&](R *fptr, long long stride, long long off, R *gptr, int e_zero, int e_b, int e_h) {
          PNB_CUDA(cudaMemsetAsync(p->d_grid, 0, p->grid_bytes, st));   // reference ndft-parallel.c:2629-2635
          if (e_zero >= 0) rec(p, e_zero);
          NodeArgs<R> na = base_args();
          na.f = fptr; na.f_stride = stride; na.f_off = off; na.grad = gptr;
          if (na.grad && na.pre_psi && !na.pre_dpsi) na.pre_psi = nullptr;
          if (fptr || gptr) launch_B_any(p, nd, na, true);
          if (e_b >= 0) rec(p, e_b);
          halo(p, true);
          if (e_h >= 0) rec(p, e_h);
        };

        if (!ik) {
          if (conv) spread(df, 1, 0, dg, 3, 4, 5); else { rec(p, 3); rec(p, 4); rec(p, 5); }
          if (!(cf & C_OMIT_FFT)) fft_backward(p);
          rec(p, 6);
        } else {
          rec(p, 3); rec(p, 4); rec(p, 5);
          if (!(cf & C_OMIT_DECONV)) PNB_CUDA(cudaMemsetAsync(p->d_g1_buffer, 0, sizeof(C) * nloc, st));
          if (cf & C_F) {
            if (conv) spread(df, 1, 0, nullptr, -1, -1, -1);
            if (!(cf & C_OMIT_FFT)) fft_backward(p);
            run_ik(p, p->d_g1, p->d_g1_buffer, 2, 0);
          }
          if (cf & C_GRAD_F)
            for (int dim = 0; dim < 3; dim++) {
              if (conv) spread(dg, 3, dim, nullptr, -1, -1, -1);
              if (!(cf & C_OMIT_FFT)) fft_backward(p);
              if (!(cf & C_OMIT_DECONV)) run_ik(p, p->d_g1, p->d_g1_buffer, 1, dim);
            }
          PNB_CUDA(cudaMemcpyAsync(p->d_g1, p->d_g1_buffer, sizeof(C) * nloc, cudaMemcpyDeviceToDevice, st));
          rec(p, 6);
        }
        // ---- D^H: f_hat += g1 * 1/phi_hat ----
        if (fh && !(cf & C_OMIT_DECONV)) run_deconv(p, nullptr, fh, true);
        if (pass) nd->swap_il();
      }
      if (!nd || !nd->spec_pending) break;
      nd->spec_pending = false;
      const unsigned long long hx = hash_device_x_result(p, nd);
      if (hx == nd->bin_hash) break;
      PNB_CUDA(cudaStreamSynchronize(st));
      { R *t = nd->d_x; nd->d_x = nd->d_x_alt; nd->d_x_alt = t; const size_t c = nd->cap_x; nd->cap_x = nd->cap_x_alt; nd->cap_x_alt = c; }
      nd->binned = false; nd->adj_same_x_last = false; nd->known_hash = hx;
      dx = nd->d_x;
    }
    p->il_pass = 0;
    rec(p, 7);
    if (fh && !fh_dev) PNB_CUDA(cudaMemcpyAsync(p->f_hat, fh, sizeof(C) * nloc, cudaMemcpyDeviceToHost, st));
    rec(p, 8);
    PNB_CUDA(cudaStreamSynchronize(st));
    peer_check(p);
    finish_timers(p, true, ik);
  }

  static void finish_timers(P *p, bool adjoint, bool ik) {
    double *s = p->stage_ms[adjoint ? 1 : 0];
    double *T = adjoint ? p->timer_adj : p->timer_trafo;
    if (!adjoint) {
      s[0] = ms(p, 6, 7); s[1] = ms(p, 4, 5); s[2] = ms(p, 5, 6); s[3] = ms(p, 2, 3); s[4] = ms(p, 1, 2);
      s[5] = ms(p, 0, 1) + ms(p, 3, 4); s[6] = ms(p, 7, 8); s[7] = ms(p, 0, 8);
      if (p->side_nodes) {   // node side on its own stream: x upload ev[11]->ev[4], binning ev[4]->ev[5], node table ev[9]->ev[10]
        s[0] = ms(p, 12, 7) + ms(p, 9, 10); s[1] = ms(p, 4, 5); s[2] = ms(p, 3, 6); s[5] = ms(p, 0, 1) + ms(p, 11, 4);
        if (p->x_first) s[5] = ms(p, 0, 1);      // coordinates, then f_hat: one after the other on the bus
      }
    } else {
      s[0] = ms(p, 3, 4); s[1] = ms(p, 1, 2); s[2] = ms(p, 2, 3) + ms(p, 4, 5); s[3] = ms(p, 5, 6); s[4] = ms(p, 6, 7);
      s[5] = ms(p, 0, 1); s[6] = ms(p, 7, 8); s[7] = ms(p, 0, 8);
    }
    if (ik) { s[0] = s[1] = s[2] = s[3] = s[4] = 0; }
    T[T_ITER] += 1;
    T[T_WHOLE] += 1e-3 * s[7];
    T[T_LOOP_B] += 1e-3 * s[0];
    T[T_SORT_NODES] += 1e-3 * s[1];
    T[T_GCELLS] += 1e-3 * s[2];
    T[T_MATRIX_B] += 1e-3 * (s[0] + s[1] + s[2]);
    T[T_MATRIX_F] += 1e-3 * s[3];
    T[T_MATRIX_D] += 1e-3 * s[4];
  }
};

}  // namespace pnb
