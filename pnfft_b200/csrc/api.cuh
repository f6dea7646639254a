// extern "C" entry points of libpnfft_b200.so for one precision.  Included twice (api_d.cu, api_f.cu) with
//   PNX(name)  -> pnfft_##name / pnfftf_##name      RT -> double / float
// Signatures are those of include/pnfft.h (= reference api/pnfft.h:51-288).
#include <pnfft.h>

#include "core.cuh"

using pnb::INT;
typedef pnb::Core<RT> CoreT;
typedef pnb::Plan<RT> PlanT;
typedef pnb::Nodes<RT> NodesT;
typedef RT CT[2];

#define AS_PLAN(p) (reinterpret_cast<PlanT *>(p))
#define AS_NODES(p) (reinterpret_cast<NodesT *>(p))

namespace {

void local_size_impl(int d, const INT *N, const INT *n, const RT *x_max, int m, MPI_Comm comm, unsigned flags, bool c2r,
                     INT *local_N, INT *local_N_start, RT *lo, RT *up) {
  if (d != 3) { fprintf(stderr, "!!! Error in PNFFT Planer: d != 3 not yet implemented !!!\n"); return; }
  pnb::Mesh mesh;
  if (!pnb::get_mesh(comm, mesh)) return;
  pnb::Layout L;
  pnb::compute_layout<RT>(L, mesh, N, n, x_max, m, c2r, flags);
  for (int t = 0; t < 3; t++) { local_N[t] = L.local_N[t]; local_N_start[t] = L.local_N_start[t]; }
  pnb::node_borders<RT>(L, x_max, lo, up);
}

void adv_defaults(const INT *N, INT *n, RT *x_max, int *m) {
  // reference api/api-adv.c:172-192
  for (int t = 0; t < 3; t++) { n[t] = 2 * N[t]; x_max[t] = (RT)0.5; }
  *m = 6;
}

}  // namespace

extern "C" {

int PNX(create_procmesh)(int rnk, MPI_Comm comm, const int *np, MPI_Comm *comm_cart) {
  int periods[3] = {1, 1, 1}, size = 0, prod = 1;
  MPI_Comm_size(comm, &size);
  for (int t = 0; t < rnk; t++) prod *= np[t];
  if (prod != size) return 1;
  return MPI_Cart_create(comm, rnk, np, periods, 1, comm_cart);
}
int PNX(create_procmesh_2d)(MPI_Comm comm, int np0, int np1, MPI_Comm *comm_cart_2d) {
  const int np[2] = {np0, np1};
  return PNX(create_procmesh)(2, comm, np, comm_cart_2d);
}

void PNX(local_size_guru)(int d, const INT *N, const INT *n, const RT *x_max, int m, MPI_Comm comm_cart, unsigned pnfft_flags,
                          INT *local_N, INT *local_N_start, RT *lo, RT *up) {
  local_size_impl(d, N, n, x_max, m, comm_cart, pnfft_flags, false, local_N, local_N_start, lo, up);
}
void PNX(local_size_guru_c2r)(int d, const INT *N, const INT *n, const RT *x_max, int m, MPI_Comm comm_cart, unsigned pnfft_flags,
                              INT *local_N, INT *local_N_start, RT *lo, RT *up) {
  local_size_impl(d, N, n, x_max, m, comm_cart, pnfft_flags, true, local_N, local_N_start, lo, up);
}
void PNX(local_size_adv)(int d, const INT *N, MPI_Comm comm_cart, unsigned pnfft_flags, INT *local_N, INT *local_N_start, RT *lo, RT *up) {
  INT n[3]; RT x_max[3]; int m;
  adv_defaults(N, n, x_max, &m);
  local_size_impl(d, N, n, x_max, m, comm_cart, pnfft_flags, false, local_N, local_N_start, lo, up);
}
void PNX(local_size_adv_c2r)(int d, const INT *N, MPI_Comm comm_cart, unsigned pnfft_flags, INT *local_N, INT *local_N_start, RT *lo, RT *up) {
  INT n[3]; RT x_max[3]; int m;
  adv_defaults(N, n, x_max, &m);
  local_size_impl(d, N, n, x_max, m, comm_cart, pnfft_flags, true, local_N, local_N_start, lo, up);
}
void PNX(local_size_3d)(const INT *N, MPI_Comm comm_cart, unsigned pnfft_flags, INT *local_N, INT *local_N_start, RT *lo, RT *up) {
  PNX(local_size_adv)(3, N, comm_cart, pnfft_flags, local_N, local_N_start, lo, up);
}
void PNX(local_size_3d_c2r)(const INT *N, MPI_Comm comm_cart, unsigned pnfft_flags, INT *local_N, INT *local_N_start, RT *lo, RT *up) {
  PNX(local_size_adv_c2r)(3, N, comm_cart, pnfft_flags, local_N, local_N_start, lo, up);
}

PNX(plan) PNX(init_guru)(int d, const INT *N, const INT *n, const RT *x_max, int m, unsigned pnfft_flags, unsigned fftw_flags, MPI_Comm comm_cart) {
  if (d != 3) { fprintf(stderr, "!!! Error in PNFFT Planer: d != 3 not yet implemented !!!\n"); return nullptr; }
  return reinterpret_cast<PNX(plan)>(CoreT::init(N, n, x_max, m, pnfft_flags, fftw_flags, comm_cart, false));
}
PNX(plan) PNX(init_guru_c2r)(int d, const INT *N, const INT *n, const RT *x_max, int m, unsigned pnfft_flags, unsigned fftw_flags, MPI_Comm comm_cart) {
  if (d != 3) { fprintf(stderr, "!!! Error in PNFFT Planer: d != 3 not yet implemented !!!\n"); return nullptr; }
  return reinterpret_cast<PNX(plan)>(CoreT::init(N, n, x_max, m, pnfft_flags, fftw_flags, comm_cart, true));
}
PNX(plan) PNX(init_adv)(int d, const INT *N, unsigned pnfft_flags, unsigned fftw_flags, MPI_Comm comm_cart) {
  INT n[3]; RT x_max[3]; int m;
  adv_defaults(N, n, x_max, &m);
  return PNX(init_guru)(d, N, n, x_max, m, pnfft_flags, fftw_flags, comm_cart);
}
PNX(plan) PNX(init_adv_c2r)(int d, const INT *N, unsigned pnfft_flags, unsigned fftw_flags, MPI_Comm comm_cart) {
  INT n[3]; RT x_max[3]; int m;
  adv_defaults(N, n, x_max, &m);
  return PNX(init_guru_c2r)(d, N, n, x_max, m, pnfft_flags, fftw_flags, comm_cart);
}
PNX(plan) PNX(init_3d)(const INT *N, MPI_Comm comm_cart) { return PNX(init_adv)(3, N, PNFFT_MALLOC_F_HAT, 0, comm_cart); }
PNX(plan) PNX(init_3d_c2r)(const INT *N, MPI_Comm comm_cart) { return PNX(init_adv_c2r)(3, N, PNFFT_MALLOC_F_HAT, 0, comm_cart); }
void PNX(finalize)(PNX(plan) ths, unsigned flags) { CoreT::finalize(AS_PLAN(ths), flags); }

PNX(nodes) PNX(init_nodes)(INT local_M, unsigned malloc_flags) { return reinterpret_cast<PNX(nodes)>(CoreT::init_nodes(local_M, malloc_flags)); }
void PNX(free_nodes)(PNX(nodes) ths, unsigned flags) { CoreT::free_nodes(AS_NODES(ths), flags); }
void PNX(precompute_psi)(PNX(plan) ths, PNX(nodes) nodes, unsigned precompute_flags) { CoreT::precompute_psi(AS_PLAN(ths), AS_NODES(nodes), precompute_flags); }

void PNX(set_f)(CT *f, PNX(nodes) nodes) { AS_NODES(nodes)->f = (RT *)f; }
void PNX(set_grad_f)(CT *grad_f, PNX(nodes) nodes) { AS_NODES(nodes)->grad_f = (RT *)grad_f; }
void PNX(set_hessian_f)(CT *h, PNX(nodes) nodes) { AS_NODES(nodes)->hessian_f = (RT *)h; }
void PNX(set_f_real)(RT *f, PNX(nodes) nodes) { AS_NODES(nodes)->f = f; }
void PNX(set_grad_f_real)(RT *grad_f, PNX(nodes) nodes) { AS_NODES(nodes)->grad_f = grad_f; }
void PNX(set_hessian_f_real)(RT *h, PNX(nodes) nodes) { AS_NODES(nodes)->hessian_f = h; }
void PNX(set_x)(RT *x, PNX(nodes) nodes) {
  // new coordinates: whatever was derived from the old ones (upload, bins) is void
  NodesT *nd = AS_NODES(nodes);
  nd->x = x;
  nd->binned = nd->il.binned = false;
  nd->d_x_bound = nd->il.d_x_bound = nullptr;
  nd->x_uploaded = false;
}
void PNX(set_f_hat)(CT *f_hat, PNX(plan) ths) { AS_PLAN(ths)->f_hat = (PlanT::C *)f_hat; }
void PNX(set_f_hat_real)(RT *f_hat, PNX(plan) ths) { AS_PLAN(ths)->f_hat = (PlanT::C *)f_hat; }
void PNX(set_b)(RT b0, RT b1, RT b2, PNX(plan) ths) {
  // reference api/api-basic.c:587-596: new shape parameters, window tables recomputed
  PlanT *p = AS_PLAN(ths);
  p->b[0] = b0; p->b[1] = b1; p->b[2] = b2;
  p->win_gen++;   // node-table rows cached for unchanged coordinates hold window values of the old shape
  CoreT::upload_window_tables(p);
}
CT *PNX(get_f)(const PNX(nodes) nodes) { return (CT *)AS_NODES(nodes)->f; }
CT *PNX(get_grad_f)(const PNX(nodes) nodes) { return (CT *)AS_NODES(nodes)->grad_f; }
CT *PNX(get_hessian_f)(const PNX(nodes) nodes) { return (CT *)AS_NODES(nodes)->hessian_f; }
RT *PNX(get_f_real)(const PNX(nodes) nodes) { return AS_NODES(nodes)->f; }
RT *PNX(get_grad_f_real)(const PNX(nodes) nodes) { return AS_NODES(nodes)->grad_f; }
RT *PNX(get_hessian_f_real)(const PNX(nodes) nodes) { return AS_NODES(nodes)->hessian_f; }
RT *PNX(get_x)(const PNX(nodes) nodes) { return AS_NODES(nodes)->x; }
CT *PNX(get_f_hat)(const PNX(plan) ths) { return (CT *)AS_PLAN(ths)->f_hat; }
RT *PNX(get_f_hat_real)(const PNX(plan) ths) { return (RT *)AS_PLAN(ths)->f_hat; }
int PNX(get_d)(const PNX(plan)) { return 3; }
int PNX(get_m)(const PNX(plan) ths) { return AS_PLAN(ths)->L.m; }
void PNX(get_x_max)(const PNX(plan) ths, RT *x_max) { for (int t = 0; t < 3; t++) x_max[t] = AS_PLAN(ths)->x_max[t]; }
void PNX(get_N)(const PNX(plan) ths, INT *N) { for (int t = 0; t < 3; t++) N[t] = AS_PLAN(ths)->L.N[t]; }
void PNX(get_n)(const PNX(plan) ths, INT *n) { for (int t = 0; t < 3; t++) n[t] = AS_PLAN(ths)->L.n[t]; }
// The reference's planner tests the plan flags against precompute-namespace constants (api/api-guru.c:150-155:
// PNFFT_PRE_HESSIAN_PSI = 1 << 3 sets PNFFT_PRE_GRAD_PSI = 1 << 2, which sets PNFFT_PRE_PSI = 1 << 1), so a plan made with
// PNFFT_PRE_LIN_PSI reports PNFFT_PRE_CONST_PSI and PNFFT_FAST_GAUSSIAN as well, one made with PNFFT_PRE_CONST_PSI reports
// PNFFT_FAST_GAUSSIAN: what pnfft_get_pnfft_flags and the timer reports show (the plan itself keeps the caller's flags; the
// interpolation order already follows the promotion, Core::init).
static unsigned reference_plan_flags(unsigned fl) {
  if (fl & (1u << 3)) fl |= 1u << 2;
  if (fl & (1u << 2)) fl |= 1u << 1;
  return fl;
}
unsigned PNX(get_pnfft_flags)(const PNX(plan) ths) { return reference_plan_flags(AS_PLAN(ths)->pnfft_flags); }
unsigned PNX(get_pfft_flags)(const PNX(plan) ths) { return AS_PLAN(ths)->pfft_flags; }
void PNX(get_b)(const PNX(plan) ths, RT *b0, RT *b1, RT *b2) { *b0 = AS_PLAN(ths)->b[0]; *b1 = AS_PLAN(ths)->b[1]; *b2 = AS_PLAN(ths)->b[2]; }

void PNX(trafo)(PNX(plan) ths, PNX(nodes) nodes, unsigned compute_flags) { CoreT::trafo(AS_PLAN(ths), AS_NODES(nodes), compute_flags); }
void PNX(adj)(PNX(plan) ths, PNX(nodes) nodes, unsigned compute_flags) { CoreT::adj(AS_PLAN(ths), AS_NODES(nodes), compute_flags); }

void PNX(init)(void) { int f = 0; MPI_Initialized(&f); if (!f) MPI_Init(nullptr, nullptr); }
void PNX(cleanup)(void) {}

void *PNX(malloc)(size_t n) {
  void *p = nullptr;
  if (n == 0) n = 1;
  if (cudaHostAlloc(&p, n, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    fprintf(stderr, "pnfft-b200: page-locked allocation of %zu bytes failed (no CUDA device?)\n", n);
    return nullptr;
  }
  return p;
}
RT *PNX(alloc_real)(size_t n) { return (RT *)PNX(malloc)(sizeof(RT) * n); }
CT *PNX(alloc_complex)(size_t n) { return (CT *)PNX(malloc)(2 * sizeof(RT) * n); }
void PNX(free)(void *p) { if (p) cudaFreeHost(p); }

// PFFT's generator is not available; the formula is the one of reference tests/check_vs_pfft.c:167-181
void PNX(init_f_hat_3d)(const INT *N, const INT *local_N, const INT *local_N_start, unsigned, CT *data) {
  INT m = 0;
  for (INT k0 = local_N_start[0]; k0 < local_N_start[0] + local_N[0]; k0++)
    for (INT k1 = local_N_start[1]; k1 < local_N_start[1] + local_N[1]; k1++)
      for (INT k2 = local_N_start[2]; k2 < local_N_start[2] + local_N[2]; k2++, m++) {
        const INT g = ((k0 + N[0] / 2) * N[1] + (k1 + N[1] / 2)) * N[2] + (k2 + N[2] / 2);
        data[m][0] = (RT)(1000.0 / (double)(2 * g + 1));
        data[m][1] = (RT)(1000.0 / (double)(2 * g + 2));
      }
}
void PNX(init_f)(INT local_M, CT *data) {
  for (INT j = 0; j < local_M; j++) {
    data[j][0] = (RT)100.0 * (RT)rand() / (RT)RAND_MAX;
    data[j][1] = (RT)100.0 * (RT)rand() / (RT)RAND_MAX;
  }
}
void PNX(init_x_3d_adv)(const RT *lo, const RT *up, const RT *x_max, INT loc_M, RT *x) {
  for (int t = 0; t < 3; t++) if (lo[t] < -x_max[t] || x_max[t] < up[t]) return;
  for (INT j = 0; j < loc_M; j++)
    for (int t = 0; t < 3; t++) {
      RT tmp;
      do {
        RT r;
        do { r = (RT)rand() / (RT)RAND_MAX; } while (r >= (RT)1.0);
        tmp = (up[t] - lo[t]) * r + lo[t];
      } while (tmp < -x_max[t] || x_max[t] <= tmp);
      x[3 * j + t] = tmp;
    }
}
void PNX(init_x_3d)(const RT *lo, const RT *up, INT loc_M, RT *x) {
  const RT x_max[3] = {(RT)0.5, (RT)0.5, (RT)0.5};
  PNX(init_x_3d_adv)(lo, up, x_max, loc_M, x);
}
void PNX(zero_f_hat)(PNX(plan) ths) {
  PlanT *p = AS_PLAN(ths);
  if (!p || !p->f_hat) return;
  const size_t bytes = sizeof(PlanT::C) * (size_t)CoreT::local_N_total(p);
  if (pnb::is_device_ptr(p->f_hat)) cudaMemset(p->f_hat, 0, bytes); else memset(p->f_hat, 0, bytes);
}

RT PNX(inv_phi_hat)(const PNX(plan) ths, int dim, INT k) {
  const PlanT *p = AS_PLAN(ths);
  return pnb::phi_hat_any<RT>(pnb::window_hat_kind(p->pnfft_flags), (long)k, (long)p->L.n[dim], p->b[dim], p->L.m, true);
}
RT PNX(phi_hat)(const PNX(plan) ths, int dim, INT k) {
  const PlanT *p = AS_PLAN(ths);
  return pnb::phi_hat_any<RT>(pnb::window_hat_kind(p->pnfft_flags), (long)k, (long)p->L.n[dim], p->b[dim], p->L.m, false);
}
// psi(x), dpsi(x), ddpsi(x): the window and its derivatives at offset x (reference kernel/ndft-parallel.c:2288-2336), by the
// formulas the kernels, the Hessian path and the PNFFT_PRE_*_PSI tables use (window.h).  window_tap takes y = l - n x.
static RT window_at(int kind, int which, RT n, RT b, int m, RT x) {
  RT psi = 0, d = 0;
  if (which == 0) {
    if (kind == pnb::WIN_BSPLINE) return pnb::bspline<RT>(2 * m, n * x + (RT)m);
    pnb::window_tap<RT>(kind, n * x, n, b, m, false, &psi, &d);
    return psi;
  }
  if (which == 1) {
    if (kind == pnb::WIN_BSPLINE) return n * (pnb::bspline<RT>(2 * m - 1, n * x + (RT)m) - pnb::bspline<RT>(2 * m - 1, n * x + (RT)m - (RT)1));
    pnb::window_tap<RT>(kind, -(n * x), n, b, m, true, &psi, &d);
    return d;
  }
  if (kind != pnb::WIN_BSPLINE) pnb::window_tap<RT>(kind, -(n * x), n, b, m, true, &psi, &d);
  return pnb::window_ddtap<RT>(kind, -(n * x), n, b, m, psi, d);
}
RT PNX(psi)(const PNX(plan) ths, int dim, RT x) {
  const PlanT *p = AS_PLAN(ths);
  return window_at(p->kind, 0, (RT)p->L.n[dim], p->b[dim], p->L.m, x);
}
RT PNX(dpsi)(const PNX(plan) ths, int dim, RT x) {
  const PlanT *p = AS_PLAN(ths);
  return window_at(p->kind, 1, (RT)p->L.n[dim], p->b[dim], p->L.m, x);
}
RT PNX(ddpsi)(const PNX(plan) ths, int dim, RT x) {
  const PlanT *p = AS_PLAN(ths);
  return window_at(p->kind, 2, (RT)p->L.n[dim], p->b[dim], p->L.m, x);
}

// every rank prints its own vector in turn (reference api/api-basic.c:705-780: formats, "Rank r, name" header, nothing for N < 1)
void PNX(vpr_complex)(CT *data, INT N, const char *name, MPI_Comm comm) {
  if (N < 1) return;
  int rank = 0, size = 1;
  MPI_Comm_size(comm, &size); MPI_Comm_rank(comm, &rank);
  fflush(stdout);
  MPI_Barrier(comm);
  for (int t = 0; t < size; t++) {
    if (t == rank) {
      printf("\nRank %d, %s", rank, name);
      for (INT k = 0; k < N; k++) { if (k % 4 == 0) printf("\n%4td.", k / 4); printf(" %.2e+%.2ei,", (double)data[k][0], (double)data[k][1]); }
      printf("\n");
      fflush(stdout);
    }
    MPI_Barrier(comm);
  }
}
void PNX(vpr_real)(RT *data, INT N, const char *name, MPI_Comm comm) {
  if (N < 1) return;
  int rank = 0, size = 1;
  MPI_Comm_size(comm, &size); MPI_Comm_rank(comm, &rank);
  fflush(stdout);
  MPI_Barrier(comm);
  for (int t = 0; t < size; t++) {
    if (t == rank) {
      printf("\nRank %d, %s", rank, name);
      for (INT k = 0; k < N; k++) { if (k % 8 == 0) printf("\n%4td.", k / 8); printf(" %e,", (double)data[k]); }
      printf("\n");
      fflush(stdout);
    }
    MPI_Barrier(comm);
  }
}

double *PNX(get_timer_trafo)(PNX(plan) ths) { return PNX(timer_copy)(AS_PLAN(ths)->timer_trafo); }
double *PNX(get_timer_adj)(PNX(plan) ths) { return PNX(timer_copy)(AS_PLAN(ths)->timer_adj); }
// reference kernel/timer.c:57-66: the slots after the iteration count are divided by it; the count itself stays
void PNX(timer_average)(double *t) { if (t[0] < 1.0) return; for (int i = 1; i < PNFFT_TIMER_LENGTH; i++) t[i] /= t[0]; }
double *PNX(timer_copy)(const double *orig) {
  double *c = (double *)malloc(sizeof(double) * PNFFT_TIMER_LENGTH);
  for (int i = 0; i < PNFFT_TIMER_LENGTH; i++) c[i] = orig[i];
  return c;
}
double *PNX(timer_reduce_max)(MPI_Comm comm, double *timer) {
  double *r = (double *)malloc(sizeof(double) * PNFFT_TIMER_LENGTH);
  MPI_Reduce(timer, r, PNFFT_TIMER_LENGTH, MPI_DOUBLE, MPI_MAX, 0, comm);
  return r;
}
double *PNX(timer_add)(const double *a, const double *b) {
  double *c = (double *)malloc(sizeof(double) * PNFFT_TIMER_LENGTH);
  for (int i = 0; i < PNFFT_TIMER_LENGTH; i++) c[i] = a[i] + b[i];
  return c;
}
void PNX(timer_free)(double *t) { free(t); }
void PNX(reset_timer)(PNX(plan) ths) {
  memset(AS_PLAN(ths)->timer_trafo, 0, sizeof(double) * PNFFT_TIMER_LENGTH);
  memset(AS_PLAN(ths)->timer_adj, 0, sizeof(double) * PNFFT_TIMER_LENGTH);
}
// ---- timer reports in the reference's Octave-readable format (kernel/timer.c:139-367): rank 0 writes a legend, one line
// with the plan's flags and sizes, and the averaged maxima over the ranks under the prefixes pnfft_trf / pnfft_adj, all
// indexed by log2(procs) + 1.  (The reference appends PFFT's and the ghost-cell plan's own reports in the _adv variants;
// there is no PFFT here.)
struct TimerReport {
  unsigned flags; INT N[3], n[3]; int m, np[3];
  const double *trafo, *adj;
};
static void timer_report_legend(FILE *f) {
  fprintf(f, "%% N  - NFFT size\n%% n  - FFT size\n%% np - process grid\n%% procs - number of processes\n"
             "%% pnfft - PNFFT runtime\n%% pfft  - PFFT runtime\n%% index(i) = log(procs(i)) + 1\n");
}
static void timer_report_run(FILE *f, const TimerReport &r, int size, int idx) {
  const unsigned fl = reference_plan_flags(r.flags);
  fprintf(f, "%% pnfft_flags == %s", (fl & PNFFT_WINDOW_GAUSSIAN) ? "PNFFT_WINDOW_GAUSSIAN" : (fl & PNFFT_WINDOW_BSPLINE) ? "PNFFT_WINDOW_BSPLINE" :
          (fl & PNFFT_WINDOW_SINC_POWER) ? "PNFFT_WINDOW_SINC_POWER" : (fl & PNFFT_WINDOW_BESSEL_I0) ? "PNFFT_WINDOW_BESSEL_I0" : "PNFFT_WINDOW_KAISER_BESSEL");
  const struct { unsigned bit; const char *on, *off; } tab[] = {
    {PNFFT_PRE_PHI_HAT, " | PNFFT_PRE_PHI_HAT", ""}, {PNFFT_FAST_GAUSSIAN, " | PNFFT_FAST_GAUSSIAN", ""},
    {PNFFT_PRE_CONST_PSI, " | PNFFT_PRE_CONST_PSI", ""}, {PNFFT_PRE_LIN_PSI, " | PNFFT_PRE_LIN_PSI", ""},
    {PNFFT_PRE_QUAD_PSI, " | PNFFT_PRE_QUAD_PSI", ""}, {PNFFT_PRE_CUB_PSI, " | PNFFT_PRE_CUB_PSI", ""},
    {PNFFT_FFT_IN_PLACE, " | PNFFT_FFT_IN_PLACE", " | PNFFT_FFT_OUT_OF_PLACE"}, {PNFFT_SORT_NODES, " | PNFFT_SORT_NODES", ""},
    {PNFFT_INTERLACED, " | PNFFT_INTERLACED", ""}, {PNFFT_SHIFTED_F_HAT, " | PNFFT_SHIFTED_F_HAT", ""},
    {PNFFT_SHIFTED_X, " | PNFFT_SHIFTED_X", ""}, {PNFFT_TRANSPOSED_F_HAT, " | PNFFT_TRANSPOSED_F_HAT", ""},
    {PNFFT_DIFF_IK, " | PNFFT_DIFF_IK", " | PNFFT_DIFF_AD"}, {PNFFT_REAL_F, " | PNFFT_REAL_F", ""},
    {PNFFT_MALLOC_F_HAT, " | PNFFT_MALLOC_F_HAT", ""}};
  for (const auto &t : tab) fputs((fl & t.bit) ? t.on : t.off, f);
  fprintf(f, "\nindex(%d) = %d;  procs(%d) = %d;  np_pnfft(%d, 1:3) = [%d %d %d];  ", idx, idx, idx, size, idx, r.np[0], r.np[1], r.np[2]);
  fprintf(f, "N_pnfft(%d, 1:3) = [%td %td %td ];  n_pnfft(%d, 1:3) = [%td %td %td ];  m_pnfft(%d) = %d;\n", idx, r.N[0], r.N[1], r.N[2],
          idx, r.n[0], r.n[1], r.n[2], idx, r.m);
}
// one direction: maxima over the ranks, averaged over the iterations; the iteration count printed is the caller's own
static void timer_report_block(FILE *f, MPI_Comm comm, const char *prefix, const double *timer, bool adv, int idx) {
  double *mt = PNX(timer_reduce_max)(comm, const_cast<double *>(timer));
  PNX(timer_average)(mt);
  if (f) {
    if (!adv) {
      fprintf(f, "%s_iter(%d)    = %d;  %s(%d)   = %.3e;\n", prefix, idx, (int)timer[PNFFT_TIMER_ITER], prefix, idx, mt[PNFFT_TIMER_WHOLE]);
    } else {
      fprintf(f, "%s_matrix_D(%d)   = %.3e;  %s_matrix_F(%d)   = %.3e;\n", prefix, idx, mt[PNFFT_TIMER_MATRIX_D], prefix, idx, mt[PNFFT_TIMER_MATRIX_F]);
      fprintf(f, "%s_matrix_B(%d)   = %.3e;  %s_gcells(%d)     = %.3e;\n", prefix, idx, mt[PNFFT_TIMER_MATRIX_B], prefix, idx, mt[PNFFT_TIMER_GCELLS]);
      fprintf(f, "%s_sort_nodes(%d) = %.3e;  %s_loop_B(%d)     = %.3e;\n", prefix, idx, mt[PNFFT_TIMER_SORT_NODES], prefix, idx, mt[PNFFT_TIMER_LOOP_B]);
      fprintf(f, "%s_shift_in(%d)   = %.3e;  %s_shift_out(%d)  = %.3e;\n", prefix, idx, mt[PNFFT_TIMER_SHIFT_INPUT], prefix, idx, mt[PNFFT_TIMER_SHIFT_OUTPUT]);
    }
  }
  free(mt);
}
// f is the destination on rank 0 and NULL elsewhere (the other ranks only take part in the reductions)
static void timer_report(FILE *f, MPI_Comm comm, const TimerReport &r, bool legend, bool basic, bool adv) {
  int size = 1;
  MPI_Comm_size(comm, &size);
  const int idx = (int)lrint(round(log((double)size) / log(2.0))) + 1;
  if (f && legend) timer_report_legend(f);
  if (f && basic) timer_report_run(f, r, size, idx);
  if (basic) { timer_report_block(f, comm, "pnfft_trf", r.trafo, false, idx); timer_report_block(f, comm, "pnfft_adj", r.adj, false, idx); }
  if (adv) { timer_report_block(f, comm, "pnfft_trf", r.trafo, true, idx); timer_report_block(f, comm, "pnfft_adj", r.adj, true, idx); }
  if (f) fflush(f);
}
static TimerReport timer_report_of(const PlanT *p) {
  TimerReport r;
  r.flags = p->pnfft_flags; r.m = p->L.m;
  for (int t = 0; t < 3; t++) { r.N[t] = p->L.N[t]; r.n[t] = p->L.n[t]; }
  r.np[0] = p->mesh.np[0]; r.np[1] = p->mesh.np[1]; r.np[2] = 1;
  r.trafo = p->timer_trafo; r.adj = p->timer_adj;
  return r;
}
static FILE *timer_report_file(const char *name, MPI_Comm comm, bool *is_new) {
  int rank = 0;
  MPI_Comm_rank(comm, &rank);
  if (rank) return nullptr;
  FILE *probe = fopen(name, "r");
  *is_new = probe == nullptr;
  if (probe) fclose(probe);
  FILE *f = fopen(name, "a+");
  if (!f) { fprintf(stderr, "Error: Cannot open file %s.\n", name); exit(1); }
  return f;
}
static void timer_report_to(const TimerReport &r, const char *name, MPI_Comm comm, bool adv) {
  int rank = 0;
  MPI_Comm_rank(comm, &rank);
  if (!name) { timer_report(rank ? nullptr : stdout, comm, r, true, true, adv); return; }
  bool is_new = false;
  FILE *f = timer_report_file(name, comm, &is_new);
  timer_report(f, comm, r, is_new, true, adv);
  if (f) fclose(f);
}
void PNX(print_average_timer)(const PNX(plan) ths, MPI_Comm comm) { timer_report_to(timer_report_of(AS_PLAN(ths)), nullptr, comm, false); }
void PNX(print_average_timer_adv)(const PNX(plan) ths, MPI_Comm comm) { timer_report_to(timer_report_of(AS_PLAN(ths)), nullptr, comm, true); }
void PNX(write_average_timer)(const PNX(plan) ths, const char *name, MPI_Comm comm) { timer_report_to(timer_report_of(AS_PLAN(ths)), name, comm, false); }
void PNX(write_average_timer_adv)(const PNX(plan) ths, const char *name, MPI_Comm comm) { timer_report_to(timer_report_of(AS_PLAN(ths)), name, comm, true); }
// host-only: the report pnfft_write_average_timer(_adv) would append to `name` for a plan with these flags, sizes, process mesh
// and timer arrays (no plan, no GPU; for the CPU test suite)
void PNX(b200_timer_report_host)(const char *name, unsigned pnfft_flags, const INT *N, const INT *n, int m, const int *np3,
                                 const double *timer_trafo, const double *timer_adj, int adv, MPI_Comm comm) {
  TimerReport r;
  r.flags = pnfft_flags; r.m = m;
  for (int t = 0; t < 3; t++) { r.N[t] = N[t]; r.n[t] = n[t]; r.np[t] = np3[t]; }
  r.trafo = timer_trafo; r.adj = timer_adj;
  timer_report_to(r, name, comm, adv != 0);
}

// every rank in turn prints its block; k_t = local_N_start[t] + i_t (reference api/pnfft.h:233-238)
static void apr_block(const RT *data, int ncomp, const INT *local_N, const INT *local_N_start, const char *name, MPI_Comm comm) {
  int rank = 0, size = 1;
  MPI_Comm_rank(comm, &rank); MPI_Comm_size(comm, &size);
  for (int r = 0; r < size; r++) {
    if (r == rank) {
      printf("rank %d: %s", rank, name);
      INT l = 0;
      for (INT k0 = 0; k0 < local_N[0]; k0++)
        for (INT k1 = 0; k1 < local_N[1]; k1++) {
          for (INT k2 = 0; k2 < local_N[2]; k2++, l++) {
            if (ncomp == 2) printf("  [%td,%td,%td] %.4e%+.4ei", k0 + local_N_start[0], k1 + local_N_start[1], k2 + local_N_start[2], (double)data[2 * l], (double)data[2 * l + 1]);
            else printf("  [%td,%td,%td] %.4e", k0 + local_N_start[0], k1 + local_N_start[1], k2 + local_N_start[2], (double)data[l]);
          }
          printf("\n");
        }
      fflush(stdout);
    }
    MPI_Barrier(comm);
  }
}
// with PNFFT_TRANSPOSED_F_HAT the block is stored (k1, k2, k0): extents and starts are rotated into memory order before the
// walk (reference apr_3d, api/api-basic.c:783-800; the element format is PFFT's there, which is not available)
static void apr_rotated(const RT *data, int ncomp, const INT *local_N, const INT *local_N_start, unsigned pnfft_flags, const char *name, MPI_Comm comm) {
  const int shift = (pnfft_flags & PNFFT_TRANSPOSED_F_HAT) ? 1 : 0;
  INT lN[3], lNs[3];
  for (int t = 0; t < 3; t++) { lN[t] = local_N[(t + shift) % 3]; lNs[t] = local_N_start[(t + shift) % 3]; }
  apr_block(data, ncomp, lN, lNs, name, comm);
}
void PNX(apr_complex_3d)(CT *data, INT *local_N, INT *local_N_start, unsigned pnfft_flags, const char *name, MPI_Comm comm) {
  apr_rotated((const RT *)data, 2, local_N, local_N_start, pnfft_flags, name, comm);
}
void PNX(apr_real_3d)(RT *data, INT *local_N, INT *local_N_start, unsigned pnfft_flags, const char *name, MPI_Comm comm) {
  apr_rotated(data, 1, local_N, local_N_start, pnfft_flags, name, comm);
}
void PNX(get_args)(int argc, char **argv, const char *name, int neededArgs, unsigned type, void *parameter) {
  pfft_get_args(argc, argv, name, neededArgs, type, parameter);
}
// defaults and option names of the reference's test drivers (api/api-basic.c:820-938)
void PNX(check_init_parameters)(int argc, char **argv, INT *N, INT *n, INT *local_M, int *m, unsigned *pnfft_flags,
                                unsigned *compute_flags, double *x_max, int *np, int *compare_direct, int *debug) {
  int window = 4, fast_gaussian = 0, intpol = -1, interlaced = 0, diff_ik = 0, tr_f_hat = 0;
  int cf = 1, cg = 1, ch = 1;      // the reference's defaults (api/api-basic.c:834-836)
  N[0] = N[1] = N[2] = 16; n[0] = n[1] = n[2] = 0; *local_M = 0; *m = 6;
  x_max[0] = x_max[1] = x_max[2] = 0.5;
  np[0] = np[1] = np[2] = 2;
  struct { const char *name; int cnt; unsigned type; void *dst; } opt[] = {
    {"-pnfft_local_M", 1, PFFT_PTRDIFF_T, local_M}, {"-pnfft_N", 3, PFFT_PTRDIFF_T, N}, {"-pnfft_n", 3, PFFT_PTRDIFF_T, n},
    {"-pnfft_np", 3, PFFT_INT, np}, {"-pnfft_m", 1, PFFT_INT, m}, {"-pnfft_window", 1, PFFT_INT, &window},
    {"-pnfft_fast_gaussian", 1, PFFT_INT, &fast_gaussian}, {"-pnfft_intpol", 1, PFFT_INT, &intpol},
    {"-pnfft_interlaced", 1, PFFT_INT, &interlaced}, {"-pnfft_diff_ik", 1, PFFT_INT, &diff_ik},
    {"-pnfft_tr_f_hat", 1, PFFT_INT, &tr_f_hat}, {"-pnfft_x_max", 3, PFFT_DOUBLE, x_max}, {"-pnfft_debug", 1, PFFT_INT, debug},
    {"-pnfft_compare_direct", 1, PFFT_INT, compare_direct}, {"-pnfft_compute_f", 1, PFFT_INT, &cf},
    {"-pnfft_compute_grad_f", 1, PFFT_INT, &cg}, {"-pnfft_compute_hessian_f", 1, PFFT_INT, &ch}};
  for (auto &o : opt) pfft_get_args(argc, argv, o.name, o.cnt, o.type, o.dst);
  if (*local_M == 0) *local_M = N[0] * N[1] * N[2] / ((INT)np[0] * np[1] * np[2]);
  for (int t = 0; t < 3; t++) if (n[t] == 0) n[t] = 2 * N[t];
  static const unsigned win_flag[] = {PNFFT_WINDOW_GAUSSIAN, PNFFT_WINDOW_BSPLINE, PNFFT_WINDOW_SINC_POWER, PNFFT_WINDOW_BESSEL_I0,
                                      PNFFT_WINDOW_KAISER_BESSEL, PNFFT_WINDOW_GAUSSIAN_T};
  static const unsigned ip_flag[] = {PNFFT_PRE_CONST_PSI, PNFFT_PRE_LIN_PSI, PNFFT_PRE_QUAD_PSI, PNFFT_PRE_CUB_PSI};
  const unsigned wf = (window >= 0 && window <= 5) ? win_flag[window] : PNFFT_WINDOW_KAISER_BESSEL;
  const unsigned ipf = (intpol >= 0 && intpol <= 3) ? ip_flag[intpol] : 0u;
  *compute_flags = (cf ? PNFFT_COMPUTE_F : 0u) | (cg ? PNFFT_COMPUTE_GRAD_F : 0u) | (ch ? PNFFT_COMPUTE_HESSIAN_F : 0u);
  *pnfft_flags = wf | (fast_gaussian ? PNFFT_FAST_GAUSSIAN : 0u) | ipf | (diff_ik ? PNFFT_DIFF_IK : PNFFT_DIFF_AD) |
                 (tr_f_hat ? PNFFT_TRANSPOSED_F_HAT : 0u) | (interlaced ? PNFFT_INTERLACED : 0u);
  pfft_printf(MPI_COMM_WORLD, "pnfft-b200 test set-up: N = %td x %td x %td (-pnfft_N), n = %td x %td x %td (-pnfft_n), local_M = %td "
              "(-pnfft_local_M), m = %d (-pnfft_m), window = %d (-pnfft_window), fast_gaussian = %d, intpol = %d, interlaced = %d, "
              "diff_ik = %d, tr_f_hat = %d, compute f/grad/hessian = %d/%d/%d, compare_direct = %d, np = %d x %d x %d (-pnfft_np)\n",
              N[0], N[1], N[2], n[0], n[1], n[2], *local_M, *m, window, fast_gaussian, intpol, interlaced, diff_ik, tr_f_hat,
              cf, cg, ch, *compare_direct, np[0], np[1], np[2]);
}

// ---------------------------------------------------------------------------------------------
// extensions
// ---------------------------------------------------------------------------------------------
void PNX(b200_get_local_no)(const PNX(plan) ths, INT *local_no, INT *local_no_start, INT *no) {
  const PlanT *p = AS_PLAN(ths);
  for (int t = 0; t < 3; t++) { local_no[t] = p->L.local_no[t]; local_no_start[t] = p->L.local_no_start[t]; no[t] = p->L.no[t]; }
}

static void grid_io(PlanT *p, RT *compact, bool set) {
  const pnb::Layout &L = p->L;
  const int nc = L.c2r ? 1 : 2;
  const size_t cnt = (size_t)L.local_no[0] * L.local_no[1] * L.local_no[2] * nc;
  RT *d = nullptr;
  const bool dev = pnb::is_device_ptr(compact);
  if (dev) d = compact;
  else {
    PNB_CUDA(cudaMalloc((void **)&d, sizeof(RT) * (cnt ? cnt : 1)));
    if (set) PNB_CUDA(cudaMemcpyAsync(d, compact, sizeof(RT) * cnt, cudaMemcpyHostToDevice, p->stream));
  }
  pnb::BoxMap bm = pnb::dense_map(L.local_no[0], L.local_no[1], L.local_no[2], L.ngc[1], L.pitch2, L.gcb[0], L.gcb[1], L.gcb[2], 0);
  if (L.c2r) pnb::box_copy<RT>(p->stream, (RT *)p->d_grid, d, bm, set ? pnb::BOX_C2A : pnb::BOX_A2C, false, nullptr);
  else pnb::box_copy<PlanT::C>(p->stream, (PlanT::C *)p->d_grid, (PlanT::C *)d, bm, set ? pnb::BOX_C2A : pnb::BOX_A2C, false, nullptr);
  if (!dev) {
    if (!set) PNB_CUDA(cudaMemcpyAsync(compact, d, sizeof(RT) * cnt, cudaMemcpyDeviceToHost, p->stream));
    PNB_CUDA(cudaStreamSynchronize(p->stream));
    cudaFree(d);
  } else PNB_CUDA(cudaStreamSynchronize(p->stream));
}
void PNX(b200_set_grid)(PNX(plan) ths, const RT *compact_grid) { grid_io(AS_PLAN(ths), const_cast<RT *>(compact_grid), true); }
void PNX(b200_get_grid)(PNX(plan) ths, RT *compact_grid) { grid_io(AS_PLAN(ths), compact_grid, false); }

void PNX(b200_set_g1)(PNX(plan) ths, const CT *g1) {
  PlanT *p = AS_PLAN(ths);
  PNB_CUDA(cudaMemcpy(p->d_g1, g1, sizeof(PlanT::C) * (size_t)CoreT::local_N_total(p), cudaMemcpyDefault));
}
void PNX(b200_get_g1)(PNX(plan) ths, CT *g1) {
  PlanT *p = AS_PLAN(ths);
  PNB_CUDA(cudaMemcpy(g1, p->d_g1, sizeof(PlanT::C) * (size_t)CoreT::local_N_total(p), cudaMemcpyDefault));
}

void PNX(b200_node_grid_index)(PNX(plan) ths, PNX(nodes) nodes, INT *u_and_m0) {
  PlanT *p = AS_PLAN(ths); NodesT *nd = AS_NODES(nodes);
  const size_t M = (size_t)nd->local_M;
  if (!M) return;
  const RT *dx = CoreT::dev_in(p, nd->x, &nd->d_x, &nd->cap_x, 3 * M, true);
  long long *d = nullptr;
  PNB_CUDA(cudaMalloc((void **)&d, sizeof(long long) * 4 * M));
  pnb::k_node_grid_index<RT><<<(unsigned)((M + 255) / 256), 256, 0, p->stream>>>(CoreT::geom(p), dx, (int)M, d);
  PNB_CUDA(cudaMemcpyAsync(u_and_m0, d, sizeof(long long) * 4 * M, cudaMemcpyDeviceToHost, p->stream));
  PNB_CUDA(cudaStreamSynchronize(p->stream));
  cudaFree(d);
}

void PNX(b200_sort_nodes)(PNX(plan) ths, PNX(nodes) nodes, INT *keys, INT *perm) {
  PlanT *p = AS_PLAN(ths); NodesT *nd = AS_NODES(nodes);
  const size_t M = (size_t)nd->local_M;
  if (!M) return;
  const RT *dx = CoreT::dev_in(p, nd->x, &nd->d_x, &nd->cap_x, 3 * M, true);
  unsigned long long *k_in = nullptr, *k_out = nullptr;
  int *i_in = nullptr, *i_out = nullptr;
  PNB_CUDA(cudaMalloc((void **)&k_in, 8 * M)); PNB_CUDA(cudaMalloc((void **)&k_out, 8 * M));
  PNB_CUDA(cudaMalloc((void **)&i_in, 4 * M)); PNB_CUDA(cudaMalloc((void **)&i_out, 4 * M));
  pnb::k_sort_keys<RT><<<(unsigned)((M + 255) / 256), 256, 0, p->stream>>>(CoreT::geom(p), p->L.n[0], p->L.n[1], p->L.n[2], dx, (int)M, k_in, i_in);
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, i_in, i_out, (int)M, 0, 64, p->stream);
  void *d_tmp = nullptr;
  PNB_CUDA(cudaMalloc(&d_tmp, tmp ? tmp : 1));
  // LSD radix sort is stable: equal keys keep their original order, like the reference's radix_lsdf (util/util.c:206-277)
  cub::DeviceRadixSort::SortPairs(d_tmp, tmp, k_in, k_out, i_in, i_out, (int)M, 0, 64, p->stream);
  std::vector<unsigned long long> hk(M);
  std::vector<int> hi(M);
  PNB_CUDA(cudaMemcpyAsync(hk.data(), k_out, 8 * M, cudaMemcpyDeviceToHost, p->stream));
  PNB_CUDA(cudaMemcpyAsync(hi.data(), i_out, 4 * M, cudaMemcpyDeviceToHost, p->stream));
  PNB_CUDA(cudaStreamSynchronize(p->stream));
  for (size_t i = 0; i < M; i++) { keys[i] = (INT)hk[i]; perm[i] = (INT)hi[i]; }
  cudaFree(k_in); cudaFree(k_out); cudaFree(i_in); cudaFree(i_out); cudaFree(d_tmp);
}

void PNX(b200_window_tensor)(PNX(plan) ths, PNX(nodes) nodes, RT *psi, RT *dpsi) {
  PlanT *p = AS_PLAN(ths); NodesT *nd = AS_NODES(nodes);
  const size_t M = (size_t)nd->local_M;
  if (!M) return;
  const size_t cnt = M * 3 * (size_t)p->L.cutoff;
  const RT *dx = CoreT::dev_in(p, nd->x, &nd->d_x, &nd->cap_x, 3 * M, true);
  RT *d_psi = nullptr, *d_dpsi = nullptr;
  PNB_CUDA(cudaMalloc((void **)&d_psi, sizeof(RT) * cnt));
  if (dpsi) PNB_CUDA(cudaMalloc((void **)&d_dpsi, sizeof(RT) * cnt));
  pnb::k_window_tensor<RT><<<(unsigned)((M * 32 + 255) / 256), 256, 0, p->stream>>>(CoreT::geom(p), dx, (int)M, d_psi, d_dpsi);
  PNB_CUDA(cudaMemcpyAsync(psi, d_psi, sizeof(RT) * cnt, cudaMemcpyDeviceToHost, p->stream));
  if (dpsi) PNB_CUDA(cudaMemcpyAsync(dpsi, d_dpsi, sizeof(RT) * cnt, cudaMemcpyDeviceToHost, p->stream));
  PNB_CUDA(cudaStreamSynchronize(p->stream));
  cudaFree(d_psi); cudaFree(d_dpsi);
}

// variant bit 0: generic global-memory kernels; bit 1: exact window evaluation instead of the fitted polynomials;
// 8: z-march v1 (CTA-synchronous) instead of the warp-autonomous v2
void PNX(b200_set_kernel_variant)(PNX(plan) ths, int variant) { AS_PLAN(ths)->kernel_variant = variant & 13; AS_PLAN(ths)->use_poly = (variant & 2) ? 0 : 1; }
// "the node coordinates will not change until I call pnfft_set_x again": the upload of x and its binning are reused by
// every following pnfft_trafo / pnfft_adj on these nodes (the reference re-reads x every call, api/api-basic.c:199-244)
void PNX(b200_nodes_x_static)(PNX(nodes) nodes, int on) {
  NodesT *nd = AS_NODES(nodes);
  nd->x_static = on != 0;
  if (!on) { nd->binned = nd->il.binned = false; nd->x_uploaded = false; }
}
int PNX(b200_get_poly_degree)(PNX(plan) ths) { return AS_PLAN(ths)->poly_deg; }
#ifdef ZM2_TIMING
// development aid: per-warp cycle accounting of k_gather_zm2 (build with -DZM2_TIMING)
void PNX(b200_gather_timing)(long long *out96, int reset) {
  static long long *d = nullptr;
  if (!d) { cudaMalloc(&d, 112 * 8); cudaMemset(d, 0, 112 * 8); cudaMemcpyToSymbol(pnb::g_zm2_timing, &d, sizeof(d)); }
  cudaDeviceSynchronize();
  if (out96) cudaMemcpy(out96, d, 112 * 8, cudaMemcpyDeviceToHost);
  if (reset) { cudaMemset(d, 0, 96 * 8); cudaMemset(d + 96, 0x3f, 16 * 8); }
}
#endif
// Host evaluation of the Kaiser-Bessel taps exactly as k_node_table2 computes them in double (csrc/window.h: kb_tap_fast,
// small-argument taps through window_tap): psi / dpsi are [M][3][2m+1].  Lets the CPU tests pin the one-exponential
// arithmetic against the reference's tensors without a GPU.
void PNX(b200_kb_taps_host)(const double *x, INT M, const INT *n, const double *b, int m, double *psi, double *dpsi) {
  const int c = 2 * m + 1;
  for (INT j = 0; j < M; j++)
    for (int t = 0; t < 3; t++) {
      const double nxv = (double)n[t] * x[3 * j + t], flv = floor(nxv);
      for (int s = 0; s < c; s++) {
        const double y = flv - nxv - (double)m + (double)s;
        double a = 0, d = 0;
        if (!pnb::kb_tap_fast(y, (double)n[t], b[t], m, true, &a, &d))
          pnb::window_tap<double>(pnb::WIN_KAISER_BESSEL, y, (double)n[t], b[t], m, true, &a, &d);
        psi[((size_t)j * 3 + t) * c + s] = a;
        if (dpsi) dpsi[((size_t)j * 3 + t) * c + s] = d;
      }
    }
}
// ... and of the window itself: out[i] = psi (which = 0), dpsi (1) or ddpsi (2) at offset x[i], as pnfft_psi / pnfft_dpsi /
// pnfft_ddpsi of a plan with these flags, sizes and shape parameter return them (b <= 0: the window's default shape)
void PNX(b200_psi_host)(unsigned pnfft_flags, INT N, INT n, RT b, int m, int which, const RT *x, INT len, RT *out) {
  const int kind = pnb::window_kind(pnfft_flags);
  const RT bb = b > 0 ? b : pnb::window_shape<RT>(kind, m, (RT)n / (RT)N);
  for (INT i = 0; i < len; i++) out[i] = window_at(kind, which, (RT)n, bb, m, x[i]);
}
// Host evaluation of the window's Fourier coefficients exactly as the D tables (Core::upload_window_tables) and
// pnfft_phi_hat / pnfft_inv_phi_hat compute them: out[i] = phi_hat(k[i]) (inverse = 0) or 1 / phi_hat(k[i]) for the window
// the plan flags select, oversampled size n, shape parameter b (<= 0: the default of that window at sigma = n / N).
void PNX(b200_phi_hat_host)(unsigned pnfft_flags, INT N, INT n, RT b, int m, const INT *k, INT len, int inverse, RT *out) {
  const int kind = pnb::window_kind(pnfft_flags), hat = pnb::window_hat_kind(pnfft_flags);
  const RT bb = b > 0 ? b : pnb::window_shape<RT>(kind, m, (RT)n / (RT)N);
  for (INT i = 0; i < len; i++) out[i] = pnb::phi_hat_any<RT>(hat, (long)k[i], (long)n, bb, m, inverse != 0);
}
// Host-only: the f_hat block the direct NDFT (csrc/direct.cuh) broadcasts for / reduces to rank pid of a p0 x p1 mesh, in
// memory order: out = { len[3], start[3], axis[3], N_of_axis[3] }.  For the CPU test suite (against the golden layouts).
void PNX(b200_direct_block)(const INT *N, const INT *n, int m, int p0, int p1, int pid, unsigned pnfft_flags, int c2r, int *out12) {
  pnb::Mesh M;
  M.np[0] = p0; M.np[1] = p1; M.size = p0 * p1;
  const RT xm[3] = {(RT)0.5, (RT)0.5, (RT)0.5};
  const pnb::DirectBlock B = pnb::Direct<RT>::block_of_mesh(M, N, n, xm, m, c2r != 0, pnfft_flags, pid);
  for (int q = 0; q < 3; q++) { out12[q] = B.len[q]; out12[3 + q] = B.start[q]; out12[6 + q] = B.axis[q]; out12[9 + q] = B.Nax[q]; }
}
// Host-only check of the pencil FFT's composed self maps (fftpipe.cuh: compose_self_map) for one rank of a p0 x p1 mesh:
// every re-distribution stage is emulated on index arrays, once through pack -> chunk -> unpack and once through the
// composed map, forward and backward.  Returns the number of self transfers checked, -1 on a mismatch, -2 if a self
// transfer did not compose.  Needs no GPU (tests/test_abi.py).
// Host evaluation of the work-item cutter of the gridding kernels (zmarch2.cuh: zm_segment), for the CPU test suite.
void PNX(b200_column_piece)(const int *prefix, int nt2, int sub, int seg, int nseg, int target, int fill, int *tz0, int *tz1) {
  int a = 0, b = 0;
  pnb::zm_segment(prefix, nt2, sub, seg, nseg, target, fill, a, b);
  *tz0 = a; *tz1 = b;
}

int PNX(b200_check_self_maps)(const INT *N, const INT *n, int m, int p0, int p1, int c0, int c1, int c2r) {
  pnb::Mesh M;
  M.np[0] = p0; M.np[1] = p1; M.co[0] = c0; M.co[1] = c1; M.size = p0 * p1; M.rank = M.rank_of(c0, c1);
  pnb::Layout L;
  const RT xm[3] = {(RT)0.5, (RT)0.5, (RT)0.5};
  pnb::compute_layout<RT>(L, M, N, n, xm, m, c2r != 0, 0u);
  const pnb::PipeGeom G = pnb::build_pipe(L, M);
  auto apply = [](std::vector<long long> &A, std::vector<long long> &Cb, const pnb::BoxMap &bm, bool a2c, bool sign) {
    for (long long i0 = 0; i0 < bm.dims[0]; i0++)
      for (long long i1 = 0; i1 < bm.dims[1]; i1++)
        for (long long i2 = 0; i2 < bm.dims[2]; i2++) {
          const long long ia = bm.a_off + i0 * bm.a_str[0] + i1 * bm.a_str[1] + i2 * bm.a_str[2];
          const long long ic = bm.c_off + i0 * bm.c_str[0] + i1 * bm.c_str[1] + i2 * bm.c_str[2];
          const bool neg = sign && ((i0 + i1 + i2 + bm.parity) & 1);
          if (a2c) Cb.at((size_t)ic) = neg ? -A.at((size_t)ia) : A.at((size_t)ia);
          else A.at((size_t)ia) = neg ? -Cb.at((size_t)ic) : Cb.at((size_t)ic);
        }
  };
  int checked = 0;
  for (int s = 0; s < 3; s++) {
    const pnb::Stage &S = G.st[s];
    for (const auto &T : S.tr) {
      if (T.peer != M.rank || T.send_elems == 0) continue;
      if (T.self_maps.empty()) return -2;
      const size_t ns = (size_t)std::max(S.src_elems, G.buf_elems), nd = (size_t)std::max(S.dst_elems, G.buf_elems);
      // forward: src -> chunk -> dst  against  src -> dst
      std::vector<long long> src(ns), chunk((size_t)T.send_elems, 0), d1(nd, 0), d2(nd, 0);
      for (size_t i = 0; i < ns; i++) src[i] = (long long)i + 1;
      for (const auto &bm : T.send_maps) apply(src, chunk, bm, true, T.send_sign);
      for (const auto &bm : T.recv_maps) apply(d1, chunk, bm, false, false);
      for (const auto &bm : T.self_maps) apply(d2, src, bm, false, T.send_sign);
      if (d1 != d2) return -1;
      // backward: dst-side array -> chunk -> src-side array  against  the composed copy
      std::vector<long long> arr(nd), o1(ns, 0), o2(ns, 0);
      for (size_t i = 0; i < nd; i++) arr[i] = (long long)i + 1;
      std::fill(chunk.begin(), chunk.end(), 0);
      for (const auto &bm : T.recv_maps) apply(arr, chunk, bm, true, false);
      for (const auto &bm : T.send_maps) apply(o1, chunk, bm, false, T.send_sign);
      for (const auto &bm : T.self_maps) apply(arr, o2, bm, true, T.send_sign);
      if (o1 != o2) return -1;
      checked++;
    }
  }
  return checked;
}
void PNX(b200_get_stage_ms)(PNX(plan) ths, int adjoint, double *ms8) { for (int i = 0; i < 8; i++) ms8[i] = AS_PLAN(ths)->stage_ms[adjoint ? 1 : 0][i]; }
long long PNX(b200_kernel_launches)(PNX(plan) ths) { return AS_PLAN(ths)->launches; }
long long PNX(b200_library_calls)(PNX(plan) ths) { return AS_PLAN(ths)->lib_launches; }
void *PNX(b200_get_stream)(PNX(plan) ths) { return (void *)AS_PLAN(ths)->stream; }

}  // extern "C"
