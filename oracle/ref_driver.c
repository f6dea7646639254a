/* TEST INFRASTRUCTURE ONLY -- never linked into the product (pnfft_b200/).
 *
 * Driver around the UNMODIFIED reference sources (compiled where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/libpnfft_ref.so).  It runs the
 * reference's own API sequence (tests/simple_test.c:27-90 of the reference:
 * create_procmesh_2d -> local_size_guru -> init_guru -> init_nodes -> trafo/adj) on P
 * "virtual MPI ranks" (threads, oracle/shim/shim_mpi.c) and moves data between global
 * arrays handed in by the caller (ctypes/numpy) and the ranks' local blocks.
 *
 * This translation unit #includes kernel/ndft-parallel.c so that the file-static
 * functions of the hot path (lowest_summation_index, sort_nodes_for_better_cache_handle)
 * can be probed for the bit-exact integer parity tests.
 *
 * Compiled twice: double (prefix refdrv_) and, with -DPNFFT_PREC_SINGLE, float
 * (prefix refdrvf_).
 */
#include "kernel/ndft-parallel.c"

#include <string.h>

#if defined(PNFFT_PREC_SINGLE)
#define DRV(name) refdrvf_##name
#else
#define DRV(name) refdrv_##name
#endif

enum { OP_TRAFO = 0, OP_ADJ = 1, OP_LAYOUT = 2 };

typedef struct {
  int np0, np1;
  INT N[3], n[3];
  double x_max[3];
  int m;
  unsigned pnfft_flags, precompute_flags, compute_flags;
  int c2r;
  int op;
  int repeat;            /* how often trafo/adj is executed (timing) */
  INT M;                 /* global number of nodes */
  const R *x;            /* [M][3] */
  R *f_hat;              /* global, logical order [N0][N1][N2c], complex interleaved */
  R *f;                  /* [M] complex interleaved (c2c) or real (c2r) */
  R *grad_f;             /* [M][3] complex interleaved or real */
  R *grid;               /* optional global FFT-output array [no0][no1][no2] (complex or real) */
  R *g1;                 /* optional global FFT-input array, layout like f_hat */
  int set_grid, get_grid, set_g1, get_g1;
  int *owner;            /* out [M]: owning rank of each node (-1: none) */
  INT *layout;           /* out [P][15]: local_N, local_N_start, local_no, local_no_start, no */
  R *borders;            /* out [P][6]: lo[3], up[3] */
  double *timers;        /* out [P][2][PNFFT_TIMER_LENGTH]: trafo, adj */
  INT *node_index;       /* optional out [M][4]: u_j[0..2] in the padded local array, m0 */
  INT *sort_perm;        /* optional out [M]: global node index in the rank's processing order */
  double b_out[3];       /* out: window shape parameters */
  R *hessian_f;          /* optional out [M][6] complex interleaved or real (PNFFT_COMPUTE_HESSIAN_F) */
  double b_in[3];        /* b_in[0] > 0: pnfft_set_b(b_in) right after the plan is made (api/api-basic.c:587-596) */
} DRV(job);

typedef struct {
  DRV(job) *job;
  int fail;
} run_ctx;

static INT N2c(const DRV(job) *J) { return J->c2r ? J->N[2] / 2 + 1 : J->N[2]; }

/* offset (in complex units) of logical (k0,k1,k2) local indices inside a local f_hat block */
static INT fhat_off(const INT *lN, unsigned flags, INT i0, INT i1, INT i2)
{
  if (flags & PNFFT_TRANSPOSED_F_HAT) return (i1 * lN[2] + i2) * lN[0] + i0;
  return (i0 * lN[1] + i1) * lN[2] + i2;
}

static void copy_fhat(const DRV(job) *J, const INT *lN, const INT *lNs, R *glob, R *loc, int to_local)
{
  const INT G1 = J->N[1], G2 = N2c(J);
  for (INT i0 = 0; i0 < lN[0]; i0++)
    for (INT i1 = 0; i1 < lN[1]; i1++)
      for (INT i2 = 0; i2 < lN[2]; i2++) {
        INT g0 = lNs[0] + i0 + J->N[0] / 2, g1 = lNs[1] + i1 + J->N[1] / 2, g2 = lNs[2] + i2 + J->N[2] / 2;
        INT go = 2 * ((g0 * G1 + g1) * G2 + g2), lo = 2 * fhat_off(lN, J->pnfft_flags, i0, i1, i2);
        if (to_local) { loc[lo] = glob[go]; loc[lo + 1] = glob[go + 1]; }
        else { glob[go] = loc[lo]; glob[go + 1] = loc[lo + 1]; }
      }
}

static void copy_grid(const DRV(job) *J, const INT *no, const INT *lno, const INT *lnos, R *glob, R *loc, int to_local)
{
  const int tup = J->c2r ? 1 : 2;
  for (INT i0 = 0; i0 < lno[0]; i0++)
    for (INT i1 = 0; i1 < lno[1]; i1++) {
      INT g0 = lnos[0] + i0 + no[0] / 2, g1 = lnos[1] + i1 + no[1] / 2;
      R *g = glob + ((g0 * no[1] + g1) * no[2]) * tup, *l = loc + ((i0 * lno[1] + i1) * lno[2]) * tup;
      if (to_local) memcpy(l, g, sizeof(R) * (size_t)(lno[2] * tup));
      else memcpy(g, l, sizeof(R) * (size_t)(lno[2] * tup));
    }
}

static void rank_main(int rank, void *arg)
{
  run_ctx *ctx = (run_ctx *)arg;
  DRV(job) *J = ctx->job;
  const int tup = J->c2r ? 1 : 2;
  MPI_Comm comm_cart;
  INT local_N[3], local_N_start[3];
  R lo[3], up[3], x_max[3];

  for (int t = 0; t < 3; t++) x_max[t] = (R)J->x_max[t];

  if (PNX(create_procmesh_2d)(MPI_COMM_WORLD, J->np0, J->np1, &comm_cart)) { ctx->fail = 1; return; }

  if (J->c2r)
    PNX(local_size_guru_c2r)(3, J->N, J->n, x_max, J->m, comm_cart, J->pnfft_flags, local_N, local_N_start, lo, up);
  else
    PNX(local_size_guru)(3, J->N, J->n, x_max, J->m, comm_cart, J->pnfft_flags, local_N, local_N_start, lo, up);

  PNX(plan) ths = J->c2r
    ? PNX(init_guru_c2r)(3, J->N, J->n, x_max, J->m, J->pnfft_flags | PNFFT_MALLOC_F_HAT, PFFT_ESTIMATE, comm_cart)
    : PNX(init_guru)(3, J->N, J->n, x_max, J->m, J->pnfft_flags | PNFFT_MALLOC_F_HAT, PFFT_ESTIMATE, comm_cart);

  if (J->layout) {
    INT *L = J->layout + 15 * rank;
    for (int t = 0; t < 3; t++) {
      L[t] = local_N[t]; L[3 + t] = local_N_start[t];
      L[6 + t] = ths->local_no[t]; L[9 + t] = ths->local_no_start[t]; L[12 + t] = ths->no[t];
    }
  }
  if (J->borders) for (int t = 0; t < 3; t++) { J->borders[6 * rank + t] = lo[t]; J->borders[6 * rank + 3 + t] = up[t]; }
  if (J->b_in[0] > 0) PNX(set_b)((R)J->b_in[0], (R)J->b_in[1], (R)J->b_in[2], ths);
  if (rank == 0) for (int t = 0; t < 3; t++) J->b_out[t] = (double)ths->b[t];

  /* node ownership: the caller of PNFFT has to hand every rank the nodes with
   * lo <= x < up (kernel/ndft-parallel.c:734-775, api/api-adv.c:35-56) */
  INT local_M = 0;
  INT *mine = (INT *)malloc(sizeof(INT) * (size_t)(J->M > 0 ? J->M : 1));
  for (INT j = 0; j < J->M; j++) {
    int in = 1;
    for (int t = 0; t < 3; t++) if (!(lo[t] <= J->x[3 * j + t] && J->x[3 * j + t] < up[t])) in = 0;
    if (in) { mine[local_M++] = j; if (J->owner) J->owner[j] = rank; }
  }

  if (J->op != OP_LAYOUT) {
    PNX(nodes) nodes = PNX(init_nodes)(local_M, PNFFT_MALLOC_X | PNFFT_MALLOC_F | PNFFT_MALLOC_GRAD_F
                                                 | (J->hessian_f ? PNFFT_MALLOC_HESSIAN_F : 0u));
    for (INT p = 0; p < local_M; p++)
      for (int t = 0; t < 3; t++) nodes->x[3 * p + t] = J->x[3 * mine[p] + t];

    if (J->precompute_flags) PNX(precompute_psi)(ths, nodes, J->precompute_flags);

    if (J->node_index || J->sort_perm) {
      INT gcb[3], gca[3], ngc[3];
      get_size_gcells(ths->m, ths->cutoff, ths->pnfft_flags, gcb, gca);
      local_array_size(ths->local_no, gcb, gca, ngc);
      if (J->node_index)
        for (INT p = 0; p < local_M; p++) {
          R fl[3]; INT u[3];
          lowest_summation_index(ths->n, ths->m, nodes->x + 3 * p, ths->local_no_start, gcb, fl, u);
          INT *o = J->node_index + 4 * mine[p];
          o[0] = u[0]; o[1] = u[1]; o[2] = u[2]; o[3] = PNFFT_PLAIN_INDEX_3D(u, ngc);
        }
      if (J->sort_perm && local_M > 0) {
        INT *ar = (INT *)malloc(sizeof(INT) * 2 * (size_t)local_M);
        sort_nodes_for_better_cache_handle(3, ths->n, ths->m, local_M, nodes->x, ar);
        /* ranks write disjoint slots: position = (number of nodes of lower ranks) + p is not
         * known here, so the permutation is stored at the slots of this rank's own nodes */
        for (INT p = 0; p < local_M; p++) J->sort_perm[mine[p]] = mine[ar[2 * p + 1]];
        free(ar);
      }
    }

    if (J->op == OP_TRAFO) {
      if (J->f_hat) copy_fhat(J, local_N, local_N_start, J->f_hat, (R *)ths->f_hat, 1);
      if (J->set_g1 && J->g1) copy_fhat(J, local_N, local_N_start, J->g1, ths->g1, 1);
      if (J->set_grid && J->grid) copy_grid(J, ths->no, ths->local_no, ths->local_no_start, J->grid, ths->g2, 1);
      if (J->compute_flags & PNFFT_COMPUTE_ACCUMULATED) {
        for (INT p = 0; p < local_M; p++) {
          for (int c = 0; c < tup; c++) nodes->f[tup * p + c] = J->f[tup * mine[p] + c];
          for (int c = 0; c < 3 * tup; c++) nodes->grad_f[3 * tup * p + c] = J->grad_f ? J->grad_f[3 * tup * mine[p] + c] : 0;
        }
      }
      for (int r = 0; r < (J->repeat > 0 ? J->repeat : 1); r++)
        PNX(trafo)(ths, nodes, J->compute_flags);
      if (J->f && (J->compute_flags & PNFFT_COMPUTE_F))
        for (INT p = 0; p < local_M; p++)
          for (int c = 0; c < tup; c++) J->f[tup * mine[p] + c] = nodes->f[tup * p + c];
      if (J->grad_f && (J->compute_flags & PNFFT_COMPUTE_GRAD_F))
        for (INT p = 0; p < local_M; p++)
          for (int c = 0; c < 3 * tup; c++) J->grad_f[3 * tup * mine[p] + c] = nodes->grad_f[3 * tup * p + c];
      if (J->hessian_f && (J->compute_flags & PNFFT_COMPUTE_HESSIAN_F))
        for (INT p = 0; p < local_M; p++)
          for (int c = 0; c < 6 * tup; c++) J->hessian_f[6 * tup * mine[p] + c] = nodes->hessian_f[6 * tup * p + c];
      if (J->get_g1 && J->g1) copy_fhat(J, local_N, local_N_start, J->g1, ths->g1, 0);
      if (J->get_grid && J->grid) copy_grid(J, ths->no, ths->local_no, ths->local_no_start, J->grid, ths->g2, 0);
    } else {
      if (J->f)
        for (INT p = 0; p < local_M; p++)
          for (int c = 0; c < tup; c++) nodes->f[tup * p + c] = J->f[tup * mine[p] + c];
      if (J->grad_f)
        for (INT p = 0; p < local_M; p++)
          for (int c = 0; c < 3 * tup; c++) nodes->grad_f[3 * tup * p + c] = J->grad_f[3 * tup * mine[p] + c];
      if ((J->compute_flags & PNFFT_COMPUTE_ACCUMULATED) && J->f_hat)
        copy_fhat(J, local_N, local_N_start, J->f_hat, (R *)ths->f_hat, 1);
      if (J->set_g1 && J->g1) copy_fhat(J, local_N, local_N_start, J->g1, ths->g1, 1);
      if (J->set_grid && J->grid) copy_grid(J, ths->no, ths->local_no, ths->local_no_start, J->grid, ths->g2, 1);
      for (int r = 0; r < (J->repeat > 0 ? J->repeat : 1); r++)
        PNX(adj)(ths, nodes, J->compute_flags);
      if (J->f_hat) copy_fhat(J, local_N, local_N_start, J->f_hat, (R *)ths->f_hat, 0);
      if (J->get_g1 && J->g1) copy_fhat(J, local_N, local_N_start, J->g1, ths->g1, 0);
      if (J->get_grid && J->grid) copy_grid(J, ths->no, ths->local_no, ths->local_no_start, J->grid, ths->g2, 0);
    }

    if (J->timers) {
      double *T = J->timers + 2 * PNFFT_TIMER_LENGTH * rank;
      for (int t = 0; t < PNFFT_TIMER_LENGTH; t++) { T[t] = ths->timer_trafo[t]; T[PNFFT_TIMER_LENGTH + t] = ths->timer_adj[t]; }
    }
    PNX(free_nodes)(nodes, PNFFT_FREE_X | PNFFT_FREE_F | PNFFT_FREE_GRAD_F | (J->hessian_f ? PNFFT_FREE_HESSIAN_F : 0u));
  }

  free(mine);
  PNX(finalize)(ths, PNFFT_FREE_F_HAT);
  MPI_Comm_free(&comm_cart);
}

int DRV(run)(DRV(job) *job)
{
  run_ctx ctx;
  ctx.job = job; ctx.fail = 0;
  if (job->owner) for (INT j = 0; j < job->M; j++) job->owner[j] = -1;
  shim_mpi_run(job->np0 * job->np1, rank_main, &ctx);
  return ctx.fail;
}

/* ---- scalar probes of the reference's window functions (single rank) ---- */
typedef struct {
  INT N[3], n[3];
  double x_max[3];
  int m;
  unsigned pnfft_flags;
  int c2r;
} DRV(probe_cfg);

static PNX(plan) probe_plan(const DRV(probe_cfg) *P, MPI_Comm *comm)
{
  R x_max[3];
  for (int t = 0; t < 3; t++) x_max[t] = (R)P->x_max[t];
  PNX(create_procmesh_2d)(MPI_COMM_WORLD, 1, 1, comm);
  return P->c2r ? PNX(init_guru_c2r)(3, P->N, P->n, x_max, P->m, P->pnfft_flags, PFFT_ESTIMATE, *comm)
                : PNX(init_guru)(3, P->N, P->n, x_max, P->m, P->pnfft_flags, PFFT_ESTIMATE, *comm);
}

/* which: 0 psi, 1 dpsi, 2 inv_phi_hat, 3 phi_hat, 4 ddpsi, 5 pnfft_get_pnfft_flags.  arg: x (real) for 0/1/4, k (as R) for 2/3 */
void DRV(probe)(const DRV(probe_cfg) *P, int which, int dim, INT count, const R *arg, R *out)
{
  MPI_Comm comm;
  PNX(plan) ths = probe_plan(P, &comm);
  for (INT i = 0; i < count; i++) {
    switch (which) {
      case 0: out[i] = PNX(psi)(ths, dim, arg[i]); break;
      case 1: out[i] = PNX(dpsi)(ths, dim, arg[i]); break;
      case 2: out[i] = PNX(inv_phi_hat)(ths, dim, (INT)arg[i]); break;
      case 4: out[i] = PNX(ddpsi)(ths, dim, arg[i]); break;
      case 5: out[i] = (R)PNX(get_pnfft_flags)(ths); break;
      default: out[i] = PNX(phi_hat)(ths, dim, (INT)arg[i]); break;
    }
  }
  PNX(finalize)(ths, 0);
  MPI_Comm_free(&comm);
}

/* 3*(2m+1) tensor-product window values (and derivatives) of one node, as the hot loop
 * evaluates them (pre_psi_tensor / pre_dpsi_tensor, kernel/ndft-parallel.c:1621-1953) */
void DRV(probe_tensor)(const DRV(probe_cfg) *P, INT count, const R *x, R *psi, R *dpsi)
{
  MPI_Comm comm;
  PNX(plan) ths = probe_plan(P, &comm);
  const int c = ths->cutoff;
  for (INT j = 0; j < count; j++) {
    R fl[3]; INT u[3];
    project_node_to_grid(ths->n, ths->m, x + 3 * j, fl, u);
    pre_psi_tensor(ths->n, ths->b, ths->m, c, x + 3 * j, fl, ths->exp_const, ths->spline_coeffs, ths->pnfft_flags,
                   ths->intpol_order, ths->intpol_num_nodes, ths->intpol_tables_psi, psi + 3 * c * j);
    if (dpsi)
      pre_dpsi_tensor(ths->n, ths->b, ths->m, c, x + 3 * j, fl, ths->spline_coeffs,
                      ths->intpol_order, ths->intpol_num_nodes, ths->intpol_tables_dpsi,
                      psi + 3 * c * j, ths->pnfft_flags, dpsi + 3 * c * j);
  }
  PNX(finalize)(ths, 0);
  MPI_Comm_free(&comm);
}

/* sort keys of kernel/ndft-parallel.c:2131-2142 and the stable radix order */
void DRV(probe_sort)(const DRV(probe_cfg) *P, INT count, const R *x, INT *keys_and_perm)
{
  sort_nodes_for_better_cache_handle(3, P->n, P->m, count, x, keys_and_perm);
}

double DRV(bessel_i0)(double x) { return (double)PNX(bessel_i0)((R)x); }
double DRV(bessel_i1)(double x) { return (double)PNX(bessel_i1)((R)x); }

#if defined(PNFFT_PREC_SINGLE)
/* The reference's float build still calls the double-mangled pnfft_get_args from
 * pnfftf_check_init_parameters (api/api-basic.c:844-860, a name-mangling slip); the float oracle
 * library is linked without the double objects, so the symbol is supplied here. */
void pnfft_get_args(int argc, char **argv, const char *name, const int neededArgs, const unsigned type, void *parameter)
{
  PNX(get_args)(argc, argv, name, neededArgs, type, parameter);
}
/* likewise the un-mangled pfft_printf used by api/api-basic.c's parameter printer */
#include <stdarg.h>
void pfft_printf(MPI_Comm comm, const char *format, ...)
{
  int rank = 0;
  MPI_Comm_rank(comm, &rank);
  if (rank) return;
  va_list ap; va_start(ap, format); vfprintf(stdout, format, ap); va_end(ap);
}
#endif
