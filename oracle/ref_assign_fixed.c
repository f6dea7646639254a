/* TEST INFRASTRUCTURE ONLY -- the one documented deviation of the oracle build from
 * the unmodified reference.
 *
 * kernel/assign.c:166-176 of the reference calls its file-static helpers
 *   spread_f_r2r_pre_psi(f, pre_psi, m0, grid_size, cutoff, ostride, use_interlacing, grid)
 * while they are declared (kernel/assign.c:33-40, 517-521) as
 *   (..., int cutoff, int use_interlacing, INT ostride, R *grid),
 * so with ostride=1, use_interlacing=0 the real-valued adjoint spreads everything
 * into grid[m0] scaled by 0.5 (defect 1).  The public entry point is
 * renamed while the reference file is compiled and re-defined below with the two
 * arguments in the declared order; nothing else in assign.c is touched.
 * Build with -DORACLE_KEEP_C2R_SPREAD_BUG to get the reference's literal behaviour.
 */
#ifndef ORACLE_KEEP_C2R_SPREAD_BUG
#  if defined(PNFFT_PREC_SINGLE)
#    define pnfftf_spread_f_r2r pnfftf_spread_f_r2r_as_shipped
#  else
#    define pnfft_spread_f_r2r pnfft_spread_f_r2r_as_shipped
#  endif
#endif

#include "kernel/assign.c"

#ifndef ORACLE_KEEP_C2R_SPREAD_BUG
#  undef pnfft_spread_f_r2r
#  undef pnfftf_spread_f_r2r

void PNX(spread_f_r2r)(
    PNX(plan) ths, PNX(nodes) nodes, INT ind,
    R f, R *pre_psi,
    INT m0, const INT *grid_size, int cutoff, INT ostride,
    int use_interlacing, int interlaced,
    R *grid)
{
  (void)ths;
  R *tab = interlaced ? nodes->pre_psi_il : nodes->pre_psi;
  if (~nodes->precompute_flags & PNFFT_PRE_PSI)
    spread_f_r2r_pre_psi(f, pre_psi, m0, grid_size, cutoff, use_interlacing, ostride, grid);
  else if (nodes->precompute_flags & PNFFT_PRE_FULL)
    spread_f_r2r_pre_full_psi(f, tab + ind * PNFFT_POW3(cutoff), m0, grid_size, cutoff, use_interlacing, ostride, grid);
  else
    spread_f_r2r_pre_psi(f, tab + ind * 3 * cutoff, m0, grid_size, cutoff, use_interlacing, ostride, grid);
}
#endif
