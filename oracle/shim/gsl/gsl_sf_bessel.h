/* TEST INFRASTRUCTURE ONLY -- stand-in for GSL's <gsl/gsl_sf_bessel.h>.
 * The reference forwards pnfft_bessel_i0/i1 to GSL (kernel/bessel_i0.c:375,
 * kernel/bessel_i1.c:570); GSL is not installed here and its version is not pinned
 * by the reference ("parity unpinned" at this boundary).  shim_gsl.c evaluates the
 * defining series in long double, which is at least as accurate as GSL's
 * Chebyshev fits (checked against scipy.special.i0/i1 in tests/test_oracle.py). */
#ifndef ORACLE_SHIM_GSL_SF_BESSEL_H
#define ORACLE_SHIM_GSL_SF_BESSEL_H 1
#ifdef __cplusplus
extern "C" {
#endif
double gsl_sf_bessel_I0(double x);
double gsl_sf_bessel_I1(double x);
#ifdef __cplusplus
}
#endif
#endif
