/* TEST INFRASTRUCTURE ONLY -- see gsl/gsl_sf_bessel.h in this directory. */
#include <math.h>
#include "gsl/gsl_sf_bessel.h"

/* modified Bessel functions of the first kind, order nu in {0,1}, long double */
static long double bessel_i_series(int nu, long double x)
{
  long double q = 0.25L * x * x, term = 1.0L, sum = 1.0L;
  for (int k = 1; k < 500; k++) {
    term *= q / ((long double)k * (long double)(k + nu));
    sum += term;
    if (term < 1e-22L * sum) break;
  }
  return nu ? 0.5L * x * sum : sum;
}

static long double bessel_i_asymptotic(int nu, long double x)
{
  const long double pi = 3.141592653589793238462643383279502884L;
  long double mu = 4.0L * nu * nu, term = 1.0L, sum = 1.0L;
  for (int k = 1; k < 200; k++) {
    long double t = term * -(mu - (2.0L * k - 1.0L) * (2.0L * k - 1.0L)) / (8.0L * k * x);
    if (fabsl(t) >= fabsl(term)) break;
    term = t;
    sum += term;
    if (fabsl(term) < 1e-22L * fabsl(sum)) break;
  }
  return expl(x) / sqrtl(2.0L * pi * x) * sum;
}

double gsl_sf_bessel_I0(double x)
{
  long double ax = fabsl((long double)x);
  return (double)(ax < 40.0L ? bessel_i_series(0, ax) : bessel_i_asymptotic(0, ax));
}

double gsl_sf_bessel_I1(double x)
{
  long double ax = fabsl((long double)x);
  long double v = ax < 40.0L ? bessel_i_series(1, ax) : bessel_i_asymptotic(1, ax);
  return (double)(x < 0 ? -v : v);
}
