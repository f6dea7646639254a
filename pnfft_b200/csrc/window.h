// Window functions of the PNFFT B matrix and their Fourier coefficients (D matrix), written once for
// host and device.  Formulas restate SURVEY.md Appendix A, i.e. reference
//   kernel/ndft-parallel.c:1678-1797 (psi per tap), :1850-1953 (dpsi per tap), :2195-2269 (1-d windows),
//   kernel/matrix_D.c:30-132 (1/phi_hat), kernel/bspline.c:45-91 (cardinal B-spline), kernel/sinc.c:26-53.
// Everything is expressed in grid units: tap s of a node sits at y_s = floor(n x) - n x - m + s.
#pragma once
#include <cmath>
#include <cfloat>

#if defined(__CUDACC__)
#define PNB_HD __host__ __device__ __forceinline__
#else
#define PNB_HD inline
#endif

namespace pnb {

enum WindowKind { WIN_KAISER_BESSEL = 0, WIN_GAUSSIAN = 1, WIN_BSPLINE = 2, WIN_SINC_POWER = 3, WIN_BESSEL_I0 = 4,
                  // Fourier coefficients only (phi_hat_any): the Gaussian window with the coefficients of its truncation to
                  // [-m/n, m/n] (PNFFT_WINDOW_GAUSSIAN_T); psi itself is WIN_GAUSSIAN
                  WIN_GAUSSIAN_T = 5 };

constexpr int kMaxM = 16;               // 2m+1 <= 33 taps per axis
constexpr int kMaxCutoff = 2 * kMaxM + 1;

// math wrappers so that one template body serves float and double
PNB_HD double m_sqrt(double v) { return sqrt(v); }
PNB_HD float m_sqrt(float v) { return sqrtf(v); }
PNB_HD double m_exp(double v) { return exp(v); }
PNB_HD float m_exp(float v) { return expf(v); }
PNB_HD double m_sinh(double v) { return sinh(v); }
PNB_HD float m_sinh(float v) { return sinhf(v); }
PNB_HD double m_cosh(double v) { return cosh(v); }
PNB_HD float m_cosh(float v) { return coshf(v); }
PNB_HD double m_sin(double v) { return sin(v); }
PNB_HD float m_sin(float v) { return sinf(v); }
PNB_HD double m_cos(double v) { return cos(v); }
PNB_HD float m_cos(float v) { return cosf(v); }
PNB_HD double m_tan(double v) { return tan(v); }
PNB_HD float m_tan(float v) { return tanf(v); }
PNB_HD double m_floor(double v) { return floor(v); }
PNB_HD float m_floor(float v) { return floorf(v); }
PNB_HD double m_fabs(double v) { return fabs(v); }
PNB_HD float m_fabs(float v) { return fabsf(v); }
template <class R> PNB_HD R m_eps();
template <> PNB_HD double m_eps<double>() { return DBL_EPSILON; }
template <> PNB_HD float m_eps<float>() { return FLT_EPSILON; }

template <class R> PNB_HD R m_pi() { return (R)3.14159265358979323846; }

// x^(2m) for integer m >= 0 (the reference calls pow(x, 2.0*m); the exponent is always an even integer)
template <class R> PNB_HD R pow_2m(R x, int m) {
  R x2 = x * x, r = (R)1;
  int e = m;
  while (e) { if (e & 1) r *= x2; x2 *= x2; e >>= 1; }
  return r;
}

// sinc with the Taylor branch near 0 (reference kernel/sinc.c:26-53)
template <class R> PNB_HD R sinc(R x) {
  const R b = m_eps<R>();
  const R bs = m_sqrt(b), bs2 = m_sqrt(bs);
  const R ax = m_fabs(x);
  if (ax >= bs2) return m_sin(x) / x;
  R r = (R)1;
  if (ax >= b) {
    const R x2 = x * x;
    r -= x2 / (R)6;
    if (ax >= bs) r += (x2 * x2) / (R)120;
  }
  return r;
}

// Modified Bessel functions I0, I1 by their (all-positive) power series; the reference forwards to GSL
// (kernel/bessel_i0.c:375, kernel/bessel_i1.c:570).  Arguments on this path are <= m*b ~ 50.
template <class R> PNB_HD R bessel_i(int nu, R x) {
  const double ax = fabs((double)x);
  if (ax > 60.0) {  // asymptotic expansion, never reached by valid window parameters
    const double mu = 4.0 * nu * nu;
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 60; k++) {
      const double t = term * -(mu - (2.0 * k - 1.0) * (2.0 * k - 1.0)) / (8.0 * k * ax);
      if (fabs(t) >= fabs(term)) break;
      term = t; sum += term;
    }
    const double v = exp(ax) / sqrt(2.0 * 3.14159265358979323846 * ax) * sum;
    return (R)((nu && x < 0) ? -v : v);
  }
  const double q = 0.25 * ax * ax;
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 400; k++) {
    term *= q / ((double)k * (double)(k + nu));
    sum += term;
    if (term < 1e-17 * sum) break;
  }
  double v = nu ? 0.5 * ax * sum : sum;
  if (nu && x < 0) v = -v;
  return (R)v;
}

// Cardinal B-spline of order k (degree k-1, support [0,k]) at x -- scalar, O(k^2); the reference's
// de Boor scheme kernel/bspline.c:45-91 evaluates the same function.
template <class R> PNB_HD R bspline(int k, R x) {
  if (!(x > (R)0 && x < (R)k)) return (R)0;
  if ((R)k - x < x) x = (R)k - x;  // symmetry
  int r = (int)ceil((double)x) - 1;  // x in (r, r+1]
  const R v = x - (R)r;              // in (0,1]
  // a[j] = B_q(v + j), j = 0..q-1
  R a[2 * kMaxM + 2];
  for (int j = 0; j <= k; j++) a[j] = (R)0;
  a[0] = (R)1;
  for (int q = 2; q <= k; q++) {
    const R inv = (R)1 / (R)(q - 1);
    for (int j = q - 1; j >= 0; j--) {
      const R t = v + (R)j;
      const R lo = j > 0 ? a[j - 1] : (R)0;
      a[j] = (t * a[j] + ((R)q - t) * lo) * inv;
    }
  }
  return a[r];
}

// All 2m+1 taps of the B-spline window of one axis at once.  frac = n x - floor(n x) in [0,1).
// psi[s] = B_{2m}(s - frac); dpsi[s] = n (B_{2m-1}(s-1-frac) - B_{2m-1}(s-frac))  (= -n B'_{2m}(s-frac))
// (reference kernel/ndft-parallel.c:1716-1742, :1867-1893)
template <class R> PNB_HD void bspline_taps(int m, R frac, R n, R *psi, R *dpsi) {
  const int k = 2 * m;
  const R v = (R)1 - frac;  // in (0,1]
  R a[2 * kMaxM + 2];
  for (int j = 0; j <= k; j++) a[j] = (R)0;
  a[0] = (R)1;
  for (int q = 2; q <= k; q++) {
    if (q == k && dpsi) {
      // a[j] = B_{k-1}(v + j), j = 0..k-2.  tap s (>=1) sits at s - frac = v + (s-1)
      dpsi[0] = (R)0;  // B_{k-1}(-1-frac) - B_{k-1}(-frac) = 0
      for (int s = 1; s <= k; s++) {
        const R hi = (s - 1 <= k - 2) ? a[s - 1] : (R)0;  // B_{k-1}(s - frac)   = a[s-1]
        const R lo = (s - 2 >= 0) ? a[s - 2] : (R)0;      // B_{k-1}(s-1-frac)  = a[s-2]
        dpsi[s] = n * (lo - hi);
      }
    }
    const R inv = (R)1 / (R)(q - 1);
    for (int j = q - 1; j >= 0; j--) {
      const R t = v + (R)j;
      const R lo = j > 0 ? a[j - 1] : (R)0;
      a[j] = (t * a[j] + ((R)q - t) * lo) * inv;
    }
  }
  psi[0] = (R)0;
  for (int s = 1; s <= k; s++) psi[s] = a[s - 1];
  if (k == 1 && dpsi) { dpsi[0] = dpsi[1] = (R)0; }
}

// The same recursion with the order known at compile time: every loop unrolls, the triangle stays in registers and the
// level reciprocals are constants (bit-identical to bspline_taps: same operations in the same order).
template <class R, int M, bool WANT_D> PNB_HD void bspline_taps_fixed(R frac, R n, R (&psi)[2 * M + 1], R (&dpsi)[WANT_D ? 2 * M + 1 : 1]) {
  constexpr int k = 2 * M;
  const R v = (R)1 - frac;  // in (0,1]
  R a[k + 1];
#pragma unroll
  for (int j = 0; j <= k; j++) a[j] = (R)0;
  a[0] = (R)1;
#pragma unroll
  for (int q = 2; q <= k; q++) {
    if (q == k && WANT_D) {
      dpsi[0] = (R)0;
#pragma unroll
      for (int s = 1; s <= k; s++) {
        const R hi = (s - 1 <= k - 2) ? a[s - 1] : (R)0;
        const R lo = (s - 2 >= 0) ? a[s - 2] : (R)0;
        dpsi[s] = n * (lo - hi);
      }
    }
    const R inv = (R)1 / (R)(q - 1);
#pragma unroll
    for (int j = q - 1; j >= 0; j--) {
      const R t = v + (R)j;
      const R lo = j > 0 ? a[j - 1] : (R)0;
      a[j] = (t * a[j] + ((R)q - t) * lo) * inv;
    }
  }
  psi[0] = (R)0;
#pragma unroll
  for (int s = 1; s <= k; s++) psi[s] = a[s - 1];
}

// One tap of a point-wise window: y = l - n x (grid units), z = -y.  want_d: also the AD gradient weight.
// exp(x) for 0 <= x < 700 without the range / special-case handling of the library routine: x = k ln2 + r with
// |r| <= ln2 / 2 (two-term Cody-Waite), degree-13 Taylor polynomial (truncation < 2^-57), 2^k through the exponent field.
#if defined(__CUDACC__)
// Taylor coefficients 1/13! ... 1/0! in constant memory: a DFMA takes them as a c[bank][offset] operand; as literals each
// costs two UMOVs per use (18 % of the node-table kernel's instructions, profiles/r2_ncu_full.md)
static __constant__ double kExpTaylor[14] = {
    1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0};
#endif
PNB_HD double exp_pos(double x) {
  const double k = rint(x * 1.4426950408889634074);
  double r = fma(-k, 6.93147180369123816490e-01, x);
  r = fma(-k, 1.90821492927058770002e-10, r);
#if defined(__CUDA_ARCH__)
  double p = kExpTaylor[0];
#pragma unroll
  for (int i = 1; i < 14; i++) p = fma(p, r, kExpTaylor[i]);
  return __longlong_as_double(__double_as_longlong(p) + ((long long)(int)k << 52));
#else
  double p = 1.0 / 6227020800.0;
  p = fma(p, r, 1.0 / 479001600.0);
  p = fma(p, r, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return ldexp(p, (int)k);
#endif
}

#if defined(__CUDACC__)
// exp(x) for -700 < x < 700 (0 below): exp_pos's reduction and polynomial; the exponent is added as a multiple of 2^52
__device__ __forceinline__ double exp_mid(double x) {
  const double k = rint(x * 1.4426950408889634074);
  double r = fma(-k, 6.93147180369123816490e-01, x);
  r = fma(-k, 1.90821492927058770002e-10, r);
  double p = kExpTaylor[0];
#pragma unroll
  for (int i = 1; i < 14; i++) p = fma(p, r, kExpTaylor[i]);
  const double v = __longlong_as_double(__double_as_longlong(p) + (long long)(int)k * 4503599627370496LL);
  return x > -700.0 ? v : 0.0;
}
#endif
#if defined(__CUDA_ARCH__)
// sqrt(d) and 1 / sqrt(d) for a normal, positive d: MUFU.RSQ64H seed and two coupled Newton steps (no special-case branch;
// the library sqrt and the IEEE division each carry one).  Both results within 1 ulp.
__device__ __forceinline__ void sqrt_rsqrt_pos(double d, double *s, double *rs) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double g = d * y, h = 0.5 * y;
  double e = fma(-g, h, 0.5);
  g = fma(g, e, g); h = fma(h, e, h);
  e = fma(-g, h, 0.5);
  g = fma(g, e, g); h = fma(h, e, h);
  *s = g; *rs = h + h;
}
// 1 / p for a normal p: MUFU.RCP64H seed and two Newton steps
__device__ __forceinline__ double rcp_pos(double p) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
  double e = fma(-p, y, 1.0);
  y = fma(y, e, y);
  e = fma(-p, y, 1.0);
  y = fma(y, e, y);
  return y;
}
#endif

// Kaiser-Bessel tap and derivative in double with ONE exponential and ONE division (kernel/ndft-parallel.c:2241-2269
// evaluates sinh, cosh and three quotients per tap).  Same formulas, a few ulp from the library-call version; taps with
// b r < 0.5 (sinh would cancel) fall through to window_tap.
PNB_HD bool kb_tap_fast(double y, double n, double b, int m, bool want_d, double *psi_out, double *dpsi_out) {
  const double d = (double)m * (double)m - y * y;
  const double inv_pi = 0.31830988618379067154;
#if defined(__CUDA_ARCH__)
  const double ad = fabs(d);
  if (!(ad >= 1e-300)) return false;             // on the edge of the support: the library-call path
  double r, inv_r;
  sqrt_rsqrt_pos(ad, &r, &inv_r);
  const double x = b * r;
  if (!(x >= 0.5)) return false;
  if (d < 0.0) {                                   // the tap beyond the main lobe: sin / cos, one argument reduction
    double sn, cs;
    sincos(x, &sn, &cs);
    const double psi = sn * inv_r * inv_pi;
    *psi_out = psi;
    if (want_d) *dpsi_out = n * y * (inv_r * inv_r) * (psi - (b * inv_pi) * cs);
    return true;
  }
  const double e = exp_pos(x);
  const double ei = rcp_pos(e);
#else
  const double r = sqrt(fabs(d)), x = b * r;
  if (!(x >= 0.5)) return false;
  if (d < 0.0) {                                   // the tap beyond the main lobe: sin / cos, one argument reduction
    double sn, cs;
    sincos(x, &sn, &cs);
    const double inv_r = 1.0 / r;
    const double psi = sn * inv_r * inv_pi;
    *psi_out = psi;
    if (want_d) *dpsi_out = n * y * (inv_r * inv_r) * (psi - (b * inv_pi) * cs);
    return true;
  }
  const double e = exp_pos(x);
  const double q = 1.0 / (e * r);
  const double ei = q * r, inv_r = q * e;          // 1 / e, 1 / r
#endif
  const double psi = (0.5 * (e - ei)) * inv_r * inv_pi;
  *psi_out = psi;
  if (want_d) *dpsi_out = n * (-y) * (inv_r * inv_r) * (psi - (b * inv_pi) * (0.5 * (e + ei)));
  return true;
}

template <class R> PNB_HD void window_tap(int kind, R y, R n, R b, int m, bool want_d, R *psi_out, R *dpsi_out) {
  const R pi = m_pi<R>();
  R psi = (R)0, dpsi = (R)0;
  const R z = -y;
  switch (kind) {
    case WIN_GAUSSIAN: {
      psi = m_exp(-(y * y) / b) / m_sqrt(pi * b);
      if (want_d) dpsi = (R)(-2.0) * n / b * z * psi;
    } break;
    case WIN_SINC_POWER: {
      psi = pow_2m(sinc(pi * y / b), m) / b;
      if (want_d) {
        const R w = pi * z / b;
        dpsi = (m_fabs(w) > m_eps<R>()) ? (R)2 * (R)m * pi * n / b * ((R)1 / m_tan(w) - (R)1 / w) * psi : (R)0;
      }
    } break;
    case WIN_BSPLINE: {
      psi = bspline<R>(2 * m, y + (R)m);
      if (want_d) dpsi = n * (bspline<R>(2 * m - 1, y + (R)m - (R)1) - bspline<R>(2 * m - 1, y + (R)m));
    } break;
    case WIN_BESSEL_I0: {
      const R d = (R)m * (R)m - y * y;
      if (d < 0) { psi = (R)0; dpsi = (R)0; }
      else {
        const R r = m_sqrt(d);
        psi = (R)0.5 * bessel_i<R>(0, b * r);
        if (want_d) dpsi = (d > 0) ? (R)(-0.5) * b * n * z * bessel_i<R>(1, b * r) / r : -((R)0.25 * b * b * n) * z;
      }
    } break;
    default: {  // Kaiser-Bessel
      const R d = (R)m * (R)m - y * y;
      const R r = m_sqrt(m_fabs(d));
      if (d < 0) {
        psi = m_sin(b * r) / (pi * r);
        if (want_d) dpsi = n * z / d * (psi - b * m_cos(b * r) / pi);
      } else if (d > 0) {
        psi = m_sinh(b * r) / (pi * r);
        if (want_d) dpsi = n * z / d * (psi - b * m_cosh(b * r) / pi);
      } else {
        psi = b / pi;
        if (want_d) dpsi = -n * (R)m * b * b * b / ((R)3 * pi);
      }
    } break;
  }
  *psi_out = psi;
  if (want_d) *dpsi_out = dpsi;
}

// second derivative of the window with respect to x (reference kernel/ndft-parallel.c:2010-2105, 2220-2285), from the
// tap's value and first derivative; y as in window_tap (z = -y = n x - l)
template <class R> PNB_HD R window_ddtap(int kind, R y, R n, R b, int m, R psi, R dpsi) {
  const R pi = m_pi<R>();
  const R z = -y;
  switch (kind) {
    case WIN_GAUSSIAN:
      return (R)2 * n * n / b * ((R)2 / b * z * z - (R)1) * psi;
    case WIN_SINC_POWER: {
      const R w = pi * z / b, c = pi * n / b;
      if (m_fabs(w) > m_eps<R>()) {
        const R ct = (R)1 / m_tan(w);
        return (R)2 * (R)m * c * (ct - (R)1 / w) * dpsi + (R)2 * (R)m * c * c * ((R)1 / (w * w) - (R)1 - ct * ct) * psi;
      }
      return (R)(-2.0) * (R)m * c * c / ((R)3 * b);
    }
    case WIN_BSPLINE:
      return n * n * (bspline<R>(2 * m - 2, y + (R)m) - (R)2 * bspline<R>(2 * m - 2, y + (R)m - (R)1) + bspline<R>(2 * m - 2, y + (R)m - (R)2));
    case WIN_BESSEL_I0: {
      const R d = (R)m * (R)m - z * z;
      if (d < 0) return (R)0;
      if (d > 0) {
        const R r = m_sqrt(d), yy = z * z;
        return (R)0.5 * b * n * n / d * (b * yy * bessel_i<R>(0, b * r) - bessel_i<R>(1, b * r) / r * (yy + (R)m * (R)m));
      }
      return (b * b * (R)m * n) * (b * b * (R)m * n) / (R)16 - (b * n) * (b * n) / (R)4;
    }
    default: {   // Kaiser-Bessel
      const R d = (R)m * (R)m - z * z;
      const R r = m_sqrt(m_fabs(d));
      if (d < 0) return (R)3 * n * z * dpsi / d + n * n * psi / d * ((R)1 + (b * z) * (b * z)) - b * n * n / (pi * d) * m_cos(b * r);
      if (d > 0) return (R)3 * n * z * dpsi / d + n * n * psi / d * ((R)1 + (b * z) * (b * z)) - b * n * n / (pi * d) * m_cosh(b * r);
      return b * (b * n) * (b * n) / ((R)15 * pi) * ((b * (R)m) * (b * (R)m) - (R)5);
    }
  }
}

// ---- Fourier coefficients of the window (host only in practice: 3 tables per plan) ----
// Re erf(a + i c) for real a, c -- the factor that turns the Gaussian's Fourier coefficient into the one of the Gaussian
// truncated at the cutoff (reference kernel/matrix_D.c:37-65, which calls libcerf's cerf).  Own evaluation: along the
// vertical path t = a + i s,  erf(a + i c) = erf(a) + 2/sqrt(pi) * Int_0^c exp(s^2 - a^2) (sin(2 a s) + i cos(2 a s)) ds,
// the real part integrated by composite 32-point Gauss-Legendre in long double (the integrand is entire; panels of at
// most 0.5 in s and 8 rad in phase).  Even in c.
struct GaussLegendre32 {     // nodes and weights on [-1, 1]: Newton's iteration on the Legendre recurrence, once per process
  long double x[32], w[32];
  GaussLegendre32() {
    const int n = 32;
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < n; i++) {
      long double z = cosl(pi * ((long double)i + 0.75L) / ((long double)n + 0.5L)), dp = 1.0L;
      for (int it = 0; it < 100; it++) {
        const long double pn = legendre(n, z, &dp);
        const long double dz = pn / dp;
        z -= dz;
        if (fabsl(dz) < 1e-19L) break;
      }
      (void)legendre(n, z, &dp);
      x[i] = z;
      w[i] = 2.0L / ((1.0L - z * z) * dp * dp);
    }
  }
  static long double legendre(int n, long double z, long double *deriv) {      // P_n(z) and P_n'(z)
    long double p0 = 1.0L, p1 = z;
    for (int k = 2; k <= n; k++) { const long double pk = (((long double)(2 * k - 1)) * z * p1 - ((long double)(k - 1)) * p0) / (long double)k; p0 = p1; p1 = pk; }
    *deriv = (long double)n * (z * p1 - p0) / (z * z - 1.0L);
    return p1;
  }
};
inline const GaussLegendre32 &gauss_legendre_32() { static const GaussLegendre32 t; return t; }
inline long double re_erf_complex(long double a, long double c) {
  const long double *xs = gauss_legendre_32().x, *ws = gauss_legendre_32().w;
  const long double ac = fabsl(c);
  long panels_l = (long)ceill(ac / 0.5L), panels_p = (long)ceill(2.0L * fabsl(a) * ac / 8.0L);
  long panels = panels_l > panels_p ? panels_l : panels_p;
  if (panels < 1) panels = 1;
  const long double h = ac / (long double)panels;
  long double sum = 0.0L;
  for (long q = 0; q < panels; q++) {
    const long double mid = ((long double)q + 0.5L) * h;
    for (int i = 0; i < 32; i++) {
      const long double s = mid + 0.5L * h * xs[i];
      sum += ws[i] * expl(s * s - a * a) * sinl(2.0L * a * s);
    }
  }
  sum *= 0.5L * h;
  return erfl(a) + 1.12837916709551257389615890312154517L * sum;   // 2 / sqrt(pi)
}
template <class R> inline R phi_hat_any(int kind, long k, long n, R b, int m, bool inverse) {
  const R pi = m_pi<R>();
  switch (kind) {
    case WIN_GAUSSIAN: {
      const R e = (pi * (R)k / (R)n) * (pi * (R)k / (R)n) * b;
      return inverse ? m_exp(e) : m_exp(-e);
    }
    case WIN_GAUSSIAN_T: {   // reference kernel/matrix_D.c:37-65: exp(-(pi k / n)^2 b) * Re erf(m / sqrt(b) + i pi k sqrt(b) / n)
      const R sqrtb = m_sqrt(b);
      const R w = (R)re_erf_complex((long double)((R)m / sqrtb), (long double)(pi * (R)k * sqrtb / (R)n));
      const R e = (pi * (R)k / (R)n) * (pi * (R)k / (R)n) * b;
      const R coeff = m_exp(-e) * w;
      return inverse ? (R)1 / coeff : coeff;
    }
    case WIN_BSPLINE: {
      const R s = sinc<R>((R)k * pi / (R)n);
      return (R)std::pow((double)s, (inverse ? -2.0 : 2.0) * m);
    }
    case WIN_SINC_POWER: {
      const R d = m_fabs((R)k * b / (R)n);
      if (inverse) return (d < (R)m) ? (R)1 / bspline<R>(2 * m, d + (R)m) : (R)0;
      return bspline<R>(2 * m, d + (R)m);
    }
    case WIN_BESSEL_I0: {
      const R t = (R)2 * pi * (R)k / (R)n;
      const R d = b * b - t * t;
      const R r = m_sqrt(m_fabs(d));
      if (d < 0) return inverse ? r / m_sin((R)m * r) : m_sin((R)m * r) / r;
      if (d > 0) return inverse ? r / m_sinh((R)m * r) : m_sinh((R)m * r) / r;
      return inverse ? (R)1 / (R)m : (R)m;
    }
    default: {
      const R t = (R)2 * pi * (R)k / (R)n;
      const R d = b * b - t * t;
      if (d < 0) return (R)0;
      if (d > 0) { const R v = bessel_i<R>(0, (R)m * m_sqrt(d)); return inverse ? (R)1 / v : v; }
      return (R)1;
    }
  }
}

// window shape parameter b (reference kernel/ndft-parallel.c:1013-1064)
template <class R> inline R window_shape(int kind, int m, R sigma) {
  const R pi = m_pi<R>();
  switch (kind) {
    case WIN_GAUSSIAN: return ((R)m / pi) * (R)2 * sigma / ((R)2 * sigma - (R)1);
    case WIN_BSPLINE: return (R)0;
    case WIN_SINC_POWER: return (R)m * ((R)2 * sigma) / ((R)2 * sigma - (R)1);
    default: return pi * ((R)2 - (R)1 / sigma);
  }
}

}  // namespace pnb
