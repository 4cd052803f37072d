// Log-space upper-tail probabilities for the GWAS statistics.
//
// The reference forms p-values only inside GenomicBreedingCore.plot (call sites
// /root/reference/src/gwas.jl:252 `plot(fit, TDist(n-1))` and :392 `plot(fit, Normal())`)
// as 1 - cdf(dist, |stat|), which is exactly 0 in Float64 beyond |t| ~ 8.3.  Here the
// survival function is evaluated in log space so -log10 p stays finite.
#pragma once
#include <math.h>

namespace gbm {

// Continued fraction of the incomplete beta function (modified Lentz).
__host__ __device__ inline double betacf(double a, double b, double x) {
  const double tiny = 1e-300, eps = 1e-16;
  double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 5000; ++m) {
    double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < eps) break;
  }
  return h;
}

// ln Gamma(a + 1/2) - ln Gamma(a), accurate for large a (asymptotic series in 1/a) and
// through lgamma for small a.
__host__ __device__ inline double lgamma_half_diff(double a) {
  if (a < 64.0) return lgamma(a + 0.5) - lgamma(a);
  // ln G(a+1/2) - ln G(a) = 1/2 ln a - 1/(8a) + 1/(192 a^3) - 1/(640 a^5) + 17/(14336 a^7) - ...
  double r = 1.0 / a, r2 = r * r;
  return 0.5 * log(a) + r * (-0.125 + r2 * (1.0 / 192.0 + r2 * (-1.0 / 640.0 + r2 * (17.0 / 14336.0))));
}

__host__ __device__ inline double log_sf_normal(double z);

// ln P(T_nu > t) for t >= 0.
__host__ __device__ inline double log_sf_t(double t, double nu) {
  if (!(t == t)) return t;  // NaN
  t = fabs(t);
  if (isinf(t)) return -INFINITY;
  if (nu >= 200.0 && t <= 8.0) {
    // Hill (1970, CACM Algorithm 395): the normal deviate z with P(Z > z) = P(T_nu > t), an expansion in
    // 1/(nu - 1/2) of w = (nu - 1/2) ln(1 + t^2/nu).  In this region its error in -log10 p is below 3e-12
    // (checked against the continued fraction for nu = 200 .. 1e6), and it replaces the slowest stretch of the
    // continued fraction (up to 53 iterations around |t| = 1.75, which every warp of a marker batch pays).
    const double a = nu - 0.5, b = 48.0 * a * a;
    const double y = a * log1p(t * t / nu);
    const double z = (((((-0.4 * y - 3.3) * y - 24.0) * y - 85.5) / (0.8 * y * y + 100.0 + b) + y + 3.0) / b + 1.0) * sqrt(y);
    return log_sf_normal(z);
  }
  const double a = 0.5 * nu, b = 0.5;
  const double t2 = t * t;
  const double omx = t2 / (nu + t2);  // 1 - x, x = nu/(nu+t^2)
  const double lnx = -log1p(t2 / nu);
  const double x = nu / (nu + t2);
  // ln B(a, 1/2) = ln G(a) + ln G(1/2) - ln G(a + 1/2)
  const double lnB = 0.5723649429247001 /* ln sqrt(pi) */ - lgamma_half_diff(a);
  if (x < (a + 1.0) / (a + b + 2.0)) {
    // I_x(a,b) = x^a (1-x)^b / (a B) * cf(a,b,x)
    double lnI = a * lnx + b * log(omx) - log(a) - lnB + log(betacf(a, b, x));
    return lnI + (-0.6931471805599453);
  }
  // small |t|: I_x(a,b) = 1 - I_{1-x}(b,a)
  double lnJ = b * log(omx) + a * lnx - log(b) - lnB + log(betacf(b, a, omx));
  double J = (t == 0.0) ? 0.0 : exp(lnJ);
  return log(0.5 * (1.0 - J));
}

// ln P(Z > z) for z >= 0, standard normal.
__host__ __device__ inline double log_sf_normal(double z) {
  if (!(z == z)) return z;
  z = fabs(z);
  if (isinf(z)) return -INFINITY;
  if (z < 20.0) return log(0.5 * erfc(z * 0.7071067811865476));
  // Mills-ratio asymptotic series: sf = phi(z)/z * (1 - 1/z^2 + 3/z^4 - 15/z^6 + ...)
  double r = 1.0 / (z * z), term = 1.0, sum = 1.0;
  for (int k = 1; k <= 12; ++k) {
    term *= -(2.0 * k - 1.0) * r;
    sum += term;
  }
  return -0.5 * z * z - log(z) - 0.9189385332046727 /* ln sqrt(2 pi) */ + log(sum);
}

}  // namespace gbm
