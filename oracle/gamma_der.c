/* TEST INFRASTRUCTURE (oracle) -- plain C, float64.
 *
 * dg/dalpha of a standard Gamma(alpha,1) draw g at fixed uniform quantile: the implicit
 * reparameterisation gradient tf.random.gamma supplies to the reference's ADVI step
 * [EXT: Figurnov et al. 2018; Eigen igamma_der_a / gamma_sample_der_alpha], which
 * mederrata_spmf/poisson.py:438-539 relies on through build_trainable_InverseGamma_dist.
 *
 * Same two expansions as oracle/spmf_oracle.py::gamma_sample_der_alpha (the vectorised python
 * version, which stays the definition; tests/test_oracle.py checks the two against each other and
 * against mpmath) -- here as scalar loops so that bench-scale parity tests (D*K*S ~ 1e7 draws)
 * finish in seconds:
 *   series   (g <= 1 or g <= a+1):  dg/da = (g/a) [sum T_n H_n - (log g - psi(a+1)) sum T_n]
 *   fraction (otherwise)         :  dg/da = g [dans/da + ans (log g - psi(a))]   (Cephes igamc)
 */
#include <math.h>

static double digamma_d(double x) {
  double r = 0.0;
  while (x < 12.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  /* asymptotic series with Bernoulli numbers up to B14 */
  const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 +
                   f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}

static double der_one(double a, double x) {
  if (x <= 1.0 || x <= a + 1.0) {
    double T = 1.0, H = 0.0, sT = 1.0, sTH = 0.0;
    for (int n = 1; n < 4000; ++n) {
      T *= x / (a + n);
      H += 1.0 / (a + n);
      sT += T;
      sTH += T * H;
      if (T * (1.0 + H) < 1e-18 * sT) break;
    }
    return (x / a) * (sTH - (log(x) - digamma_d(a + 1.0)) * sT);
  }
  double y = 1.0 - a, z = x + y + 1.0;
  const double dy = -1.0, dz = -1.0;
  double pkm2 = 1.0, qkm2 = x, pkm1 = x + 1.0, qkm1 = z * x;
  double dpkm2 = 0.0, dqkm2 = 0.0, dpkm1 = 0.0, dqkm1 = dz * x;
  double ans = pkm1 / qkm1, dans = (dpkm1 - ans * dqkm1) / qkm1;
  for (int c = 1; c < 5000; ++c) {
    y += 1.0;
    z += 2.0;
    const double yc = y * c, dyc = dy * c;
    const double pk = pkm1 * z - pkm2 * yc, qk = qkm1 * z - qkm2 * yc;
    const double dpk = dpkm1 * z + pkm1 * dz - dpkm2 * yc - pkm2 * dyc;
    const double dqk = dqkm1 * z + qkm1 * dz - dqkm2 * yc - qkm2 * dyc;
    const double nans = pk / qk, ndans = (dpk - nans * dqk) / qk;
    const double delta = fabs(ndans - dans), dval = fabs(nans - ans);
    ans = nans; dans = ndans;
    pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;
    dpkm2 = dpkm1; dpkm1 = dpk; dqkm2 = dqkm1; dqkm1 = dqk;
    if (fabs(pk) > 1e150) {
      const double sc = 1e-150;
      pkm2 *= sc; pkm1 *= sc; qkm2 *= sc; qkm1 *= sc;
      dpkm2 *= sc; dpkm1 *= sc; dqkm2 *= sc; dqkm1 *= sc;
    }
    if (c > 4 && delta < 1e-17 * (1.0 + fabs(dans)) && dval < 1e-17 * fabs(ans)) break;
  }
  return x * (dans + ans * (log(x) - digamma_d(a)));
}

void spmf_oracle_gamma_der(const double* a, const double* x, double* out, long long n) {
#pragma omp parallel for schedule(static, 4096)
  for (long long i = 0; i < n; ++i) out[i] = der_one(a[i], x[i]);
}
