// lrr_oracle_muse.hpp — CPU restatement of rrel_muset_obj(0,false)::scale (MUSE), TEST INFRASTRUCTURE ONLY.
// Shared by the oracle (lrr_oracle_core.hpp) and by the VXL API stand-in used to compile the reference's own
// estimator sources (vxl_shim/rrel/rrel_muset_obj.h): rrel is third-party code absent from /root/reference and
// unpinned upstream, so both sides use this one restatement (MUSE stays "parity unpinned").
#ifndef LRR_ORACLE_MUSE_HPP_
#define LRR_ORACLE_MUSE_HPP_
#include <algorithm>
#include <cmath>
#include <vector>
namespace lrro {
// ----------------------------------------------------------------------------------------------
// MUSE (rrel_muset_obj(0,false)::scale; call sites cdcl :721-725, rrl :630-634).  UNPINNED upstream.
// Restated from Miller & Stewart, "MUSE: Robust surface fitting using unbiased scale estimates",
// CVPR'96, trimmed variant: over inlier fractions 0.25..0.95 step 0.05 minimise
//   s_k * stddev(k-th order stat) / E(k-th order stat),  s_k = sum of k smallest |r| / divisor(k,n)
// for half-normal order statistics; return the minimising s_k (no sk refinement).
inline double gauss_inv_cdf(double p) {
  // Acklam's rational approximation + one Halley step on erfc: |rel err| < 1e-15.
  static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                             1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                             6.680131188771972e+01, -1.328068155288572e+01};
  static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                             -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
  static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                             3.754408661907416e+00};
  double x;
  if (p < 0.02425) {
    double q = std::sqrt(-2 * std::log(p));
    x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
        ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  } else if (p <= 1 - 0.02425) {
    double q = p - 0.5, r = q * q;
    x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
        (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
  } else {
    double q = std::sqrt(-2 * std::log(1 - p));
    x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
        ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
  }
  double e = 0.5 * std::erfc(-x / std::sqrt(2.0)) - p;
  double u = e * std::sqrt(2 * M_PI) * std::exp(x * x / 2);
  x = x - u / (1 + x * u / 2);
  return x;
}
inline double muse_expected_kth(unsigned k, unsigned n) {
  return gauss_inv_cdf(0.5 * (1.0 + double(k) / double(n + 1)));
}
inline double muse_stddev_kth(unsigned k, unsigned n) {
  double pk = double(k) / double(n + 1), qk = 1.0 - pk;
  double Qk = gauss_inv_cdf(0.5 * (1.0 + pk));
  double pQk = std::sqrt(2 / M_PI) * std::exp(-0.5 * Qk * Qk);  // half-normal density at Qk
  double Qk_prime = 1.0 / pQk;
  double Qk_dprime = Qk / (pQk * pQk);
  double Qk_tprime = (1.0 + 2.0 * Qk * Qk) / (pQk * pQk * pQk);
  double vrk = (pk * qk / double(n + 2)) * Qk_prime * Qk_prime +
               (pk * qk / (double(n + 2) * double(n + 2))) *
                   (2.0 * (qk - pk) * Qk_prime * Qk_dprime + pk * qk * (Qk_prime * Qk_tprime + 0.5 * Qk_dprime * Qk_dprime));
  return std::sqrt(vrk);
}
inline double muse_divisor(unsigned k, unsigned n) {
  double r = double(k) / double(n);
  double mean = gauss_inv_cdf(0.5 * (1.0 + r));
  return double(n) * std::sqrt(2 / M_PI) * (1.0 - std::exp(-0.5 * mean * mean));
}
// Given |r| sorted ascending.  Shared driver so that the objective loop is stated once.
inline double muse_scale_sorted(const std::vector<double>& abs_sorted) {
  const unsigned n = (unsigned)abs_sorted.size();
  if (n == 0) return 0.0;
  const double min_frac = 0.25, max_frac = 0.95, frac_inc = 0.05, min_ratio = 3.0;
  bool at_start = true;
  double best_sk = 0, best_obj = 0, sum = 0;
  unsigned prev_k = 0;
  for (double frac = min_frac; frac <= max_frac + 0.00001; frac += frac_inc) {
    long kk = std::lround(frac * n);
    unsigned k = (kk <= 0) ? 1u : (unsigned)kk;
    if (k > n) k = n;
    double ek = muse_expected_kth(k, n), sd = muse_stddev_kth(k, n);
    if (ek / sd < min_ratio) continue;
    for (unsigned i = prev_k; i < k; ++i) sum += abs_sorted[i];
    double sk = sum / muse_divisor(k, n);
    double obj = sk * sd / ek;
    if (at_start || obj < best_obj) { at_start = false; best_sk = sk; best_obj = obj; }
    prev_k = k;
  }
  return best_sk;
}
inline double muse_scale(std::vector<double> r) {
  for (auto& x : r) x = std::fabs(x);
  std::sort(r.begin(), r.end());
  return muse_scale_sorted(r);
}

}  // namespace lrro
#endif
