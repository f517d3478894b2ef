// TEST INFRASTRUCTURE (oracle/ref_shim): boost::math::chi_squared + quantile(), the two names larvio.cpp:354-356 uses to fill
// its gating table.  Written from the definition: quantile(p) solves P(k/2, x/2) = p with the regularised lower incomplete
// gamma function P (series for x < a + 1, Lentz continued fraction otherwise), by bisection to the last bit.
#ifndef LVB_REF_SHIM_BOOST_CHI_SQUARED
#define LVB_REF_SHIM_BOOST_CHI_SQUARED
#include <cmath>
namespace boost { namespace math {
class chi_squared {
 public:
  explicit chi_squared(double k) : k_(k) {}
  double degrees_of_freedom() const { return k_; }
 private:
  double k_;
};
namespace shim_detail {
inline double gamma_p(double a, double x) {
  if (x <= 0.0) return 0.0;
  const double lg = std::lgamma(a);
  if (x < a + 1.0) {
    double term = 1.0 / a, sum = term, ap = a;
    for (int n = 0; n < 100000; ++n) { ap += 1.0; term *= x / ap; sum += term; if (std::fabs(term) < std::fabs(sum) * 1e-17) break; }
    return sum * std::exp(-x + a * std::log(x) - lg);
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 100000; ++i) {
    double an = -i * (i - a); b += 2.0;
    d = an * d + b; if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c; if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d; double del = d * c; h *= del;
    if (std::fabs(del - 1.0) < 1e-17) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - lg) * h;
}
}  // namespace shim_detail
inline double cdf(const chi_squared& d, double x) { return shim_detail::gamma_p(0.5 * d.degrees_of_freedom(), 0.5 * x); }
inline double quantile(const chi_squared& d, double p) {
  double lo = 0.0, hi = d.degrees_of_freedom() + 10.0;
  while (cdf(d, hi) < p) hi *= 2.0;
  for (int it = 0; it < 200; ++it) {
    double mid = 0.5 * (lo + hi);
    if (mid == lo || mid == hi) break;
    if (cdf(d, mid) < p) lo = mid; else hi = mid;
  }
  return 0.5 * (lo + hi);
}
}}  // namespace boost::math
#endif
