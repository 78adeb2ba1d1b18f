// Host-side accuracy check of the kernel math (csrc/vbhem_math.cuh compiles for the host as well).
// Prints: exp_max_rel log_max_abs_small log_max_rel_large  (see tests/test_math.py)
#include "vbhem_math.cuh"
#include <cstdio>
#include <random>
using namespace vbhem;
int main() {
  static const double et[] = VBHEM_EXP_TABLE;
  static const double lt[] = VBHEM_LOG_TABLE;
  std::mt19937_64 g(1);
  std::uniform_real_distribution<double> U(0, 1);
  double we = 0, wl_small = 0, wl_rel = 0;
  for (int n = 0; n < 3000000; n++) {
    const double x = (0.001 - U(g)) * (n % 3 == 0 ? 708.0 : (n % 3 == 1 ? 40.0 : 1.0));
    const double a = exp_nonpos_tab(x, et), b = std::exp(x);
    const double e = std::fabs(a - b) / b;
    if (e > we) we = e;
    const double s = std::exp((U(g) - 0.9) * (n % 2 ? 640.0 : 30.0));
    const double add = (U(g) - 0.5) * 2000.0;
    const double la = log_pos_plus(s, lt, add), lb = add + std::log(s);
    const double ea = std::fabs(la - lb);
    // (a) plain log (add = 0) in absolute terms for |log s| < 32; (b) c + log s relative to the operands' scale
    const double l0 = std::fabs(log_pos_plus(s, lt, 0.0) - std::log(s));
    if (std::fabs(std::log(s)) < 32 && l0 > wl_small) wl_small = l0;
    const double er = ea / std::fmax(std::fmax(std::fabs(add), std::fabs(std::log(s))), 1.0);
    if (er > wl_rel) wl_rel = er;
  }
  std::printf("%.6e %.6e %.6e\n", we, wl_small, wl_rel);
  // special values: clamp, range test, NaN screens, order key
  std::printf("%d %d %d %d %d %d %d %d\n", exp_nonpos_tab(-1e9, et) < 1e-300, exp_nonpos_tab(-INFINITY, et) < 1e-300,
              exp_nonpos_tab(0.0, et) == 1.0, range_excess(1.0) < kRangeBound, range_excess(1e-281) >= kRangeBound,
              range_excess(NAN) >= kRangeBound && range_excess(INFINITY) >= kRangeBound && range_excess(-1.0) >= kRangeBound,
              nan_or_posinf(NAN) && nan_or_posinf(-NAN) && nan_or_posinf(INFINITY) && !nan_or_posinf(-INFINITY) && !nan_or_posinf(-3.0),
              nonfinite(NAN) && nonfinite(-INFINITY) && !nonfinite(1e308));
  int bad = 0;
  double slack = 0;
  for (int n = 0; n < 1000000; n++) {
    const double a = (U(g) - 0.5) * std::exp((U(g) - 0.5) * 60), b = (U(g) - 0.5) * std::exp((U(g) - 0.5) * 60);
    if ((a < b) && (order_key(a) > order_key(b))) bad++;
    const double c = key_to_double(order_key(a));
    const double sl = std::fabs(c - a) / std::fabs(a);
    if (sl > slack) slack = sl;
  }
  std::printf("%d %.3e\n", bad, slack);
  return 0;
}
