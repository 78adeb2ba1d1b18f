// Host-side accuracy check of the kernel math (csrc/vbhem_math.cuh compiles for the host as well).
// Prints three lines, see tests/test_math.py:
//   exp_max_rel  log_max_abs_small  log_max_rel  lse_max_err
//   eight special-value flags
//   range-test flags
#include "vbhem_math.cuh"
#include <cstdio>
#include <random>
using namespace vbhem;

static const double et[] = VBHEM_EXP_TABLE;
static const double lt[] = VBHEM_LOG_TABLE;

static const long double UL = 0.693147180559945309417232121458L / VBHEM_EXP_BINS;   // u in long double

static double exp_full(double y) {            // what the kernel does with an un-shifted element (y in units)
  int e;
  const double m = exp_split(y, et, e);
  return std::ldexp(m, e);
}

int main() {
  std::mt19937_64 g(1);
  std::uniform_real_distribution<double> U(0, 1);
  double we = 0, wl_small = 0, wl_rel = 0, wlse = 0;
  for (int n = 0; n < 3000000; n++) {
    // exp over [-700, 700]: relative error of mant * 2^e against libm
    const double y = (U(g) - 0.5) * (n % 3 == 0 ? 1400.0 : (n % 3 == 1 ? 80.0 : 2.0)) * kPerUnit;
    const double a = exp_full(y);
    const long double b = expl((long double)y * UL);
    const double e = (double)(fabsl((long double)a - b) / b);
    if (e > we) we = e;
    // c ln2 + log s against libm (long double for the product)
    const double s = std::exp((U(g) - 0.9) * (n % 2 ? 640.0 : 30.0));
    const int c = (int)((U(g) - 0.5) * 4000.0);
    const long double want = (long double)c * 0.693147180559945309417232121458L + std::log((long double)s);
    const long double la = (long double)log_shift(s, lt, c) * UL;
    const double l0 = (double)fabsl((long double)log_shift(s, lt, 0) * UL - logl((long double)s));
    if (std::fabs(std::log(s)) < 32 && l0 > wl_small) wl_small = l0;
    const double er = (double)fabsl(la - want) /
                      std::fmax(std::fmax(std::fabs(c * 0.6931), std::fabs(std::log(s))), 1.0);
    if (er > wl_rel) wl_rel = er;
  }
  // a whole log-sum-exp row the way the kernel does it, at large un-shifted magnitudes
  for (int n = 0; n < 300000; n++) {
    const double base = -std::exp(U(g) * 16.0);            // down to -9e6
    double G[8], w[8];
    long double ref_max = -1e300L;
    for (int k = 0; k < 8; k++) { G[k] = (base + (U(g) - 0.5) * (n % 2 ? 20.0 : 1500.0)) * kPerUnit; w[k] = U(g); }
    for (int k = 0; k < 8; k++) if (G[k] * UL > ref_max) ref_max = G[k] * UL;
    long double ref = 0;
    for (int k = 0; k < 8; k++) ref += (long double)w[k] * expl((long double)G[k] * UL - ref_max);
    ref = ref_max + logl(ref);
    int e[8], c = -2147483647;
    double m[8], S = 0;
    for (int k = 0; k < 8; k++) { m[k] = exp_split(G[k], et, e[k]); if (e[k] > c) c = e[k]; }
    for (int k = 0; k < 8; k++) S += w[k] * exp_scale(m[k], e[k] - c);
    const long double got = (long double)log_shift(S, lt, c) * UL;
    const double er = (double)(fabsl(got - ref) / std::fmax(1.0L, fabsl(ref)));
    if (er > wlse) wlse = er;
  }
  std::printf("%.6e %.6e %.6e %.6e\n", we, wl_small, wl_rel, wlse);
  // special values
  int e3;
  const double m3 = exp_split(0.0, et, e3);
  std::printf("%d %d %d %d %d %d %d %d\n", m3 == 1.0 && e3 == 0, exp_scale(1.5, -5000) < 1e-300 && exp_scale(1.5, -5000) > 0,
              exp_scale(1.25, -3) == 1.25 / 8,
              nan_or_large(NAN) && nan_or_large(-NAN) && nan_or_large(INFINITY) && !nan_or_large(-INFINITY) && !nan_or_large(-3.0),
              nan_or_large(1.01 * kUnitsCeil) && !nan_or_large(0.99 * kUnitsCeil) && !nan_or_large(-1e300),
              nonfinite(NAN) && nonfinite(-INFINITY) && !nonfinite(1e308),
              std::fabs(kUnit * kPerUnit - 1.0) < 1e-15, kUnitsFloor == -std::ldexp(1.0, 36));
  // range tracking: row shifts are trusted from the emission floor's neighbourhood up to 1.5e9 orders
  int ea, eb, ec, ed;
  exp_split(-std::ldexp(1.0, 35) + 512, et, ea); exp_split(3.8e11, et, eb); exp_split(kUnitsFloor, et, ec); exp_split(4.0e11, et, ed);
  std::printf("%d %d %d %d\n", shift_excess(ea) <= kShiftBound, shift_excess(eb) <= kShiftBound,
              shift_excess(ec) > kShiftBound, shift_excess(ed) > kShiftBound);
  return 0;
}
