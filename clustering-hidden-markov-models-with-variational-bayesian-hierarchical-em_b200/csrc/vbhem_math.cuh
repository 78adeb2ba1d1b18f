// vbhem_math.cuh -- lean fp64 exp / log / reciprocal for the lattice kernels.
//
// The E-step is bound by the fp64 pipe and libdevice's exp / log cost 19 / 47 DFMA-equivalents
// per call on B200 (profiles/r01_fp64_peaks.json), most of it argument-range and special-case
// handling these call sites never need:
//   exp_nonpos(x): x = G - max(G) is <= 0 (up to the float rounding of the shift) or -inf
//   log_pos(s)   : s is a sum of non-negative products, already checked >= 1e-280
// NaN / +Inf never reach them un-flagged: the caller tests every step's inputs and sends such
// pairs to the general log-space kernel.  Constants: csrc/gen_mathtab.py -> vbhem_mathtab.h.
// The same code compiles for the host (g++) so tests/test_math.py can check it against libm.
#pragma once
#include "vbhem_mathtab.h"

#ifdef __CUDACC__
#define VB_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#include <cstdint>
#include <cstring>
#define VB_HD inline
#endif

namespace vbhem {

// Polynomial / reduction constants.  On the device they live in constant memory so that DFMA reads
// them as c[bank][offset] operands; as literals ptxas re-materialises each one with two moves per use.
#define VB_MATH_CONSTS { 1.0 / 120, 1.0 / 24, 1.0 / 6, /*3*/ VBHEM_EXP_SCALE, -VBHEM_EXP_STEP_HI, \
                         -VBHEM_EXP_STEP_LO, 6755399441055744.0, /*7*/ 0.2, 1.0 / 3, /*9*/ 0.69314718055994530942 }
#ifdef __CUDACC__
__constant__ double g_mc[10] = VB_MATH_CONSTS;
#endif
#ifdef __CUDA_ARCH__
#define VB_MC(i) g_mc[i]
#else
static const double h_mc[10] = VB_MATH_CONSTS;
#define VB_MC(i) h_mc[i]
#endif

VB_HD int hi_word(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32);
#endif
}
VB_HD int lo_word(double x) {
#ifdef __CUDA_ARCH__
  return __double2loint(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return (int)(u & 0xffffffffu);
#endif
}
VB_HD double make_double(int hi, int lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, lo);
#else
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x;
#endif
}

// e^x for x <~ 0 (or -inf) with a VBHEM_EXP_BINS-entry table of 2^(j/BINS) (etab): 10 fp64
// instructions.  x = k (ln2/BINS) + r;  e^x = 2^(k >> SHIFT) * T[k & (BINS-1)] * (1 + r + ... + r^5/120).
// Below -708.5 it returns ~1e-308 instead of the reference's denormal / zero: the value only ever
// enters sums whose largest term is ~1.  NaN inputs are the caller's business (see pair_inputs_bad).
VB_HD double exp_nonpos_tab(double xin, const double *etab) {
#if defined(VB_EXPERIMENT) && (VB_EXPERIMENT == 2 || VB_EXPERIMENT == 3 || VB_EXPERIMENT == 5)   // timing experiment only
  return fma(xin, 1e-9, 1.0) + etab[0];
#endif
  // clamp at about -708.5 with one unsigned min on the high word (negative doubles order by magnitude
  // there; -inf and anything below the clamp become ~1e-308, i.e. nothing next to a row maximum of 1)
  const unsigned hx = (unsigned)hi_word(xin);
  const double x = make_double((int)(hx < 0xC0862400u ? hx : 0xC0862400u), lo_word(xin));
  const double t = fma(x, VB_MC(3), VB_MC(6));             // BINS x / ln2 + 1.5 * 2^52
  const int k = lo_word(t);                                // rint(BINS x / ln2)
  const double kf = t - VB_MC(6);
  double r = fma(kf, VB_MC(4), x);
  r = fma(kf, VB_MC(5), r);
  const double T = etab[k & (VBHEM_EXP_BINS - 1)];
  double q = fma(r, VB_MC(0), VB_MC(1));
  q = fma(q, r, VB_MC(2));
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  const double e = fma(T * r, q, T);                       // T * e^r
  return make_double(hi_word(e) + ((k >> VBHEM_EXP_SHIFT) << 20), lo_word(e));
}

// Range tracking on the integer pipe: a lattice sum s is trusted iff 2^-930 <= s < inf.
// range_excess(s) is < kRangeBound exactly then (NaN, Inf, negatives, zero and tiny values all map to
// large unsigned numbers), so one running unsigned max per pair replaces per-step fp64 compares.
constexpr unsigned kRangeBound = 0x7FF00000u - 0x05D00000u;
VB_HD unsigned range_excess(double s) { return (unsigned)(hi_word(s) - 0x05D00000); }

// true for NaN and +Inf (-Inf is a legitimate "impossible state" value of an emission term)
VB_HD bool nan_or_posinf(double x) {
  const int h = hi_word(x);
  return (h >= 0x7ff00000) | ((unsigned)h > 0xfff00000u) | (((unsigned)h == 0xfff00000u) & (lo_word(x) != 0));
}
// true for NaN and +-Inf
VB_HD bool nonfinite(double x) { return ((unsigned)hi_word(x) & 0x7ff00000u) == 0x7ff00000u; }

// Order-preserving map double -> int32 that looks at the high word only, and its inverse.
VB_HD int order_key(double x) {
  const int h = hi_word(x);
  return h ^ ((h >> 31) & 0x7fffffff);
}
VB_HD double key_to_double(int key) { return make_double(key ^ ((key >> 31) & 0x7fffffff), 0); }

// add + log(s) for normal positive s, 8 fp64 instructions.  tab = VBHEM_LOG_TABLE as (invc, logc)
// pairs: s = 2^e m, r = m invc - 1 (|r| < 2^-8.9), log s = e ln2 + logc + (r - r^2/2 + ... + r^5/5).
// Absolute error <= ~4e-16 + |e| 2.4e-17 (ln2 is a single double: the result is always added to a
// lattice shift of magnitude >> 1, so absolute accuracy is what matters).
VB_HD double log_pos_plus(double s, const double *tab, double add) {
#if defined(VB_EXPERIMENT) && (VB_EXPERIMENT == 2 || VB_EXPERIMENT == 4 || VB_EXPERIMENT == 5)   // timing experiment only
  return fma(s, 1e-9, add) + tab[0];
#endif
  const int hi = hi_word(s);
  const int e = (hi >> 20) - 1023;
  const int i = (hi >> VBHEM_LOG_SHIFT) & (VBHEM_LOG_BINS - 1);
  const double m = make_double((hi & 0x000fffff) | 0x3ff00000, lo_word(s));
  const double invc = tab[2 * i], logc = tab[2 * i + 1];
  const double r = fma(m, invc, -1.0);
  const double r2 = r * r;
  double q = fma(r, VB_MC(7), -0.25);
  q = fma(q, r, VB_MC(8));
  q = fma(q, r, -0.5);
  const double p = fma(r2, q, r);
  return fma((double)e, VB_MC(9), add) + (logc + p);
}
VB_HD double log_pos(double s, const double *tab) { return log_pos_plus(s, tab, 0.0); }

}  // namespace vbhem
