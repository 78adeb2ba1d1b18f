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
#define VB_MATH_CONSTS { 1.0 / 720, 1.0 / 120, 1.0 / 24, 1.0 / 6, /*4*/ VBHEM_32_LN2, -VBHEM_LN2_32_HI, \
                         -VBHEM_LN2_32_LO, 6755399441055744.0, /*8*/ -1.0 / 6, 0.2, 1.0 / 3, VBHEM_LN2_HI, \
                         VBHEM_LN2_LO }
#ifdef __CUDACC__
__constant__ double g_mc[13] = VB_MATH_CONSTS;
#endif
#ifdef __CUDA_ARCH__
#define VB_MC(i) g_mc[i]
#else
static const double h_mc[13] = VB_MATH_CONSTS;
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

// e^x for x <= ~0.  Returns 0 below -708 (the reference's exp would return a denormal < 2.3e-308
// there; it only ever enters sums whose largest term is 1).
VB_HD double exp_nonpos(double x) {
  constexpr double c[VBHEM_EXP_NCOEF] = VBHEM_EXP_COEF;
  constexpr double kMagic = 6755399441055744.0;            // 1.5 * 2^52
  const double t = fma(x, VBHEM_LOG2E, kMagic);
  const int k = lo_word(t);                                // rint(x / ln2)
  const double kf = t - kMagic;
  double r = fma(kf, -VBHEM_LN2_HI, x);
  r = fma(kf, -VBHEM_LN2_LO, r);
  double p = c[VBHEM_EXP_NCOEF - 1];
#pragma unroll
  for (int i = VBHEM_EXP_NCOEF - 2; i >= 0; i--) p = fma(p, r, c[i]);
  const double e = fma(p, r, 1.0);
  const double s = make_double(hi_word(e) + (k << 20), lo_word(e));
  return x < -708.0 ? 0.0 : s;
}

// Same function with a 32-entry table of 2^(j/32) (etab): 11 fp64 instructions instead of 16.
VB_HD double exp_nonpos_tab(double x, const double *etab) {
  const double t = fma(x, VB_MC(4), VB_MC(7));             // 32 x / ln2 + 1.5 * 2^52
  const int k = lo_word(t);                                // rint(32 x / ln2)
  const double kf = t - VB_MC(7);
  double r = fma(kf, VB_MC(5), x);
  r = fma(kf, VB_MC(6), r);
  const double T = etab[k & (VBHEM_EXP_BINS - 1)];
  double q = fma(r, VB_MC(0), VB_MC(1));
  q = fma(q, r, VB_MC(2));
  q = fma(q, r, VB_MC(3));
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  const double e = fma(T * r, q, T);                       // T * e^r
  const double s = make_double(hi_word(e) + ((k >> 5) << 20), lo_word(e));
  // x < -708  <=>  its high word, read as unsigned, exceeds that of -708.0 (x <= 0 here; -inf too)
  return (unsigned)hi_word(x) > 0xC0862000u ? 0.0 : s;
}

// tiny <= s < inf, with tiny = 2^-930 ~ 1.1e-280, on the integer pipe (false for NaN, Inf, negatives).
VB_HD bool sum_is_safe(double s) {
  return (unsigned)(hi_word(s) - 0x05D00000) < (unsigned)(0x7FF00000 - 0x05D00000);
}

// Order-preserving map double -> int32 that looks at the high word only, and its inverse.
VB_HD int order_key(double x) {
  const int h = hi_word(x);
  return h ^ ((h >> 31) & 0x7fffffff);
}
VB_HD double key_to_double(int key) { return make_double(key ^ ((key >> 31) & 0x7fffffff), 0); }

// log(s) for normal positive s.  tab = VBHEM_LOG_TABLE as (invc, logc) pairs.
VB_HD double log_pos(double s, const double *tab) {
  const int hi = hi_word(s);
  const int e = (hi >> 20) - 1023;
  const int i = (hi >> 13) & (VBHEM_LOG_BINS - 1);
  const double m = make_double((hi & 0x000fffff) | 0x3ff00000, lo_word(s));
  const double invc = tab[2 * i], logc = tab[2 * i + 1];
  const double r = fma(m, invc, -1.0);
  const double r2 = r * r;
  double q = fma(r, VB_MC(8), VB_MC(9));
  q = fma(q, r, -0.25);
  q = fma(q, r, VB_MC(10));
  q = fma(q, r, -0.5);
  const double p = fma(r2, q, r);
  const double ed = (double)e;
  return fma(ed, VB_MC(11), logc) + fma(ed, VB_MC(12), p);
}

}  // namespace vbhem
