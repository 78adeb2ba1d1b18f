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

// log(s) for normal positive s.  tab = VBHEM_LOG_TABLE as (invc, logc) pairs.
VB_HD double log_pos(double s, const double *tab) {
  const int hi = hi_word(s);
  const int e = (hi >> 20) - 1023;
  const int i = (hi >> 13) & (VBHEM_LOG_BINS - 1);
  const double m = make_double((hi & 0x000fffff) | 0x3ff00000, lo_word(s));
  const double invc = tab[2 * i], logc = tab[2 * i + 1];
  const double r = fma(m, invc, -1.0);
  const double r2 = r * r;
  double q = fma(-1.0 / 6, r, 0.2);
  q = fma(q, r, -0.25);
  q = fma(q, r, 1.0 / 3);
  q = fma(q, r, -0.5);
  const double p = fma(r2, q, r);
  const double ed = (double)e;
  return fma(ed, VBHEM_LN2_HI, logc) + fma(ed, VBHEM_LN2_LO, p);
}

}  // namespace vbhem
