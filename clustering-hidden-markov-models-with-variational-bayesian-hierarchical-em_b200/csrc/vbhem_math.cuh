// vbhem_math.cuh -- lean fp64 exp / log for the lattice kernels.
//
// The E-step is bound by the fp64 pipe AND by register-file read bandwidth (a DFMA with three
// distinct register operands issues every 3 cycles, not 2: profiles/r01b_pipe_overlap.json), and
// libdevice's exp / log cost 19 / 47 DFMA-equivalents per call on B200
// (profiles/r01_fp64_peaks.json), most of it range and special-case handling these call sites never
// need.  Two ideas remove nearly all of that:
//
//  1. UNITS.  Every log-domain quantity of the lattice (emission terms, lse, L, G) is kept in units
//     of u = ln2 / BINS:  y = x / u.  Then the argument reduction of e^x is exact and free of
//     constants: k = rint(y), r = y - k, e^x = 2^(k / BINS) e^(r u).  The tables and polynomial
//     coefficients absorb u (csrc/gen_mathtab.py); the kernels convert once on the way in (the
//     emission term) and once on the way out (L_elbo).
//  2. INTEGER SHIFT.  A lattice row needs lse_sigma(G) = c + log sum_sigma w exp(G - c) for ANY c
//     near the row maximum, so no floating-point maximum is ever taken:
//       exp_split(y)    -> (mant, e) with e^(y u) = mant 2^e, mant in [1, 2)     7 fp64 instructions
//       row shift       =  integer max of e over the row (32-bit shuffles)
//       exp_scale()     -> mant 2^(e - c): two integer instructions on the high word
//       log_shift(s, c) =  ((c + exponent(s)) ln2 + log(mantissa(s))) / u         6 fp64 instructions + I2F
//
// Domain: exp_split takes |y| < 2^39 (no clamp inside: the kernels clamp the emission terms
// once per pair and the recursion cannot leave the range while its row shifts stay in
// [-2^27, 1.5e9] binary orders, which shift_excess() tracks; pairs outside go to the general
// log-space kernel).  NaN / +Inf never reach these functions un-flagged (pair_inputs_bad in
// vbhem_lattice.cu).  The same code compiles for the host (g++) so that tests/test_math.py can check
// it against libm.
#pragma once
#ifndef VBHEM_MATHTAB
#define VBHEM_MATHTAB "vbhem_mathtab.h"
#endif
#include VBHEM_MATHTAB
// Polynomial degrees: the generator's choice keeps both functions within 2 ulp; -DVB_EXP_DEG / -DVB_LOG_DEG override
// them for accuracy / speed experiments (scripts/sweep_math_degrees.sh).
#ifdef VB_EXP_DEG
#undef VBHEM_EXP_DEG
#define VBHEM_EXP_DEG VB_EXP_DEG
#endif
#ifdef VB_LOG_DEG
#undef VBHEM_LOG_DEG
#define VBHEM_LOG_DEG VB_LOG_DEG
#endif

#ifdef __CUDACC__
#define VB_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#include <cstdint>
#include <cstring>
#define VB_HD inline
#endif

namespace vbhem {

constexpr double kUnit = VBHEM_UNIT;          // u: natural-log value of one lattice unit
constexpr double kPerUnit = VBHEM_PER_UNIT;   // 1 / u
// Emission terms are clamped to >= kUnitsFloor once per pair (also the value of padded / impossible
// states); rows whose shift falls below -2^27 orders (= -2^35 units) are not trusted.
constexpr double kUnitsFloor = -68719476736.0;          // -2^36 units (~ -1.9e8 natural)
constexpr double kUnitsCeil = 34359738368.0;            //  2^35 units (~  9.3e7 natural)

// Polynomial constants.  On the device they live in constant memory: ptxas loads them into uniform registers
// (LDCU) or, for a few, hoists them into vector registers; as literals it re-materialises each one with two
// moves per use.  (Passing them as a __grid_constant__ kernel parameter gives the same SASS: sm_100 DFMA has
// no c[bank][offset] operand form, only R / UR / 32-bit immediate.)   [0..4] exp: u^n/n!;  [5..9] log: (-1)^(n+1)/(n kappa^n u);  [10] 1.5 * 2^52;  [11] 1;  [12] -kappa
#define VB_MATH_CONSTS { VBHEM_EXP_COEF, VBHEM_LOG_COEF, 6755399441055744.0, 1.0, -(VBHEM_LOG_KAPPA) }
#ifdef __CUDACC__
static __constant__ double g_mc[13] = VB_MATH_CONSTS;   // one copy per translation unit
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
// A lane's view of a lookup table.  Under nvcc it is the 32-bit shared-space byte address and the
// loads are explicit ld.shared (address arithmetic stays in 32 bits: mask, shift-add, load); for the
// host build (tests) it is a plain pointer.
#ifdef __CUDACC__
typedef unsigned tab_t;
__host__ __device__ __forceinline__ double tab_ld(tab_t base, unsigned byte_off) {
  double v = 0.0;
#ifdef __CUDA_ARCH__
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + byte_off));
#endif
  return v;
}
__host__ __device__ __forceinline__ void tab_ld2(tab_t base, unsigned byte_off, double &a, double &b) {
  a = b = 0.0;
#ifdef __CUDA_ARCH__
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(base + byte_off));
#endif
}
// a & mask as an opaque instruction, so that the compiler keeps "mask, then scale-and-add" (LOP3 + IMAD/LEA)
// instead of re-associating table addresses into shift, mask, add
template <unsigned MASK>
__host__ __device__ __forceinline__ unsigned and_opaque(unsigned a) {
  unsigned d = a & MASK;
#ifdef __CUDA_ARCH__
  asm("and.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "n"(MASK));
#endif
  return d;
}
// high word of the mantissa-only double: (hi & 0x000fffff) | one, one = 0x3ff00000 held in a register (one LOP3)
__host__ __device__ __forceinline__ int mant_hi(int hi, int one) {
  int d = (hi & 0x000fffff) | one;
#ifdef __CUDA_ARCH__
  asm("lop3.b32 %0, %1, 0x000fffff, %2, 0xEA;" : "=r"(d) : "r"(hi), "r"(one));
#endif
  return d;
}
#else
inline int mant_hi(int hi, int one) { return (hi & 0x000fffff) | one; }
template <unsigned MASK>
inline unsigned and_opaque(unsigned a) { return a & MASK; }
typedef const double *tab_t;
inline double tab_ld(tab_t base, unsigned byte_off) { return base[byte_off / 8]; }
inline void tab_ld2(tab_t base, unsigned byte_off, double &a, double &b) { a = base[byte_off / 8]; b = base[byte_off / 8 + 1]; }
#endif

// low 32 bits of (hi:lo) >> s, 0 < s < 32
VB_HD int funnel_right(int lo, int hi, int s) {
#ifdef __CUDA_ARCH__
  return (int)__funnelshift_r((unsigned)lo, (unsigned)hi, s);
#else
  return (int)((((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo) >> s);
#endif
}

// e^(y u) = mant * 2^e with mant in [1 - 2^-9, 2).  k = rint(y) is read off the low bits of
// y + 1.5 2^52 (k may have up to 42 bits: e is taken from both words), r = y - k exactly, e^(r u) by
// its Taylor polynomial.  ES = stride of the table in doubles (the first-pass kernel keeps one copy
// per shared-memory bank pair).
template <int ES = 1>
VB_HD double exp_split(double y, tab_t etab, int &e) {
  const double t = y + VB_MC(10);
  const double r = y - (t - VB_MC(10));
  e = funnel_right(lo_word(t), hi_word(t), VBHEM_EXP_SHIFT);   // k >> SHIFT
  const double T = tab_ld(etab, and_opaque<VBHEM_EXP_BINS - 1>((unsigned)lo_word(t)) * (8u * ES));
#if VBHEM_EXP_DEG == 3
  double q = fma(r, VB_MC(2), VB_MC(1));
#elif VBHEM_EXP_DEG == 4
  double q = fma(r, VB_MC(3), VB_MC(2));
  q = fma(q, r, VB_MC(1));
#else
  double q = fma(r, VB_MC(4), VB_MC(3));
  q = fma(q, r, VB_MC(2));
  q = fma(q, r, VB_MC(1));
#endif
  q = fma(q, r, VB_MC(0));
#ifdef VB_EXP_TAIL_3REG
  return fma(T * r, q, T);                                 // T * e^(r u): a DFMA with three register operands (3 cycles)
#else
  return T * fma(q, r, 1.0);                               // T * e^(r u), every instruction with at most two register operands
#endif
}

// mant * 2^de for de <= 0 (clamped at -1000: ~1e-301, nothing next to the row maximum's [1, 2)).
VB_HD double exp_scale(double mant, int de) {
  const int d = de > -1000 ? de : -1000;
  return make_double(hi_word(mant) + (int)((unsigned)d << 20), lo_word(mant));
}

// Validity of a row shift c (binary orders): trusted iff -2^27 <= c <= 1.5e9.  shift_excess(c) is
// <= kShiftBound exactly then, so one running unsigned max per pair covers every row of every step.
constexpr unsigned kShiftBound = 1500000000u + 134217728u;
VB_HD unsigned shift_excess(int c) { return (unsigned)c + 134217728u; }

// true for NaN, +Inf and anything above kUnitsCeil = 2^35 (-Inf is a legitimate "impossible state" term)
VB_HD bool nan_or_large(double y) {
  const int h = hi_word(y);
  return (h >= 0x42200000) | ((unsigned)h > 0xfff00000u) | (((unsigned)h == 0xfff00000u) & (lo_word(y) != 0));
}
// true for NaN and +-Inf
VB_HD bool nonfinite(double x) { return ((unsigned)hi_word(x) & 0x7ff00000u) == 0x7ff00000u; }

// (c ln2 + log s) / u for normal positive s; cb = c - 1023.  tab = VBHEM_LOG_TABLE as (invc, logc/u)
// pairs: s = 2^e m, r = m invc - 1, log s = e ln2 + logc + log1p(r); in units, e ln2 / u = e BINS
// exactly.  The table and the polynomial are in the scaled residual kappa r (gen_mathtab.py: the quadratic
// coefficient becomes the DFMA immediate -128).  LS = stride of the table in pairs.
// Error <= ~3e-12 units = 8e-15 natural (Taylor degree 4).
VB_HD constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x >> 1); }
template <int LS = 1>
VB_HD double log_shift_biased(double s, tab_t tab, int cb, int one = 0x3ff00000) {
  const int hi = hi_word(s);
  const int e = (hi >> 20) + cb;
  // byte offset of pair i = top LOG_BITS mantissa bits: i * 16 * LS, as one mask and one shift
  constexpr int sh = 20 - VBHEM_LOG_BITS - 4 - ilog2(LS);
  static_assert(sh >= 0 && (1 << ilog2(LS)) == LS, "table stride");
  const unsigned off = and_opaque<(VBHEM_LOG_BINS - 1u) << (20 - VBHEM_LOG_BITS)>((unsigned)hi) >> sh;
  const double m = make_double(mant_hi(hi, one), lo_word(s));
  double invc, logc;
  tab_ld2(tab, off, invc, logc);
  const double r = fma(m, invc, VB_MC(12));               // kappa (m invc_i - 1)
#if VBHEM_LOG_DEG == 4
  double q = fma(r, VB_MC(8), VB_MC(7));
#else
  double q = fma(r, VB_MC(9), VB_MC(8));
  q = fma(q, r, VB_MC(7));
#endif
  q = fma(q, r, VBHEM_LOG_C2);
  q = fma(q, r, VB_MC(5));
  return fma((double)e, (double)VBHEM_EXP_BINS, fma(q, r, logc));
}
template <int LS = 1>
VB_HD double log_shift(double s, tab_t tab, int c) { return log_shift_biased<LS>(s, tab, c - 1023); }

}  // namespace vbhem
