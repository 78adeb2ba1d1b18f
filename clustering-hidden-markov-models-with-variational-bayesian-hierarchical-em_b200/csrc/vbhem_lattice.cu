// vbhem_lattice.cu -- the tensor-core lattice kernels (sm_100a, fp64 DMMA).
//
// One warp owns one (base i, centre j) pair whose state lattice fits an 8 x 8
// tile (Sb, Sr <= 8; smaller HMMs are padded).  Every 8x8 matrix of the
// recursion lives in the accumulator layout of mma.sync.m8n8k4.f64 -- lane
// (g = lane/4, q = lane%4) holds elements [g][2q] and [g][2q+1] -- and every
// contraction is two DMMAs:  D(m|n) = sum_c A[m][c] * X[n][c]   ("mma_nt").
// Feeding the *variable* matrix as the B operand makes each product come out
// transposed, which is exactly what the alternation "contract over centre
// states, then over base states" of the recursion needs, so the backward pass
// runs with no data movement between lanes except one 4-lane max and one
// 2-value shuffle per step.
//
// Arithmetic: the reference evaluates, per time step and per rho, a max-shifted
// log-sum-exp over sigma of  logA[rho,sigma] + E[beta,sigma] + L[beta,sigma]
// (/root/reference/src/vbhem/vbhem_hmm_bwd_fwd_mex.c:920-1000: 2*Sb*Sr^2 exp per
// step).  Here the shift is taken per beta only,
//     c[beta] = max_sigma G[beta,sigma],  G = E + L,
//     Bm[beta,sigma] = exp(G[beta,sigma] - c[beta]),
//     S[beta,rho] = sum_sigma expA[rho,sigma] * Bm[beta,sigma]      (DMMA)
//     lse[beta,rho] = c[beta] + log S[beta,rho]
//     Theta[rho,sigma,beta] = expA[rho,sigma] * Bm[beta,sigma] / S[beta,rho]
// which needs Sb*Sr exp + Sb*Sr log per step and is the same number up to
// rounding whenever S does not underflow.  The shift never exists as a double:
// exp_split() returns e^G as mantissa x 2^e and c[beta] is the integer maximum
// of e over the row (vbhem_math.cuh), so S >= min expA > 0 by construction.
// Pairs with NaN/Inf in the data, centre weights below 2^-900, or a row shift
// outside +-1.5e9 binary orders are flagged and redone by the general
// log-space kernel, so every result keeps the reference's semantics.
// The forward pass (vbhem_hmm_bwd_fwd_mex.c:1100-1298) needs only Bm and 1/S of
// each step (2 x 64 doubles instead of the 512-double Theta slice), kept in
// shared memory and read back transposed.
#include "vbhem_internal.cuh"
#include "vbhem_math.cuh"

#include <type_traits>


namespace vbhem {

namespace {

constexpr int kLd = 10;           // padded row stride of an 8x8 tile in shared memory (doubles)
constexpr int kTile = 8 * kLd;    // doubles per tile

struct Frag { double x0, x1; };

// (invc, logc) pairs of log_shift() and the 2^(j/BINS) table of exp_split().  Every CTA copies them to
// shared memory: the per-lane index is random, so a plain table costs ~3 bank-conflict replays per
// lookup.  The first-pass kernel therefore keeps one copy per bank group -- 16 interleaved copies of
// the 8-byte exp table, 8 of the 16-byte log table -- and lane l reads copy l % 16 (l % 8): every
// lookup is conflict-free (LDS.64 is served per half-warp, LDS.128 per quarter-warp).  The second-pass
// kernel needs its shared memory for the stored lattice and keeps four copies of each.
__device__ const double g_log_table[2 * VBHEM_LOG_BINS] = VBHEM_LOG_TABLE;
__device__ const double g_exp_table[VBHEM_EXP_BINS] = VBHEM_EXP_TABLE;
template <int LS, int ES>
struct Tables {
  static constexpr int kLogDoubles = 2 * VBHEM_LOG_BINS * LS;
  static constexpr int kDoubles = kLogDoubles + VBHEM_EXP_BINS * ES;
  static __device__ __forceinline__ void load(double *dst) {
    for (int x = threadIdx.x; x < kLogDoubles; x += blockDim.x) {
      const int e = x & 1, i = (x >> 1) / LS;                               // dst[(i * LS + copy) * 2 + e]
      dst[x] = g_log_table[2 * i + e];
    }
    for (int x = threadIdx.x; x < VBHEM_EXP_BINS * ES; x += blockDim.x) dst[kLogDoubles + x] = g_exp_table[x / ES];
  }
  // the calling lane's views of the two tables (shared-space byte addresses)
  static __device__ __forceinline__ tab_t log_view(const double *base, int lane) {
    return (tab_t)__cvta_generic_to_shared(base + 2 * (lane % LS));
  }
  static __device__ __forceinline__ tab_t exp_view(const double *base, int lane) {
    return (tab_t)__cvta_generic_to_shared(base + kLogDoubles + (lane % ES));
  }
};
using TabFirst = Tables<8, 16>;   // first pass: bank-replicated
#ifndef VB_FULL_LS
#define VB_FULL_LS 4
#endif
#ifndef VB_FULL_ES
#define VB_FULL_ES 4
#endif
constexpr int kFullLS = VB_FULL_LS, kFullES = VB_FULL_ES;
using TabFull = Tables<kFullLS, kFullES>;   // second pass: the stored lattice needs the space, 24 KB of tables

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// D(m|n) = sum_c A[m][c] * X[n][c], all three in accumulator layout.
__device__ __forceinline__ void mma_nt_acc(Frag &D, const Frag &A, const Frag &X) {
  // one accumulator chain: splitting it costs two DADDs and is slower (profiles/r01b_sweep_occupancy.log)
  dmma(D.x0, D.x1, A.x0, X.x0);
  dmma(D.x0, D.x1, A.x1, X.x1);
}
__device__ __forceinline__ Frag mma_nt(const Frag &A, const Frag &X) {
  Frag D{0.0, 0.0};
  mma_nt_acc(D, A, X);
  return D;
}

// 1 / x for a normal positive x: MUFU.RCP64H seed + one cubic correction (r (1 + e + e^2), e = 1 - x r, |e| <= 2^-20),
// 3 fp64 instructions where the IEEE division takes ~ 10 and a branch; the lattice sums S qualify: every S is bounded
// below by a centre weight >= 2^-900 (k_build_records) times a row maximum in [1, 2).
__device__ __forceinline__ double rcp_pos(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
}

__device__ __forceinline__ double quad_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ double across_sum(double v) {   // over g, for one q
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

__device__ __forceinline__ void tile_store(double *tile, int g, int q, const Frag &f) {
  *reinterpret_cast<double2 *>(tile + g * kLd + 2 * q) = make_double2(f.x0, f.x1);
}
__device__ __forceinline__ Frag tile_load_t(const double *tile, int g, int q) {   // transposed fragment
  return Frag{tile[(2 * q) * kLd + g], tile[(2 * q + 1) * kLd + g]};
}

// ---- per-pair constants -----------------------------------------------------
template <int D, bool FULL>
struct BaseConst {
  static constexpr int DD = FULL ? D * D : D;
  Frag Ab;        // Ab(beta' = g | beta = 2q+e)  = A[from g, to 2q+e]
  Frag AbT;       // [g][2q+e] = A[from 2q+e, to g]
  double prior;   // prior[g]
  double mu[D];   // centres of state g
  double cov[DD]; // covars of state g
  int Sb;
};

template <int D, bool FULL>
__device__ __forceinline__ void load_base(BaseConst<D, FULL> &b, const BaseView &B, int i, int g, int q) {
  constexpr int DD = BaseConst<D, FULL>::DD;
  const int ld = B.maxSb;
  const double *A = B.A + (size_t)ld * ld * i;
  b.Sb = B.Sb[i];
  const int c0 = 2 * q, c1 = 2 * q + 1;
  b.Ab.x0 = (g < ld && c0 < ld) ? A[g + ld * c0] : 0.0;
  b.Ab.x1 = (g < ld && c1 < ld) ? A[g + ld * c1] : 0.0;
  b.AbT.x0 = (g < ld && c0 < ld) ? A[c0 + ld * g] : 0.0;
  b.AbT.x1 = (g < ld && c1 < ld) ? A[c1 + ld * g] : 0.0;
  b.prior = g < ld ? B.prior[g + (size_t)ld * i] : 0.0;
#pragma unroll
  for (int k = 0; k < D; k++) b.mu[k] = g < ld ? B.mu[k + D * (g + (size_t)ld * i)] : 0.0;
#pragma unroll
  for (int k = 0; k < DD; k++) b.cov[k] = g < ld ? B.cov[k + DD * (g + (size_t)ld * i)] : 0.0;
}

// Centre record (built by k_build_records in vbhem_aux.cu), doubles:
//   [0,64) expA[rho][sigma] | [64,72) expPi | [72,80) c0 = d*log(2pi)+cterm (+inf when padded)
//   [80, 80+8D) m[sigma][k] | [80+8D, 80+8D+8DD) vW[sigma][kl]
template <int D, bool FULL>
struct CentreConst {
  Frag expA;      // (rho = g | sigma = 2q+e)
  Frag expAT;     // [g][2q+e] = expA[rho = 2q+e][sigma = g]
  Frag expPi;     // sigma = 2q+e
};

template <int D, bool FULL>
__device__ __forceinline__ void load_centre(CentreConst<D, FULL> &c, const double *rec, int g, int q) {
  c.expA = Frag{rec[g * 8 + 2 * q], rec[g * 8 + 2 * q + 1]};
  c.expAT = Frag{rec[(2 * q) * 8 + g], rec[(2 * q + 1) * 8 + g]};
  c.expPi = Frag{rec[64 + 2 * q], rec[64 + 2 * q + 1]};
}

// E(beta = g | sigma = 2q+e) / u, vbhem_hmm_bwd_fwd_mex.c:796-843 (full) / :744-759 (diag).
template <int D, bool FULL>
__device__ __forceinline__ Frag emission(const BaseConst<D, FULL> &b, const double *rec, int g, int q) {
  constexpr int DD = BaseConst<D, FULL>::DD;
  double out[2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const int sg = 2 * q + e;
    const double *m = rec + 80 + sg * D;
    const double *P = rec + 80 + 8 * D + sg * DD;
    double acc = rec[72 + sg];
    if (FULL) {
#pragma unroll
      for (int k = 0; k < DD; k++) acc += P[k] * b.cov[k];
#pragma unroll
      for (int tj = 0; tj < D; tj++) {
        double t = 0.0;
#pragma unroll
        for (int ti = 0; ti < D; ti++) t += (b.mu[ti] - m[ti]) * P[ti + D * tj];
        acc += t * (b.mu[tj] - m[tj]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < D; k++) {
        const double df = b.mu[k] - m[k];
        acc += P[k] * b.cov[k];
        acc += P[k] * (df * df);
      }
    }
    // In lattice units (vbhem_math.cuh), clamped below at kUnitsFloor, which is also the value of an
    // impossible (padded) centre state; NaN and +Inf pass through for pair_inputs_bad.  Padded base
    // states: any finite row works (their weights are zero); 0 keeps S = rowsum(expA).
    const double v = (-0.5 * kPerUnit) * acc;
    out[e] = (g < b.Sb) ? (v < kUnitsFloor ? kUnitsFloor : v) : (isinf(rec[72 + sg]) ? kUnitsFloor : 0.0);
  }
  return Frag{out[0], out[1]};
}

// First-pass versions on a centre record staged in shared memory (32-bit shared address `rec`): the two
// elements of a lane (sigma = 2q, 2q+1) sit next to each other in every block of the record, so every
// quantity arrives in 16-byte loads.
__device__ __forceinline__ Frag lds_frag(unsigned saddr) {
  Frag f;
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(f.x0), "=d"(f.x1) : "r"(saddr));
  return f;
}
template <int D, bool FULL>
__device__ __forceinline__ Frag emission_smem(const BaseConst<D, FULL> &b, unsigned rec, int g, int q) {
  constexpr int DD = BaseConst<D, FULL>::DD;
  const Frag c0 = lds_frag(rec + 8u * (72 + 2 * q));
  double m[2 * D], P[2 * DD];
#pragma unroll
  for (int k = 0; k < D; k++) {
    const Frag f = lds_frag(rec + 8u * (80 + 2 * D * q + 2 * k));
    m[2 * k] = f.x0; m[2 * k + 1] = f.x1;
  }
#pragma unroll
  for (int k = 0; k < DD; k++) {
    const Frag f = lds_frag(rec + 8u * (80 + 8 * D + 2 * DD * q + 2 * k));
    P[2 * k] = f.x0; P[2 * k + 1] = f.x1;
  }
  double out[2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const double *me = m + e * D, *Pe = P + e * DD;
    double acc = e ? c0.x1 : c0.x0;
    if (FULL) {
#pragma unroll
      for (int k = 0; k < DD; k++) acc += Pe[k] * b.cov[k];
#pragma unroll
      for (int tj = 0; tj < D; tj++) {
        double t = 0.0;
#pragma unroll
        for (int ti = 0; ti < D; ti++) t += (b.mu[ti] - me[ti]) * Pe[ti + D * tj];
        acc += t * (b.mu[tj] - me[tj]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < D; k++) {
        const double df = b.mu[k] - me[k];
        acc += Pe[k] * b.cov[k];
        acc += Pe[k] * (df * df);
      }
    }
    // see emission(); a padded base row takes 0 everywhere (its weights are zero, any finite row works)
    const double v = (-0.5 * kPerUnit) * acc;
    out[e] = (g < b.Sb) ? (v < kUnitsFloor ? kUnitsFloor : v) : 0.0;
  }
  return Frag{out[0], out[1]};
}

// Shift of one lattice row = the largest binary exponent of e^G over the row (integer pipe, 32-bit
// shuffles over the four lanes that hold the row).
// SP ("centre pairing"): 0 = one centre per tile.  1 = the tile holds TWO centres of at most four states each --
// columns 0-3 and 4-7, a block-diagonal expA -- so a row has one shift per half, shared by the two lanes that
// hold that half.  2 = two centres split at any column k (1 <= k <= 7; columns [0,k) and [k,8)): a row has a
// left and a right shift, every lane ends up with both (row_shift_split).
template <int SP = 0>
__device__ __forceinline__ int row_shift(int e0, int e1) {
  int m = max(e0, e1);
  m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
  if (SP == 0) m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
  return m;
}
// Which of a lane's two columns (2q, 2q+1) lie in the left segment of a split tile; lane q = 0 always holds a
// left column and publishes the left shift to the transposed read, the others publish the right one.
struct Split { bool l0, l1, pubL; };
__device__ __forceinline__ void row_shift_split(const Split &sp, int e0, int e1, int &cL, int &cR) {
  constexpr int kNone = -2147483647 - 1;
  cL = max(sp.l0 ? e0 : kNone, sp.l1 ? e1 : kNone);
  cR = max(sp.l0 ? kNone : e0, sp.l1 ? kNone : e1);
  cL = max(cL, __shfl_xor_sync(0xffffffffu, cL, 1));
  cR = max(cR, __shfl_xor_sync(0xffffffffu, cR, 1));
  cL = max(cL, __shfl_xor_sync(0xffffffffu, cL, 2));
  cR = max(cR, __shfl_xor_sync(0xffffffffu, cR, 2));
}

// The linear-domain recursion cannot create a NaN out of clean inputs (sums of non-negative
// products, logs of sums that are bounded below by a centre weight), so the inputs are screened once
// per pair, not per step: the pair is bad if a transition / prior / centre weight is non-finite, the
// base transition rows sum (in magnitude) to more than 1.001 (growth bound of the recursion), or an
// emission term is NaN, +Inf or above 2^35 units (9e7).  k_build_records poisons the emission constant of a
// centre whose weights are non-finite or below 2^-900, which lands here as a NaN.  Such pairs go to the general
// log-space kernel, which reproduces the reference.
__device__ __forceinline__ bool base_bad(const Frag &Ab, double prior) {
  const double rs = quad_sum(fabs(Ab.x0) + fabs(Ab.x1));
  return nonfinite((rs <= 1.001 ? 0.0 : INFINITY) + prior);   // NaN / Inf anywhere => non-finite
}
__device__ __forceinline__ bool pair_inputs_bad(bool base_is_bad, const Frag &E) {
  return base_is_bad | nan_or_large(E.x0) | nan_or_large(E.x1);
}

// One backward step (t = T-1 .. 1).  In: G(beta|sigma) = E + L of this step.  Out: G of the
// previous time step, G(beta'|rho) = E + sum_beta Ab[beta',beta] lse[beta,rho].  With STORE the
// step's Bm(beta|sigma) and 1/S^T(rho|beta) go to shared memory for the forward pass.  `rng`
// accumulates the excess of the step's row shifts (see shift_excess).
template <bool STORE>
__device__ __forceinline__ void bwd_step(const Frag &E, Frag &G, const Frag &expA, const Frag &Ab,
                                         tab_t ltab, tab_t etab, double *sBm, double *sInvS,
                                         unsigned &rng, int g, int q) {
  int e0, e1;
  const double m0 = exp_split<kFullES>(G.x0, etab, e0), m1 = exp_split<kFullES>(G.x1, etab, e1);
  const int c = row_shift(e0, e1);                             // shift of row beta = g
  rng = max(rng, shift_excess(c));
  const Frag Bm{exp_scale(m0, e0 - c), exp_scale(m1, e1 - c)};
  const Frag ST = mma_nt(expA, Bm);                            // S^T(rho = g | beta = 2q+e)
  const int cb0 = __shfl_sync(0xffffffffu, c - 1023, (2 * q) * 4);
  const int cb1 = __shfl_sync(0xffffffffu, c - 1023, (2 * q + 1) * 4);
  const Frag ell{log_shift_biased<kFullLS>(ST.x0, ltab, cb0), log_shift_biased<kFullLS>(ST.x1, ltab, cb1)};   // lse^T(rho | beta)
  G = E;
  mma_nt_acc(G, Ab, ell);                                      // G(beta' = g | rho = 2q+e)
  if (STORE) {
    tile_store(sBm, g, q, Bm);
    tile_store(sInvS, g, q, Frag{rcp_pos(ST.x0), rcp_pos(ST.x1)});
  }
}

// Termination (vbhem_hmm_bwd_fwd_mex.c:1029-1080).  Returns L_elbo (all lanes) and
// nu_1(beta = g | sigma = 2q+e) = prior * Theta_1.
template <int LS, int ES>
__device__ __forceinline__ double terminate(const Frag &G, const Frag &expPi, double prior, tab_t ltab,
                                            tab_t etab, int q, Frag &nu1, unsigned &rng) {
  int e0, e1;
  const double m0 = exp_split<ES>(G.x0, etab, e0), m1 = exp_split<ES>(G.x1, etab, e1);
  const int c = row_shift(e0, e1);
  rng = max(rng, shift_excess(c));
  const Frag Bm{exp_scale(m0, e0 - c), exp_scale(m1, e1 - c)};
  const double s = quad_sum(expPi.x0 * Bm.x0 + expPi.x1 * Bm.x1);
  const double ell = log_shift<LS>(s, ltab, c);
  const double Lq = across_sum(q == 0 ? (prior * kUnit) * ell : 0.0);    // valid in lanes with q == 0
  const double w = prior / s;
  nu1 = Frag{w * expPi.x0 * Bm.x0, w * expPi.x1 * Bm.x1};
  return __shfl_sync(0xffffffffu, Lq, 0);
}

// ---- first pass: L_elbo for every pair --------------------------------------
// Each warp keeps NP pairs (one base, NP consecutive centres) in flight: the recursion of one pair
// is a single dependent chain (shift -> exp -> DMMA -> log -> DMMA), so independent pairs are the
// only instruction-level parallelism there is; every phase below is written across the NP pairs.
#ifndef VB_BWD_WARPS
#define VB_BWD_WARPS 16     // warps per CTA (one 16-warp CTA per SM measured best: profiles/r01_sweep_occupancy.log)
#endif
#ifndef VB_BWD_MINB
#define VB_BWD_MINB 1       // CTAs per SM the register budget is set for
#endif
#ifndef VB_NP
#define VB_NP 2             // pairs in flight per warp
#endif
#ifndef VB_KC_BYTES
#define VB_KC_BYTES (64 * 1024)   // shared-memory budget for centre records per CTA
#endif
constexpr int kBwdWarps = VB_BWD_WARPS;
#ifndef VB_UNIT_CENTRES
#define VB_UNIT_CENTRES 16
#endif
constexpr int kUnitCentres = VB_UNIT_CENTRES;   // centres per work unit of the first pass (multiple of VB_NP)
constexpr int kNP = VB_NP;

template <int NP, int LS, int ES, int SP>
__device__ __forceinline__ void bwd_step_n(const Frag (&E)[NP], Frag (&G)[NP], const Frag (&expA)[NP],
                                           const Frag &Ab, tab_t ltab, tab_t etab,
                                           unsigned (&rng)[NP], int q, int one, int src0, int src1, const Split &sp) {
  int e0[NP], e1[NP], c[NP];
  Frag Bm[NP], ST[NP], ell[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    Bm[p].x0 = exp_split<ES>(G[p].x0, etab, e0[p]);
    Bm[p].x1 = exp_split<ES>(G[p].x1, etab, e1[p]);
  }
  if constexpr (SP == 2) {
    int cL[NP], cR[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) row_shift_split(sp, e0[p], e1[p], cL[p], cR[p]);
#pragma unroll
    for (int p = 0; p < NP; p++) {
      rng[p] = max(rng[p], max(shift_excess(cL[p]), shift_excess(cR[p])));
      Bm[p] = Frag{exp_scale(Bm[p].x0, e0[p] - (sp.l0 ? cL[p] : cR[p])), exp_scale(Bm[p].x1, e1[p] - (sp.l1 ? cL[p] : cR[p]))};
      c[p] = sp.pubL ? cL[p] : cR[p];
    }
  } else {
#pragma unroll
    for (int p = 0; p < NP; p++) c[p] = row_shift<SP>(e0[p], e1[p]);
#pragma unroll
    for (int p = 0; p < NP; p++) {
      rng[p] = max(rng[p], shift_excess(c[p]));
      Bm[p] = Frag{exp_scale(Bm[p].x0, e0[p] - c[p]), exp_scale(Bm[p].x1, e1[p] - c[p])};
    }
  }
#pragma unroll
  for (int p = 0; p < NP; p++) ST[p] = mma_nt(expA[p], Bm[p]);
#pragma unroll
  for (int p = 0; p < NP; p++) {
    // shift of row beta = 2q+e (and, with SP, of the segment that this lane's rho = g belongs to)
    const int cb0 = __shfl_sync(0xffffffffu, c[p] - 1023, src0);
    const int cb1 = __shfl_sync(0xffffffffu, c[p] - 1023, src1);
    ell[p] = Frag{log_shift_biased<LS>(ST[p].x0, ltab, cb0, one), log_shift_biased<LS>(ST[p].x1, ltab, cb1, one)};
  }
#pragma unroll
  for (int p = 0; p < NP; p++) {
    G[p] = E[p];
    mma_nt_acc(G[p], Ab, ell[p]);
  }
}

// Termination of NP pairs side by side (first pass: L_elbo only; L[p] is valid in lane p when NP == 2,
// in lane 0 otherwise).
template <int NP, int LS, int ES, int SP>
__device__ __forceinline__ void terminate_n(const Frag (&G)[NP], const Frag (&expPi)[NP], double prior,
                                            tab_t ltab, tab_t etab, int q, unsigned (&rng)[NP],
                                            double (&L)[NP], const Split &sp) {
  int e0[NP], e1[NP], c[NP];
  double m0[NP], m1[NP], s[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    m0[p] = exp_split<ES>(G[p].x0, etab, e0[p]);
    m1[p] = exp_split<ES>(G[p].x1, etab, e1[p]);
  }
  const double pu = prior * kUnit;
  if constexpr (SP == 2) {
    // two centres per tile split at any column, two tiles in flight: four segment sums per row; a transposing
    // reduction over the quad leaves lane q with the sum of (tile q & 1, segment q >> 1), one logarithm per lane
    static_assert(NP == 2, "centre pairing is written for two tiles in flight");
    int cL[2], cR[2];
    double tL[2], tR[2];
#pragma unroll
    for (int p = 0; p < 2; p++) row_shift_split(sp, e0[p], e1[p], cL[p], cR[p]);
#pragma unroll
    for (int p = 0; p < 2; p++) {
      rng[p] = max(rng[p], max(shift_excess(cL[p]), shift_excess(cR[p])));
      const double a0 = expPi[p].x0 * exp_scale(m0[p], e0[p] - (sp.l0 ? cL[p] : cR[p]));
      const double a1 = expPi[p].x1 * exp_scale(m1[p], e1[p] - (sp.l1 ? cL[p] : cR[p]));
      tL[p] = (sp.l0 ? a0 : 0.0) + (sp.l1 ? a1 : 0.0);
      tR[p] = (sp.l0 ? 0.0 : a0) + (sp.l1 ? 0.0 : a1);
    }
    const bool odd = q & 1, hi = q & 2;
    // xor 1: even lanes keep tile 0's sums, odd lanes tile 1's
    const double kL = (odd ? tL[1] : tL[0]) + __shfl_xor_sync(0xffffffffu, odd ? tL[0] : tL[1], 1);
    const double kR = (odd ? tR[1] : tR[0]) + __shfl_xor_sync(0xffffffffu, odd ? tR[0] : tR[1], 1);
    // xor 2: lanes q < 2 keep the left segment, q >= 2 the right one
    const double v = (hi ? kR : kL) + __shfl_xor_sync(0xffffffffu, hi ? kL : kR, 2);
    const int cs = odd ? (hi ? cR[1] : cL[1]) : (hi ? cR[0] : cL[0]);
    double w = pu * log_shift<LS>(v, ltab, cs);
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    L[0] = w;
    L[1] = w;
  } else {
#pragma unroll
  for (int p = 0; p < NP; p++) c[p] = row_shift<SP>(e0[p], e1[p]);
#pragma unroll
  for (int p = 0; p < NP; p++) {
    rng[p] = max(rng[p], shift_excess(c[p]));
    s[p] = expPi[p].x0 * exp_scale(m0[p], e0[p] - c[p]) + expPi[p].x1 * exp_scale(m1[p], e1[p] - c[p]);
  }
  if constexpr (SP == 1) {
    // two centres per tile, two tiles in flight: the half-row sums of both tiles, then lane q takes the
    // logarithm for (tile q & 1, half q >> 1); after the sum over base states lane q of row 0 holds that L_elbo
    static_assert(NP == 2, "centre pairing is written for two tiles in flight");
    const double v0 = s[0] + __shfl_xor_sync(0xffffffffu, s[0], 1);
    const double v1 = s[1] + __shfl_xor_sync(0xffffffffu, s[1], 1);
    const bool odd = q & 1;
    double w = pu * log_shift<LS>(odd ? v1 : v0, ltab, odd ? c[1] : c[0]);
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    L[0] = w;
    L[1] = w;
  } else if constexpr (NP == 2) {
    // Transposing reduction over the quad: lanes with even q end up with pair 0's sum, odd q with pair
    // 1's, so each lane takes ONE logarithm and the sum over base states runs once for both pairs.
    // L[0] is valid in lane 0, L[1] in lane 1.
    const bool odd = q & 1;
    double v = (odd ? s[1] : s[0]) + __shfl_xor_sync(0xffffffffu, odd ? s[0] : s[1], 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    double w = q < 2 ? pu * log_shift<LS>(v, ltab, odd ? c[1] : c[0]) : 0.0;
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    L[0] = w;
    L[NP - 1] = w;
  } else {
#pragma unroll
    for (int p = 0; p < NP; p++) s[p] += __shfl_xor_sync(0xffffffffu, s[p], 1);
#pragma unroll
    for (int p = 0; p < NP; p++) s[p] += __shfl_xor_sync(0xffffffffu, s[p], 2);
#pragma unroll
    for (int p = 0; p < NP; p++) L[p] = q == 0 ? pu * log_shift<LS>(s[p], ltab, c[p]) : 0.0;
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
      for (int p = 0; p < NP; p++) L[p] += __shfl_xor_sync(0xffffffffu, L[p], o);
    }
  }
  }
}

// rec_src / list: without SP, tile r of this launch is centre list[r] and its record is rec_src + RS * list[r];
// with SP, tile r is the pair (list[2r], list[2r+1]) (-1 = no partner) and its record is rec_src + RS * r.
// ksplit (SP == 2): the column where the second centre of every tile of this launch starts.
template <int D, bool FULL, int SP>
__global__ void __launch_bounds__(kBwdWarps * 32, VB_BWD_MINB)
k_lattice_bwd(BaseView B, const double *__restrict__ rec_src, const int *__restrict__ list, int ntiles, int T, int kc,
              double *L_elbo, Fallback fb, int i0, int i1, int ksplit) {
  extern __shared__ __align__(16) double smem[];
  constexpr int CPT = SP ? 2 : 1;                                           // centres per tile
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int warp = threadIdx.x >> 5;
  const int j0 = blockIdx.y * kc;
  const int nj = min(kc, ntiles - j0);
  const tab_t ltab = TabFirst::log_view(smem, lane), etab = TabFirst::exp_view(smem, lane);
  int one;   // 0x3ff00000 pinned in a register (see mant_hi)
  asm volatile("mov.b32 %0, 0x3ff00000;" : "=r"(one));
  double *recs = smem + TabFirst::kDoubles;
  constexpr int RS = 80 + 8 * D + 8 * BaseConst<D, FULL>::DD;              // lattice_rec_stride(d, dd)
  int *sidx = reinterpret_cast<int *>(recs + (size_t)kc * RS);             // centre index of every tile slot
  const unsigned recs_s = (unsigned)__cvta_generic_to_shared(recs);
  pdl_trigger();
  TabFirst::load(smem);             // constants: overlaps the tail of the kernel before (PDL)
  pdl_wait();                       // the centre records and tile lists come from the kernels before
  for (int x = threadIdx.x; x < nj * CPT; x += blockDim.x) sidx[x] = list[(size_t)j0 * CPT + x];
  for (int x = threadIdx.x; x < nj * RS; x += blockDim.x) {
    const int r = x / RS, o = x - r * RS;
    recs[x] = rec_src[(size_t)RS * (SP ? j0 + r : list[j0 + r]) + o];
  }
  __syncthreads();
  // lanes that hold the shift of row beta = 2q+e (with SP: of the segment this lane's rho = g lies in)
  const int src0 = (2 * q) * 4 + (SP == 1 ? (g >> 2) * 2 : (SP == 2 ? (g < ksplit ? 0 : 3) : 0)), src1 = src0 + 4;
  const Split sp{2 * q < ksplit, 2 * q + 1 < ksplit, q == 0};

  // Work unit = one base x kUnitCentres tiles (re-reading the base per unit costs < 1 %; smaller units were measured
  // slower throughout: 4 tiles per unit cost 4 % at 12 500 bases).  What is left after the last FULL round of the
  // grid's warps -- up to a whole round for a 12 500-base shard (8-GPU strong scaling), 4 % of its first pass -- is
  // cut in quarter units, so the ragged end of the kernel is a quarter as long.  Warp ids are SM-interleaved
  // (warp w of CTA b is number w * gridDim.x + b): the warps with one piece more than the others are spread evenly
  // over the SMs instead of filling the first few.
  constexpr int unit_centres = kUnitCentres, sub = kUnitCentres / 4;
  static_assert(sub % kNP == 0 && sub >= kNP, "quarter units hold whole groups of pairs in flight");
  const int nuc = (nj + unit_centres - 1) / unit_centres;
  const long long nunits = (long long)(i1 - i0) * nuc, nw = (long long)gridDim.x * kBwdWarps;
  const long long nfull = nunits / nw * nw, nwork = nfull + 4 * (nunits - nfull);
  for (long long w = (long long)warp * gridDim.x + blockIdx.x; w < nwork; w += nw) {
    long long u = w;
    int t0 = 0, tn = unit_centres;
    if (w >= nfull) { u = nfull + ((w - nfull) >> 2); t0 = (int)((w - nfull) & 3) * sub; tn = sub; }
    const int i = i0 + (int)(u / nuc), ju = (int)(u % nuc) * unit_centres, jc = ju + t0;
    const int je = min(min(ju + unit_centres, nj), jc + tn);
    BaseConst<D, FULL> b;
    load_base(b, B, i, g, q);
    const bool base_chk = base_bad(b.Ab, b.prior);
    for (int jj = jc; jj < je; jj += kNP) {
      Frag E[kNP], G[kNP], expA[kNP], expPi[kNP];
      unsigned rng[kNP];
      bool bad[kNP];
#pragma unroll
      for (int p = 0; p < kNP; p++) {
        const unsigned rec = recs_s + 8u * RS * (unsigned)min(jj + p, nj - 1);   // odd tail: redo the last tile, not stored
        expA[p] = lds_frag(rec + 8u * (g * 8 + 2 * q));
        expPi[p] = lds_frag(rec + 8u * (64 + 2 * q));
        E[p] = emission_smem<D, FULL>(b, rec, g, q);
        G[p] = E[p];
        rng[p] = 0u;
        bad[p] = pair_inputs_bad(base_chk, E[p]);
      }
      for (int t = T - 1; t >= 1; t--)
        bwd_step_n<kNP, 8, 16, SP>(E, G, expA, b.Ab, ltab, etab, rng, q, one, src0, src1, sp);
      // termination (vbhem_hmm_bwd_fwd_mex.c:1029-1080); running it as one more lattice step (logPi for every
      // row of logA, the prior for every row of A_b) needs no special code but twice the fp64 work: slower
      double Lt[kNP];
      terminate_n<kNP, 8, 16, SP>(G, expPi, b.prior, ltab, etab, q, rng, Lt, sp);
      if constexpr (kNP == 2) {
        // Lt[0] == Lt[1] holds, in lane l of row 0, the value of tile l (without SP) or of (tile l & 1, segment
        // l >> 1) (with SP): one store for all; the (rare) flagged pairs take the slow path
        const bool ok = (max(rng[0], rng[1]) <= kShiftBound) & !(bad[0] | bad[1]);
        const bool all_ok = __all_sync(0xffffffffu, ok);
        const int tile = jj + (lane & 1), slot = SP ? 2 * tile + ((lane >> 1) & 1) : tile;
        const int j = (lane < 2 * CPT && tile < nj) ? sidx[slot] : -1;
        const size_t pi = (size_t)i + (size_t)B.N * (j < 0 ? 0 : j);
        VB_ASSERT(j < 0 || (i >= i0 && i < i1 && i < B.N && slot < CPT * kc && j < (1 << 30)));
        if (j >= 0) L_elbo[pi] = Lt[0];
        if (!all_ok) {
#pragma unroll
          for (int p = 0; p < 2; p++) {
            const bool good = __all_sync(0xffffffffu, (rng[p] <= kShiftBound) & !bad[p]);
            if (!good && j >= 0 && (lane & 1) == p) { fb.flags[pi] = 1; atomicAdd(fb.count, 1); }
          }
        }
      } else {
        static_assert(!SP || kNP == 2, "centre pairing needs two tiles in flight");
#pragma unroll
        for (int p = 0; p < kNP; p++) {
          const double Le = Lt[p];
          const bool good = __all_sync(0xffffffffu, (rng[p] <= kShiftBound) & !bad[p]);
          if (lane == 0 && jj + p < nj) {
            const size_t pi = (size_t)i + (size_t)B.N * sidx[jj + p];
            L_elbo[pi] = Le;
            if (!good) { fb.flags[pi] = 1; atomicAdd(fb.count, 1); }
          }
        }
      }
    }
  }
}

// One block per pair of centres that share a tile: the tile record out of the two centres' own records
// (k_build_records): slots [0, ks) <- centre a's states 0.., slots [ks, 8) <- centre b's, expA block-diagonal.
__global__ void k_pair_records(const double *rec, const int *pairs, int rs, int d, int dd, int ks, double *out) {
  pdl_trigger();
  pdl_wait();
  const int a = pairs[2 * blockIdx.x], b = pairs[2 * blockIdx.x + 1];
  double *o = out + (size_t)rs * blockIdx.x;
  for (int x = threadIdx.x; x < rs; x += blockDim.x) {
    // (block, slot, k): which per-state block of the record, which of the 8 tile slots, which entry
    int base, width, slot, k;
    if (x < 64) { base = 0; width = 8; slot = x / 8; k = x % 8; }          // expA[rho = slot][sigma = k]
    else if (x < 72) { base = 64; width = 1; slot = x - 64; k = 0; }       // expPi
    else if (x < 80) { base = 72; width = 1; slot = x - 72; k = 0; }       // c0
    else if (x < 80 + 8 * d) { base = 80; width = d; slot = (x - 80) / d; k = (x - 80) % d; }
    else { base = 80 + 8 * d; width = dd; slot = (x - base) / dd; k = (x - base) % dd; }
    const int src = slot < ks ? a : b, sl = slot < ks ? slot : slot - ks;   // the centre's own state index
    double v;
    if (base == 0) {
      const int kl = k < ks ? k : k - ks;
      if ((k < ks) != (slot < ks)) v = 0.0;                                // off-diagonal block
      else if (src >= 0) v = rec[(size_t)rs * src + sl * 8 + kl];
      else v = kl == 0 ? 1.0 : 0.0;                                        // no partner: a harmless one-state centre
    } else if (src >= 0) {
      v = rec[(size_t)rs * src + base + sl * width + k];
    } else {
      v = base == 64 ? (sl == 0 ? 1.0 : 0.0) : (base == 72 ? (sl == 0 ? 0.0 : INFINITY) : 0.0);
    }
    o[x] = v;
  }
}

// ---- full pass: backward + forward + statistics -----------------------------
template <int D, bool FULL>
struct PairStats {
  static constexpr int DD = FULL ? D * D : D;
  double nu1, pr, mu[D], Mu[DD];   // row sigma = g, replicated over q
  Frag xi;                         // (sigma = g | rho = 2q+e)
};

// smem (per warp): (T-1) x {Bm tile, invS tile} + one exchange tile.
template <int D, bool FULL>
__device__ __forceinline__ bool pair_full(const BaseView &B, int i, const BaseConst<D, FULL> &b,
                                          const CentreConst<D, FULL> &c, const double *rec, int T,
                                          tab_t ltab, tab_t etab, double *sm, int g, int q, double &Lout,
                                          PairStats<D, FULL> &st) {
  constexpr int DD = BaseConst<D, FULL>::DD;
  double *xch = sm + (size_t)2 * kTile * (T - 1);
  const Frag E = emission(b, rec, g, q);
  Frag G = E;
  unsigned rng = 0u;
  const bool bad = pair_inputs_bad(base_bad(b.Ab, b.prior), E);
  for (int t = T - 1; t >= 1; t--) {
    double *s = sm + (size_t)2 * kTile * (t - 1);
    bwd_step<true>(E, G, c.expA, b.Ab, ltab, etab, s, s + kTile, rng, g, q);
  }
  Frag nu1;
  Lout = terminate<kFullLS, kFullES>(G, c.expPi, b.prior, ltab, etab, q, nu1, rng);
  if (!__all_sync(0xffffffffu, (rng <= kShiftBound) & !bad)) return false;

  // nu_1 (beta|sigma) -> nuT (sigma = g | gamma = 2q+e)
  tile_store(xch, g, q, nu1);
  __syncwarp();
  Frag nuT = tile_load_t(xch, g, q);
  __syncwarp();
  st.nu1 = quad_sum(nuT.x0 + nuT.x1);
  Frag n = nuT;
  Frag Xi{0.0, 0.0};
  for (int t = 1; t < T; t++) {
    const double *s = sm + (size_t)2 * kTile * (t - 1);
    const Frag BmT = tile_load_t(s, g, q);                 // (sigma' = g | gamma' = 2q+e)
    const Frag invS = tile_load_t(s + kTile, g, q);        // (gamma' = g | rho = 2q+e)
    const Frag foo = mma_nt(b.AbT, nuT);                   // (gamma' | rho): sum_gamma A[gamma,gamma'] nu[rho,gamma]
    const Frag h{foo.x0 * invS.x0, foo.x1 * invS.x1};
    tile_store(xch, g, q, h);
    const Frag M = mma_nt(c.expAT, h);                     // (sigma' | gamma'): sum_rho expA[rho,sigma'] h[gamma',rho]
    nuT = Frag{BmT.x0 * M.x0, BmT.x1 * M.x1};
    n.x0 += nuT.x0; n.x1 += nuT.x1;
    __syncwarp();
    const Frag hT = tile_load_t(xch, g, q);                // (rho = g | gamma' = 2q+e)
    __syncwarp();
    mma_nt_acc(Xi, BmT, hT);                               // (sigma | rho): sum_gamma Bm[gamma,sigma] h[gamma,rho]
  }
  st.xi = Frag{c.expAT.x0 * Xi.x0, c.expAT.x1 * Xi.x1};    // xi_hat[rho = 2q+e, sigma = g]

  // emission statistics (vbhem_hmm_bwd_fwd_mex.c:1348-1469): row sigma = g, gamma = 2q+e
  st.pr = quad_sum(n.x0 + n.x1);
  const int ld = B.maxSb;
  const int s0 = 2 * q, s1 = 2 * q + 1;
  double m0[D], m1[D], c0[DD], c1[DD];
#pragma unroll
  for (int k = 0; k < D; k++) {
    m0[k] = s0 < ld ? B.mu[k + D * (s0 + (size_t)ld * i)] : 0.0;
    m1[k] = s1 < ld ? B.mu[k + D * (s1 + (size_t)ld * i)] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < DD; k++) {
    c0[k] = s0 < ld ? B.cov[k + DD * (s0 + (size_t)ld * i)] : 0.0;
    c1[k] = s1 < ld ? B.cov[k + DD * (s1 + (size_t)ld * i)] : 0.0;
  }
#pragma unroll
  for (int k = 0; k < D; k++) st.mu[k] = quad_sum(n.x0 * m0[k] + n.x1 * m1[k]);
#pragma unroll
  for (int kl = 0; kl < DD; kl++) {
    const int k = FULL ? kl / D : kl, l = FULL ? kl % D : kl;
    st.Mu[kl] = quad_sum(n.x0 * (m0[k] * m0[l] + c0[kl]) + n.x1 * (m1[k] * m1[l] + c1[kl]));
  }
  return true;
}

template <int D, bool FULL>
__global__ void __launch_bounds__(512, 1)
k_lattice_full(BaseView B, CentreView C, int T, PairWork W, PairOut O, Fallback fb, int warps_per_block) {
  extern __shared__ __align__(16) double smem[];
  constexpr int DD = BaseConst<D, FULL>::DD;
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int warp = threadIdx.x >> 5;
  const tab_t ltab = TabFull::log_view(smem, lane), etab = TabFull::exp_view(smem, lane);
  pdl_trigger();
  TabFull::load(smem);
  pdl_wait();
  __syncthreads();
  double *sm = smem + TabFull::kDoubles + (size_t)warp * ((size_t)2 * kTile * (T - 1) + kTile);
  const long long gw = (long long)blockIdx.x * warps_per_block + warp;
  const long long nw = (long long)gridDim.x * warps_per_block;
  const int ro = C.maxSr;

  if (W.items == nullptr) {
    // every pair, literal outputs (mode 1) or L_elbo only (mode 0 is served by k_lattice_bwd)
    for (long long p = gw; p < W.npairs; p += nw) {
      const int i = (int)(p % B.N), j = (int)(p / B.N);
      BaseConst<D, FULL> b;
      CentreConst<D, FULL> c;
      const double *rec = C.rec + (size_t)C.rec_stride * j;
      load_base(b, B, i, g, q);
      load_centre(c, rec, g, q);
      PairStats<D, FULL> st;
      double Le;
      const bool ok = pair_full(B, i, b, c, rec, T, ltab, etab, sm, g, q, Le, st);
      if (!ok) {
        if (lane == 0) { fb.flags[p] = 1; atomicAdd(fb.count, 1); }
        continue;
      }
      if (lane == 0 && O.L_elbo) O.L_elbo[p] = Le;
      if (g < ro) {
        if (2 * q < ro) O.xi[(2 * q) + ro * (g + (size_t)ro * p)] = st.xi.x0;
        if (2 * q + 1 < ro) O.xi[(2 * q + 1) + ro * (g + (size_t)ro * p)] = st.xi.x1;
        if (q == 0) {
          O.nu1[g + (size_t)ro * p] = st.nu1;
          O.pr[g + (size_t)ro * p] = st.pr;
#pragma unroll
          for (int k = 0; k < D; k++) O.mu[g + ro * (k + (size_t)D * p)] = st.mu[k];
#pragma unroll
          for (int k = 0; k < DD; k++) O.Mu[g + ro * (k + (size_t)DD * p)] = st.Mu[k];
        }
      }
    }
    return;
  }

  // second pass of the fused E-step: one warp per range of the device-side list (PairWork), the weighted sums
  // of one centre in registers, written to slot (range + centre) whenever the centre changes
  const int nitems = W.totals[0], ipr = W.totals[1];
  for (int w = (int)gw; w < W.nranges; w += (int)nw) {
    const int e0 = w * ipr, e1 = min(e0 + ipr, nitems);
    if (e0 >= e1) break;                                   // ranges ascend: nothing further for this warp either
    int jcur = -1;
    CentreConst<D, FULL> c;
    const double *rec = nullptr;
    PairStats<D, FULL> acc{};
    auto flush = [&]() {
      VB_ASSERT(jcur >= 0 && jcur < C.K && w + jcur < W.nranges + C.K);
      double *part = O.partial + (size_t)O.per * (w + jcur);
      double *p_nu1 = part, *p_xi = part + ro, *p_pr = p_xi + ro * ro, *p_mu = p_pr + ro, *p_Mu = p_mu + ro * D;
      if (g < ro) {
        if (2 * q < ro) p_xi[(2 * q) + ro * g] = acc.xi.x0;
        if (2 * q + 1 < ro) p_xi[(2 * q + 1) + ro * g] = acc.xi.x1;
        if (q == 0) {
          p_nu1[g] = acc.nu1;
          p_pr[g] = acc.pr;
#pragma unroll
          for (int k = 0; k < D; k++) p_mu[g + ro * k] = acc.mu[k];
#pragma unroll
          for (int k = 0; k < DD; k++) p_Mu[g + ro * k] = acc.Mu[k];
        }
      }
    };
    for (int e = e0; e < e1; e++) {
      const PairItem it = W.items[e];
      VB_ASSERT(e >= 0 && (long long)e < W.npairs && it.i >= 0 && it.i < B.N && it.j >= jcur && it.j < C.K);
      if (it.j != jcur) {
        if (jcur >= 0) flush();
        jcur = it.j;
        rec = C.rec + (size_t)C.rec_stride * jcur;
        load_centre(c, rec, g, q);
        acc.nu1 = acc.pr = 0.0;
        acc.xi = Frag{0.0, 0.0};
#pragma unroll
        for (int k = 0; k < D; k++) acc.mu[k] = 0.0;
#pragma unroll
        for (int k = 0; k < DD; k++) acc.Mu[k] = 0.0;
      }
      const double wgt = W.zhat[it.i + (size_t)B.N * it.j] * W.Ntilde[it.i];
      BaseConst<D, FULL> b;
      load_base(b, B, it.i, g, q);
      PairStats<D, FULL> st;
      double Le;
      const bool ok = pair_full(B, it.i, b, c, rec, T, ltab, etab, sm, g, q, Le, st);
      if (lane == 0) fb.flags[e] = ok ? 0 : 1;             // every item: the flag array needs no clearing
      if (!ok) {
        if (lane == 0) atomicAdd(fb.count, 1);
        continue;
      }
      acc.nu1 += wgt * st.nu1;
      acc.pr += wgt * st.pr;
      acc.xi.x0 += wgt * st.xi.x0;
      acc.xi.x1 += wgt * st.xi.x1;
#pragma unroll
      for (int k = 0; k < D; k++) acc.mu[k] += wgt * st.mu[k];
#pragma unroll
      for (int k = 0; k < DD; k++) acc.Mu[k] += wgt * st.Mu[k];
    }
    flush();
  }
}

constexpr size_t kSmemBudget = 200 * 1024;

size_t full_smem_per_warp(int T) { return sizeof(double) * ((size_t)2 * kTile * (T > 1 ? T - 1 : 0) + kTile); }

template <typename F>
int dispatch(int d, int covmode, F f) {
  if (covmode) {
    if (d == 1) return f(std::integral_constant<int, 1>{}, std::true_type{});
    if (d == 2) return f(std::integral_constant<int, 2>{}, std::true_type{});
    if (d == 3) return f(std::integral_constant<int, 3>{}, std::true_type{});
  } else {
    if (d == 1) return f(std::integral_constant<int, 1>{}, std::false_type{});
    if (d == 2) return f(std::integral_constant<int, 2>{}, std::false_type{});
    if (d == 3) return f(std::integral_constant<int, 3>{}, std::false_type{});
  }
  return -1;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace

int lattice_rec_stride(int d, int dd) { return 80 + 8 * d + 8 * dd; }

int lattice_bwd_wave() { return VB_BWD_MINB * sm_count() * kBwdWarps; }

bool lattice_supported(const BaseView &B, const CentreView &C, int T, bool need_forward) {
  if (B.maxSb > kS8 || C.maxSr > kS8 || B.d < 1 || B.d > 3 || T < 1) return false;
  if (need_forward && full_smem_per_warp(T) > kSmemBudget) return false;
  return true;
}

namespace {
template <int SP>
int launch_bwd_list(const BaseView &B, const double *rec_src, const int *list, int ntiles, int rs, int T,
                    double *L_elbo, Fallback fb, cudaStream_t st, int i0, int n, int ksplit) {
  if (ntiles <= 0) return 0;
  int kc = (int)((VB_KC_BYTES) / (sizeof(double) * rs));
  if (kc > ntiles) kc = ntiles;
  const size_t smem = sizeof(double) * ((size_t)kc * rs + TabFirst::kDoubles) + sizeof(int) * 2 * (size_t)kc + 16;
  const int ny = (ntiles + kc - 1) / kc;
  const int cap = VB_BWD_MINB * sm_count();
  const int kcu = (kc + kUnitCentres - 1) / kUnitCentres;
  long long nxl = ((long long)n * kcu + kBwdWarps - 1) / kBwdWarps;
  int nx = nxl > (1 << 30) ? (1 << 30) : (int)nxl;
  if (nx > cap) nx = cap;
  return dispatch(B.d, B.covmode, [&](auto Dc, auto Fc) {
    auto kern = k_lattice_bwd<decltype(Dc)::value, decltype(Fc)::value, SP>;
    static size_t smem_set[64] = {};                            // per instantiation and device (the attribute is per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > smem_set[dev & 63]) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      smem_set[dev & 63] = smem;
    }
    launch_chain<1>(kern, dim3(nx, ny), dim3(kBwdWarps * 32), smem, st, B, rec_src, list, ntiles, T, kc, L_elbo, fb, i0,
                 i0 + n, ksplit);
    return 1;
  });
}
}  // namespace

// First pass over bases [i0, i0 + n): the centres of C.big_list one per tile, the pairs of the tile classes two per
// tile (class 0: halves, split at column 4; classes 1-3: split at that column).
int launch_lattice_bwd(const BaseView &B, const CentreView &C, int T, double *L_elbo, Fallback fb,
                       cudaStream_t st, int i0, int n) {
  if (n < 0) n = B.N - i0;
  if (n <= 0) return 0;
  int launched = 0;
  launched += launch_bwd_list<0>(B, C.rec, C.big_list, C.nbig, C.rec_stride, T, L_elbo, fb, st, i0, n, 8);
  int off = 0;
  for (int c = 0; c < kTileClasses; c++) {
    const double *rec = C.pair_rec + (size_t)C.rec_stride * off;
    const int *list = C.pair_list + 2 * (size_t)off;
    if (c == 0) launched += launch_bwd_list<1>(B, rec, list, C.npairs[c], C.rec_stride, T, L_elbo, fb, st, i0, n, 4);
    else launched += launch_bwd_list<2>(B, rec, list, C.npairs[c], C.rec_stride, T, L_elbo, fb, st, i0, n, c);
    off += C.npairs[c];
  }
  return launched;
}

int launch_pair_records(const CentreView &C, int d, int dd, double *out, cudaStream_t st) {
  int launched = 0, off = 0;
  for (int c = 0; c < kTileClasses; c++) {
    if (C.npairs[c] <= 0) continue;
    launch_chain(k_pair_records, dim3(C.npairs[c]), dim3(128), 0, st, C.rec, C.pair_list + 2 * (size_t)off,
                 C.rec_stride, d, dd, c == 0 ? 4 : c, out + (size_t)C.rec_stride * off);
    off += C.npairs[c];
    launched++;
  }
  return launched;
}

int launch_lattice_full(const BaseView &B, const CentreView &C, int T, const PairWork &W, const PairOut &O,
                        Fallback fb, cudaStream_t st) {
  const size_t pw = full_smem_per_warp(T);
  int wpb = (int)(kSmemBudget / pw);
  if (wpb > 16) wpb = 16;
  if (wpb < 1) return -1;
  long long units = W.items ? W.nranges : W.npairs;       // list work: its length is on the device; every range gets a warp
  if (W.items && W.npairs < units) units = W.npairs;
  if (units <= 0) return 0;
  long long blocks = (units + wpb - 1) / wpb;
  const long long cap = sm_count();
  if (blocks > cap) blocks = cap;
  const size_t smem = pw * wpb + sizeof(double) * TabFull::kDoubles;
  return dispatch(B.d, B.covmode, [&](auto Dc, auto Fc) {
    auto kern = k_lattice_full<decltype(Dc)::value, decltype(Fc)::value>;
    static size_t smem_set[64] = {};                            // per instantiation and device (the attribute is per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > smem_set[dev & 63]) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      smem_set[dev & 63] = smem;
    }
    launch_chain<1>(kern, dim3((unsigned)blocks), dim3(wpb * 32), smem, st, B, C, T, W, O, fb, wpb);
    return 1;
  });
}

}  // namespace vbhem
