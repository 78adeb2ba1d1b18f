// vbhmm_fb.cu -- the VB-HMM forward-backward pass of the toolbox's HMM learning (SURVEY.md 8f #3,
// BASELINE.json configs[4]) on the device.
//
// Reference: /root/reference/src/hmm/vbhmm_fb.m:264-380 and its native twin
// /root/reference/src/hmm/vbhmm_fb_mex.c:314-950, selected by vbhmm_fb.m:99-145:
//   [logrho_Saved, gamma_all, xi_sum, phi_norm] =
//       vbhmm_fb_mex(data, K, N, dim, maxT, m, W, v, beta, logLambdaTilde, const_denominator, t_pz1, t_tpztzt1)
// Per sequence: Normal-Wishart expected emission log-densities (Bishop 10.64 / 10.46), max-shifted
// emissions, scaled forward recursion, backward recursion, responsibilities gamma and summed pair
// marginals xi, and phi_norm = sum_t log c_t + sum_t max_k logrho.
//
// This path is HBM-bound (about 3 flop per byte), unlike the VBHEM E-step.  Design:
//   * one thread per sequence; the observations are repacked once into X[dim, N, maxT], so that every
//     load and every store of a warp at time step t is one contiguous run of 32 x (dim or K) doubles;
//   * the model (m, W, v, dim/beta, logLambdaTilde, p(z1), p(z_t | z_t-1)) travels as a by-value kernel
//     parameter: it sits in the constant bank and every use is a DFMA operand, never a load;
//   * the forward pass writes logrho and parks alpha_hat in the gamma output; the backward pass reads
//     it back in place, recomputes the emissions of step t+1 from X (cheaper than reading logrho back)
//     and the scale c_{t+1} from alpha_t (bit-identical to the forward value, so no scale array);
//     (taking c_{t+1} = sum_j alpha_hat_t(j) eta(j) from the backward sums and log c from the lean log saves 70 fp64
//     instructions per step but costs registers: 2.30 -> 2.36 ms here, while the batched kernel, which is not held
//     to 128 registers, gains 3 % from it -- vbhmm_batch.inl)
//   * outputs beyond a sequence's own length are written as zeros by the same thread (no memset pass).
#include "../../include/vbhem_b200.h"
#include "vbhem_math.cuh"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

namespace vbhem {
int fail_fb(int code, const char *fmt, ...);
}

namespace {

using vbhem::fail_fb;

#define CU(x)                                                                                                 \
  do {                                                                                                        \
    cudaError_t e_ = (x);                                                                                     \
    if (e_ != cudaSuccess)                                                                                    \
      return fail_fb(e_ == cudaErrorMemoryAllocation ? VBHEM_ENOMEM : VBHEM_ECUDA, "%s: %s (%s:%d)", #x,      \
                     cudaGetErrorString(e_), __FILE__, __LINE__);                                             \
  } while (0)

constexpr int kFbThreads = 128;
#ifndef VB_FB_MINB
#define VB_FB_MINB 4
#endif
#ifndef VB_FB_LEAN
#define VB_FB_LEAN 0       // 1: k_fb takes c(t+1) from the backward sums and log c from the lean log (see the header note)
#endif
#ifndef VB_FB_BLOCKS
#define VB_FB_BLOCKS 0     // > 0: cap the resident blocks per SM (L2 residency of the parked alpha_hat)
#endif
constexpr int kFbMaxK = 8, kFbMaxDim = 3;         // template instantiations; beyond: the generic kernel
constexpr int kFbGenK = 32, kFbGenDim = 8;

template <int K, int DIM>
struct FbPar {
  double m[K][DIM];          // m(:,k)
  double W[K][DIM][DIM];     // W[k][i][j] = W(j,i,k): the MEX walks W sequentially (vbhmm_fb_mex.c:378-391)
  double v[K], dob[K];       // v(k), dim / beta(k)
  double lLT[K];             // logLambdaTilde(k)
  double pz1[K];             // exp(logPiTilde)
  double A[K][K];            // exp(logATilde)(from, to)
  double cden;
};

// logrho(k) of one observation, in the reference's order of operations (vbhmm_fb_mex.c:334-458).
template <int K, int DIM>
__device__ __forceinline__ void emissions(const FbPar<K, DIM> &P, const double (&x)[DIM], double (&lr)[K]) {
#pragma unroll
  for (int k = 0; k < K; k++) {
    double diff[DIM];
#pragma unroll
    for (int i = 0; i < DIM; i++) diff[i] = x[i] - P.m[k][i];
    double mterm = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; i++) {
      double tmp = 0.0;
#pragma unroll
      for (int j = 0; j < DIM; j++) tmp += P.W[k][i][j] * diff[j];
      mterm += tmp * diff[i];
    }
    const double delta = P.dob[k] + P.v[k] * mterm;
    lr[k] = 0.5 * (P.lLT[k] - delta) - P.cden;
  }
}

// 2^(j/256) table of the lattice kernels' exp and the (1/c, -log(1/c)/u) table of their log (vbhem_math.cuh), staged
// per block
__device__ const double g_fb_exp_table[VBHEM_EXP_BINS] = VBHEM_EXP_TABLE;
__device__ const double g_fb_log_table[2 * VBHEM_LOG_BINS] = VBHEM_LOG_TABLE;

// exp table | log table in shared memory; returns the two views
template <int THREADS>
__device__ __forceinline__ void stage_tables(double *s_tab, vbhem::tab_t &etab, vbhem::tab_t &ltab) {
  for (int j = threadIdx.x; j < VBHEM_EXP_BINS; j += THREADS) s_tab[j] = g_fb_exp_table[j];
  for (int j = threadIdx.x; j < 2 * VBHEM_LOG_BINS; j += THREADS) s_tab[VBHEM_EXP_BINS + j] = g_fb_log_table[j];
  __syncthreads();
  etab = (vbhem::tab_t)__cvta_generic_to_shared(s_tab);
  ltab = (vbhem::tab_t)__cvta_generic_to_shared(s_tab + VBHEM_EXP_BINS);
}
constexpr int kFbTabDoubles = VBHEM_EXP_BINS + 2 * VBHEM_LOG_BINS;

// p(k) = exp(logrho(k) - max_k logrho): arguments are <= 0, so the lean exp of the lattice kernels applies
// (8 fp64 instructions against libdevice's ~19): y = x / u, e^x = mant 2^e, e <= 0; below 2^-1000 it returns
// ~1e-301 where libdevice returns 0 -- both nothing next to the state that holds the maximum (p = 1).
template <int K>
__device__ __forceinline__ double shifted_exp(const double (&lr)[K], double (&p)[K], vbhem::tab_t etab) {
  double mx = lr[0];
#pragma unroll
  for (int k = 1; k < K; k++) mx = lr[k] > mx ? lr[k] : mx;
#pragma unroll
  for (int k = 0; k < K; k++) {
    double y = (lr[k] - mx) * vbhem::kPerUnit;
    y = y > -4.0e8 ? y : -4.0e8;                  // exp_split's domain; far below the 2^-1000 clamp anyway
    int e;
    const double mant = vbhem::exp_split<1>(y, etab, e);
    p[k] = vbhem::exp_scale(mant, e);
  }
  return mx;
}

// Delta(k) = (alpha A)(k) p(k) and its sum c: the forward step, also the backward pass's way of
// getting c_{t+1} back (same inputs, same order => same bits).
template <int K, int DIM>
__device__ __forceinline__ double forward_step(const FbPar<K, DIM> &P, const double (&alpha)[K], const double (&p)[K],
                                               double (&delta)[K]) {
  double c = 0.0;
#pragma unroll
  for (int k = 0; k < K; k++) {
    double tmp = 0.0;
#pragma unroll
    for (int j = 0; j < K; j++) tmp += alpha[j] * P.A[j][k];
    delta[k] = tmp * p[k];
    c += delta[k];
  }
  return c;
}

template <int K, int DIM>
__global__ void __launch_bounds__(kFbThreads, VB_FB_MINB)
k_fb(const __grid_constant__ FbPar<K, DIM> P, int N, int maxT, const int *__restrict__ Tn,
     const double *__restrict__ X, double *__restrict__ logrho, double *__restrict__ gamma,
     double *__restrict__ xi_sum, double *__restrict__ phi) {
#if VB_FB_LEAN
  __shared__ __align__(16) double s_tab[kFbTabDoubles];
  vbhem::tab_t etab, ltab;
  stage_tables<kFbThreads>(s_tab, etab, ltab);
#else
  __shared__ __align__(16) double s_etab[VBHEM_EXP_BINS];
  for (int j = threadIdx.x; j < VBHEM_EXP_BINS; j += kFbThreads) s_etab[j] = g_fb_exp_table[j];
  __syncthreads();
  const vbhem::tab_t etab = (vbhem::tab_t)__cvta_generic_to_shared(s_etab);
#endif
  const int n = blockIdx.x * kFbThreads + threadIdx.x;
  if (n >= N) return;
  const int T = Tn[n];
  const size_t NK = (size_t)N * K, ND = (size_t)N * DIM;
  const double *xp = X + (size_t)DIM * n;
  double *lp = logrho ? logrho + (size_t)K * n : nullptr, *gp = gamma + (size_t)K * n;
  double alpha[K], p[K], lr[K], x[DIM], xn[DIM];
  double sumlogc = 0.0, summax = 0.0;

  // ---- forward (vbhmm_fb.m:316-347).  The next observation is requested one step ahead; logrho is a
  // streaming store (never read again here), alpha_hat a normal one (read back below, ideally from L2).
#pragma unroll
  for (int i = 0; i < DIM; i++) xn[i] = T > 0 ? xp[i] : 0.0;
  for (int t = 0; t < T; t++) {
#pragma unroll
    for (int i = 0; i < DIM; i++) x[i] = xn[i];
    if (t + 1 < T) {
#pragma unroll
      for (int i = 0; i < DIM; i++) xn[i] = xp[ND * (t + 1) + i];
    }
    emissions<K, DIM>(P, x, lr);
    if (logrho) {
#pragma unroll
      for (int k = 0; k < K; k++) __stcs(lp + NK * t + k, lr[k]);
    }
    summax += shifted_exp<K>(lr, p, etab);
    double delta[K], c;
    if (t == 0) {
      c = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) { delta[k] = P.pz1[k] * p[k]; c += delta[k]; }
    } else {
      c = forward_step<K, DIM>(P, alpha, p, delta);
    }
    const double rc = 1.0 / c;
#pragma unroll
    for (int k = 0; k < K; k++) { alpha[k] = delta[k] * rc; gp[NK * t + k] = alpha[k]; }
#if VB_FB_LEAN
    sumlogc += vbhem::log_shift<1>(c, ltab, 0);
  }
  phi[n] = sumlogc * vbhem::kUnit + summax;
#else
    sumlogc += log(c);
  }
  phi[n] = sumlogc + summax;
#endif
  // zeros beyond the sequence's own length (the reference's outputs start as zeros)
  for (int t = T; t < maxT; t++) {
#pragma unroll
    for (int k = 0; k < K; k++) { if (logrho) __stcs(lp + NK * t + k, 0.0); __stcs(gp + NK * t + k, 0.0); }
  }

  // ---- backward, responsibilities, summed pair marginals (vbhmm_fb.m:349-377).  The K x K accumulator of
  // each thread lives in shared memory ([j][k][thread]: conflict-free), which keeps the register count low
  // enough for three to four resident blocks per SM.
  extern __shared__ double xis[];
  double *xi = xis + threadIdx.x;
  double beta[K], an[K];
#pragma unroll
  for (int j = 0; j < K; j++) {
    beta[j] = 1.0;
#pragma unroll
    for (int k = 0; k < K; k++) xi[(j * K + k) * kFbThreads] = 0.0;
  }
  // gamma(:, T) = alpha(:, T) is already in place; x holds the last observation
  if (T >= 2) {
#pragma unroll
    for (int k = 0; k < K; k++) an[k] = gp[NK * (T - 2) + k];
  }
  for (int t = T - 2; t >= 0; t--) {
    // x = observation t+1 (from the previous iteration / the forward pass), alpha = alpha_hat(t)
#pragma unroll
    for (int k = 0; k < K; k++) alpha[k] = an[k];
#pragma unroll
    for (int i = 0; i < DIM; i++) xn[i] = xp[ND * t + i];             // observation t, used next iteration
    if (t > 0) {
#pragma unroll
      for (int k = 0; k < K; k++) an[k] = gp[NK * (t - 1) + k];      // alpha_hat(t-1), parked by the forward pass
    }
    emissions<K, DIM>(P, x, lr);
    shifted_exp<K>(lr, p, etab);                                      // p(x_{t+1} | z) / max
    double bpi[K];
#if VB_FB_LEAN
#pragma unroll
    for (int k = 0; k < K; k++) bpi[k] = beta[k] * p[k];
    double c = 0.0;   // c(t+1) = sum_j alpha_hat_t(j) eta(j): sum_k alpha_hat(k) beta_hat(k) = 1 at every step
#pragma unroll
    for (int j = 0; j < K; j++) {
      double eta = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) eta += bpi[k] * P.A[j][k];
      beta[j] = eta;
      c += alpha[j] * eta;
    }
    const double rc = 1.0 / c;
#pragma unroll
    for (int j = 0; j < K; j++) beta[j] *= rc;
#else
    double delta[K];
    const double rc = 1.0 / forward_step<K, DIM>(P, alpha, p, delta);   // 1 / c(t+1)
#pragma unroll
    for (int k = 0; k < K; k++) bpi[k] = beta[k] * p[k];
#pragma unroll
    for (int j = 0; j < K; j++) {
      double eta = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) eta += bpi[k] * P.A[j][k];
      beta[j] = eta * rc;
    }
#endif
#pragma unroll
    for (int k = 0; k < K; k++) __stcs(gp + NK * t + k, alpha[k] * beta[k]);
#pragma unroll
    for (int j = 0; j < K; j++) {
      const double aj = alpha[j] * rc;
#pragma unroll
      for (int k = 0; k < K; k++) xi[(j * K + k) * kFbThreads] += P.A[j][k] * (aj * bpi[k]);
    }
#pragma unroll
    for (int i = 0; i < DIM; i++) x[i] = xn[i];
  }
  double *xo = xi_sum + (size_t)K * K * n;
#pragma unroll
  for (int k = 0; k < K; k++)
#pragma unroll
    for (int j = 0; j < K; j++) __stcs(xo + j + K * k, xi[(j * K + k) * kFbThreads]);
}

// ---- M-step statistics of vbhmm_em.m:158-246 and the fb-dependent lower-bound terms (vbhmm_em_lb.m:214,217)
// straight from the responsibilities in HBM, so that gamma_all / logrho_Saved never cross PCIe:
//   Nk1(k) = sum_n gamma(k,n,1), Nk(k) = sum_{n,t} gamma, M = sum_n xi_sum,
//   first / second moments of the observations about the CURRENT means m_k (no cancellation),
//   Lt63 = sum gamma .* logrho (logrho recomputed from X), Lt64 = sum phi_norm.
// One thread per (n, t) element, grid-stride; fixed grid => fixed-order (deterministic) sums.
constexpr int kMomThreads = 256;
template <int K, int DIM>
struct MomLayout {
  static constexpr int DD = DIM * DIM;
  static constexpr int nk1 = 0, nk = K, sx = 2 * K, sxx = 2 * K + K * DIM, M = 2 * K + K * DIM + K * DD,
                       lt63 = M + K * K, lt64 = lt63 + 1, total = lt64 + 1;
};

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&acc)[NV], double *out) {
  __shared__ double sh[kMomThreads / 32][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < NV; v++) {
    double a = acc[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) sh[warp][v] = a;
  }
  __syncthreads();
  for (int v = threadIdx.x; v < NV; v += kMomThreads) {
    double a = 0.0;
    for (int w = 0; w < kMomThreads / 32; w++) a += sh[w][v];
    out[v] = a;
  }
}

template <int K, int DIM>
__global__ void __launch_bounds__(kMomThreads)
k_fb_moments(const __grid_constant__ FbPar<K, DIM> P, int N, int maxT, const double *__restrict__ X,
             const double *__restrict__ gamma, const double *__restrict__ xi_sum, const double *__restrict__ phi,
             double *__restrict__ partial) {
  using L = MomLayout<K, DIM>;
  double acc[L::total];
#pragma unroll
  for (int v = 0; v < L::total; v++) acc[v] = 0.0;
  const long long tot = (long long)N * maxT, stride = (long long)gridDim.x * kMomThreads;
  for (long long e = (long long)blockIdx.x * kMomThreads + threadIdx.x; e < tot; e += stride) {
    double g[K], x[DIM], lr[K];
#pragma unroll
    for (int k = 0; k < K; k++) g[k] = __ldcs(gamma + (size_t)K * e + k);
#pragma unroll
    for (int i = 0; i < DIM; i++) x[i] = X[(size_t)DIM * e + i];
    emissions<K, DIM>(P, x, lr);
    const bool first = e < N;                                   // t == 0
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (first) acc[L::nk1 + k] += g[k];
      acc[L::nk + k] += g[k];
      acc[L::lt63] += g[k] * lr[k];
      double diff[DIM];
#pragma unroll
      for (int i = 0; i < DIM; i++) diff[i] = x[i] - P.m[k][i];
#pragma unroll
      for (int i = 0; i < DIM; i++) {
        const double gd = g[k] * diff[i];
        acc[L::sx + k * DIM + i] += gd;
#pragma unroll
        for (int j = 0; j < DIM; j++) acc[L::sxx + (k * DIM + i) * DIM + j] += gd * diff[j];
      }
    }
  }
  for (long long n = (long long)blockIdx.x * kMomThreads + threadIdx.x; n < N; n += stride) {
    acc[L::lt64] += phi[n];
#pragma unroll
    for (int q = 0; q < K * K; q++) acc[L::M + q] += __ldcs(xi_sum + (size_t)K * K * n + q);   // [from + K * to]
  }
  block_reduce_store<L::total>(acc, partial + (size_t)L::total * blockIdx.x);
}

// Fixed-order sum over the blocks, then the normalisation of vbhmm_em.m:161-246.
// out: Nk1[K] | Nk[K] | M[K,K] | xbar[K,dim] (row k) | S[dim,dim,K] | Lt63 | Lt64
template <int K, int DIM>
__global__ void k_fb_moments_finish(const __grid_constant__ FbPar<K, DIM> P, int nblocks,
                                    const double *__restrict__ partial, double *__restrict__ out) {
  using L = MomLayout<K, DIM>;
  __shared__ double s[L::total];
  for (int v = threadIdx.x; v < L::total; v += blockDim.x) {
    double a = 0.0;
    for (int b = 0; b < nblocks; b++) a += partial[(size_t)L::total * b + v];
    s[v] = a;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  double *o = out;
  for (int k = 0; k < K; k++) o[k] = s[L::nk1 + k] + 1e-50;                       // vbhmm_em.m:163
  o += K;
  for (int k = 0; k < K; k++) o[k] = s[L::nk + k] + 1e-50;                        // :171-172
  o += K;
  for (int q = 0; q < K * K; q++) o[q] = s[L::M + q];                             // :178
  o += K * K;
  double *xb = o, *S = o + K * DIM;
  for (int k = 0; k < K; k++) {
    const double nk = s[L::nk + k], Nk = nk + 1e-50;
    double del[DIM];                                                              // xbar - m_k
    for (int i = 0; i < DIM; i++) {
      xb[k + K * i] = (P.m[k][i] * nk + s[L::sx + k * DIM + i]) / Nk;             // :216-219, [K x dim] column-major
      del[i] = xb[k + K * i] - P.m[k][i];
    }
    for (int i = 0; i < DIM; i++)
      for (int j = 0; j < DIM; j++)                                               // :240-245 about xbar
        S[i + DIM * (j + DIM * k)] = (s[L::sxx + (k * DIM + i) * DIM + j] - s[L::sx + k * DIM + i] * del[j] -
                                      del[i] * s[L::sx + k * DIM + j] + nk * del[i] * del[j]) / Nk;
  }
  o = S + DIM * DIM * K;
  o[0] = s[L::lt63];
  o[1] = s[L::lt64];
}

// Any K <= 32, dim <= 8: the same recursion with the model in global memory and local arrays.
struct FbParGen { const double *m, *W, *v, *dob, *lLT, *pz1, *A; double cden; int K, dim; };

__device__ double gen_emit(const FbParGen &P, const double *x, double *lr, double *p) {
  const int K = P.K, dim = P.dim;
  double mx = 0.0;
  for (int k = 0; k < K; k++) {
    double diff[kFbGenDim];
    for (int i = 0; i < dim; i++) diff[i] = x[i] - P.m[i + dim * k];
    double mterm = 0.0;
    for (int i = 0; i < dim; i++) {
      double tmp = 0.0;
      for (int j = 0; j < dim; j++) tmp += P.W[j + dim * (i + dim * k)] * diff[j];
      mterm += tmp * diff[i];
    }
    lr[k] = 0.5 * (P.lLT[k] - (P.dob[k] + P.v[k] * mterm)) - P.cden;
    if (k == 0 || lr[k] > mx) mx = lr[k];
  }
  for (int k = 0; k < K; k++) p[k] = exp(lr[k] - mx);
  return mx;
}
__device__ double gen_forward(const FbParGen &P, const double *alpha, const double *p, double *delta) {
  double c = 0.0;
  for (int k = 0; k < P.K; k++) {
    double tmp = 0.0;
    for (int j = 0; j < P.K; j++) tmp += alpha[j] * P.A[j + P.K * k];
    delta[k] = tmp * p[k];
    c += delta[k];
  }
  return c;
}

__global__ void __launch_bounds__(kFbThreads)
k_fb_generic(FbParGen P, int N, int maxT, const int *__restrict__ Tn, const double *__restrict__ X,
             double *__restrict__ logrho, double *__restrict__ gamma, double *__restrict__ xi_sum,
             double *__restrict__ phi) {
  const int n = blockIdx.x * kFbThreads + threadIdx.x;
  if (n >= N) return;
  const int K = P.K, dim = P.dim, T = Tn[n];
  const size_t NK = (size_t)N * K, ND = (size_t)N * dim;
  const double *xp = X + (size_t)dim * n;
  double *lp = logrho + (size_t)K * n, *gp = gamma + (size_t)K * n, *xo = xi_sum + (size_t)K * K * n;
  double alpha[kFbGenK], p[kFbGenK], lr[kFbGenK], delta[kFbGenK], beta[kFbGenK], bpi[kFbGenK], x[kFbGenDim];
  double sumlogc = 0.0, summax = 0.0;
  for (int t = 0; t < T; t++) {
    for (int i = 0; i < dim; i++) x[i] = xp[ND * t + i];
    summax += gen_emit(P, x, lr, p);
    for (int k = 0; k < K; k++) lp[NK * t + k] = lr[k];
    double c;
    if (t == 0) { c = 0.0; for (int k = 0; k < K; k++) { delta[k] = P.pz1[k] * p[k]; c += delta[k]; } }
    else c = gen_forward(P, alpha, p, delta);
    const double rc = 1.0 / c;
    for (int k = 0; k < K; k++) { alpha[k] = delta[k] * rc; gp[NK * t + k] = alpha[k]; }
    sumlogc += log(c);
  }
  phi[n] = sumlogc + summax;
  for (int t = T; t < maxT; t++)
    for (int k = 0; k < K; k++) { lp[NK * t + k] = 0.0; gp[NK * t + k] = 0.0; }
  for (int q = 0; q < K * K; q++) xo[q] = 0.0;
  for (int k = 0; k < K; k++) beta[k] = 1.0;
  for (int t = T - 2; t >= 0; t--) {
    for (int i = 0; i < dim; i++) x[i] = xp[ND * (t + 1) + i];
    gen_emit(P, x, lr, p);
    for (int k = 0; k < K; k++) alpha[k] = gp[NK * t + k];
    const double rc = 1.0 / gen_forward(P, alpha, p, delta);
    for (int k = 0; k < K; k++) bpi[k] = beta[k] * p[k];
    for (int j = 0; j < K; j++) {
      double eta = 0.0;
      for (int k = 0; k < K; k++) eta += bpi[k] * P.A[j + K * k];
      delta[j] = eta * rc;
    }
    for (int k = 0; k < K; k++) { beta[k] = delta[k]; gp[NK * t + k] = alpha[k] * beta[k]; }
    for (int k = 0; k < K; k++)
      for (int j = 0; j < K; j++) xo[j + K * k] += P.A[j + K * k] * ((alpha[j] * rc) * bpi[k]);
  }
}

// flat[offs[n] + t + d * T_n]  ->  X[d + dim * (n + N * t)]
__global__ void k_repack(int N, int dim, int maxT, const int *Tn, const long long *offs, const double *flat,
                         double *X) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int T = Tn[n];
  const double *src = flat + offs[n];
  for (int t = 0; t < maxT; t++)
    for (int d = 0; d < dim; d++) X[d + (size_t)dim * (n + (size_t)N * t)] = t < T ? src[t + (size_t)d * T] : 0.0;
}

struct Buf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 8);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T *as() const { return static_cast<T *>(p); }
};

}  // namespace

struct vbhmm_data {
  int device = 0, N = 0, dim = 0, maxT = 0;
  long long total = 0;         // sum_n T_n
  cudaStream_t st = nullptr;
  Buf Tn, X, par, logrho, gamma, xi, phi, mom_partial, mom_out;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  double last_ms = 0.0;
  int lastK = 0;
};

namespace {

// stats == false: the literal pass (all four outputs).  stats == true: no logrho store, then the moment kernels
// (results in h->mom_out, layout of k_fb_moments_finish).
template <int K, int DIM>
int launch_fb_t(vbhmm_data *h, const double *m, const double *W, const double *v, const double *beta,
                const double *lLT, double cden, const double *pz1, const double *A, bool stats) {
  FbPar<K, DIM> P;
  for (int k = 0; k < K; k++) {
    for (int i = 0; i < DIM; i++) {
      P.m[k][i] = m[i + DIM * k];
      for (int j = 0; j < DIM; j++) P.W[k][i][j] = W[j + DIM * (i + DIM * k)];
    }
    P.v[k] = v[k]; P.dob[k] = DIM / beta[k]; P.lLT[k] = lLT[k]; P.pz1[k] = pz1[k];
    for (int j = 0; j < K; j++) P.A[k][j] = A[k + K * j];
  }
  P.cden = cden;
  const int blocks = (h->N + kFbThreads - 1) / kFbThreads;
  size_t smem = sizeof(double) * K * K * kFbThreads;
#if VB_FB_BLOCKS > 0
  if (smem < (size_t)(220 * 1024) / VB_FB_BLOCKS - 1024) smem = (size_t)(220 * 1024) / VB_FB_BLOCKS - 1024;
#endif
  cudaFuncSetAttribute(k_fb<K, DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_fb<K, DIM><<<blocks, kFbThreads, smem, h->st>>>(P, h->N, h->maxT, h->Tn.as<int>(), h->X.as<double>(),
                                                 stats ? nullptr : h->logrho.as<double>(), h->gamma.as<double>(),
                                                 h->xi.as<double>(), h->phi.as<double>());
  if (!stats) return 1;
  using L = MomLayout<K, DIM>;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  const long long tot = (long long)h->N * h->maxT;
  long long want = (tot + kMomThreads - 1) / kMomThreads;
  const int mb = (int)(want < 1 ? 1 : (want > 8LL * sms ? 8LL * sms : want));
  if (h->mom_partial.ensure(sizeof(double) * (size_t)L::total * mb) != cudaSuccess ||
      h->mom_out.ensure(sizeof(double) * L::total) != cudaSuccess)
    return -1;
  k_fb_moments<K, DIM><<<mb, kMomThreads, 0, h->st>>>(P, h->N, h->maxT, h->X.as<double>(), h->gamma.as<double>(),
                                                      h->xi.as<double>(), h->phi.as<double>(),
                                                      h->mom_partial.as<double>());
  k_fb_moments_finish<K, DIM><<<1, 128, 0, h->st>>>(P, mb, h->mom_partial.as<double>(), h->mom_out.as<double>());
  return 3;
}

template <int K>
int launch_fb_k(vbhmm_data *h, const double *m, const double *W, const double *v, const double *beta,
                const double *lLT, double cden, const double *pz1, const double *A, bool stats) {
  switch (h->dim) {
    case 1: return launch_fb_t<K, 1>(h, m, W, v, beta, lLT, cden, pz1, A, stats);
    case 2: return launch_fb_t<K, 2>(h, m, W, v, beta, lLT, cden, pz1, A, stats);
    case 3: return launch_fb_t<K, 3>(h, m, W, v, beta, lLT, cden, pz1, A, stats);
  }
  return 0;
}

int run_fb(vbhmm_data *h, int K, const double *m, const double *W, const double *v, const double *beta,
           const double *lLT, double cden, const double *pz1, const double *A, bool stats = false) {
  const int N = h->N, dim = h->dim;
  const size_t nkt = (size_t)K * N * (h->maxT > 0 ? h->maxT : 1);
  if (stats && !(K <= kFbMaxK && dim <= kFbMaxDim))
    return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: vbhmm_fb_stats takes K <= %d, dim <= %d (use vbhmm_fb)",
                   kFbMaxK, kFbMaxDim);
  if (!stats) CU(h->logrho.ensure(sizeof(double) * nkt));
  CU(h->gamma.ensure(sizeof(double) * nkt));
  CU(h->xi.ensure(sizeof(double) * (size_t)K * K * N));
  CU(h->phi.ensure(sizeof(double) * (size_t)N));
  CU(cudaEventRecord(h->ev[0], h->st));
  int launched = 0;
  if (K <= kFbMaxK && dim <= kFbMaxDim) {
    switch (K) {
      case 1: launched = launch_fb_k<1>(h, m, W, v, beta, lLT, cden, pz1, A, stats); break;
      case 2: launched = launch_fb_k<2>(h, m, W, v, beta, lLT, cden, pz1, A, stats); break;
      case 3: launched = launch_fb_k<3>(h, m, W, v, beta, lLT, cden, pz1, A, stats); break;
      case 4: launched = launch_fb_k<4>(h, m, W, v, beta, lLT, cden, pz1, A, stats); break;
      case 5: launched = launch_fb_k<5>(h, m, W, v, beta, lLT, cden, pz1, A, stats); break;
      case 6: launched = launch_fb_k<6>(h, m, W, v, beta, lLT, cden, pz1, A, stats); break;
      case 7: launched = launch_fb_k<7>(h, m, W, v, beta, lLT, cden, pz1, A, stats); break;
      case 8: launched = launch_fb_k<8>(h, m, W, v, beta, lLT, cden, pz1, A, stats); break;
    }
  }
  if (launched < 0) return fail_fb(VBHEM_ENOMEM, "out of device memory for the moment partials");
  if (!launched) {
    // model to HBM: m | W | v | dob | lLT | pz1 | A
    const size_t nd = (size_t)dim * K + (size_t)dim * dim * K + 4 * (size_t)K + (size_t)K * K;
    std::vector<double> hp(nd);
    double *q = hp.data();
    memcpy(q, m, sizeof(double) * dim * K); q += (size_t)dim * K;
    memcpy(q, W, sizeof(double) * dim * dim * K); q += (size_t)dim * dim * K;
    memcpy(q, v, sizeof(double) * K); q += K;
    for (int k = 0; k < K; k++) q[k] = dim / beta[k];
    q += K;
    memcpy(q, lLT, sizeof(double) * K); q += K;
    memcpy(q, pz1, sizeof(double) * K); q += K;
    memcpy(q, A, sizeof(double) * K * K);
    CU(h->par.ensure(sizeof(double) * nd));
    CU(cudaMemcpyAsync(h->par.p, hp.data(), sizeof(double) * nd, cudaMemcpyHostToDevice, h->st));
    CU(cudaStreamSynchronize(h->st));     // hp goes out of scope
    FbParGen P;
    const double *d = h->par.as<double>();
    P.m = d; d += (size_t)dim * K;
    P.W = d; d += (size_t)dim * dim * K;
    P.v = d; d += K;
    P.dob = d; d += K;
    P.lLT = d; d += K;
    P.pz1 = d; d += K;
    P.A = d;
    P.cden = cden; P.K = K; P.dim = dim;
    k_fb_generic<<<(N + kFbThreads - 1) / kFbThreads, kFbThreads, 0, h->st>>>(
        P, N, h->maxT, h->Tn.as<int>(), h->X.as<double>(), h->logrho.as<double>(), h->gamma.as<double>(),
        h->xi.as<double>(), h->phi.as<double>());
  }
  CU(cudaEventRecord(h->ev[1], h->st));
  CU(cudaGetLastError());
  h->lastK = stats ? -K : K;   // after a statistics pass logrho is not available for vbhmm_fb_fetch
  return VBHEM_OK;
}

int check_model(vbhmm_data *h, int K, const double *m, const double *W, const double *v, const double *beta,
                const double *lLT, const double *pz1, const double *A) {
  if (!h) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: data handle is NULL");
  if (K < 1 || K > kFbGenK) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: K must be 1..%d", kFbGenK);
  if (!m || !W || !v || !beta || !lLT || !pz1 || !A)
    return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: a model array is NULL");
  return VBHEM_OK;
}

}  // namespace

extern "C" {

int vbhmm_pack_data(int device, int N, int dim, const int32_t *T, const long long *offs, const double *flat,
                    vbhmm_data_t **out) {
  if (!out) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: out is NULL");
  *out = nullptr;
  if (N < 1 || dim < 1 || dim > kFbGenDim)
    return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: N >= 1 and dim in 1..%d required", kFbGenDim);
  if (!T || !offs || !flat) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: data is not a cell array");
  int maxT = 0;
  long long total = 0, span = 0;
  for (int n = 0; n < N; n++) {
    if (T[n] < 0) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: T[%d] < 0", n);
    if (T[n] > maxT) maxT = T[n];
    total += T[n];
    const long long end = offs[n] + (long long)T[n] * dim;
    if (offs[n] < 0) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: offs[%d] < 0", n);
    if (end > span) span = end;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail_fb(VBHEM_ENODEVICE, "no CUDA device (this library has no CPU path)");
  if (device < 0 || device >= ndev) return fail_fb(VBHEM_EINVAL, "device %d out of range (have %d)", device, ndev);
  CU(cudaSetDevice(device));
  vbhmm_data *h = new (std::nothrow) vbhmm_data;
  if (!h) return fail_fb(VBHEM_ENOMEM, "out of host memory");
  h->device = device; h->N = N; h->dim = dim; h->maxT = maxT; h->total = total;
  Buf dflat, doffs;
  auto bail = [&](int code) { dflat.release(); doffs.release(); vbhmm_free_data(h); return code; };
#define CUB(x)                                                                                               \
  do {                                                                                                       \
    cudaError_t e_ = (x);                                                                                    \
    if (e_ != cudaSuccess)                                                                                   \
      return bail(fail_fb(e_ == cudaErrorMemoryAllocation ? VBHEM_ENOMEM : VBHEM_ECUDA, "%s: %s", #x,        \
                          cudaGetErrorString(e_)));                                                          \
  } while (0)
  CUB(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
  CUB(cudaEventCreate(&h->ev[0]));
  CUB(cudaEventCreate(&h->ev[1]));
  CUB(h->Tn.ensure(sizeof(int) * (size_t)N));
  CUB(h->X.ensure(sizeof(double) * (size_t)dim * N * (maxT > 0 ? maxT : 1)));
  CUB(dflat.ensure(sizeof(double) * (size_t)(span > 0 ? span : 1)));
  CUB(doffs.ensure(sizeof(long long) * (size_t)N));
  CUB(cudaMemcpyAsync(h->Tn.p, T, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, h->st));
  CUB(cudaMemcpyAsync(doffs.p, offs, sizeof(long long) * (size_t)N, cudaMemcpyHostToDevice, h->st));
  if (span > 0) CUB(cudaMemcpyAsync(dflat.p, flat, sizeof(double) * (size_t)span, cudaMemcpyHostToDevice, h->st));
  k_repack<<<(N + 127) / 128, 128, 0, h->st>>>(N, dim, maxT, h->Tn.as<int>(), doffs.as<long long>(),
                                               dflat.as<double>(), h->X.as<double>());
  CUB(cudaStreamSynchronize(h->st));
  CUB(cudaGetLastError());
#undef CUB
  dflat.release();
  doffs.release();
  *out = h;
  return VBHEM_OK;
}

int vbhmm_free_data(vbhmm_data_t *h) {
  if (!h) return VBHEM_OK;
  cudaSetDevice(h->device);
  if (h->st) { cudaStreamSynchronize(h->st); cudaStreamDestroy(h->st); }
  for (auto &e : h->ev) if (e) cudaEventDestroy(e);
  Buf *all[] = {&h->Tn, &h->X, &h->par, &h->logrho, &h->gamma, &h->xi, &h->phi, &h->mom_partial, &h->mom_out};
  for (Buf *b : all) b->release();
  delete h;
  return VBHEM_OK;
}

int vbhmm_data_shape(vbhmm_data_t *h, int *N, int *dim, int *maxT, long long *total) {
  if (!h) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: data handle is NULL");
  if (N) *N = h->N;
  if (dim) *dim = h->dim;
  if (maxT) *maxT = h->maxT;
  if (total) *total = h->total;
  return VBHEM_OK;
}

int vbhmm_fb_dev(vbhmm_data_t *h, int K, const double *m, const double *W, const double *v, const double *beta,
                 const double *logLambdaTilde, double const_denominator, const double *pz1, const double *tpztzt1) {
  int rc = check_model(h, K, m, W, v, beta, logLambdaTilde, pz1, tpztzt1);
  if (rc) return rc;
  CU(cudaSetDevice(h->device));
  return run_fb(h, K, m, W, v, beta, logLambdaTilde, const_denominator, pz1, tpztzt1);
}

int vbhmm_fb_fetch(vbhmm_data_t *h, double *logrho_Saved, double *gamma_all, double *xi_sum, double *phi_norm) {
  if (!h) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: data handle is NULL");
  if (h->lastK < 1) return fail_fb(VBHEM_EINVAL, "no literal forward-backward pass (vbhmm_fb / vbhmm_fb_dev) has run last");
  CU(cudaSetDevice(h->device));
  const int K = h->lastK;
  const size_t nkt = (size_t)K * h->N * h->maxT;
  if (logrho_Saved && nkt) CU(cudaMemcpyAsync(logrho_Saved, h->logrho.p, sizeof(double) * nkt, cudaMemcpyDeviceToHost, h->st));
  if (gamma_all && nkt) CU(cudaMemcpyAsync(gamma_all, h->gamma.p, sizeof(double) * nkt, cudaMemcpyDeviceToHost, h->st));
  if (xi_sum) CU(cudaMemcpyAsync(xi_sum, h->xi.p, sizeof(double) * (size_t)K * K * h->N, cudaMemcpyDeviceToHost, h->st));
  if (phi_norm) CU(cudaMemcpyAsync(phi_norm, h->phi.p, sizeof(double) * (size_t)h->N, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  return VBHEM_OK;
}

int vbhmm_fb(vbhmm_data_t *h, int K, const double *m, const double *W, const double *v, const double *beta,
             const double *logLambdaTilde, double const_denominator, const double *pz1, const double *tpztzt1,
             double *logrho_Saved, double *gamma_all, double *xi_sum, double *phi_norm) {
  int rc = vbhmm_fb_dev(h, K, m, W, v, beta, logLambdaTilde, const_denominator, pz1, tpztzt1);
  if (rc) return rc;
  return vbhmm_fb_fetch(h, logrho_Saved, gamma_all, xi_sum, phi_norm);
}

int vbhmm_fb_stats(vbhmm_data_t *h, int K, const double *m, const double *W, const double *v, const double *beta,
                   const double *logLambdaTilde, double const_denominator, const double *pz1, const double *tpztzt1,
                   double *Nk1, double *Nk, double *M, double *xbar, double *S, double *lb2) {
  int rc = check_model(h, K, m, W, v, beta, logLambdaTilde, pz1, tpztzt1);
  if (rc) return rc;
  CU(cudaSetDevice(h->device));
  rc = run_fb(h, K, m, W, v, beta, logLambdaTilde, const_denominator, pz1, tpztzt1, true);
  if (rc) return rc;
  const int d = h->dim;
  const size_t tot = 2 * (size_t)K + (size_t)K * K + (size_t)K * d + (size_t)d * d * K + 2;
  std::vector<double> host(tot);
  CU(cudaMemcpyAsync(host.data(), h->mom_out.p, sizeof(double) * tot, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  const double *q = host.data();
  if (Nk1) memcpy(Nk1, q, sizeof(double) * K);
  q += K;
  if (Nk) memcpy(Nk, q, sizeof(double) * K);
  q += K;
  if (M) memcpy(M, q, sizeof(double) * K * K);
  q += (size_t)K * K;
  if (xbar) memcpy(xbar, q, sizeof(double) * K * d);
  q += (size_t)K * d;
  if (S) memcpy(S, q, sizeof(double) * d * d * K);
  q += (size_t)d * d * K;
  if (lb2) { lb2[0] = q[0]; lb2[1] = q[1]; }
  return VBHEM_OK;
}

int vbhmm_fb_last_ms(vbhmm_data_t *h, double *ms) {
  if (!h || !ms) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: NULL argument");
  CU(cudaSetDevice(h->device));
  CU(cudaEventSynchronize(h->ev[1]));
  float f = 0.f;
  CU(cudaEventElapsedTime(&f, h->ev[0], h->ev[1]));
  *ms = f;
  return VBHEM_OK;
}

}  // extern "C"

#include "vbhmm_batch.inl"
#include "vbhmm_gmm.inl"
