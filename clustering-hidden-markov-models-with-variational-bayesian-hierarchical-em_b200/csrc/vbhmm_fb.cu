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
// Unlike the VBHEM E-step this path has 5 KB of compulsory traffic per sequence against ~ 14 000 fp64 instructions:
// HBM, the fp64 pipe, issue slots and the LSU all matter.  Design (DESIGN.md 3.7 has the measurements):
//   * one thread per sequence, a warp takes 32 consecutive sequences at a time from a device counter (persistent
//     grid); the observations are repacked once into X[dim, N, maxT] and arrive through a per-lane cp.async ring;
//   * the model (m, the folded quadratic forms of the emission densities, p(z1), p(z_t | z_t-1)) travels as a
//     by-value kernel parameter: it sits in the constant bank and every use is a DFMA operand, never a load;
//   * alpha_hat waits for the backward pass ON CHIP: the last steps in shared memory, the rest in an L2-resident
//     scratch buffer whose size depends on the number of resident threads only, read back by cp.async;
//   * the backward pass does not evaluate the emissions again: p_{t+1} / c_{t+1} = alpha_hat_{t+1} / (alpha_hat_t A);
//   * every output row (the reference's [K, N, maxT] layout) is transposed through shared memory by the warp and
//     written as full 256-byte runs; rows beyond a sequence's own length are written as zeros by the same warp
//     (no memset pass).
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

#ifndef VB_FB_THREADS
#define VB_FB_THREADS 128
#endif
constexpr int kFbThreads = VB_FB_THREADS;
#ifndef VB_FB_MINB
#define VB_FB_MINB 2       // resident CTAs per SM: 256 sequences in flight per SM, <= 255 registers (Xi lives in registers)
#endif
#ifndef VB_FB_RING
#define VB_FB_RING 4       // slots of the observation ring (cp.async RING - 1 steps ahead); a power of two
#endif
#ifndef VB_FB_AHEAD
#define VB_FB_AHEAD 4      // the backward pass copies parked alpha_hat from the scratch buffer this many steps ahead
#endif
#ifndef VB_FB_HINT
#define VB_FB_HINT 1       // 1: L2 evict_last on the parked alpha_hat (scripts/sweep_fb.sh A/Bs it)
#endif
constexpr int kFbMaxK = 8, kFbMaxDim = 3;         // template instantiations; beyond: the generic kernel
constexpr int kFbGenK = 32, kFbGenDim = 8;

template <int K, int DIM>
struct FbPar {
  static constexpr int NQ = DIM * (DIM + 1) / 2;
  double m[K][DIM];          // m(:,k)
  double W[K][DIM][DIM];     // W[k][i][j] = W(j,i,k): the MEX walks W sequentially (vbhmm_fb_mex.c:378-391)
  double v[K], dob[K];       // v(k), dim / beta(k)
  double lLT[K];             // logLambdaTilde(k)
  double pz1[K];             // exp(logPiTilde)
  double A[K][K];            // exp(logATilde)(from, to)
  double cden;
  // what the kernels evaluate (fb_finish_par): logrho(k) = c0(k) - sum_{i<=j} q(k,ij) d_i d_j with d = x - m(:,k),
  // q(k,ii) = v W_ii / 2, q(k,ij) = v (W_ij + W_ji) / 2, c0 = (logLambdaTilde - dim/beta) / 2 - const_denominator:
  // vbhmm_fb_mex.c:334-458 with the constants folded -- 7 instead of 11 fp64 instructions per state at dim = 2 and six
  // constants per state instead of ten; differs from the reference's order of operations by rounding only.
  double q[K][NQ], c0[K];
};

template <int K, int DIM>
__host__ __device__ inline void fb_finish_par(FbPar<K, DIM> &P) {
  for (int k = 0; k < K; k++) {
    int n = 0;
    for (int i = 0; i < DIM; i++)
      for (int j = i; j < DIM; j++)
        P.q[k][n++] = i == j ? 0.5 * P.v[k] * P.W[k][i][i] : 0.5 * P.v[k] * (P.W[k][i][j] + P.W[k][j][i]);
    P.c0[k] = 0.5 * (P.lLT[k] - P.dob[k]) - P.cden;
  }
}

// logrho(k) of one observation, in the reference's order of operations (vbhmm_fb_mex.c:334-458).
template <int K, int DIM>
__device__ __forceinline__ void emissions(const FbPar<K, DIM> &P, const double (&x)[DIM], double (&lr)[K]) {
#pragma unroll
  for (int k = 0; k < K; k++) {
    double diff[DIM];
#pragma unroll
    for (int i = 0; i < DIM; i++) diff[i] = x[i] - P.m[k][i];
    double acc = P.c0[k];
    int n = 0;
#pragma unroll
    for (int i = 0; i < DIM; i++) {
      double u = P.q[k][n++] * diff[i];
#pragma unroll
      for (int j = i + 1; j < DIM; j++) u = fma(P.q[k][n++], diff[j], u);
      acc = fma(-u, diff[i], acc);
    }
    lr[k] = acc;
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
    // exp_split's domain: y >= -4e8 units (far below the 2^-1000 clamp anyway).  y <= 0, so its magnitude grows with
    // its high word read as unsigned: one integer minimum instead of a compare and two selects (NaN is clamped, too)
    y = __hiloint2double((int)min((unsigned)__double2hiint(y), 0xC1B7D784u), __double2loint(y));
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

// 1 / x for a normal positive x: MUFU.RCP64H seed + one cubic correction (<= 1 ulp; no special-case branch)
__device__ __forceinline__ double rcp_pos(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);                    // |e| <= 2^-22: r (1 + e + e^2) is exact to 2^-66
  return fma(r, fma(e, e, e), r);
}

// Accesses to the parked alpha_hat in HBM-backed scratch: the slots are re-used by every sequence a thread takes, so
// they are meant to live in L2 (evict_last); everything else the kernel touches is streamed (ld/st .cs).
#if VB_FB_HINT
__device__ __forceinline__ unsigned long long l2_keep_policy() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_stream_policy() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
#define VB_CP_HINT ".L2::cache_hint"
#define VB_CP_POL(p) , "l"(p)
#define VB_CP_ARG ", %2"
__device__ __forceinline__ void st_keep(double *q, double v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(q), "d"(v), "l"(pol) : "memory");
}
#else
#define VB_CP_HINT ""
#define VB_CP_POL(p)
#define VB_CP_ARG ""
__device__ __forceinline__ unsigned long long l2_stream_policy() { return 0ull; }
__device__ __forceinline__ unsigned long long l2_keep_policy() { return 0ull; }
__device__ __forceinline__ void st_keep(double *q, double v, unsigned long long) { *q = v; }
#endif

// The observations of a lane's sequence arrive through a private ring in shared memory, filled by cp.async
// VB_FB_RING - 1 steps ahead of their use: no register holds a load in flight (rotating prefetch registers makes every
// move wait for the youngest load: the first version of this kernel spent 16 % of its samples there).
template <int DIM>
struct ObsRing {
  static constexpr int kSlot = DIM * kFbThreads;       // doubles per ring slot: [slot][thread][DIM]
  const double *xp;                                    // the lane's observation at t = 0
  size_t stride;                                       // N * DIM
  unsigned base;                                       // shared address of the lane's place in slot 0
  int T;
  unsigned long long pol;                              // L2 evict_first: every observation is read once
  __device__ __forceinline__ void request(int u) const {              // observation u -> slot u % RING (a group either way)
    if (u < T) {
      const unsigned dst = base + (unsigned)(u % VB_FB_RING) * (kSlot * 8u);
      const double *src = xp + stride * u;
      if constexpr (DIM == 2) {
        asm volatile("cp.async.cg.shared.global" VB_CP_HINT " [%0], [%1], 16" VB_CP_ARG ";" ::"r"(dst), "l"(src) VB_CP_POL(pol)
                     : "memory");
      } else {
#pragma unroll
        for (int d = 0; d < DIM; d++)
          asm volatile("cp.async.ca.shared.global" VB_CP_HINT " [%0], [%1], 8" VB_CP_ARG ";" ::"r"(dst + 8u * d), "l"(src + d)
                       VB_CP_POL(pol) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  __device__ __forceinline__ void start() const {
#pragma unroll
    for (int u = 0; u < VB_FB_RING - 1; u++) request(u);
  }
  // observation u (consumed in order, each once): requests u + RING - 1 into the slot read one step ago
  __device__ __forceinline__ void take(int u, double (&x)[DIM]) const {
    request(u + VB_FB_RING - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(VB_FB_RING - 1) : "memory");
    const unsigned src = base + (unsigned)(u % VB_FB_RING) * (kSlot * 8u);
    if constexpr (DIM == 2) {
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x[0]), "=d"(x[1]) : "r"(src) : "memory");
    } else {
#pragma unroll
      for (int d = 0; d < DIM; d++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x[d]) : "r"(src + 8u * d) : "memory");
    }
  }
};

// Where alpha_hat(t) of the running sequence waits for the backward pass.  Reverse index r = T-2-t (r = 0 is read
// first).  The forward pass leaves the first `ts` of them in shared memory and the rest in the CTA's region of the
// scratch buffer, both laid out [r][k][thread] (conflict-free / one contiguous run per warp, compile-time strides).
// The backward pass reads everything from shared memory: the `ts` slots are a ring, and the entry for index
// r + VB_FB_AHEAD is copied in from the scratch buffer by cp.async while index r is in use (its slot's previous
// occupant, r + AHEAD - ts, was consumed earlier) -- no register waits on L2 or HBM.
template <int K>
struct Park {
  double *sm, *gl;             // already offset by the thread; gl is biased by -ts rows
  unsigned sbase;              // shared address of sm
  int ts, last;                // last = T - 2 of the running sequence (highest index parked)
  unsigned long long pol;
  static constexpr int kRow = K * kFbThreads;
  __device__ __forceinline__ void put(int r, const double (&a)[K]) const {
    if (r < ts) {
      double *q = sm + r * kRow;
#pragma unroll
      for (int k = 0; k < K; k++) q[k * kFbThreads] = a[k];
    } else {
      double *q = gl + (size_t)r * kRow;
#pragma unroll
      for (int k = 0; k < K; k++) st_keep(q + k * kFbThreads, a[k], pol);
    }
  }
  // backward side: indices are requested and read in order, so the ring positions are running offsets (no modulo)
  unsigned rq_off, rd_off;     // byte offsets of the slots of the next request / the next read
  int rq;                      // next index to request
  __device__ __forceinline__ void request() {                          // one cp.async group per index, empty or not
    if (rq >= ts && rq <= last) {
      const unsigned dst = sbase + rq_off;
      const double *src = gl + (size_t)rq * kRow;
#pragma unroll
      for (int k = 0; k < K; k++)
        asm volatile("cp.async.ca.shared.global" VB_CP_HINT " [%0], [%1], 8" VB_CP_ARG ";" ::"r"(dst + k * (kFbThreads * 8u)),
                     "l"(src + k * kFbThreads) VB_CP_POL(pol) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    rq++;
    rq_off += kRow * 8u;
    if (rq_off == (unsigned)ts * (kRow * 8u)) rq_off = 0;
  }
  __device__ __forceinline__ void start() {                            // after the forward pass
    rq = 0; rq_off = 0; rd_off = 0;
#pragma unroll
    for (int r = 0; r < VB_FB_AHEAD; r++) request();
  }
  __device__ __forceinline__ void get(double (&a)[K]) {                // the next index
    request();
    asm volatile("cp.async.wait_group %0;" ::"n"(VB_FB_AHEAD) : "memory");
    const unsigned src = sbase + rd_off;
#pragma unroll
    for (int k = 0; k < K; k++)
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[k]) : "r"(src + k * (kFbThreads * 8u)) : "memory");
    rd_off += kRow * 8u;
    if (rd_off == (unsigned)ts * (kRow * 8u)) rd_off = 0;
  }
};

// The outputs keep the reference's layout [K, N, maxT]: at one time step the 32 sequences of a warp own ONE contiguous
// run of 32 K doubles, but a lane's K values sit K doubles apart -- stored lane by lane, every STG would touch 32
// sectors with 8 useful bytes each (L2 write transactions, not bytes, bounded the first version of this kernel).  So
// the warp transposes through shared memory: lane l deposits its K values at [l K + k], then store i writes elements
// 32 i + l -- K full-line stores.  Two buffers alternate, so one __syncwarp per call orders everything.
template <int K>
struct RowStore {
  static constexpr int kBuf = K * 32 + K;              // + the skew of even K
  double *buf;                                         // this warp's two buffers
  int lane, nvalid;                                    // nvalid = K * (#sequences of the warp's chunk that exist)
  int flip;
  __device__ __forceinline__ static int at(int e) { return (K % 2 == 0) ? e + (e >> 5) : e; }   // odd K: conflict-free as is
  __device__ __forceinline__ void store(double *dst, const double (&v)[K]) {
    double *b = buf + flip * kBuf;
    flip ^= 1;
#pragma unroll
    for (int k = 0; k < K; k++) b[at(lane * K + k)] = v[k];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < K; i++) {
      const int e = 32 * i + lane;
      if (e < nvalid) __stcs(dst + e, b[at(e)]);
    }
  }
};

// 32 consecutive sequences, one per lane: forward (vbhmm_fb.m:316-347), then backward + responsibilities + summed
// pair marginals (:349-377).  Both loops run over the full 0..maxT-1 for the whole warp (a lane beyond its own length
// contributes the zeros the reference's outputs start with), so that every output row is written by full-line stores.
//
// The backward pass never touches the emissions again.  The reference's step is
//     beta_t(j) = sum_k A(j,k) beta_{t+1}(k) p_{t+1}(k) / c_{t+1},   xi_t(j,k) = alpha_t(j) A(j,k) beta_{t+1}(k) p_{t+1}(k) / c_{t+1}
// and the forward pass made alpha_{t+1}(k) = pred_t(k) p_{t+1}(k) / c_{t+1} with pred_t(k) = sum_j alpha_t(j) A(j,k), so
//     p_{t+1}(k) / c_{t+1} = alpha_{t+1}(k) / pred_t(k):
// one 25-term product and five reciprocals replace the second evaluation of the emission densities, their exp and the
// scale (~ 200 of the ~ 500 fp64 instructions per step) and the second read of the observations; sum_k gamma_t(k) = 1
// is preserved exactly as before (sum_j alpha_t(j) beta_t(j) = sum_k alpha_{t+1}(k) beta_{t+1}(k)).  pred_t(k) > 0 whenever
// A > 0; a column of exact zeros makes alpha_{t+1}(k) = 0 and the guarded quotient 0 (the state is unreachable).
template <int K, int DIM, bool UNI>
__device__ __forceinline__ void fb_chunk(const FbPar<K, DIM> &P, int base, int N, int maxT, int T, int Tw,
                                         const double *__restrict__ X, double *__restrict__ logrho,
                                         double *__restrict__ gamma, double *__restrict__ xi_sum,
                                         double *__restrict__ phi, const Park<K> &park_of_thread, RowStore<K> &rows,
                                         unsigned xring, vbhem::tab_t etab, vbhem::tab_t ltab) {
  Park<K> park = park_of_thread;
  park.last = T - 2;
  // UNI: every lane has a sequence and all 32 have the same length T = Tw -- no per-lane conditions in the loops
  const int n = base + rows.lane;
  const bool exists = UNI || n < N;
  const size_t NK = (size_t)N * K, ND = (size_t)N * DIM;
  double *lrow = logrho ? logrho + (size_t)K * base : nullptr, *grow = gamma + (size_t)K * base;
  double alpha[K], p[K], lr[K];
  double sumlogc = 0.0, summax = 0.0;

  // ---- forward, software-pipelined: the emission densities and their exp of step t+1 do not depend on the recursion
  // of step t, so one step does both (two independent dependency chains per lane: with 8 warps per SM the kernel is
  // bound by the latency of its chains).  Two steps per iteration with the roles of the (logrho, p) pairs swapped, so
  // that nothing is moved between registers.
  ObsRing<DIM> obs;
  obs.xp = X + (size_t)DIM * (exists ? n : base);
  obs.stride = ND;
  obs.base = xring;
  obs.T = T;
  obs.pol = l2_stream_policy();
  obs.start();
  auto next_obs = [&](int t, double (&l)[K], double (&q)[K]) -> double {   // logrho(:, t), p(:, t), the shift; t < T
    double x[DIM];
    obs.take(t, x);
    emissions<K, DIM>(P, x, l);
    return shifted_exp<K>(l, q, etab);
  };
  double mxn = 0.0;
  auto fstep = [&](int t, double (&lc)[K], double (&pc)[K], double (&ln)[K], double (&pn)[K]) {
    if (UNI || t < T) {
      if (t + 1 < T) mxn = next_obs(t + 1, ln, pn);
      double delta[K], c;
      if (t == 0) {
        c = 0.0;
#pragma unroll
        for (int k = 0; k < K; k++) { delta[k] = P.pz1[k] * pc[k]; c += delta[k]; }
      } else {
        c = forward_step<K, DIM>(P, alpha, pc, delta);
      }
      const double rc = rcp_pos(c);
#pragma unroll
      for (int k = 0; k < K; k++) alpha[k] = delta[k] * rc;
      if (t < T - 1) { park.put(T - 2 - t, alpha); summax += mxn; }      // alpha_hat(T-1) stays in registers
      sumlogc += vbhem::log_shift<1>(c, ltab, 0);
    }
    if (lrow) rows.store(lrow + NK * t, lc);
    if (!UNI && !(t + 1 < T)) {
#pragma unroll
      for (int k = 0; k < K; k++) ln[k] = 0.0;
    }
  };
  double l2[K], p2[K];
#pragma unroll
  for (int k = 0; k < K; k++) { alpha[k] = 0.0; lr[k] = 0.0; p[k] = 0.0; l2[k] = 0.0; p2[k] = 0.0; }
  if (0 < T) summax = next_obs(0, lr, p);
  for (int t = 0; t < Tw; t += 2) {
    fstep(t, lr, p, l2, p2);
    if (t + 1 < Tw) fstep(t + 1, l2, p2, lr, p);
  }
  if (exists) __stcs(phi + n, sumlogc * vbhem::kUnit + summax);
#pragma unroll
  for (int k = 0; k < K; k++) lr[k] = 0.0;
  // rows beyond the longest sequence of the warp: the zeros the reference's outputs start with
  for (int t = maxT - 1; t >= Tw; t--) {
    if (lrow) rows.store(lrow + NK * t, lr);
    rows.store(grow + NK * t, lr);
  }

  // ---- backward, software-pipelined as well: q(k) = alpha_hat_{t+1}(k) / pred_t(k) depends on parked values only, so
  // the quotients of step t-1 are formed while step t runs; what is left on the recurrence is beta -> beta.* q -> A * (.).
  // Two steps per iteration with the roles of the (alpha_hat, q) pairs swapped (nothing moves between registers).
  double Xi[K][K], beta[K], a1[K], q1[K], a2[K], q2[K];
#pragma unroll
  for (int j = 0; j < K; j++) {
    beta[j] = 1.0; a1[j] = 0.0; q1[j] = 0.0; a2[j] = 0.0; q2[j] = 0.0;
#pragma unroll
    for (int k = 0; k < K; k++) Xi[j][k] = 0.0;
  }
  auto quotients = [&](const double (&lo)[K], const double (&hi)[K], double (&out)[K]) {   // hi(k) / sum_j lo(j) A(j,k)
#pragma unroll
    for (int k = 0; k < K; k++) {
      double pred = 1e-300;                                            // keeps the quotient finite for an unreachable state
#pragma unroll
      for (int j = 0; j < K; j++) pred += lo[j] * P.A[j][k];
      out[k] = hi[k] * rcp_pos(pred);
    }
  };
  // ac = alpha_hat(t), qc = p_{t+1} / c_{t+1}; leaves alpha_hat(t-1) and the quotients of step t-1 in (an, qn)
  auto bstep = [&](int t, double (&ac)[K], double (&qc)[K], double (&an)[K], double (&qn)[K], double (&g)[K]) {
    if (t >= 1) {
      park.get(an);                                                    // alpha_hat(t-1): the next parked index
      quotients(an, ac, qn);
    }
    double bw[K];
#pragma unroll
    for (int k = 0; k < K; k++) bw[k] = beta[k] * qc[k];               // beta_{t+1}(k) p_{t+1}(k) / c_{t+1}
#pragma unroll
    for (int j = 0; j < K; j++) {
      double eta = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) eta += bw[k] * P.A[j][k];
      beta[j] = eta;
#pragma unroll
      for (int k = 0; k < K; k++) Xi[j][k] = fma(ac[j], bw[k], Xi[j][k]);
    }
#pragma unroll
    for (int k = 0; k < K; k++) g[k] = ac[k] * beta[k];
  };
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (UNI) {
    double g[K];
    if (Tw >= 1) rows.store(grow + NK * (Tw - 1), alpha);              // gamma(:, T) = alpha_hat(:, T)
    if (Tw >= 2) {
      park.start();
      park.get(a1);
      quotients(a1, alpha, q1);
    }
    for (int t = Tw - 2; t >= 0; t -= 2) {
      bstep(t, a1, q1, a2, q2, g);
      rows.store(grow + NK * t, g);
      if (t >= 1) {
        bstep(t - 1, a2, q2, a1, q1, g);
        rows.store(grow + NK * (t - 1), g);
      }
    }
  } else {
    // ragged warp: a lane joins at its own t = T - 1.  Which register pair is "current" follows the parity of t, not of
    // the lane's own step count, so the two instantiations of the step are a warp-uniform branch (with per-lane parity
    // half of the lanes would run one, half the other, one after the other).
    auto body = [&](int t, double (&ac)[K], double (&qc)[K], double (&an)[K], double (&qn)[K]) {
      double g[K];
      if (t < T - 1) {
        bstep(t, ac, qc, an, qn, g);
      } else if (t == T - 1) {
#pragma unroll
        for (int k = 0; k < K; k++) g[k] = alpha[k];
        if (T >= 2) {                                                  // what step t - 1 will find as its current pair
          park.start();
          park.get(an);
          quotients(an, alpha, qn);
        }
      } else {
#pragma unroll
        for (int k = 0; k < K; k++) g[k] = 0.0;
      }
      rows.store(grow + NK * t, g);
    };
    for (int t = Tw - 1; t >= 0; t--) {
      if (t & 1) body(t, a2, q2, a1, q1);
      else body(t, a1, q1, a2, q2);
    }
  }
  if (exists) {
    double *xo = xi_sum + (size_t)K * K * n;
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
      for (int j = 0; j < K; j++) __stcs(xo + j + K * k, P.A[j][k] * Xi[j][k]);
  }
}

// Persistent grid (every CTA resident): a warp takes 32 consecutive sequences at a time from a device counter, one
// sequence per thread; a thread's parking place is the same for every sequence it takes, so the scratch buffer is
// #resident threads x (maxT - 1 - ts) x K doubles whatever N is, and stays in L2.  A warp whose 32 sequences have one
// length runs the loops without per-lane conditions (fb_chunk<.., true>); ragged warps take the general variant.
template <int K, int DIM>
__global__ void __launch_bounds__(kFbThreads, VB_FB_MINB)
k_fb(const __grid_constant__ FbPar<K, DIM> P, int N, int maxT, int ts, const int *__restrict__ Tn,
     const double *__restrict__ X, double *__restrict__ logrho, double *__restrict__ gamma,
     double *__restrict__ xi_sum, double *__restrict__ phi, double *__restrict__ scratch, int *__restrict__ counter) {
  __shared__ __align__(16) double s_tab[kFbTabDoubles];
  __shared__ __align__(16) double s_rows[kFbThreads / 32][2 * RowStore<K>::kBuf];
  __shared__ __align__(16) double s_obs[VB_FB_RING * ObsRing<DIM>::kSlot];
  const unsigned xring = (unsigned)__cvta_generic_to_shared(s_obs) + threadIdx.x * (DIM * 8u);
  extern __shared__ double s_park[];
  vbhem::tab_t etab, ltab;
  stage_tables<kFbThreads>(s_tab, etab, ltab);
  Park<K> park;
  park.ts = ts;
  park.sm = s_park + threadIdx.x;
  park.sbase = (unsigned)__cvta_generic_to_shared(park.sm);
  park.last = -1;
  park.gl = scratch + ((size_t)blockIdx.x * (maxT > 1 + ts ? maxT - 1 - ts : 0) - ts) * Park<K>::kRow + threadIdx.x;
  park.pol = l2_keep_policy();
  RowStore<K> rows;
  rows.lane = threadIdx.x & 31;
  rows.buf = s_rows[threadIdx.x >> 5];
  rows.flip = 0;
  rows.nvalid = 0;
  for (;;) {
    int base = 0;
    if (rows.lane == 0) base = atomicAdd(counter, 32);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= N) break;
    const int n = base + rows.lane;
    const int T = n < N ? Tn[n] : 0;
    const int Tw = __reduce_max_sync(0xffffffffu, T);
    rows.nvalid = K * (N - base < 32 ? N - base : 32);
    if (__all_sync(0xffffffffu, T == Tw && n < N))
      fb_chunk<K, DIM, true>(P, base, N, maxT, Tw, Tw, X, logrho, gamma, xi_sum, phi, park, rows, xring, etab, ltab);
    else
      fb_chunk<K, DIM, false>(P, base, N, maxT, T, Tw, X, logrho, gamma, xi_sum, phi, park, rows, xring, etab, ltab);
  }
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
  Buf Tn, X, par, logrho, gamma, xi, phi, mom_partial, mom_out, scratch, counter;
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
  fb_finish_par(P);
  // persistent grid: every CTA resident; the last steps' alpha_hat parks in shared memory, the rest in the L2-sized scratch
  int sms = 148, smem_sm = 227 * 1024;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, h->device);
  const int want = (h->N + kFbThreads - 1) / kFbThreads;
  const int blocks = want < sms * VB_FB_MINB ? want : sms * VB_FB_MINB;
  const size_t per_step = sizeof(double) * K * kFbThreads;
  const size_t budget = (size_t)smem_sm / VB_FB_MINB - sizeof(double) * (kFbTabDoubles + (kFbThreads / 32) * 2 * RowStore<K>::kBuf +
                                                                                VB_FB_RING * DIM * kFbThreads) -
                        2048;                                      // 1 KB per CTA is the system's
  const int parked = h->maxT > 1 ? h->maxT - 1 : 0;
  int ts = (int)(budget / per_step);
  if (ts > parked) ts = parked;
  if (ts < 1) ts = 1;
  if (parked > ts && ts <= VB_FB_AHEAD) return -2;     // the ring must be longer than its look-ahead (never for K <= 8)
  const size_t smem = per_step * (size_t)ts;
  const size_t scratch = per_step * (size_t)(parked > ts ? parked - ts : 0) * blocks;
  if (h->scratch.ensure(scratch) != cudaSuccess || h->counter.ensure(sizeof(int)) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(k_fb<K, DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
  cudaMemsetAsync(h->counter.p, 0, sizeof(int), h->st);
  k_fb<K, DIM><<<blocks, kFbThreads, smem, h->st>>>(P, h->N, h->maxT, ts, h->Tn.as<int>(), h->X.as<double>(),
                                                 stats ? nullptr : h->logrho.as<double>(), h->gamma.as<double>(),
                                                 h->xi.as<double>(), h->phi.as<double>(), h->scratch.as<double>(),
                                                 h->counter.as<int>());
  if (!stats) return 1;
  using L = MomLayout<K, DIM>;
  const long long tot = (long long)h->N * h->maxT;
  const long long wantm = (tot + kMomThreads - 1) / kMomThreads;
  const int mb = (int)(wantm < 1 ? 1 : (wantm > 8LL * sms ? 8LL * sms : wantm));
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
  if (launched == -2) return fail_fb(VBHEM_ECUDA, "vbhmm_fb: shared memory too small for the alpha_hat ring");
  if (launched < 0) return fail_fb(VBHEM_ENOMEM, "out of device memory for the scratch buffer / moment partials");
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
  Buf *all[] = {&h->Tn, &h->X, &h->par, &h->logrho, &h->gamma, &h->xi, &h->phi, &h->mom_partial, &h->mom_out, &h->scratch,
                &h->counter};
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
