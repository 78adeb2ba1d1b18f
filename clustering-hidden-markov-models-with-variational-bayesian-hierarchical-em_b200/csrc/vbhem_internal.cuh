// vbhem_internal.cuh -- device-side views and launch entry points shared by the
// kernels (vbhem_general.cu, vbhem_lattice.cu, vbhem_aux.cu) and the C ABI
// (vbhem_capi.cu).  Not part of the public boundary (include/vbhem_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Checked build (csrc/build.sh -DVBHEM_CHECKED -o ..._checked.so): device-side bounds assertions on every index
// the kernels compute from device-resident counts (list positions, partial slots, pair indices).  compute-sanitizer
// is closed on the GPU pool this was developed on (profiles/README.md), so this is the memcheck stand-in: the
// parity tests run once under the checked library (scripts/checked_build.sh).
#ifdef VBHEM_CHECKED
#include <cassert>
#define VB_ASSERT(x) assert(x)
#else
#define VB_ASSERT(x) ((void)0)
#endif

namespace vbhem {

constexpr double kTwoPi = 6.283185307179586476925286766559;
constexpr double kTiny = 1e-280;     // below this a linear-domain lattice sum is not trusted
constexpr double kSkipZ = 1e-8;      // vbhem_compute_Statistics.m:35
constexpr int kS8 = 8;               // lattice kernel tile: 8 x 8 states

// Base HMM set in HBM: the C-ABI column-major arrays, zero padded beyond Sb[i].
struct BaseView {
  int N, maxSb, d, covmode, dd;
  const int *Sb;          // [N]
  const double *prior;    // [maxSb, N]
  const double *A;        // [maxSb(from), maxSb(to), N]
  const double *mu;       // [d, maxSb, N]
  const double *cov;      // [dd, maxSb, N]
  const double *Ntilde;   // [N]
};

// Centre set after the pre-pass (log domain), C-ABI layout.
struct CentreView {
  int K, maxSr;
  const int *Sr;          // [K]
  const int *group;       // [K]
  const double *logPi;    // [maxSr, K]
  const double *logA;     // [maxSr(from), maxSr(to), K]
  const double *m;        // [d, maxSr, K]
  const double *vW;       // [dd, maxSr, K]
  const double *cterm;    // [maxSr, K]
  const double *logOmega; // [K]
  // Linear-domain records for the lattice kernel (see build_records), one per centre:
  //   expA[8][8] (row rho, col sigma) | expPi[8] | c0[8] | m[8][d] | vW[8][dd]
  const double *rec;
  int rec_stride;         // doubles per record
  // First-pass tiling: a centre alone in its tile (big_list[nbig]) or two centres per tile, in kTileClasses
  // classes laid out one after the other in pair_list (two entries per tile, -1 = no partner) and pair_rec (one
  // tile record each): class 0 = two centres of at most four states, columns 0-3 / 4-7; class k = 1..3 = a centre
  // of at most k states in columns [0, k) and one of at most 8 - k states in columns [k, 8).
  const int *big_list, *pair_list;
  const double *pair_rec;
  int nbig, npairs[4];
};
constexpr int kTileClasses = 4;

// One unit of second-pass work: the pair (base i, centre j); its weight Z = zhat[i + N j] * Ntilde[i] is read
// where it is used.
struct PairItem { int i, j; };

// Where a pair kernel puts its results.
//   mode 0: L_elbo only
//   mode 1: literal per-pair outputs (Boundary-1)
//   mode 2: weighted accumulation of the per-pair statistics into per-range
//           partial sums (deterministic second pass of the fused E-step)
struct PairOut {
  int mode;
  double *L_elbo;                               // [N, K]
  double *nu1, *xi, *pr, *mu, *Mu;              // mode 1, leading dim maxSr, pair index p = i + N*j
  double *partial;                              // mode 2: [nranges + K][per] raw sums, slot = range + centre
  int per;                                      // doubles per centre statistic block
};

// Work description: either every (i, j) pair or an explicit list (centre-major, base-ascending) whose LENGTH
// LIVES ON THE DEVICE, so that the host never waits for it: totals = {#items, items per range}.  The list is cut
// in `nranges` contiguous ranges of totals[1] items (range w = items [w * totals[1], (w + 1) * totals[1]));
// one warp owns a range, accumulates the weighted statistics of one centre in registers and writes them to
// partial slot (w + j) whenever the centre changes -- ranges and centres both ascend along the list, so the
// slot is unique, at most nranges + K slots exist, and the slots of centre j are the consecutive ones of the
// ranges that overlap [seg_start[j], seg_start[j + 1]).
struct PairWork {
  const PairItem *items;  // nullptr => all pairs
  const int *totals;      // device {#items, items per range}; nullptr when items == nullptr
  const double *zhat;     // [N, K] and
  const double *Ntilde;   // [N]: weight of item (i, j) = zhat[i + N j] * Ntilde[i]
  int nranges;
  long long npairs;       // all-pairs work: N * K; list work: upper bound of the list length
};

// Pairs the lattice kernel could not finish in the linear domain: one byte per
// pair (first pass) or per work item (second pass) plus a device counter.
struct Fallback { unsigned char *flags; int *count; };

inline __host__ __device__ int stat_per(int r, int d, int dd) { return r + r * r + r + r * d + r * dd; }

// ---- programmatic dependent launch (PDL) --------------------------------------
// An E-step is a chain of ~12 dependent kernels, most of them a few microseconds long, so at a 12 500-base shard
// (8-GPU strong scaling) the launch latency between them is a visible share of the step.  Every kernel of the
// chain is launched with cudaLaunchAttributeProgrammaticStreamSerialization: it lets its successor's CTAs be
// scheduled at once (pdl_trigger, first statement) and waits for its predecessor's results itself (pdl_wait, before
// the first dependent read; a no-op when launched the ordinary way).  EVERY path through EVERY chain kernel must
// execute pdl_wait -- a kernel that returned without it would complete before its predecessor and let ITS
// successor run too early.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// vbhem_aux.cu: kernel class 0 = the small kernels, 1 = the pair kernels (k_lattice_*, k_general);
// VBHEM_PDL = 0 none, 1 all, 2 small only (default), 3 pair kernels only
bool pdl_enabled(int cls);
template <int CLS = 0, typename... KA, typename... A>
inline cudaError_t launch_chain(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at;
  cfg.numAttrs = pdl_enabled(CLS) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
}
#endif

// ---- launchers (each returns the number of kernels it launched) -------------
// flags (fall-back mode): only flagged pairs / items are processed; the all-pairs flags are CLEARED as they are
// served, so the flag array is all-zero again when the kernel ends (no memset per E-step)
// count (fall-back mode): device ints; all-pairs work uses count[0] = pending, count[2] = served so far; list work
// only reads count[0].  All-pairs work looks at bases [i0, i1) only (i1 < 0: all).
int launch_general(const BaseView &B, const CentreView &C, int T, const PairWork &W, const PairOut &O,
                   unsigned char *flags, int *count, double *scratch, size_t scratch_doubles, cudaStream_t st,
                   int i0 = 0, int i1 = -1);
size_t general_scratch_per_warp(int maxSb, int maxSr, int d, int dd, int T);

bool lattice_supported(const BaseView &B, const CentreView &C, int T, bool need_forward);
// bases [i0, i0 + n) only (n < 0: to the end); L_elbo keeps its leading dimension B.N
int launch_lattice_bwd(const BaseView &B, const CentreView &C, int T, double *L_elbo, Fallback fb,
                       cudaStream_t st, int i0 = 0, int n = -1);
int launch_lattice_full(const BaseView &B, const CentreView &C, int T, const PairWork &W, const PairOut &O,
                        Fallback fb, cudaStream_t st);
int lattice_rec_stride(int d, int dd);
int launch_pair_records(const CentreView &C, int d, int dd, double *out, cudaStream_t st);
int lattice_bwd_wave();   // bases one full wave of the first-pass grid takes

// aux kernels (vbhem_aux.cu)
int zhat_blocks(int N);
// with rec != nullptr the same launch also builds the lattice records of C (launch_build_records)
int launch_prepass(int K, int maxSr, int d, int covmode, const int *Sr, const int *gstart, int ngroups,
                   const double *alpha, const double *eta, const double *epsilon, const double *W, const double *v,
                   const double *lambda, double *logOmega, double *logPi, double *logA, double *logLam, double *cterm,
                   double *vW, const CentreView &C, double *rec, int *counters, cudaStream_t st);
// `counters`: optional device ints {first-pass fall-backs, second-pass fall-backs} cleared by the same launch
int launch_build_records(const CentreView &C, int d, int dd, double *rec, int *counters, cudaStream_t st);
int launch_sanitize_base(BaseView B, double *prior, double *A, double *mu, double *cov, cudaStream_t st, int i0 = 0,
                         int n = -1);
// zhat and, per block of 256 bases, the partial sums of Z (Nj_part), of the lower-bound terms (lb_part) and the
// counts of pairs with Z > 1e-8 (cnt); blocks [blk0, blk0 + nblk) of zhat_blocks(N) (nblk < 0: all)
int launch_zhat(int N, int K, int ngroups, const int *gstart, const double *L_elbo, const double *logOmega,
                const double *Ntilde, double *zhat, double *Nj_part, int *cnt, double *lb_part, cudaStream_t st,
                int blk0 = 0, int nblk = -1);
int zhat_block_size();
// deterministic (centre-major, base-ascending) list of those pairs; seg_start[K + 1] = first item of every centre,
// totals = {#active, items per range = ceil(#active / nranges), at least 1}.  Nothing here depends on a value the
// host would have to read back: every grid is sized by N and K alone.
// The same launch sums the per-block partials of launch_zhat in a fixed order: Nj[K], lb[2 * ngroups].
int launch_list_scan(int N, int K, int ngroups, int nranges, const int *cnt, const double *Nj_part,
                     const double *lb_part, int *offs, int *segcnt, int *seg_start, int *totals, double *Nj, double *lb,
                     cudaStream_t st);
int launch_list_fill(int N, int K, const double *zhat, const double *Ntilde, const int *offs, const int *seg_start,
                     PairItem *items, cudaStream_t st);
// S_*[., j] = sum of centre j's partial slots in range order (see PairWork)
int launch_partial_reduce(int K, int per, const int *seg_start, const int *totals, const double *partial,
                          double *S_nu1, double *S_xi, double *S_pr, double *S_mu, double *S_Mu, int r, int d, int dd,
                          cudaStream_t st);
int launch_finalize(int K, int maxSr, int d, int covmode, const int *Sr, double *Nj, double *S_xi, double *S_pr,
                    double *S_mu, double *S_Mu, cudaStream_t st);

// Device-resident EM state (vbhem_em.cu): prior hyper-parameters, current centre hyper-parameters
// (C-ABI layouts of vbhem_estep_hyp) and the softmax groups.
struct EmView {
  int K, maxSr, d, covmode, ngroups;
  const int *Sr, *gstart;
  const unsigned char *frozen;   // per soft-max group: 1 = leave this H3M's hyper-parameters alone (may be null)
  double alpha0, eta0, epsilon0, lambda0, v0;
  const double *m0;       // [d]
  const double *W0inv;    // [d] diagonal of inv(W0)
  // Per-group priors (hyper-parameter learning evaluates every batched H3M under its OWN candidate prior,
  // /root/reference/src/vbhem/vbhem_h3m_c_hyp.m:105-129): null, or [ngroups][kGroupPrior] =
  // alpha0, eta0, epsilon0, lambda0, v0, m0[8], W0inv[8]; overrides the scalars above.
  const double *gprior;
  const double *alpha, *eta, *epsilon, *m, *W, *v, *lambda;
};
constexpr int kGroupPrior = 5 + 8 + 8;
// doubles per group of launch_lb_deriv's output: d/d alpha0, eta0, epsilon0, v0, lambda0, W0[8] (iid: entry 0), m0[8]
constexpr int kDerivStride = 5 + 8 + 8;
// out[g][12]: Lt1..Lt10, LB, sum_j Nj of softmax group g (vbhemh3m_lb.m:64-186)
int launch_lower_bound(const EmView &E, const double *Nj, const double *lbsum, const double *logOmega,
                       const double *logPi, const double *logA, const double *logLam, double *out, cudaStream_t st);
// vbhemh3m_lb.m:202-309: derivatives of the bound with respect to the prior hyper-parameters at the current
// (pre-M-step) state, per group; iid: W0 is a scalar (the d x d identity times it).  out[g][kDerivStride]
int launch_lb_deriv(const EmView &E, bool iid, const double *logOmega, const double *logPi, const double *logA,
                    const double *logLam, double *out, cudaStream_t st);
// vbhem_mstep_component.m:42-70 in place on the hyper-parameter arrays, from the normalised statistics
int launch_mstep(const EmView &E, const double *Nj, const double *nu1, const double *xi, const double *pr,
                 const double *ybar, const double *SC, double *alpha, double *eta, double *epsilon, double *m,
                 double *W, double *v, double *lambda, cudaStream_t st);

// `baseem` initialisation (vbhemhmm_init.m:58-100) on the device: centre state (k, j) copies the emission of base
// state (randb, randg); arrays in the layouts of vbhem_estep_hyp
int launch_init_baseem(const BaseView &B, int K, int maxSr, const int *Sr, const int *gstart, int ngroups, double alpha0,
                       double eta0, double epsilon0, double lambda0, double v0, double Nv, const int *randb,
                       const int *randg, const double *omega, double *alpha, double *eta, double *epsilon, double *m,
                       double *W, double *v, double *lambda, cudaStream_t st);

// fp64 peak micro-benchmark (vbhem_aux.cu): out[0] DMMA, out[1] DFMA TFLOP/s; returns 0 on success
int measure_fp64_peak(double out[2]);

}  // namespace vbhem
