// vbhem_aux.cu -- the small kernels around the pair kernels:
//   * centre pre-pass on the device (digamma terms, log det W, v*W)
//       /root/reference/src/vbhem/vbhem_h3m_c_step_fc.m:118-165,180-191,271-273
//   * linear-domain centre records for the lattice kernel
//   * responsibilities z_hat, Nj, Lt1, Lt7      vbhem_h3m_c_step_fc.m:271-283, vbhemh3m_lb.m:90,107
//   * deterministic list of the pairs with Z > 1e-8 (vbhem_compute_Statistics.m:35),
//     centre-major and base-ascending; its length stays on the device (no host read-back)
//   * fixed-order reduction of the per-range partial sums, and the normalisation of
//     vbhem_compute_Statistics.m:57-82.
// Every sum over bases uses a fixed tree (thread -> warp -> block -> sequential over
// blocks), so results are bit-identical from run to run.
#include "vbhem_internal.cuh"

#include <cstdlib>

namespace vbhem {

namespace {

// ---- digamma ----------------------------------------------------------------
// psi(x) for x >= 0: upward recurrence to x >= 10, then the asymptotic series.
__device__ double digamma(double x) {
  if (!(x >= 0.0)) return NAN;
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  const double ser = f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132 -
                     f * (691.0 / 32760 - f * (1.0 / 12)))))));
  return r + log(x) - 0.5 / x - ser;
}

constexpr int kMaxDim = 8;

// log(det(W)) by LU with partial pivoting, W is d x d column-major.
__device__ double logdet(const double *W, int d) {
  double a[kMaxDim * kMaxDim];
  for (int q = 0; q < d * d; q++) a[q] = W[q];
  double det = 1.0;
  for (int c = 0; c < d; c++) {
    int piv = c;
    for (int r = c + 1; r < d; r++)
      if (fabs(a[r + d * c]) > fabs(a[piv + d * c])) piv = r;
    if (piv != c) {
      for (int k = 0; k < d; k++) { const double t = a[c + d * k]; a[c + d * k] = a[piv + d * k]; a[piv + d * k] = t; }
      det = -det;
    }
    det *= a[c + d * c];
    for (int r = c + 1; r < d; r++) {
      const double f = a[r + d * c] / a[c + d * c];
      for (int k = c + 1; k < d; k++) a[r + d * k] -= f * a[c + d * k];
    }
  }
  return log(det);
}

__device__ void build_record(const CentreView &C, int d, int dd, double *rec, int j);

// One block per centre.  With C.rec_stride > 0 the block goes on to build the centre's lattice record
// (k_build_records) from what it has just written; `counters` (may be null) are the fall-back counters of the
// E-step this launch opens.
__global__ void k_prepass(int K, int maxSr, int d, int covmode, const int *Sr, const int *gstart, int ngroups,
                          const double *alpha, const double *eta, const double *epsilon, const double *W,
                          const double *v, const double *lambda, double *logOmega, double *logPi, double *logA,
                          double *logLam, double *cterm, double *vW, CentreView C, double *rec, int *counters) {
  pdl_trigger();
  pdl_wait();
  const int j = blockIdx.x;
  if (counters && j == 0 && threadIdx.x < 3) counters[threadIdx.x] = 0;
  const int r = Sr[j];
  const int dd = covmode ? d * d : d;
  // per state: logLambdaTilde, cterm, v*W
  for (int k = threadIdx.x; k < maxSr; k += blockDim.x) {
    const size_t s = k + (size_t)maxSr * j;
    if (k < r) {
      double t1 = 0.0;
      for (int i = 1; i <= d; i++) t1 += digamma(0.5 * (v[s] + 1.0) - 0.5 * i);
      double ld;
      if (covmode) ld = logdet(W + (size_t)dd * s, d);
      else { ld = 0.0; for (int q = 0; q < d; q++) ld += log(W[(size_t)dd * s + q]); }
      const double ll = t1 + d * log(2.0) + ld;
      logLam[s] = ll;
      cterm[s] = -ll + d / lambda[s];
      for (int q = 0; q < dd; q++) vW[(size_t)dd * s + q] = v[s] * W[(size_t)dd * s + q];
    } else {
      logLam[s] = 0.0; cterm[s] = 0.0;
      for (int q = 0; q < dd; q++) vW[(size_t)dd * s + q] = 0.0;
    }
  }
  // transition rows: logA[k,:] = psi(eps[k,:]) - psi(sum_k' eps[k,k'])
  for (int x = threadIdx.x; x < maxSr * maxSr; x += blockDim.x) {
    const int k = x % maxSr, c = x / maxSr;
    const double *e = epsilon + (size_t)maxSr * maxSr * j;
    double out = 0.0;
    if (k < r && c < r) {
      double sum = 0.0;
      for (int cc = 0; cc < r; cc++) sum += e[k + maxSr * cc];
      out = digamma(e[x]) - digamma(sum);
    }
    logA[(size_t)maxSr * maxSr * j + x] = out;
  }
  for (int k = threadIdx.x; k < maxSr; k += blockDim.x) {
    const double *e = eta + (size_t)maxSr * j;
    double out = 0.0;
    if (k < r) {
      double sum = 0.0;
      for (int kk = 0; kk < r; kk++) sum += e[kk];
      out = digamma(e[k]) - digamma(sum);
    }
    logPi[k + (size_t)maxSr * j] = out;
  }
  if (threadIdx.x == 0) {
    int g = 0;
    while (g + 1 < ngroups && gstart[g + 1] <= j) g++;
    double sum = 0.0;
    for (int jj = gstart[g]; jj < gstart[g + 1]; jj++) sum += alpha[jj];
    logOmega[j] = digamma(alpha[j]) - digamma(sum);
  }
  if (rec) {
    __syncthreads();          // the record reads this block's own logA / logPi / cterm / vW
    build_record(C, d, dd, rec, j);
  }
}

// One block per centre: the padded linear-domain record of the lattice kernel.
__device__ void build_record(const CentreView &C, int d, int dd, double *rec, int j) {
  const int r = C.Sr[j], ld = C.maxSr;
  double *o = rec + (size_t)C.rec_stride * j;
  const double *logA = C.logA + (size_t)ld * ld * j;
  for (int x = threadIdx.x; x < 64; x += blockDim.x) {
    const int rho = x / 8, sg = x % 8;
    double val;
    if (sg >= r) val = 0.0;                       // padded to-state: never entered
    else if (rho >= r) val = 1.0;                 // padded from-state: keeps S finite
    else val = exp(logA[rho + ld * sg]);
    o[x] = val;
  }
  for (int k = threadIdx.x; k < 8; k += blockDim.x) o[64 + k] = k < r ? exp(C.logPi[k + (size_t)ld * j]) : 0.0;
  __syncthreads();
  // The lattice kernel's sums are bounded below by these weights (vbhem_lattice.cu): a centre with a
  // weight that is non-finite or under 2^-900 (log weight < -623) gets a NaN emission constant, which sends its pairs to the
  // general log-space kernel.
  __shared__ int weak;
  if (threadIdx.x == 0) weak = 0;
  __syncthreads();
  for (int x = threadIdx.x; x < 64 + 8; x += blockDim.x) {
    const bool live = x < 64 ? (x / 8 < r && x % 8 < r) : (x - 64 < r);
    if (live && !(o[x] >= 0x1p-900 && o[x] <= 1e300)) weak = 1;   // also NaN / Inf
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 8; k += blockDim.x) {
    o[72 + k] = k < r ? (weak ? NAN : d * log(kTwoPi) + C.cterm[k + (size_t)ld * j]) : INFINITY;
    for (int q = 0; q < d; q++) o[80 + k * d + q] = k < r ? C.m[q + d * (k + (size_t)ld * j)] : 0.0;
    for (int q = 0; q < dd; q++) o[80 + 8 * d + k * dd + q] = k < r ? C.vW[q + dd * (k + (size_t)ld * j)] : 0.0;
  }
}

__global__ void k_build_records(CentreView C, int d, int dd, double *rec, int *counters) {
  pdl_trigger();
  pdl_wait();
  if (counters && blockIdx.x == 0 && threadIdx.x < 3) counters[threadIdx.x] = 0;
  build_record(C, d, dd, rec, blockIdx.x);
}

// Zero everything beyond each base HMM's own state count (the kernels rely on it).
__global__ void k_sanitize_base(BaseView B, double *prior, double *A, double *mu, double *cov, int i0, int i1) {
  pdl_trigger();
  pdl_wait();
  const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= i1) return;
  const int b = B.Sb[i], ld = B.maxSb;
  for (int k = 0; k < ld; k++) {
    if (k >= b) {
      prior[k + (size_t)ld * i] = 0.0;
      for (int q = 0; q < B.d; q++) mu[q + B.d * (k + (size_t)ld * i)] = 0.0;
      for (int q = 0; q < B.dd; q++) cov[q + B.dd * (k + (size_t)ld * i)] = 0.0;
    }
    for (int c = 0; c < ld; c++)
      if (k >= b || c >= b) A[k + ld * (c + (size_t)ld * i)] = 0.0;
  }
}

// ---- responsibilities ---------------------------------------------------------
// Four lanes share one base (lane = 4 * base-in-warp + q; lane q takes the centres j = ja + q, ja + q + 4, ...), a
// block holds 64 bases: the per-thread chain of exp / shuffle latencies is a quarter of a thread-per-base layout's,
// and a 12 500-base shard still makes 196 blocks (148 SMs).  The block is also the granule of the active-pair
// list (cnt / offs) and of the slices of a streamed E-step.
constexpr int kLPB = 4;                    // lanes per base
constexpr int kZBases = 64;                // bases per block
constexpr int kZThreads = kZBases * kLPB;
constexpr int kZWarps = kZThreads / 32;
constexpr int kJT = 32;                    // centres per reduction tile: lane l of a warp ends up with tile entry l
constexpr double kLogFloor = -115.12925464970229;   // log(1e-50)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// zhat[N,K]; per block: Nj_part[blk][K] (sum of Z), cnt[blk][K] (#Z > 1e-8),
// lb_part[blk][2*ngroups] (Lt1, Lt7 partial sums).
__global__ void __launch_bounds__(kZThreads)
k_zhat(int N, int K, int ngroups, const int *gstart, const double *L_elbo, const double *logOmega,
       const double *Ntilde, double *zhat, double *Nj_part, int *cnt, double *lb_part, int blk0) {
  __shared__ double red[kZWarps][kJT];
  __shared__ int redc[kZWarps][kJT];
  __shared__ double redl[kZWarps][2];
  pdl_trigger();
  pdl_wait();
  const int blk = blk0 + blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane & (kLPB - 1), bw = lane / kLPB;
  const int i = blk * kZBases + warp * (32 / kLPB) + bw;
  const bool live = i < N;
  const double Nt = live ? Ntilde[i] : 0.0;
  const unsigned qmask = 0x11111111u << q;           // the lanes of the warp that hold the same centres
  // soft-max groups are independent: blockIdx.y strides over them (a (K,S) grid has few bases and many groups)
  for (int g = blockIdx.y; g < ngroups; g += gridDim.y) {
    const int ja = gstart[g], jb = gstart[g + 1];
    double mv = -INFINITY, lse = 0.0;
    if (live) {
      // max shift, then the sum (logtrick.m:15-18), each lane over its centres, then over the four lanes
      for (int j = ja + q; j < jb; j += kLPB) {
        const double lz = Nt * (logOmega[j] + L_elbo[i + (size_t)N * j]);
        if (lz > mv || j < ja + kLPB) mv = lz;         // the first element as it is (a NaN stays)
      }
    }
#pragma unroll
    for (int o = 1; o < kLPB; o <<= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, mv, o);
      if (other > mv) mv = other;                       // a NaN on either side: the base's sum is NaN anyway
    }
    double s = 0.0;
    if (live)
      for (int j = ja + q; j < jb; j += kLPB) s += exp(Nt * (logOmega[j] + L_elbo[i + (size_t)N * j]) - mv);
#pragma unroll
    for (int o = 1; o < kLPB; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (live) lse = mv + log(s);
    double t1 = 0.0, t7 = 0.0;
    for (int j0 = ja; j0 < jb; j0 += kJT) {
      double mine = 0.0;
      int minec = 0;
      const int nt = min(kJT, jb - j0);
#pragma unroll
      for (int c = 0; c < kJT / kLPB; c++) {
        const int jj = q + kLPB * c, j = j0 + jj;       // uniform over the lanes with the same q
        double Z = 0.0;
        int act = 0;
        if (live && jj < nt) {
          const double Le = L_elbo[i + (size_t)N * j];
          const double x = Nt * (logOmega[j] + Le) - lse;
          const double z = exp(x) + 1e-50;
          zhat[i + (size_t)N * j] = z;
          Z = z * Nt;
          act = Z > kSkipZ;
          t1 += Z * Le;
          // log(z) = x up to 1e-50 / z; below z = 1e-30 the term z log z is under 1e-28 and x is as good
          t7 += z * (x > kLogFloor ? x : kLogFloor);
        }
        // over the eight bases of the warp that share q (fixed order)
        Z += __shfl_xor_sync(0xffffffffu, Z, 4);
        Z += __shfl_xor_sync(0xffffffffu, Z, 8);
        Z += __shfl_xor_sync(0xffffffffu, Z, 16);
        const int wc = __popc(__ballot_sync(0xffffffffu, act) & qmask);
        if (bw == c) { mine = Z; minec = wc; }          // lane 4 c + q keeps tile entry q + 4 c = its own lane index
      }
      red[warp][lane] = mine;
      redc[warp][lane] = minec;
      __syncthreads();
      if (warp == 0 && lane < nt) {
        double sN = 0.0;
        int sc = 0;
        for (int w = 0; w < kZWarps; w++) { sN += red[w][lane]; sc += redc[w][lane]; }
        Nj_part[(size_t)blk * K + j0 + lane] = sN;
        cnt[(size_t)blk * K + j0 + lane] = sc;
      }
      __syncthreads();
    }
    t1 = warp_sum(t1);
    t7 = warp_sum(t7);
    if (lane == 0) { redl[warp][0] = t1; redl[warp][1] = t7; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < kZWarps; w++) { a += redl[w][0]; b += redl[w][1]; }
      lb_part[(size_t)blk * 2 * ngroups + 2 * g] = a;
      lb_part[(size_t)blk * 2 * ngroups + 2 * g + 1] = b;
    }
    __syncthreads();
  }
}

// Per centre j (warps 0 .. K-1): offs[blk][j] = exclusive prefix over the blocks of k_zhat of cnt[.][j] (32 blocks
// per step), segcnt[j] = the centre's total, Nj[j] = sum over blocks of Nj_part[.][j].  Warps K .. K + 2 ngroups - 1:
// the lower-bound sums lb[y] = sum over blocks of lb_part[.][y].  The fp64 sums run in a fixed order -- lane l adds
// blocks l, l + 32, ... in turn, then a shuffle tree over the lanes -- so they are bit-identical from run to run.
__global__ void k_list_scan_blocks(int nblocks, int K, int ngroups, const int *cnt, const double *Nj_part,
                                   const double *lb_part, int *offs, int *segcnt, double *Nj, double *lb) {
  pdl_trigger();
  pdl_wait();
  const int x = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (x >= K + 2 * ngroups) return;
  if (x >= K) {
    const int y = x - K;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += lb_part[(size_t)b * 2 * ngroups + y];
    s = warp_sum(s);
    if (lane == 0) lb[y] = s;
    return;
  }
  const int j = x;
  int run = 0;
  double s = 0.0;
  for (int b0 = 0; b0 < nblocks; b0 += 32) {
    const int b = b0 + lane;
    const int v = b < nblocks ? cnt[(size_t)b * K + j] : 0;
    if (b < nblocks) s += Nj_part[(size_t)b * K + j];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (b < nblocks) offs[(size_t)b * K + j] = run + inc - v;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  s = warp_sum(s);
  if (lane == 0) { segcnt[j] = run; Nj[j] = s; }
}

// seg_start[j] (K + 1 entries): exclusive prefix over centres; totals = {#active, items per range}.
__global__ void k_list_scan_centres(int K, int nranges, const int *segcnt, int *seg_start, int *totals) {
  pdl_trigger();
  pdl_wait();
  // one warp; lane l owns the contiguous range [l * per, (l + 1) * per) of centres
  const int lane = threadIdx.x, per = (K + 31) / 32;
  const int ja = min(K, lane * per), jb = min(K, ja + per);
  int a = 0;
  for (int j = ja; j < jb; j++) a += segcnt[j];
  int ia = a;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int ta = __shfl_up_sync(0xffffffffu, ia, o);
    if (lane >= o) ia += ta;
  }
  int ra = ia - a;   // exclusive prefix
  for (int j = ja; j < jb; j++) {
    seg_start[j] = ra;
    ra += segcnt[j];
  }
  if (lane == 31) {
    VB_ASSERT(ia >= 0);
    seg_start[K] = ia;
    totals[0] = ia;
    // the second pass cuts the list in nranges equal ranges, one warp's serial work each
    const int ipr = (ia + nranges - 1) / nranges;
    totals[1] = ipr > 0 ? ipr : 1;
  }
}

// One thread per base, one block per block of k_zhat (kZBases bases): the items of centre j of this block start
// at seg_start[j] + offs[blk][j], in base order.
__global__ void __launch_bounds__(kZBases)
k_list_fill(int N, int K, const double *zhat, const double *Ntilde, const int *offs, const int *seg_start,
            PairItem *items) {
  constexpr int kFW = kZBases / 32;
  __shared__ int wcnt[kFW][kJT];
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * kZBases + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool live = i < N;
  const double Nt = live ? Ntilde[i] : 0.0;
  for (int j0 = blockIdx.y * kJT; j0 < K; j0 += gridDim.y * kJT) {   // centre tiles are independent
    const int nt = min(kJT, K - j0);
    int minec = 0;
    for (int jj = 0; jj < nt; jj++) {
      const double Z = live ? zhat[i + (size_t)N * (j0 + jj)] * Nt : 0.0;
      const int wc = __popc(__ballot_sync(0xffffffffu, Z > kSkipZ));
      if (lane == jj) minec = wc;
    }
    wcnt[warp][lane] = minec;
    __syncthreads();
    for (int jj = 0; jj < nt; jj++) {
      const int j = j0 + jj;
      const double Z = live ? zhat[i + (size_t)N * j] * Nt : 0.0;
      const bool act = Z > kSkipZ;
      const unsigned bal = __ballot_sync(0xffffffffu, act);
      if (act) {
        int pos = seg_start[j] + offs[(size_t)blockIdx.x * K + j] + __popc(bal & ((1u << lane) - 1u));
        for (int w = 0; w < warp; w++) pos += wcnt[w][jj];
        VB_ASSERT(pos >= seg_start[j] && pos < seg_start[j + 1] && (long long)pos < (long long)N * K);
        items[pos] = PairItem{i, j};
      }
    }
    __syncthreads();
  }
}

// S_*[., j] = sum over the partial slots of centre j, in range order: the ranges that overlap the centre's
// items [seg_start[j], seg_start[j + 1]) are w_lo .. w_hi and range w left its sum in slot w + j (PairWork).
__global__ void k_partial_reduce(int K, int per, const int *seg_start, const int *totals, const double *partial,
                                 double *S_nu1, double *S_xi, double *S_pr, double *S_mu, double *S_Mu, int r, int d,
                                 int dd) {
  pdl_trigger();
  pdl_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= K * per) return;
  const int j = x / per, q = x % per;
  const int a = seg_start[j], b = seg_start[j + 1], ipr = totals[1];
  double s = 0.0;
  if (b > a) {
    const int w_hi = (b - 1) / ipr;
    VB_ASSERT(ipr >= 1 && a >= 0 && b <= totals[0] && (long long)w_hi * ipr < (long long)totals[0]);
    // sixteen loads in flight, added in slot order (a chain of dependent load-add pairs costs an L2 latency per slot;
    // a 12 500-base shard leaves ~ 38 slots per centre: three rounds)
    for (int w = a / ipr; w <= w_hi; w += 16) {
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; u++) v[u] = w + u <= w_hi ? partial[(size_t)(w + u + j) * per + q] : 0.0;
#pragma unroll
      for (int u = 0; u < 16; u++) s += v[u];
    }
  }
  if (q < r) S_nu1[q + (size_t)r * j] = s;
  else if (q < r + r * r) S_xi[(q - r) + (size_t)r * r * j] = s;
  else if (q < 2 * r + r * r) S_pr[(q - r - r * r) + (size_t)r * j] = s;
  else if (q < 2 * r + r * r + r * d) S_mu[(q - 2 * r - r * r) + (size_t)r * d * j] = s;
  else S_Mu[(q - 2 * r - r * r - r * d) + (size_t)r * dd * j] = s;
}

// vbhem_compute_Statistics.m:57-82 in place on a stats buffer; one thread per (state, centre).
__global__ void k_finalize(int K, int r, int d, int covmode, const int *Sr, double *Nj, double *S_xi, double *S_pr,
                           double *S_mu, double *S_Mu) {
  pdl_trigger();
  pdl_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= K * r) return;
  const int k = x % r, j = x / r;
  const int dd = covmode ? d * d : d;
  if (k == 0) {
    Nj[j] += 1e-50;
    if (Sr[j] == 1) S_xi[(size_t)r * r * j] = 1e-12;
  }
  if (k >= Sr[j]) return;
  const double nr = S_pr[x] + 1e-50;
  S_pr[x] = nr;
  double yb[kMaxDim];
  for (int q = 0; q < d; q++) {
    yb[q] = S_mu[k + r * (q + (size_t)d * j)] / nr;
    S_mu[k + r * (q + (size_t)d * j)] = yb[q];
  }
  for (int q = 0; q < dd; q++) {
    const size_t o = k + r * (q + (size_t)dd * j);
    const double yy = covmode ? yb[q / d] * yb[q % d] : yb[q] * yb[q];
    S_Mu[o] = S_Mu[o] / nr - yy;
  }
}

// ---- fp64 peak chains (roofline denominator measured on the device itself) ----
__global__ void k_peak_dmma(double *out, int iters, double a, double b) {
  double c0[8], c1[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c0[i] + c1[i];
  if (s == 12345.678) out[0] = s;
}
__global__ void k_peak_dfma(double *out, int iters, double a, double b) {
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

}  // namespace

int measure_fp64_peak(double out[2]) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double *buf = nullptr;
  if (cudaMalloc(&buf, 64) != cudaSuccess) return -1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int threads = 512, blocks = sms * 4, iters = 4000;
  const double nthr = (double)threads * blocks;
  for (int which = 0; which < 2; which++) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0);
      if (which == 0) k_peak_dmma<<<blocks, threads>>>(buf, iters, 0.999999, 1e-7);
      else k_peak_dfma<<<blocks, threads>>>(buf, iters, 0.999999, 1e-7);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double flop = which == 0 ? (nthr / 32) * iters * 8 * 512.0 : nthr * iters * 8 * 2.0;
    out[which] = flop / (best * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

bool pdl_enabled(int cls) {
  static const int mode = [] {
    const char *e = getenv("VBHEM_PDL");
    return e && *e ? atoi(e) : 2;
  }();
  return mode == 1 || (mode == 2 && cls == 0) || (mode == 3 && cls == 1);
}

int launch_prepass(int K, int maxSr, int d, int covmode, const int *Sr, const int *gstart, int ngroups,
                   const double *alpha, const double *eta, const double *epsilon, const double *W, const double *v,
                   const double *lambda, double *logOmega, double *logPi, double *logA, double *logLam, double *cterm,
                   double *vW, const CentreView &C, double *rec, int *counters, cudaStream_t st) {
  launch_chain(k_prepass, dim3(K), dim3(64), 0, st, K, maxSr, d, covmode, Sr, gstart, ngroups, alpha, eta, epsilon, W, v,
               lambda, logOmega, logPi, logA, logLam, cterm, vW, C, rec, counters);
  return 1;
}

int launch_build_records(const CentreView &C, int d, int dd, double *rec, int *counters, cudaStream_t st) {
  launch_chain(k_build_records, dim3(C.K), dim3(64), 0, st, C, d, dd, rec, counters);
  return 1;
}

int launch_sanitize_base(BaseView B, double *prior, double *A, double *mu, double *cov, cudaStream_t st, int i0,
                         int n) {
  if (n < 0) n = B.N - i0;
  if (n <= 0) return 0;
  launch_chain(k_sanitize_base, dim3((n + 127) / 128), dim3(128), 0, st, B, prior, A, mu, cov, i0, i0 + n);
  return 1;
}

int zhat_blocks(int N) { return (N + kZBases - 1) / kZBases; }

// gridDim.y for kernels whose blocks over the bases each loop over `parts` independent parts: enough blocks for
// eight per SM (B200: 148 SMs) when the base set alone gives too few
static int spread(int nblk, int parts) {
  int ny = (8 * 148 + nblk - 1) / (nblk > 0 ? nblk : 1);
  if (ny > parts) ny = parts;
  return ny < 1 ? 1 : ny;
}

int launch_zhat(int N, int K, int ngroups, const int *gstart, const double *L_elbo, const double *logOmega,
                const double *Ntilde, double *zhat, double *Nj_part, int *cnt, double *lb_part, cudaStream_t st, int blk0,
                int nblk) {
  const int nb = zhat_blocks(N);
  if (nblk < 0) nblk = nb - blk0;
  if (nblk <= 0) return 0;
  launch_chain(k_zhat, dim3(nblk, spread(nblk, ngroups)), dim3(kZThreads), 0, st, N, K, ngroups, gstart, L_elbo,
               logOmega, Ntilde, zhat, Nj_part, cnt, lb_part, blk0);
  return 1;
}

int zhat_block_size() { return kZBases; }

int launch_list_scan(int N, int K, int ngroups, int nranges, const int *cnt, const double *Nj_part,
                     const double *lb_part, int *offs, int *segcnt, int *seg_start, int *totals, double *Nj, double *lb,
                     cudaStream_t st) {
  const int warps = K + 2 * ngroups;
  launch_chain(k_list_scan_blocks, dim3((warps + 3) / 4), dim3(128), 0, st, zhat_blocks(N), K, ngroups, cnt, Nj_part,
               lb_part, offs, segcnt, Nj, lb);
  launch_chain(k_list_scan_centres, dim3(1), dim3(32), 0, st, K, nranges, (const int *)segcnt, seg_start, totals);
  return 2;
}

int launch_list_fill(int N, int K, const double *zhat, const double *Ntilde, const int *offs, const int *seg_start,
                     PairItem *items, cudaStream_t st) {
  const int nb = zhat_blocks(N);
  launch_chain(k_list_fill, dim3(nb, spread(nb, (K + kJT - 1) / kJT)), dim3(kZBases), 0, st, N, K, zhat, Ntilde, offs,
               seg_start, items);
  return 1;
}

int launch_partial_reduce(int K, int per, const int *seg_start, const int *totals, const double *partial,
                          double *S_nu1, double *S_xi, double *S_pr, double *S_mu, double *S_Mu, int r, int d, int dd,
                          cudaStream_t st) {
  const int tot = K * per;
  launch_chain(k_partial_reduce, dim3((tot + 63) / 64), dim3(64), 0, st, K, per, seg_start, totals, partial, S_nu1,
               S_xi, S_pr, S_mu, S_Mu, r, d, dd);
  return 1;
}

int launch_finalize(int K, int maxSr, int d, int covmode, const int *Sr, double *Nj, double *S_xi, double *S_pr,
                    double *S_mu, double *S_Mu, cudaStream_t st) {
  const int tot = K * maxSr;
  launch_chain(k_finalize, dim3((tot + 127) / 128), dim3(128), 0, st, K, maxSr, d, covmode, Sr, Nj, S_xi, S_pr, S_mu,
               S_Mu);
  return 1;
}

}  // namespace vbhem
