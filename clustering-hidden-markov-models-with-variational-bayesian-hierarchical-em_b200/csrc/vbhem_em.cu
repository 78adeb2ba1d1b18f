// vbhem_em.cu -- the O(K S d^2) remainder of one VBHEM iteration, on the device, so that an EM run
// needs no host arithmetic and no parameter traffic between iterations:
//   * lower bound          /root/reference/src/vbhem/vbhemh3m_lb.m:64-186
//   * M-step               /root/reference/src/vbhem/vbhem_mstep_component.m:42-70,
//                          /root/reference/src/vbhem/vbhem_h3m_c_step_fc.m:396 (alpha)
// Inputs are what the E-step leaves in HBM: the normalised statistics (k_finalize), the pre-pass
// results (logOmega, logPi, logA, logLambdaTilde) and the hyper-parameters the E-step ran with.
// One H3M (softmax group) per block; every sum runs in a fixed order.
#include "vbhem_internal.cuh"

namespace vbhem {

namespace {

constexpr int kMaxDim = 8;
constexpr double kLogPi = 1.1447298858494001741434273513531;
constexpr double kLog2 = 0.69314718055994530941723212145818;
constexpr double kLog2Pi = 1.8378770664093454835606594728112;

// In-place inverse of a d x d matrix (column-major) by Gauss-Jordan with partial pivoting;
// returns log|det|.
__device__ double invert(double *a, int d) {
  double inv[kMaxDim * kMaxDim];
  for (int q = 0; q < d * d; q++) inv[q] = (q % d == q / d) ? 1.0 : 0.0;
  double logdet = 0.0;
  for (int c = 0; c < d; c++) {
    int piv = c;
    for (int r = c + 1; r < d; r++)
      if (fabs(a[r + d * c]) > fabs(a[piv + d * c])) piv = r;
    if (piv != c)
      for (int k = 0; k < d; k++) {
        double t = a[c + d * k]; a[c + d * k] = a[piv + d * k]; a[piv + d * k] = t;
        t = inv[c + d * k]; inv[c + d * k] = inv[piv + d * k]; inv[piv + d * k] = t;
      }
    const double p = a[c + d * c];
    logdet += log(fabs(p));
    const double ip = 1.0 / p;
    for (int k = 0; k < d; k++) { a[c + d * k] *= ip; inv[c + d * k] *= ip; }
    for (int r = 0; r < d; r++) {
      if (r == c) continue;
      const double f = a[r + d * c];
      if (f == 0.0) continue;
      for (int k = 0; k < d; k++) { a[r + d * k] -= f * a[c + d * k]; inv[r + d * k] -= f * inv[c + d * k]; }
    }
  }
  for (int q = 0; q < d * d; q++) a[q] = inv[q];
  return logdet;
}

__device__ double logdet_lu(const double *W, int d) {
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

constexpr int kLbThreads = 256;
constexpr int kTerms = 10;

// The prior of soft-max group g: the handle's scalars, or the group's own row (EmView::gprior).
struct Prior { double a0, e0, p0, lam0, v0; const double *m0, *W0inv; };
__device__ __forceinline__ Prior prior_of(const EmView &E, int g) {
  if (E.gprior) {
    const double *p = E.gprior + (size_t)kGroupPrior * g;
    return Prior{p[0], p[1], p[2], p[3], p[4], p + 5, p + 13};
  }
  return Prior{E.alpha0, E.eta0, E.epsilon0, E.lambda0, E.v0, E.m0, E.W0inv};
}

__device__ double digamma_em(double x) {      // as in vbhem_aux.cu: recurrence to x >= 10, then the asymptotic series
  if (!(x >= 0.0)) return NAN;
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  const double ser = f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132 -
                     f * (691.0 / 32760 - f * (1.0 / 12)))))));
  return r + log(x) - 0.5 / x - ser;
}

// Fixed-order block sum of kTerms values per thread; result valid in thread 0.
__device__ void block_sum_terms(double (&t)[kTerms], double (*sh)[kTerms]) {
  for (int k = 0; k < kTerms; k++) sh[threadIdx.x][k] = t[k];
  __syncthreads();
  for (int s = kLbThreads / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
      for (int k = 0; k < kTerms; k++) sh[threadIdx.x][k] += sh[threadIdx.x + s][k];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    for (int k = 0; k < kTerms; k++) t[k] = sh[0][k];
}

// vbhemh3m_lb.m:64-186 for the H3M of softmax group blockIdx.x.  out[g][0..9] = Lt1..Lt10,
// out[g][10] = LB = Lt1+Lt2+Lt3+Lt4+Lt5+Lt6-Lt7-Lt8-Lt9-Lt10, out[g][11] = sum_j Nj.
// W0inv: diagonal of inv(W0) (length d; iid W0 repeats the scalar).
__global__ void __launch_bounds__(kLbThreads)
k_lower_bound(EmView E, const double *Nj, const double *lbsum, const double *logOmega, const double *logPi,
              const double *logA, const double *logLam, double *out) {
  __shared__ double sh[kLbThreads][kTerms];
  pdl_trigger();
  pdl_wait();
  const int g = blockIdx.x;
  const int ja = E.gstart[g], jb = E.gstart[g + 1], Kg = jb - ja;
  const int r = E.maxSr, d = E.d, dd = E.covmode ? d * d : d;
  const Prior pr = prior_of(E, g);
  const double a0 = pr.a0, e0 = pr.e0, p0 = pr.p0, lam0 = pr.lam0, v0 = pr.v0;
  double logdetW0inv = 0.0;
  for (int q = 0; q < d; q++) logdetW0inv += log(pr.W0inv[q]);
  double lg0 = 0.0;
  for (int i = 1; i <= d; i++) lg0 += lgamma(0.5 * (v0 + 1.0 - i));
  const double cB = (d * (d - 1) / 4.0) * kLogPi;
  const double logB0 = 0.5 * v0 * logdetW0inv - 0.5 * v0 * d * kLog2 - cB - lg0;
  // t[0..9] = Lt1..Lt10 contributions of this thread
  double t[kTerms];
  for (int k = 0; k < kTerms; k++) t[k] = 0.0;

  // per-centre terms
  for (int j = ja + threadIdx.x; j < jb; j += kLbThreads) {
    const int S = E.Sr[j];
    const double lO = logOmega[j], al = E.alpha[j];
    t[1] += Nj[j] * lO;                                                            // Lt2
    double sP = 0.0, sA = 0.0, se = 0.0, sg = 0.0, sel = 0.0;
    for (int k = 0; k < S; k++) {
      const double et = E.eta[k + (size_t)r * j], lp = logPi[k + (size_t)r * j];
      sP += lp; se += et; sg += lgamma(et); sel += (et - 1.0) * lp;
    }
    for (int c = 0; c < S; c++)
      for (int k = 0; k < S; k++) sA += logA[k + r * (c + (size_t)r * j)];
    t[2] += lgamma(S * e0) - S * lgamma(e0) + (e0 - 1.0) * sP;                     // Lt3
    t[3] += S * (lgamma(S * p0) - S * lgamma(p0)) + (p0 - 1.0) * sA;               // Lt4
    t[5] += (a0 - 1.0) * lO;                                                       // Lt6 (sum part)
    t[7] += -lgamma(al) + (al - 1.0) * lO;                                         // Lt8 (sum part)
    t[8] += lgamma(se) - sg + sel;                                                 // Lt9 (prior part)
    t[9] += -0.5 * d * S;                                                          // Lt10 (constant part)
  }
  // per-state terms
  const int nst = Kg * r;
  for (int x = threadIdx.x; x < nst; x += kLbThreads) {
    const int j = ja + x / r, k = x % r;
    if (k >= E.Sr[j]) continue;
    const int S = E.Sr[j];
    const size_t s = k + (size_t)r * j;
    const double lam = E.lambda[s], v = E.v[s], lL = logLam[s];
    double Wm[kMaxDim * kMaxDim];
    for (int q = 0; q < d * d; q++) Wm[q] = 0.0;
    if (E.covmode) for (int q = 0; q < dd; q++) Wm[q] = E.W[(size_t)dd * s + q];
    else for (int q = 0; q < d; q++) Wm[q + d * q] = E.W[(size_t)dd * s + q];
    double df[kMaxDim];
    for (int q = 0; q < d; q++) df[q] = E.m[q + d * s] - pr.m0[q];
    double mWm = 0.0, trW = 0.0;
    for (int a = 0; a < d; a++) {
      double u = 0.0;
      for (int b = 0; b < d; b++) u += Wm[a + d * b] * df[b];
      mWm += df[a] * u;
      trW += pr.W0inv[a] * Wm[a + d * a];
    }
    double logdetW;
    if (E.covmode) logdetW = logdet_lu(Wm, d);
    else { logdetW = 0.0; for (int q = 0; q < d; q++) logdetW += log(Wm[q + d * q]); }
    double lg = 0.0;
    for (int i = 1; i <= d; i++) lg += lgamma(0.5 * (v + 1.0 - i));
    const double logB = -0.5 * v * logdetW - 0.5 * v * d * kLog2 - cB - lg;
    const double H = -logB - 0.5 * (v - d - 1.0) * lL + 0.5 * v * d;
    t[4] += 0.5 * (d * (log(lam0) - kLog2Pi) + lL - d * lam0 / lam - lam0 * v * mWm) + logB0 +
            0.5 * (v0 - d - 1.0) * lL - 0.5 * v * trW;                             // Lt5
    double se = 0.0, sg = 0.0, sel = 0.0;
    for (int c = 0; c < S; c++) {
      const double ep = E.epsilon[k + r * (c + (size_t)r * j)], la = logA[k + r * (c + (size_t)r * j)];
      se += ep; sg += lgamma(ep); sel += (ep - 1.0) * la;
    }
    t[8] += lgamma(se) - sg + sel;                                                 // Lt9 (transition part)
    t[9] += 0.5 * (lL + d * (log(lam) - kLog2Pi)) - H;                             // Lt10
  }
  block_sum_terms(t, sh);
  if (threadIdx.x == 0) {
    double sal = 0.0, sNj = 0.0;
    for (int j = ja; j < jb; j++) { sal += E.alpha[j]; sNj += Nj[j]; }
    t[0] = lbsum[2 * g];                                                           // Lt1
    t[6] = lbsum[2 * g + 1];                                                       // Lt7
    t[5] += lgamma(Kg * a0) - Kg * lgamma(a0);                                     // Lt6
    t[7] += lgamma(sal);                                                           // Lt8
    double *o = out + 12 * g;
    for (int k = 0; k < kTerms; k++) o[k] = t[k];
    o[10] = t[0] + t[1] + t[2] + t[3] + t[4] + t[5] - t[6] - t[7] - t[8] - t[9];
    o[11] = sNj;
  }
}

// vbhem_mstep_component.m:42-70 for state x = (rho, j), in place on the hyper-parameter arrays.
// Statistics are the normalised ones (k_finalize): Nj_rho1, Nj_rho2rho, Nj_rho, y_bar, S_plus_C.
__global__ void k_mstep(EmView E, const double *Nj, const double *nu1, const double *xi, const double *pr,
                        const double *ybar, const double *SC, double *alpha, double *eta, double *epsilon, double *m,
                        double *W, double *v, double *lambda) {
  pdl_trigger();
  pdl_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = E.maxSr, d = E.d, dd = E.covmode ? d * d : d;
  if (x >= E.K * r) return;
  const int k = x % r, j = x / r;
  int g = 0;
  if (E.frozen || E.gprior) {
    while (g + 1 < E.ngroups && E.gstart[g + 1] <= j) g++;
    if (E.frozen && E.frozen[g]) return;     // batched EM runs: H3Ms that have converged keep their hyper-parameters
  }
  const Prior P0 = prior_of(E, g);
  if (k == 0) alpha[j] = P0.a0 + Nj[j];
  if (k >= E.Sr[j]) return;
  const size_t s = x;
  const double R = pr[s], lam0 = P0.lam0;
  eta[s] = P0.e0 + nu1[s];
  for (int c = 0; c < E.Sr[j]; c++) {
    const size_t o = k + r * (c + (size_t)r * j);
    epsilon[o] = P0.p0 + xi[o];
  }
  lambda[s] = lam0 + R;
  v[s] = P0.v0 + R + 1.0;
  double yb[kMaxDim], df[kMaxDim];
  for (int q = 0; q < d; q++) {
    yb[q] = ybar[k + r * (q + (size_t)d * j)];
    df[q] = yb[q] - P0.m0[q];
    m[q + d * s] = (lam0 * P0.m0[q] + R * yb[q]) / (lam0 + R);
  }
  const double mult1 = lam0 * R / (lam0 + R);
  double a[kMaxDim * kMaxDim];
  for (int c = 0; c < d; c++)
    for (int b = 0; b < d; b++) {
      const double sc = E.covmode ? SC[k + r * ((b + d * c) + (size_t)dd * j)]
                                  : (b == c ? SC[k + r * (b + (size_t)dd * j)] : 0.0);
      a[b + d * c] = (b == c ? P0.W0inv[b] : 0.0) + R * sc + mult1 * df[b] * df[c];
    }
  invert(a, d);
  if (E.covmode) {
    for (int c = 0; c < d; c++)
      for (int b = 0; b < d; b++) W[(size_t)dd * s + b + d * c] = 0.5 * (a[b + d * c] + a[c + d * b]);
  } else {
    for (int q = 0; q < d; q++) W[(size_t)dd * s + q] = a[q + d * q];
  }
}

// vbhemh3m_lb.m:202-309 for the H3M of softmax group blockIdx.x: out[g][0..4] = d LB / d {alpha0, eta0, epsilon0,
// v0, lambda0}, [5..13) d / d W0 (iid: entry 5 only, W0 = w I), [13..21) d / d m0.  Ragged centres: every "K S" of
// the reference becomes the sum over the centres of their own state counts, as in k_lower_bound.
constexpr int kDTerms = 5 + 2 * kMaxDim;
__global__ void __launch_bounds__(kLbThreads)
k_lb_deriv(EmView E, int iid, const double *logOmega, const double *logPi, const double *logA, const double *logLam,
           double *out) {
  __shared__ double sh[kLbThreads][kDTerms];
  pdl_trigger();
  pdl_wait();
  const int g = blockIdx.x;
  const int ja = E.gstart[g], jb = E.gstart[g + 1], Kg = jb - ja;
  const int r = E.maxSr, d = E.d, dd = E.covmode ? d * d : d;
  const Prior pr = prior_of(E, g);
  double t[kDTerms];
  for (int k = 0; k < kDTerms; k++) t[k] = 0.0;
  double logdetW0inv = 0.0, sps = 0.0;
  for (int q = 0; q < d; q++) logdetW0inv += log(pr.W0inv[q]);
  for (int i = 1; i <= d; i++) sps += digamma_em(0.5 * (pr.v0 + 1.0 - i));
  const double dlogB0 = 0.5 * logdetW0inv - 0.5 * d * kLog2 - 0.5 * sps;       // d logB0 / d v0
  for (int j = ja + threadIdx.x; j < jb; j += kLbThreads) {
    const int S = E.Sr[j];
    t[0] += logOmega[j];
    double sP = 0.0, sA = 0.0;
    for (int k = 0; k < S; k++) sP += logPi[k + (size_t)r * j];
    for (int c = 0; c < S; c++)
      for (int k = 0; k < S; k++) sA += logA[k + r * (c + (size_t)r * j)];
    t[1] += S * digamma_em(S * pr.e0) - S * digamma_em(pr.e0) + sP;                       // :213-217
    t[2] += S * (S * digamma_em(S * pr.p0) - S * digamma_em(pr.p0)) + sA;                 // :221-225
  }
  const int nst = Kg * r;
  for (int x = threadIdx.x; x < nst; x += kLbThreads) {
    const int j = ja + x / r, k = x % r;
    if (k >= E.Sr[j]) continue;
    const size_t s = k + (size_t)r * j;
    const double lam = E.lambda[s], v = E.v[s];
    double Wm[kMaxDim * kMaxDim];
    for (int q = 0; q < d * d; q++) Wm[q] = 0.0;
    if (E.covmode) for (int q = 0; q < dd; q++) Wm[q] = E.W[(size_t)dd * s + q];
    else for (int q = 0; q < d; q++) Wm[q + d * q] = E.W[(size_t)dd * s + q];
    double df[kMaxDim], Wdf[kMaxDim];
    for (int q = 0; q < d; q++) df[q] = E.m[q + d * s] - pr.m0[q];
    double mWm = 0.0;
    for (int a = 0; a < d; a++) {
      double u = 0.0;
      for (int b = 0; b < d; b++) u += Wm[a + d * b] * df[b];
      Wdf[a] = u;
      mWm += df[a] * u;
    }
    t[3] += dlogB0 + 0.5 * logLam[s];                                                     // :229-230
    t[4] += 0.5 * (d / pr.lam0 - d / lam - v * mWm);                                      // :233-241
    if (iid) {                                                                            // :244-267
      double tr = 0.0;
      for (int a = 0; a < d; a++) tr += Wm[a + d * a];
      t[5] += -0.5 * pr.v0 * d * pr.W0inv[0] + 0.5 * v * pr.W0inv[0] * pr.W0inv[0] * tr;
    } else {                                                                              // :270-293
      for (int a = 0; a < d; a++)
        t[5 + a] += -0.5 * pr.v0 * pr.W0inv[a] + 0.5 * v * pr.W0inv[a] * pr.W0inv[a] * Wm[a + d * a];
    }
    for (int a = 0; a < d; a++) t[5 + kMaxDim + a] += pr.lam0 * v * Wdf[a];              // :300-309
  }
  for (int k = 0; k < kDTerms; k++) sh[threadIdx.x][k] = t[k];
  __syncthreads();
  for (int s2 = kLbThreads / 2; s2 > 0; s2 >>= 1) {
    if ((int)threadIdx.x < s2)
      for (int k = 0; k < kDTerms; k++) sh[threadIdx.x][k] += sh[threadIdx.x + s2][k];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double *o = out + (size_t)kDerivStride * g;
    for (int k = 0; k < kDTerms; k++) o[k] = sh[0][k];
    o[0] += Kg * digamma_em(Kg * pr.a0) - Kg * digamma_em(pr.a0);                         // :205-209
  }
}

// vbhemhmm_init.m:58-100 (`baseem`): thread x = (state k, centre j).  NLr = Nv / Kr of the centre's own H3M.
__global__ void k_init_baseem(BaseView B, int K, int r, const int *Sr, const int *gstart, int ngroups, double alpha0,
                              double eta0, double epsilon0, double lambda0, double v0, double Nv, const int *randb,
                              const int *randg, const double *omega, double *alpha, double *eta, double *epsilon,
                              double *m, double *W, double *v, double *lambda) {
  pdl_trigger();
  pdl_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= K * r) return;
  const int k = x % r, j = x / r, d = B.d, dd = B.dd, ld = B.maxSb;
  int g = 0;
  while (g + 1 < ngroups && gstart[g + 1] <= j) g++;
  const double NLr = Nv / (double)(gstart[g + 1] - gstart[g]);
  const int s = Sr[j];
  if (k == 0) alpha[j] = alpha0 + omega[j] * Nv;
  const size_t st = x;
  for (int c = 0; c < r; c++) epsilon[k + r * (c + (size_t)r * j)] = (k < s && c < s) ? (1.0 / s) * NLr / s + epsilon0 : 0.0;
  if (k >= s) {
    eta[st] = 0.0; v[st] = 0.0; lambda[st] = 0.0;
    for (int q = 0; q < d; q++) m[q + d * st] = 0.0;
    for (int q = 0; q < dd; q++) W[(size_t)dd * st + q] = 0.0;
    return;
  }
  const double per = NLr / s, vv = v0 + per + 1.0;
  eta[st] = (1.0 / s) * NLr + eta0;
  lambda[st] = lambda0 + per;
  v[st] = vv;
  const size_t src = (size_t)randg[st] + (size_t)ld * randb[st];
  for (int q = 0; q < d; q++) m[q + d * st] = B.mu[q + d * src];
  if (B.covmode) {
    double a[kMaxDim * kMaxDim];
    for (int q = 0; q < dd; q++) a[q] = (vv - d - 1.0) * B.cov[(size_t)dd * src + q];
    invert(a, d);
    for (int q = 0; q < dd; q++) W[(size_t)dd * st + q] = a[q];
  } else {
    for (int q = 0; q < d; q++) W[(size_t)dd * st + q] = 1.0 / ((vv - d - 1.0) * B.cov[(size_t)dd * src + q]);
  }
}

}  // namespace

int launch_init_baseem(const BaseView &B, int K, int maxSr, const int *Sr, const int *gstart, int ngroups, double alpha0,
                       double eta0, double epsilon0, double lambda0, double v0, double Nv, const int *randb,
                       const int *randg, const double *omega, double *alpha, double *eta, double *epsilon, double *m,
                       double *W, double *v, double *lambda, cudaStream_t st) {
  const int tot = K * maxSr;
  launch_chain(k_init_baseem, dim3((tot + 127) / 128), dim3(128), 0, st, B, K, maxSr, Sr, gstart, ngroups, alpha0, eta0,
               epsilon0, lambda0, v0, Nv, randb, randg, omega, alpha, eta, epsilon, m, W, v, lambda);
  return 1;
}

int launch_lb_deriv(const EmView &E, bool iid, const double *logOmega, const double *logPi, const double *logA,
                    const double *logLam, double *out, cudaStream_t st) {
  static_assert(kDTerms == kDerivStride, "derivative layout");
  launch_chain(k_lb_deriv, dim3(E.ngroups), dim3(kLbThreads), 0, st, E, iid ? 1 : 0, logOmega, logPi, logA, logLam, out);
  return 1;
}

int launch_lower_bound(const EmView &E, const double *Nj, const double *lbsum, const double *logOmega,
                       const double *logPi, const double *logA, const double *logLam, double *out, cudaStream_t st) {
  launch_chain(k_lower_bound, dim3(E.ngroups), dim3(kLbThreads), 0, st, E, Nj, lbsum, logOmega, logPi, logA, logLam, out);
  return 1;
}

int launch_mstep(const EmView &E, const double *Nj, const double *nu1, const double *xi, const double *pr,
                 const double *ybar, const double *SC, double *alpha, double *eta, double *epsilon, double *m,
                 double *W, double *v, double *lambda, cudaStream_t st) {
  const int tot = E.K * E.maxSr;
  launch_chain(k_mstep, dim3((tot + 127) / 128), dim3(128), 0, st, E, Nj, nu1, xi, pr, ybar, SC, alpha, eta, epsilon, m,
               W, v, lambda);
  return 1;
}

}  // namespace vbhem
