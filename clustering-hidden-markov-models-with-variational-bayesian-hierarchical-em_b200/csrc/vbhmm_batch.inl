// vbhmm_batch.inl -- many VB-HMM learning runs at once, device-resident (included at the end of vbhmm_fb.cu).
//
// The toolbox learns one HMM per subject (/root/reference/src/hmm/vbhmm_learn_batch.m loops vbhmm_learn over the
// subjects; vbhmm_learn.m:465-557 loops vbhmm_em over random initialisations and candidate state counts) and then
// clusters the HMMs with VBHEM.  A subject's EM run touches tens of short sequences -- far too little for a GPU --
// but all those runs are independent.  Here a RUN is (a range of the packed sequences, its own K-state model);
// every run of a batch executes the loop of /root/reference/src/hmm/vbhmm_em.m:108-418 on the device:
//   k_batch_prepass   thread per run: logLambdaTilde, logATilde, logPiTilde (vbhmm_fb.m:62-80) -> the run's FbPar
//   k_fb_batch        thread per (run, sequence): the forward-backward pass of k_fb, model read through L1
//   k_batch_sums      block per run: the sums of vbhmm_em.m:158-246 in a fixed order
//   k_batch_update    thread per run: the lower bound (vbhmm_em_lb.m:74-258), the convergence test
//                     (vbhmm_em.m:293-330) and the M-step (:353-389)
// and the host only reads back how many runs are still going.  Runs that have stopped cost nothing further.

namespace {

constexpr int kBatchThreads = 128;

// Variational parameters of one run, C-ABI (MATLAB) layouts.
template <int K, int DIM>
struct VarPar {
  double alpha[K];
  double eps[K][K];        // eps[to][from] = epsilon(from, to): column-major [K(from), K(to)]
  double beta[K], v[K];
  double m[K][DIM];        // m(:, k)
  double W[K][DIM][DIM];   // W[k][j][i] = W(i, j, k): column-major
};

// What the lower bound needs from the pre-pass of the E-step it belongs to.
template <int K>
struct PreOut { double logPi[K], logA[K][K] /* [from][to] */, lLT[K], logdetW[K]; };

struct BatchPrior { double alpha0, epsilon0, beta0, v0, m0[kFbMaxDim], W0inv[kFbMaxDim], logdetW0inv; };

__device__ double dev_digamma(double x) {   // psi(x), x > 0: recurrence to x >= 10, then the asymptotic series
  if (!(x > 0.0)) return NAN;
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  const double ser = f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132 -
                     f * (691.0 / 32760 - f * (1.0 / 12)))))));
  return r + log(x) - 0.5 / x - ser;
}

// In-place inverse of a DIM x DIM matrix by Gauss-Jordan with partial pivoting; returns log|det| of the input.
template <int DIM>
__device__ double dev_inverse(double (&a)[DIM][DIM]) {
  double inv[DIM][DIM];
  for (int i = 0; i < DIM; i++)
    for (int j = 0; j < DIM; j++) inv[i][j] = i == j ? 1.0 : 0.0;
  double logdet = 0.0;
  for (int c = 0; c < DIM; c++) {
    int piv = c;
    for (int r = c + 1; r < DIM; r++)
      if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    if (piv != c)
      for (int k = 0; k < DIM; k++) {
        double t = a[c][k]; a[c][k] = a[piv][k]; a[piv][k] = t;
        t = inv[c][k]; inv[c][k] = inv[piv][k]; inv[piv][k] = t;
      }
    const double p = a[c][c];
    logdet += log(fabs(p));
    const double ip = 1.0 / p;
    for (int k = 0; k < DIM; k++) { a[c][k] *= ip; inv[c][k] *= ip; }
    for (int r = 0; r < DIM; r++) {
      if (r == c) continue;
      const double f = a[r][c];
      for (int k = 0; k < DIM; k++) { a[r][k] -= f * a[c][k]; inv[r][k] -= f * inv[c][k]; }
    }
  }
  for (int i = 0; i < DIM; i++)
    for (int j = 0; j < DIM; j++) a[i][j] = inv[i][j];
  return logdet;
}

template <int K, int DIM>
__global__ void k_batch_prepass(int R, const unsigned char *__restrict__ done, const VarPar<K, DIM> *__restrict__ vp,
                                FbPar<K, DIM> *__restrict__ par, PreOut<K> *__restrict__ pre) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R || done[r]) return;
  const VarPar<K, DIM> &V = vp[r];
  FbPar<K, DIM> &P = par[r];
  PreOut<K> &O = pre[r];
  double sa = 0.0;
  for (int k = 0; k < K; k++) sa += V.alpha[k];
  const double psa = dev_digamma(sa);
  for (int k = 0; k < K; k++) {
    O.logPi[k] = dev_digamma(V.alpha[k]) - psa;                       // vbhmm_fb.m:76-77
    P.pz1[k] = exp(O.logPi[k]);
  }
  for (int j = 0; j < K; j++) {                                       // row j = from-state (vbhmm_fb.m:68-74)
    double se = 0.0;
    for (int k = 0; k < K; k++) se += V.eps[k][j];
    const double pse = dev_digamma(se);
    for (int k = 0; k < K; k++) {
      O.logA[j][k] = dev_digamma(V.eps[k][j]) - pse;
      P.A[j][k] = exp(O.logA[j][k]);
    }
  }
  for (int k = 0; k < K; k++) {
    double Wm[DIM][DIM];
    for (int i = 0; i < DIM; i++)
      for (int j = 0; j < DIM; j++) Wm[i][j] = V.W[k][j][i];
    const double ld = dev_inverse<DIM>(Wm);                           // log det W(:,:,k)
    double t1 = 0.0;
    for (int i = 1; i <= DIM; i++) t1 += dev_digamma(0.5 * (V.v[k] + 1.0) - 0.5 * i);   // vbhmm_fb.m:62-66
    O.logdetW[k] = ld;
    O.lLT[k] = t1 + DIM * 0.6931471805599453094 + ld;
    P.lLT[k] = O.lLT[k];
    P.v[k] = V.v[k];
    P.dob[k] = DIM / V.beta[k];
    for (int i = 0; i < DIM; i++) {
      P.m[k][i] = V.m[k][i];
      for (int j = 0; j < DIM; j++) P.W[k][i][j] = V.W[k][i][j];      // FbPar::W[k][i][j] = W(j,i,k), see vbhmm_fb.cu
    }
  }
  P.cden = DIM * 0.5 * 1.8378770664093454836;                         // dim * log(2 pi) / 2
  fb_finish_par(P);
}

// k_fb for work item w = (run, sequence): the model of the item's run comes from HBM through L1 (the items of a
// run are neighbours, so a warp reads one or two models), gamma / xi / phi are indexed by the work item.
template <int K, int DIM>
__global__ void __launch_bounds__(kFbThreads, K <= 6 ? 3 : 2)      // the K x K pair marginals live in registers
k_fb_batch(const FbPar<K, DIM> *__restrict__ pars, const unsigned char *__restrict__ done, int Wn, int N, int maxT,
           const int *__restrict__ run_of, const int *__restrict__ seq_of, const int *__restrict__ Tn,
           const double *__restrict__ X, double *__restrict__ gamma, double *__restrict__ xi_sum,
           double *__restrict__ phi) {
  __shared__ __align__(16) double s_tab[kFbTabDoubles];
  vbhem::tab_t etab, ltab;
  stage_tables<kFbThreads>(s_tab, etab, ltab);
  const int w = blockIdx.x * kFbThreads + threadIdx.x;
  if (w >= Wn) return;
  const int run = run_of[w];
  if (done[run]) return;
  const FbPar<K, DIM> &P = pars[run];
  const int n = seq_of[w], T = Tn[n];
  const size_t WK = (size_t)Wn * K, ND = (size_t)N * DIM;
  const double *xp = X + (size_t)DIM * n;
  double *gp = gamma + (size_t)K * w;
  double alpha[K], p[K], lr[K], x[DIM];
  double sumlogc = 0.0, summax = 0.0;
  for (int t = 0; t < T; t++) {
#pragma unroll
    for (int i = 0; i < DIM; i++) x[i] = xp[ND * t + i];
    emissions<K, DIM>(P, x, lr);
    summax += shifted_exp<K>(lr, p, etab);
    double delta[K], c;
    if (t == 0) {
      c = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) { delta[k] = P.pz1[k] * p[k]; c += delta[k]; }
    } else {
      c = forward_step<K, DIM>(P, alpha, p, delta);
    }
    const double rc = rcp_pos(c);
#pragma unroll
    for (int k = 0; k < K; k++) { alpha[k] = delta[k] * rc; gp[WK * t + k] = alpha[k]; }
    sumlogc += vbhem::log_shift<1>(c, ltab, 0);
  }
  phi[w] = sumlogc * vbhem::kUnit + summax;
  // backward pass as in k_fb (vbhmm_fb.cu): the emissions are not evaluated again -- p_{t+1}(k) / c_{t+1} =
  // alpha_hat_{t+1}(k) / sum_j alpha_hat_t(j) A(j,k) -- and the pair marginals accumulate in registers as
  // sum_t alpha_hat_t (x) (beta_hat p / c), multiplied by A once at the end.  alpha = alpha_hat(t+1) on entry.
  double Xi[K][K], beta[K];
#pragma unroll
  for (int j = 0; j < K; j++) {
    beta[j] = 1.0;
#pragma unroll
    for (int k = 0; k < K; k++) Xi[j][k] = 0.0;
  }
  for (int t = T - 2; t >= 0; t--) {
    double a[K], bw[K];
#pragma unroll
    for (int k = 0; k < K; k++) a[k] = gp[WK * t + k];
#pragma unroll
    for (int k = 0; k < K; k++) {
      double pred = 1e-300;
#pragma unroll
      for (int j = 0; j < K; j++) pred += a[j] * P.A[j][k];
      bw[k] = beta[k] * (alpha[k] * rcp_pos(pred));
    }
#pragma unroll
    for (int j = 0; j < K; j++) {
      double eta = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) eta += bw[k] * P.A[j][k];
      beta[j] = eta;
#pragma unroll
      for (int k = 0; k < K; k++) Xi[j][k] = fma(a[j], bw[k], Xi[j][k]);
    }
#pragma unroll
    for (int k = 0; k < K; k++) { gp[WK * t + k] = a[k] * beta[k]; alpha[k] = a[k]; }
  }
  double *xo = xi_sum + (size_t)K * K * w;
#pragma unroll
  for (int j = 0; j < K; j++)
#pragma unroll
    for (int k = 0; k < K; k++) xo[j * K + k] = P.A[j][k] * Xi[j][k];   // [from][to]
}

// Per-run state of the EM loop (vbhmm_em.m:108-418).
struct RunState { double L, lastL; int it, done, unstable, pad; };

template <int K, int DIM>
struct BatchStats { double Nk1[K], Nk[K], M[K][K] /* [from][to] */, xbar[K][DIM], S[K][DIM][DIM]; };

// Block per run: the raw sums of vbhmm_em.m:158-246 (moments about the current means, as k_fb_moments) in a fixed order.
template <int K, int DIM>
__global__ void __launch_bounds__(kBatchThreads)
k_batch_sums(int Wn, int N, int maxT, const int *__restrict__ wstart, const int *__restrict__ seq_of,
             const int *__restrict__ Tn, const double *__restrict__ X, const double *__restrict__ gamma,
             const double *__restrict__ xi_sum, const double *__restrict__ phi,
             const FbPar<K, DIM> *__restrict__ pars, const unsigned char *__restrict__ done,
             double *__restrict__ sums) {
  using L = MomLayout<K, DIM>;
  const int r = blockIdx.x;
  if (done[r]) return;
  const FbPar<K, DIM> &P = pars[r];
  const int w0 = wstart[r], w1 = wstart[r + 1];
  const size_t WK = (size_t)Wn * K, ND = (size_t)N * DIM;
  double acc[L::total];
#pragma unroll
  for (int v = 0; v < L::total; v++) acc[v] = 0.0;
  // thread -> (work item, time step) of this run, items fastest: consecutive threads read consecutive gamma
  const int nw = w1 - w0;
  for (long long e = threadIdx.x; e < (long long)nw * maxT; e += kBatchThreads) {
    const int w = w0 + (int)(e % nw), t = (int)(e / nw);
    const int n = seq_of[w];
    if (t >= Tn[n]) continue;
    double g[K], x[DIM], lr[K];
#pragma unroll
    for (int k = 0; k < K; k++) g[k] = gamma[WK * t + (size_t)K * w + k];
#pragma unroll
    for (int i = 0; i < DIM; i++) x[i] = X[ND * t + (size_t)DIM * n + i];
    emissions<K, DIM>(P, x, lr);
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (t == 0) acc[L::nk1 + k] += g[k];
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
  for (int w = w0 + threadIdx.x; w < w1; w += kBatchThreads) {
    acc[L::lt64] += phi[w];
#pragma unroll
    for (int q = 0; q < K * K; q++) acc[L::M + q] += xi_sum[(size_t)K * K * w + q];   // [from][to]
  }
  // block reduction in a fixed order
  __shared__ double sh[kBatchThreads / 32][L::total];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < L::total; v++) {
    double a = acc[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) sh[warp][v] = a;
  }
  __syncthreads();
  for (int v = threadIdx.x; v < L::total; v += kBatchThreads) {
    double a = 0.0;
    for (int q = 0; q < kBatchThreads / 32; q++) a += sh[q][v];
    sums[(size_t)L::total * r + v] = a;
  }
}

// Thread per run: normalisation, lower bound, convergence test, M-step.
template <int K, int DIM>
__global__ void __launch_bounds__(kBatchThreads)
k_batch_update(int R, const double *__restrict__ sums, const FbPar<K, DIM> *__restrict__ pars,
               const PreOut<K> *__restrict__ pre, BatchPrior pr, VarPar<K, DIM> *__restrict__ vp,
               BatchStats<K, DIM> *__restrict__ stats, RunState *__restrict__ state, unsigned char *__restrict__ done,
               double *__restrict__ LLs, int ld, int max_iter, double min_diff, int *__restrict__ n_left) {
  using L = MomLayout<K, DIM>;
  const int r = blockIdx.x * kBatchThreads + threadIdx.x;
  if (r >= R || done[r]) return;
  const FbPar<K, DIM> &P = pars[r];
  const double *s = sums + (size_t)L::total * r;

  // ---- normalisation (vbhmm_em.m:161-246), as k_fb_moments_finish
  BatchStats<K, DIM> &B = stats[r];
  VarPar<K, DIM> &V = vp[r];
  const PreOut<K> &O = pre[r];
  RunState &st = state[r];
  for (int k = 0; k < K; k++) {
    const double nk = s[L::nk + k], Nk = nk + 1e-50;
    B.Nk1[k] = s[L::nk1 + k] + 1e-50;
    B.Nk[k] = Nk;
    double del[DIM];
    for (int i = 0; i < DIM; i++) {
      B.xbar[k][i] = (P.m[k][i] * nk + s[L::sx + k * DIM + i]) / Nk;
      del[i] = B.xbar[k][i] - P.m[k][i];
    }
    for (int i = 0; i < DIM; i++)
      for (int j = 0; j < DIM; j++)
        B.S[k][i][j] = (s[L::sxx + (k * DIM + i) * DIM + j] - s[L::sx + k * DIM + i] * del[j] -
                        del[i] * s[L::sx + k * DIM + j] + nk * del[i] * del[j]) / Nk;
    for (int j = 0; j < K; j++) B.M[k][j] = s[L::M + k * K + j];
  }

  // ---- lower bound (vbhmm_em_lb.m:74-258) of the parameters this E-step ran with
  const double kLog2 = 0.6931471805599453094, kLogPi = 1.1447298858494001741, kLog2Pi = 1.8378770664093454836;
  const double a0 = pr.alpha0, e0 = pr.epsilon0, b0 = pr.beta0, v0 = pr.v0;
  auto lgam_d = [&](double vv) {   // sum_i gammaln(0.5 (vv + 1 - i))
    double t = 0.0;
    for (int i = 1; i <= DIM; i++) t += lgamma(0.5 * (vv + 1.0 - i));
    return t;
  };
  const double cB = DIM * (DIM - 1) * 0.25 * kLogPi;
  const double logB0 = 0.5 * v0 * pr.logdetW0inv - 0.5 * v0 * DIM * kLog2 - cB - lgam_d(v0);
  double H = 0.0, Lt1 = 0.0, Lt51 = 0.0, sum_lLT = 0.0, sum_vtr = 0.0, Lt8a = 0.0;
  for (int k = 0; k < K; k++) {
    const double vk = V.v[k], bk = V.beta[k], lLT = O.lLT[k];
    const double logBk = -0.5 * vk * O.logdetW[k] - 0.5 * vk * DIM * kLog2 - cB - lgam_d(vk);
    H += -logBk - 0.5 * (vk - DIM - 1) * lLT + 0.5 * vk * DIM;
    double trSW = 0.0, xWx = 0.0, mWm = 0.0, trW0W = 0.0;
    for (int i = 0; i < DIM; i++) {
      for (int j = 0; j < DIM; j++) {
        const double Wij = V.W[k][j][i];
        trSW += B.S[k][i][j] * Wij;                                   // trace(S W), both symmetric up to rounding
        xWx += (B.xbar[k][i] - V.m[k][i]) * Wij * (B.xbar[k][j] - V.m[k][j]);
        mWm += (V.m[k][i] - pr.m0[i]) * Wij * (V.m[k][j] - pr.m0[j]);
      }
      trW0W += pr.W0inv[i] * V.W[k][i][i];
    }
    Lt1 += B.Nk[k] * (lLT - DIM / bk - vk * trSW - vk * xWx - DIM * kLog2Pi);
    Lt51 += DIM * log(b0 / (2.0 * 3.14159265358979323846)) + lLT - DIM * b0 / bk - b0 * vk * mWm;
    sum_lLT += lLT;
    sum_vtr += vk * trW0W;
    Lt8a += lLT + DIM * log(bk / (2.0 * 3.14159265358979323846));
  }
  Lt1 *= 0.5;
  const double Lt5 = 0.5 * Lt51 + K * logB0 + 0.5 * (v0 - DIM - 1) * sum_lLT - 0.5 * sum_vtr;
  double Lt2a = 0.0, Lt2b = 0.0, sumPi = 0.0, sumA = 0.0, Lt71 = 0.0, Lt72 = 0.0, sal = 0.0;
  for (int k = 0; k < K; k++) {
    Lt2a += B.Nk1[k] * O.logPi[k];
    sumPi += O.logPi[k];
    Lt71 += (V.alpha[k] - 1.0) * O.logPi[k] - lgamma(V.alpha[k]);
    sal += V.alpha[k];
    double se = 0.0;
    for (int j = 0; j < K; j++) {   // row k = from
      Lt2b += B.M[k][j] * O.logA[k][j];
      sumA += O.logA[k][j];
      Lt72 += (V.eps[j][k] - 1.0) * O.logA[k][j] - lgamma(V.eps[j][k]);
      se += V.eps[j][k];
    }
    Lt72 += lgamma(se);
  }
  Lt71 += lgamma(sal);
  const double Lt2 = Lt2a + Lt2b;
  const double Lt3 = lgamma(K * a0) - K * lgamma(a0) + (a0 - 1.0) * sumPi;
  const double Lt4 = K * (lgamma(K * e0) - K * lgamma(e0)) + (e0 - 1.0) * sumA;
  const double Lt6 = Lt2a + Lt2b + s[L::lt63] - s[L::lt64];
  const double Lt7 = Lt71 + Lt72;
  const double Lt8 = 0.5 * Lt8a - 0.5 * DIM * K - H;
  double Lb = Lt1 + Lt2 + Lt3 + Lt4 + Lt5 - Lt6 - Lt7 - Lt8;

  // ---- convergence test (vbhmm_em.m:278-330); st.it counts the passes made so far
  const int iter = st.it + 1;
  if (iter > 1) st.lastL = st.L;
  bool do_break = false;
  if (iter > 1 && fabs((Lb - st.lastL) / st.lastL) <= min_diff) do_break = true;
  if (iter == max_iter) do_break = true;
  if (Lb != Lb) { do_break = true; st.unstable = 1; Lb = -INFINITY; }
  st.L = Lb;
  st.it = iter;
  LLs[(size_t)ld * r + (iter - 1)] = Lb;
  if (!(do_break && st.unstable)) {
    // ---- M-step (vbhmm_em.m:353-389)
    for (int k = 0; k < K; k++) {
      V.alpha[k] = a0 + B.Nk1[k];
      for (int j = 0; j < K; j++) V.eps[j][k] = e0 + B.M[k][j];
      const double Nk = B.Nk[k];
      V.beta[k] = b0 + Nk;
      V.v[k] = v0 + Nk + 1.0;
      const double mult1 = b0 * Nk / (b0 + Nk);
      double A[DIM][DIM];
      for (int i = 0; i < DIM; i++) {
        V.m[k][i] = (b0 * pr.m0[i] + Nk * B.xbar[k][i]) / V.beta[k];
        for (int j = 0; j < DIM; j++)
          A[i][j] = (i == j ? pr.W0inv[i] : 0.0) + Nk * B.S[k][i][j] +
                    mult1 * (B.xbar[k][i] - pr.m0[i]) * (B.xbar[k][j] - pr.m0[j]);
      }
      dev_inverse<DIM>(A);
      for (int i = 0; i < DIM; i++)
        for (int j = 0; j < DIM; j++) V.W[k][j][i] = 0.5 * (A[i][j] + A[j][i]);
    }
  }
  if (do_break) {
    st.done = 1;
    done[r] = 1;
    atomicSub(n_left, 1);
  }
}

}  // namespace

struct vbhmm_batch {
  vbhmm_data *data = nullptr;
  int R = 0, K = 0, Wn = 0;
  BatchPrior prior{};
  Buf run_of, seq_of, wstart, vp, par, pre, stats, sums, state, done, LLs, gamma, xi, phi, nleft;
  int ld = 0;
  int *h_left = nullptr;   // pinned
  size_t vp_bytes = 0, stats_bytes = 0;
  double last_ms = 0.0;
  int passes = 0;
};

namespace {

template <int K, int DIM>
int batch_alloc(vbhmm_batch *b) {
  CU(b->vp.ensure(sizeof(VarPar<K, DIM>) * (size_t)b->R));
  CU(b->par.ensure(sizeof(FbPar<K, DIM>) * (size_t)b->R));
  CU(b->pre.ensure(sizeof(PreOut<K>) * (size_t)b->R));
  CU(b->stats.ensure(sizeof(BatchStats<K, DIM>) * (size_t)b->R));
  CU(b->sums.ensure(sizeof(double) * MomLayout<K, DIM>::total * (size_t)b->R));
  b->vp_bytes = sizeof(VarPar<K, DIM>);
  b->stats_bytes = sizeof(BatchStats<K, DIM>);
  return VBHEM_OK;
}

template <int K, int DIM>
int batch_pass(vbhmm_batch *b, int max_iter, double min_diff) {
  vbhmm_data *h = b->data;
  const int R = b->R, Wn = b->Wn;
  k_batch_prepass<K, DIM><<<(R + 127) / 128, 128, 0, h->st>>>(R, b->done.as<unsigned char>(), b->vp.as<VarPar<K, DIM>>(),
                                                           b->par.as<FbPar<K, DIM>>(), b->pre.as<PreOut<K>>());
  k_fb_batch<K, DIM><<<(Wn + kFbThreads - 1) / kFbThreads, kFbThreads, 0, h->st>>>(
      b->par.as<FbPar<K, DIM>>(), b->done.as<unsigned char>(), Wn, h->N, h->maxT, b->run_of.as<int>(),
      b->seq_of.as<int>(), h->Tn.as<int>(), h->X.as<double>(), b->gamma.as<double>(), b->xi.as<double>(),
      b->phi.as<double>());
  k_batch_sums<K, DIM><<<R, kBatchThreads, 0, h->st>>>(
      Wn, h->N, h->maxT, b->wstart.as<int>(), b->seq_of.as<int>(), h->Tn.as<int>(), h->X.as<double>(),
      b->gamma.as<double>(), b->xi.as<double>(), b->phi.as<double>(), b->par.as<FbPar<K, DIM>>(),
      b->done.as<unsigned char>(), b->sums.as<double>());
  k_batch_update<K, DIM><<<(R + kBatchThreads - 1) / kBatchThreads, kBatchThreads, 0, h->st>>>(
      R, b->sums.as<double>(), b->par.as<FbPar<K, DIM>>(), b->pre.as<PreOut<K>>(), b->prior,
      b->vp.as<VarPar<K, DIM>>(), b->stats.as<BatchStats<K, DIM>>(), b->state.as<RunState>(),
      b->done.as<unsigned char>(), b->LLs.as<double>(), b->ld, max_iter, min_diff, b->nleft.as<int>());
  return VBHEM_OK;
}

template <typename F>
int batch_dispatch(int K, int dim, F f) {
#define VB_BD(KK)                                                   \
  case KK:                                                          \
    switch (dim) {                                                  \
      case 1: return f(std::integral_constant<int, KK>{}, std::integral_constant<int, 1>{}); \
      case 2: return f(std::integral_constant<int, KK>{}, std::integral_constant<int, 2>{}); \
      case 3: return f(std::integral_constant<int, KK>{}, std::integral_constant<int, 3>{}); \
    }                                                               \
    break;
  switch (K) { VB_BD(1) VB_BD(2) VB_BD(3) VB_BD(4) VB_BD(5) VB_BD(6) VB_BD(7) VB_BD(8) }
#undef VB_BD
  return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: batched learning takes K <= %d, dim <= %d", kFbMaxK, kFbMaxDim);
}

}  // namespace

extern "C" {

int vbhmm_batch_free(vbhmm_batch_t *b) {
  if (!b) return VBHEM_OK;
  if (b->data) cudaSetDevice(b->data->device);
  Buf *all[] = {&b->run_of, &b->seq_of, &b->wstart, &b->vp, &b->par, &b->pre, &b->stats, &b->sums, &b->state, &b->done, &b->LLs,
                &b->gamma, &b->xi, &b->phi, &b->nleft};
  for (Buf *x : all) x->release();
  if (b->h_left) cudaFreeHost(b->h_left);
  delete b;
  return VBHEM_OK;
}

int vbhmm_batch_begin(vbhmm_data_t *h, int R, const int32_t *seq_lo, const int32_t *seq_hi, int K, double alpha0,
                      double epsilon0, double beta0, double v0, const double *m0, const double *W0, int nW0,
                      const double *alpha, const double *epsilon, const double *beta, const double *v, const double *m,
                      const double *W, int max_iter, vbhmm_batch_t **out) {
  if (!out) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: out is NULL");
  *out = nullptr;
  if (!h) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: data handle is NULL");
  if (R < 1 || !seq_lo || !seq_hi) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: R >= 1 runs with seq_lo / seq_hi required");
  if (K < 1 || K > kFbMaxK || h->dim > kFbMaxDim)
    return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: batched learning takes K <= %d, dim <= %d", kFbMaxK, kFbMaxDim);
  if (!m0 || !W0 || (nW0 != 1 && nW0 != h->dim) || !alpha || !epsilon || !beta || !v || !m || !W)
    return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: a prior or model array is NULL");
  if (!(v0 > h->dim - 1)) return fail_fb(VBHEM_EINVAL, "v0 not large enough");   // vbhmm_init.m:149-151
  if (max_iter < 1) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: max_iter must be >= 1");
  const int dim = h->dim;
  long long Wn = 0;
  for (int r = 0; r < R; r++) {
    if (seq_lo[r] < 0 || seq_hi[r] > h->N || seq_hi[r] <= seq_lo[r])
      return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: run %d covers sequences [%d, %d) of %d", r, seq_lo[r], seq_hi[r], h->N);
    Wn += seq_hi[r] - seq_lo[r];
  }
  if (Wn > 2000000000LL) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: too many (run, sequence) work items");
  CU(cudaSetDevice(h->device));
  vbhmm_batch *b = new (std::nothrow) vbhmm_batch;
  if (!b) return fail_fb(VBHEM_ENOMEM, "out of host memory");
  b->data = h; b->R = R; b->K = K; b->Wn = (int)Wn; b->ld = max_iter;
  auto bail = [&](int code) { vbhmm_batch_free(b); return code; };
#define CUB(x)                                                                                               \
  do {                                                                                                       \
    cudaError_t e_ = (x);                                                                                    \
    if (e_ != cudaSuccess)                                                                                   \
      return bail(fail_fb(e_ == cudaErrorMemoryAllocation ? VBHEM_ENOMEM : VBHEM_ECUDA, "%s: %s", #x,        \
                          cudaGetErrorString(e_)));                                                          \
  } while (0)
  std::vector<int> run_of((size_t)Wn), seq_of((size_t)Wn), wstart(R + 1, 0);
  {
    size_t w = 0;
    for (int r = 0; r < R; r++) {
      wstart[r] = (int)w;
      for (int n = seq_lo[r]; n < seq_hi[r]; n++, w++) { run_of[w] = r; seq_of[w] = n; }
    }
    wstart[R] = (int)w;
  }
  BatchPrior &P = b->prior;
  P.alpha0 = alpha0; P.epsilon0 = epsilon0; P.beta0 = beta0; P.v0 = v0; P.logdetW0inv = 0.0;
  for (int i = 0; i < kFbMaxDim; i++) { P.m0[i] = 0.0; P.W0inv[i] = 1.0; }
  for (int i = 0; i < dim; i++) {
    P.m0[i] = m0[i];
    P.W0inv[i] = 1.0 / W0[nW0 == 1 ? 0 : i];
    P.logdetW0inv += log(P.W0inv[i]);                                 // vbhmm_em_lb.m:75-82
  }
  int rc = batch_dispatch(K, dim, [&](auto Kc, auto Dc) { return batch_alloc<decltype(Kc)::value, decltype(Dc)::value>(b); });
  if (rc) return bail(rc);
  const int maxT = h->maxT > 0 ? h->maxT : 1;
  CUB(b->run_of.ensure(sizeof(int) * (size_t)Wn));
  CUB(b->seq_of.ensure(sizeof(int) * (size_t)Wn));
  CUB(b->wstart.ensure(sizeof(int) * (size_t)(R + 1)));
  CUB(b->state.ensure(sizeof(RunState) * (size_t)R));
  CUB(b->done.ensure((size_t)R));
  CUB(b->LLs.ensure(sizeof(double) * (size_t)R * max_iter));
  CUB(b->gamma.ensure(sizeof(double) * (size_t)K * Wn * maxT));
  CUB(b->xi.ensure(sizeof(double) * (size_t)K * K * Wn));
  CUB(b->phi.ensure(sizeof(double) * (size_t)Wn));
  CUB(b->nleft.ensure(sizeof(int)));
  CUB(cudaMallocHost(&b->h_left, sizeof(int)));
  // variational parameters: the C-ABI arrays are exactly VarPar<K, DIM> run after run when interleaved per run
  std::vector<double> host((size_t)R * b->vp_bytes / sizeof(double));
  const size_t per = b->vp_bytes / sizeof(double);
  for (int r = 0; r < R; r++) {
    double *q = host.data() + per * r;
    memcpy(q, alpha + (size_t)K * r, sizeof(double) * K); q += K;
    memcpy(q, epsilon + (size_t)K * K * r, sizeof(double) * K * K); q += (size_t)K * K;
    memcpy(q, beta + (size_t)K * r, sizeof(double) * K); q += K;
    memcpy(q, v + (size_t)K * r, sizeof(double) * K); q += K;
    memcpy(q, m + (size_t)dim * K * r, sizeof(double) * dim * K); q += (size_t)dim * K;
    memcpy(q, W + (size_t)dim * dim * K * r, sizeof(double) * dim * dim * K);
  }
  CUB(cudaMemcpyAsync(b->vp.p, host.data(), b->vp_bytes * (size_t)R, cudaMemcpyHostToDevice, h->st));
  CUB(cudaMemcpyAsync(b->run_of.p, run_of.data(), sizeof(int) * (size_t)Wn, cudaMemcpyHostToDevice, h->st));
  CUB(cudaMemcpyAsync(b->seq_of.p, seq_of.data(), sizeof(int) * (size_t)Wn, cudaMemcpyHostToDevice, h->st));
  CUB(cudaMemcpyAsync(b->wstart.p, wstart.data(), sizeof(int) * (size_t)(R + 1), cudaMemcpyHostToDevice, h->st));
  CUB(cudaMemsetAsync(b->state.p, 0, sizeof(RunState) * (size_t)R, h->st));
  CUB(cudaMemsetAsync(b->done.p, 0, (size_t)R, h->st));
  CUB(cudaMemsetAsync(b->LLs.p, 0, sizeof(double) * (size_t)R * max_iter, h->st));
  CUB(cudaMemcpyAsync(b->nleft.p, &R, sizeof(int), cudaMemcpyHostToDevice, h->st));
  CUB(cudaStreamSynchronize(h->st));
#undef CUB
  *out = b;
  return VBHEM_OK;
}

int vbhmm_batch_run(vbhmm_batch_t *b, double min_diff, int *passes) {
  if (!b) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: batch handle is NULL");
  vbhmm_data *h = b->data;
  CU(cudaSetDevice(h->device));
  const int max_iter = b->ld;
  CU(cudaEventRecord(h->ev[0], h->st));
  int left = b->R, n = 0;
  while (left > 0 && n < max_iter) {
    int rc = batch_dispatch(b->K, h->dim, [&](auto Kc, auto Dc) {
      return batch_pass<decltype(Kc)::value, decltype(Dc)::value>(b, max_iter, min_diff);
    });
    if (rc) return rc;
    CU(cudaMemcpyAsync(b->h_left, b->nleft.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    CU(cudaGetLastError());
    left = *b->h_left;
    n++;
  }
  CU(cudaEventRecord(h->ev[1], h->st));
  CU(cudaEventSynchronize(h->ev[1]));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
  b->last_ms = ms;
  b->passes = n;
  if (passes) *passes = n;
  return VBHEM_OK;
}

int vbhmm_batch_fetch(vbhmm_batch_t *b, double *alpha, double *epsilon, double *beta, double *v, double *m, double *W,
                      double *Nk1, double *Nk, double *M, double *LLs, int32_t *iters, int32_t *unstable) {
  if (!b) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: batch handle is NULL");
  vbhmm_data *h = b->data;
  CU(cudaSetDevice(h->device));
  const int R = b->R, K = b->K, dim = h->dim;
  std::vector<double> hv((size_t)R * b->vp_bytes / sizeof(double)), hs((size_t)R * b->stats_bytes / sizeof(double));
  std::vector<RunState> st(R);
  CU(cudaMemcpyAsync(hv.data(), b->vp.p, b->vp_bytes * (size_t)R, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(hs.data(), b->stats.p, b->stats_bytes * (size_t)R, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(st.data(), b->state.p, sizeof(RunState) * (size_t)R, cudaMemcpyDeviceToHost, h->st));
  if (LLs) CU(cudaMemcpyAsync(LLs, b->LLs.p, sizeof(double) * (size_t)R * b->ld, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  const size_t per = b->vp_bytes / sizeof(double), pers = b->stats_bytes / sizeof(double);
  for (int r = 0; r < R; r++) {
    const double *q = hv.data() + per * r;
    if (alpha) memcpy(alpha + (size_t)K * r, q, sizeof(double) * K);
    q += K;
    if (epsilon) memcpy(epsilon + (size_t)K * K * r, q, sizeof(double) * K * K);
    q += (size_t)K * K;
    if (beta) memcpy(beta + (size_t)K * r, q, sizeof(double) * K);
    q += K;
    if (v) memcpy(v + (size_t)K * r, q, sizeof(double) * K);
    q += K;
    if (m) memcpy(m + (size_t)dim * K * r, q, sizeof(double) * dim * K);
    q += (size_t)dim * K;
    if (W) memcpy(W + (size_t)dim * dim * K * r, q, sizeof(double) * dim * dim * K);
    const double *s = hs.data() + pers * r;                           // Nk1 | Nk | M[from][to]
    if (Nk1) memcpy(Nk1 + (size_t)K * r, s, sizeof(double) * K);
    if (Nk) memcpy(Nk + (size_t)K * r, s + K, sizeof(double) * K);
    if (M)   // BatchStats::M is [from][to]; the ABI is column-major [K(from), K(to), R] like epsilon
      for (int from = 0; from < K; from++)
        for (int to = 0; to < K; to++) M[from + K * (to + (size_t)K * r)] = s[2 * K + from * K + to];
    if (iters) iters[r] = st[r].it;
    if (unstable) unstable[r] = st[r].unstable;
  }
  return VBHEM_OK;
}

int vbhmm_batch_last_ms(vbhmm_batch_t *b, double *ms) {
  if (!b || !ms) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: NULL argument");
  *ms = b->last_ms;
  return VBHEM_OK;
}

}  // extern "C"
