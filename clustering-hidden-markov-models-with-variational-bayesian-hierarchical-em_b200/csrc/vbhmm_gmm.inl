// vbhmm_gmm.inl -- batched Gaussian-mixture fits for the random initialisation of VB-HMM learning (included at the
// end of vbhmm_fb.cu).
//
// /root/reference/src/hmm/vbhmm_init.m:43-61 initialises every vbhmm_em run from a K-component full-covariance GMM
// fitted to the run's pooled observations, `gmdistribution.fit(data, K, 'Options', struct('TolFun', 1e-5), 'Start',
// 'randSample')` -- EM from K randomly chosen observations -- and vbhmm_learn.m:465-557 does that once per random
// trial.  With thousands of subjects x tens of trials these fits are the host-side bottleneck of batched learning,
// so they run here: ONE BLOCK PER RUN, the run's observations streamed from the packed X[dim, N, maxT] every EM
// iteration (they are L2-resident: a subject has a few thousand observations).
//   start:  component k = the observation (start_seq, start_t)[k] as mean, the diagonal of the pooled covariance as
//           covariance, weight 1/K                       (MATLAB's 'randSample' start uses the pooled covariance too)
//   E-step: responsibilities by max-shifted log-sum-exp; log-likelihood = sum of the log-sum-exps
//   M-step: weights Nk / n, means, covariances about the means (moments accumulated about the pooled mean, so
//           nothing cancels) + reg on the diagonal
//   stop:   |ll - ll_old| < tol |ll|  (TolFun) or max_iter iterations; a component whose mass vanishes keeps its
//           previous parameters.
// The statistics toolbox's own implementation is proprietary; parity is with oracle/numpy_ref.gmm_em, the same
// algorithm on the host.

namespace {

constexpr int kGmmThreads = 128;

template <int K, int DIM>
struct GmmComp {
  double logw[K], mu[K][DIM], P[K][DIM][DIM] /* inverse covariance */, logdet[K];
};

template <int K, int DIM>
__global__ void __launch_bounds__(kGmmThreads)
k_gmm_fit(int N, int maxT, const int *__restrict__ seq_lo, const int *__restrict__ seq_hi, const int *__restrict__ Tn,
          const double *__restrict__ X, const int *__restrict__ start_seq, const int *__restrict__ start_t, int max_iter,
          double tol, double reg, double *__restrict__ weights, double *__restrict__ means, double *__restrict__ covs,
          double *__restrict__ loglik, int *__restrict__ iters) {
  constexpr int DD = DIM * DIM;
  constexpr int NV = 1 + K * (1 + DIM + DD);          // ll | Nk | Sx | Sxx
  const int r = blockIdx.x;
  const int n0 = seq_lo[r], n1 = seq_hi[r], ns = n1 - n0;
  const size_t ND = (size_t)N * DIM;
  __shared__ double sh[kGmmThreads / 32][NV > 1 + DIM + DD ? NV : 1 + DIM + DD];
  __shared__ double tot[NV > 1 + DIM + DD ? NV : 1 + DIM + DD];
  __shared__ GmmComp<K, DIM> C;
  __shared__ double s_mu[K][DIM], s_cov[K][DIM][DIM], s_w[K], gmean[DIM], gvar[DIM];
  __shared__ int s_stop;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  auto block_sum = [&](const double *acc, int nv) {
    for (int v = 0; v < nv; v++) {
      double a = acc[v];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) sh[warp][v] = a;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < nv; v += kGmmThreads) {
      double a = 0.0;
      for (int q = 0; q < kGmmThreads / 32; q++) a += sh[q][v];
      tot[v] = a;
    }
    __syncthreads();
  };

  // ---- pooled mean and variance (two passes: exact centring)
  double acc[NV > 1 + DIM + DD ? NV : 1 + DIM + DD];
  for (int v = 0; v < 1 + DIM; v++) acc[v] = 0.0;
  for (long long e = threadIdx.x; e < (long long)ns * maxT; e += kGmmThreads) {
    const int n = n0 + (int)(e % ns), t = (int)(e / ns);
    if (t >= Tn[n]) continue;
    acc[0] += 1.0;
    for (int i = 0; i < DIM; i++) acc[1 + i] += X[ND * t + (size_t)DIM * n + i];
  }
  block_sum(acc, 1 + DIM);
  const double npts = tot[0];
  if (threadIdx.x < DIM) gmean[threadIdx.x] = tot[1 + threadIdx.x] / (npts > 0 ? npts : 1.0);
  __syncthreads();
  for (int v = 0; v < DIM; v++) acc[v] = 0.0;
  for (long long e = threadIdx.x; e < (long long)ns * maxT; e += kGmmThreads) {
    const int n = n0 + (int)(e % ns), t = (int)(e / ns);
    if (t >= Tn[n]) continue;
    for (int i = 0; i < DIM; i++) {
      const double dlt = X[ND * t + (size_t)DIM * n + i] - gmean[i];
      acc[i] += dlt * dlt;
    }
  }
  block_sum(acc, DIM);
  if (threadIdx.x < DIM) gvar[threadIdx.x] = tot[threadIdx.x] / (npts > 1 ? npts - 1.0 : 1.0) + reg;
  __syncthreads();

  // ---- start
  if (threadIdx.x < K) {
    const int k = threadIdx.x;
    const int n = start_seq[k + K * r], t = start_t[k + K * r];
    s_w[k] = 1.0 / K;
    for (int i = 0; i < DIM; i++) {
      s_mu[k][i] = X[ND * t + (size_t)DIM * n + i];
      for (int j = 0; j < DIM; j++) s_cov[k][i][j] = i == j ? gvar[i] : 0.0;
    }
  }
  if (threadIdx.x == 0) s_stop = 0;
  __syncthreads();

  double ll_old = 0.0, ll = 0.0;
  int it = 0;
  for (it = 1; it <= max_iter; it++) {
    if (threadIdx.x < K) {   // component constants of this E-step
      const int k = threadIdx.x;
      double A[DIM][DIM];
      for (int i = 0; i < DIM; i++)
        for (int j = 0; j < DIM; j++) A[i][j] = s_cov[k][i][j];
      C.logdet[k] = dev_inverse<DIM>(A);
      C.logw[k] = log(s_w[k]);
      for (int i = 0; i < DIM; i++) {
        C.mu[k][i] = s_mu[k][i];
        for (int j = 0; j < DIM; j++) C.P[k][i][j] = A[i][j];
      }
    }
    __syncthreads();
    for (int v = 0; v < NV; v++) acc[v] = 0.0;
    for (long long e = threadIdx.x; e < (long long)ns * maxT; e += kGmmThreads) {
      const int n = n0 + (int)(e % ns), t = (int)(e / ns);
      if (t >= Tn[n]) continue;
      double x[DIM], lp[K], mx = -INFINITY;
      for (int i = 0; i < DIM; i++) x[i] = X[ND * t + (size_t)DIM * n + i];
#pragma unroll
      for (int k = 0; k < K; k++) {
        double q = 0.0;
        for (int i = 0; i < DIM; i++) {
          double tq = 0.0;
          for (int j = 0; j < DIM; j++) tq += C.P[k][i][j] * (x[j] - C.mu[k][j]);
          q += tq * (x[i] - C.mu[k][i]);
        }
        lp[k] = C.logw[k] - 0.5 * (DIM * 1.8378770664093454836 + C.logdet[k] + q);
        mx = lp[k] > mx ? lp[k] : mx;
      }
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) { lp[k] = exp(lp[k] - mx); s += lp[k]; }
      acc[0] += mx + log(s);
      const double rs = 1.0 / s;
#pragma unroll
      for (int k = 0; k < K; k++) {
        const double g = lp[k] * rs;
        double *a = acc + 1 + k * (1 + DIM + DD);
        a[0] += g;
        for (int i = 0; i < DIM; i++) {
          const double di = x[i] - gmean[i];
          a[1 + i] += g * di;
          for (int j = 0; j < DIM; j++) a[1 + DIM + i * DIM + j] += g * di * (x[j] - gmean[j]);
        }
      }
    }
    block_sum(acc, NV);
    ll = tot[0];
    if (threadIdx.x < K) {   // M-step
      const int k = threadIdx.x;
      const double *a = tot + 1 + k * (1 + DIM + DD);
      const double Nk = a[0];
      if (Nk > 1e-300) {
        s_w[k] = Nk / npts;
        double d[DIM];
        for (int i = 0; i < DIM; i++) { d[i] = a[1 + i] / Nk; s_mu[k][i] = gmean[i] + d[i]; }
        for (int i = 0; i < DIM; i++)
          for (int j = 0; j < DIM; j++)
            s_cov[k][i][j] = a[1 + DIM + i * DIM + j] / Nk - d[i] * d[j] + (i == j ? reg : 0.0);
      }
    }
    if (threadIdx.x == 0) s_stop = it > 1 && fabs(ll - ll_old) < tol * fabs(ll);
    __syncthreads();
    ll_old = ll;
    if (s_stop) break;
  }
  if (threadIdx.x < K) {
    const int k = threadIdx.x;
    weights[k + K * r] = s_w[k];
    for (int i = 0; i < DIM; i++) {
      means[i + DIM * (k + (size_t)K * r)] = s_mu[k][i];
      for (int j = 0; j < DIM; j++) covs[i + DIM * (j + DIM * (k + (size_t)K * r))] = 0.5 * (s_cov[k][i][j] + s_cov[k][j][i]);
    }
  }
  if (threadIdx.x == 0) { loglik[r] = ll; iters[r] = it > max_iter ? max_iter : it; }
}

}  // namespace

extern "C" int vbhmm_gmm_fit(vbhmm_data_t *h, int R, const int32_t *seq_lo, const int32_t *seq_hi, int K,
                             const int32_t *start_seq, const int32_t *start_t, int max_iter, double tol, double reg,
                             double *weights, double *means, double *covs, double *loglik, int32_t *iters) {
  if (!h) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: data handle is NULL");
  if (R < 1 || !seq_lo || !seq_hi || !start_seq || !start_t || !weights || !means || !covs)
    return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: runs, start points and output arrays are required");
  if (K < 1 || K > kFbMaxK || h->dim > kFbMaxDim)
    return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: GMM fits take K <= %d, dim <= %d", kFbMaxK, kFbMaxDim);
  if (max_iter < 1) return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: max_iter must be >= 1");
  const int dim = h->dim;
  for (int r = 0; r < R; r++) {
    if (seq_lo[r] < 0 || seq_hi[r] > h->N || seq_hi[r] <= seq_lo[r])
      return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: run %d covers sequences [%d, %d) of %d", r, seq_lo[r], seq_hi[r], h->N);
    for (int k = 0; k < K; k++) {
      const int n = start_seq[k + K * r], t = start_t[k + K * r];
      if (n < seq_lo[r] || n >= seq_hi[r] || t < 0 || t >= h->maxT)
        return fail_fb(VBHEM_EINVAL, "vbhmm_fb_mex:invalidinput: start point %d of run %d is outside the run", k, r);
    }
  }
  CU(cudaSetDevice(h->device));
  Buf dlo, dhi, dss, dst, dw, dm, dc, dl, di;
  auto release = [&]() { for (Buf *b : {&dlo, &dhi, &dss, &dst, &dw, &dm, &dc, &dl, &di}) b->release(); };
#define CUG(x)                                                                                               \
  do {                                                                                                       \
    cudaError_t e_ = (x);                                                                                    \
    if (e_ != cudaSuccess) {                                                                                 \
      release();                                                                                             \
      return fail_fb(e_ == cudaErrorMemoryAllocation ? VBHEM_ENOMEM : VBHEM_ECUDA, "%s: %s", #x, cudaGetErrorString(e_)); \
    }                                                                                                        \
  } while (0)
  const size_t nK = (size_t)K * R;
  CUG(dlo.ensure(sizeof(int) * R)); CUG(dhi.ensure(sizeof(int) * R));
  CUG(dss.ensure(sizeof(int) * nK)); CUG(dst.ensure(sizeof(int) * nK));
  CUG(dw.ensure(sizeof(double) * nK)); CUG(dm.ensure(sizeof(double) * nK * dim));
  CUG(dc.ensure(sizeof(double) * nK * dim * dim)); CUG(dl.ensure(sizeof(double) * R)); CUG(di.ensure(sizeof(int) * R));
  CUG(cudaMemcpyAsync(dlo.p, seq_lo, sizeof(int) * R, cudaMemcpyHostToDevice, h->st));
  CUG(cudaMemcpyAsync(dhi.p, seq_hi, sizeof(int) * R, cudaMemcpyHostToDevice, h->st));
  CUG(cudaMemcpyAsync(dss.p, start_seq, sizeof(int) * nK, cudaMemcpyHostToDevice, h->st));
  CUG(cudaMemcpyAsync(dst.p, start_t, sizeof(int) * nK, cudaMemcpyHostToDevice, h->st));
  int rc = batch_dispatch(K, dim, [&](auto Kc, auto Dc) {
    k_gmm_fit<decltype(Kc)::value, decltype(Dc)::value><<<R, kGmmThreads, 0, h->st>>>(
        h->N, h->maxT, dlo.as<int>(), dhi.as<int>(), h->Tn.as<int>(), h->X.as<double>(), dss.as<int>(), dst.as<int>(),
        max_iter, tol, reg, dw.as<double>(), dm.as<double>(), dc.as<double>(), dl.as<double>(), di.as<int>());
    return VBHEM_OK;
  });
  if (rc) { release(); return rc; }
  CUG(cudaMemcpyAsync(weights, dw.p, sizeof(double) * nK, cudaMemcpyDeviceToHost, h->st));
  CUG(cudaMemcpyAsync(means, dm.p, sizeof(double) * nK * dim, cudaMemcpyDeviceToHost, h->st));
  CUG(cudaMemcpyAsync(covs, dc.p, sizeof(double) * nK * dim * dim, cudaMemcpyDeviceToHost, h->st));
  if (loglik) CUG(cudaMemcpyAsync(loglik, dl.p, sizeof(double) * R, cudaMemcpyDeviceToHost, h->st));
  if (iters) CUG(cudaMemcpyAsync(iters, di.p, sizeof(int) * R, cudaMemcpyDeviceToHost, h->st));
  CUG(cudaStreamSynchronize(h->st));
  CUG(cudaGetLastError());
#undef CUG
  release();
  return VBHEM_OK;
}
