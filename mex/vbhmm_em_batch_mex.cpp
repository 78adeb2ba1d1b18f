// mex/vbhmm_em_batch_mex.cpp -- MEX gateway of the batched, device-resident VB-HMM learning runs.
//
//   [alpha, epsilon, beta, v, m, W, N1, N, M, LLs, info] = vbhmm_em_batch_mex(data, run_subject, K, prior, m0, W0, ...
//        alpha, epsilon, beta, v, m, W, maxIter, minDiff)
//
// One call runs R independent vbhmm_em loops (/root/reference/src/hmm/vbhmm_em.m:108-418) to convergence on the GPU:
// what /root/reference/src/hmm/vbhmm_learn_batch.m (subjects) and vbhmm_learn.m:465-557 (random initialisations) do
// one after the other.
//   data         cell {S x 1}, data{s} = cell {n_s x 1} of [T x dim] matrices: subject s's sequences (the argument of
//                vbhmm_learn_batch)
//   run_subject  [R x 1], 1-based: the subject run r learns from (several runs may share a subject)
//   K            states of every run;  prior = [alpha0 epsilon0 beta0 v0];  m0 [dim x 1];  W0 [1 x 1] or [dim x 1]
//   alpha [K x R], epsilon [K x K x R] (row = from), beta [K x R], v [K x R], m [dim x K x R], W [dim x dim x K x R]:
//                the runs' initial variational parameters (vbhmm_init.m:189-201 per run, concatenated)
//   maxIter, minDiff (vbhmm_em.m:293-316)
// Outputs: the parameters after learning (same shapes), N1 [K x R], N [K x R], M [K x K x R] (row = from) of the last
// E-step (vbhmm_em.m:469-471), LLs [maxIter x R] (zero beyond a run's passes), info [2 x R] = [passes; unstable].
// From these vbhmm_em.m:426-491 builds hmm.prior / trans / pdf / LL / varpar unchanged.
#include "mex.h"

#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/vbhem_b200.h"

namespace {

void fail_rc(int rc) {
  if (rc != VBHEM_OK) mexErrMsgIdAndTxt("vbhmm_fb_mex:cuda", "%s", vbhem_last_error());
}
const double *dbl_arg(const mxArray *a, size_t n, const char *what) {
  if (!mxIsDouble(a) || mxIsComplex(a) || mxGetNumberOfElements(a) != n)
    mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "%s must be a real double array with %d elements", what, (int)n);
  return mxGetPr(a);
}
double scalar_arg(const mxArray *a) {
  if (!mxIsDouble(a) || mxIsComplex(a) || mxGetNumberOfElements(a) != 1)
    mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "arg must be scalar.");
  return mxGetScalar(a);
}
enum Arg { A_DATA, A_RUNSUB, A_K, A_PRIOR, A_M0, A_W0, A_ALPHA, A_EPS, A_BETA, A_V, A_M, A_W, A_MAXIT, A_MINDIFF, A_COUNT };

mxArray *array3(int a, int b, int c) {
  const mwSize d[3] = {(mwSize)a, (mwSize)b, (mwSize)c};
  return mxCreateNumericArray(3, d, mxDOUBLE_CLASS, mxREAL);
}

}  // namespace

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs != A_COUNT) mexErrMsgIdAndTxt("MyToolbox:arrayProduct:nrhs", "14 inputs required.");
  if (nlhs != 11) mexErrMsgIdAndTxt("MyToolbox:arrayProduct:nlhs", "11 output required.");
  if (!mxIsCell(prhs[A_DATA])) mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "1st arg must be cell");
  const int S = (int)mxGetNumberOfElements(prhs[A_DATA]);
  const int R = (int)mxGetNumberOfElements(prhs[A_RUNSUB]);
  const int K = (int)scalar_arg(prhs[A_K]);
  const int dim = (int)mxGetNumberOfElements(prhs[A_M0]);
  if (S < 1 || R < 1 || K < 1 || dim < 1) mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "empty data, runs, K or m0");
  const double *runsub = dbl_arg(prhs[A_RUNSUB], R, "run_subject"), *prior = dbl_arg(prhs[A_PRIOR], 4, "prior");
  const double *m0 = dbl_arg(prhs[A_M0], dim, "m0");
  const int nW0 = (int)mxGetNumberOfElements(prhs[A_W0]);
  if (nW0 != 1 && nW0 != dim) mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "W0 must have 1 or dim elements");
  const double *W0 = dbl_arg(prhs[A_W0], nW0, "W0");
  const size_t nK = (size_t)K * R;
  const double *alpha = dbl_arg(prhs[A_ALPHA], nK, "alpha"), *eps = dbl_arg(prhs[A_EPS], nK * K, "epsilon");
  const double *beta = dbl_arg(prhs[A_BETA], nK, "beta"), *v = dbl_arg(prhs[A_V], nK, "v");
  const double *m = dbl_arg(prhs[A_M], nK * dim, "m"), *W = dbl_arg(prhs[A_W], nK * dim * dim, "W");
  const int max_iter = (int)scalar_arg(prhs[A_MAXIT]);
  const double min_diff = scalar_arg(prhs[A_MINDIFF]);

  // all subjects' sequences into one packed set; subject s = sequences [start[s], start[s + 1])
  std::vector<int32_t> T, start(S + 1, 0);
  std::vector<long long> offs;
  std::vector<double> flat;
  for (int s = 0; s < S; s++) {
    const mxArray *sub = mxGetCell(prhs[A_DATA], s);
    if (!sub || !mxIsCell(sub)) mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "data{%d} must be a cell of sequences", s + 1);
    const int n = (int)mxGetNumberOfElements(sub);
    for (int q = 0; q < n; q++) {
      const mxArray *x = mxGetCell(sub, q);
      if (!x || !mxIsDouble(x) || mxIsComplex(x) || (mxGetNumberOfElements(x) && (int)mxGetN(x) != dim))
        mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "data{%d}{%d} must be a real [T x %d] matrix", s + 1, q + 1, dim);
      const int t = mxGetNumberOfElements(x) ? (int)mxGetM(x) : 0;
      T.push_back(t);
      offs.push_back((long long)flat.size());
      flat.insert(flat.end(), mxGetPr(x), mxGetPr(x) + (size_t)t * dim);     // [T x dim] column-major, as vbhmm_pack_data takes it
    }
    start[s + 1] = (int32_t)T.size();
  }
  if (T.empty()) mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "no sequences");
  if (flat.empty()) flat.push_back(0.0);
  std::vector<int32_t> lo(R), hi(R);
  for (int r = 0; r < R; r++) {
    const int s = (int)runsub[r];
    if (s < 1 || s > S) mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "run_subject(%d) = %d outside 1..%d", r + 1, s, S);
    lo[r] = start[s - 1];
    hi[r] = start[s];
  }

  vbhmm_data_t *data = nullptr;
  fail_rc(vbhmm_pack_data(0, (int)T.size(), dim, T.data(), offs.data(), flat.data(), &data));
  vbhmm_batch_t *b = nullptr;
  int rc = vbhmm_batch_begin(data, R, lo.data(), hi.data(), K, prior[0], prior[1], prior[2], prior[3], m0, W0, nW0, alpha,
                             eps, beta, v, m, W, max_iter, &b);
  int passes = 0;
  if (rc == VBHEM_OK) rc = vbhmm_batch_run(b, min_diff, &passes);
  if (rc == VBHEM_OK) {
    plhs[0] = mxCreateDoubleMatrix(K, R, mxREAL);
    plhs[1] = array3(K, K, R);
    plhs[2] = mxCreateDoubleMatrix(K, R, mxREAL);
    plhs[3] = mxCreateDoubleMatrix(K, R, mxREAL);
    plhs[4] = array3(dim, K, R);
    const mwSize dw[4] = {(mwSize)dim, (mwSize)dim, (mwSize)K, (mwSize)R};
    plhs[5] = mxCreateNumericArray(4, dw, mxDOUBLE_CLASS, mxREAL);
    plhs[6] = mxCreateDoubleMatrix(K, R, mxREAL);
    plhs[7] = mxCreateDoubleMatrix(K, R, mxREAL);
    plhs[8] = array3(K, K, R);
    plhs[9] = mxCreateDoubleMatrix(max_iter, R, mxREAL);
    plhs[10] = mxCreateDoubleMatrix(2, R, mxREAL);
    std::vector<int32_t> it(R), bad(R);
    rc = vbhmm_batch_fetch(b, mxGetPr(plhs[0]), mxGetPr(plhs[1]), mxGetPr(plhs[2]), mxGetPr(plhs[3]), mxGetPr(plhs[4]),
                           mxGetPr(plhs[5]), mxGetPr(plhs[6]), mxGetPr(plhs[7]), mxGetPr(plhs[8]), mxGetPr(plhs[9]),
                           it.data(), bad.data());
    for (int r = 0; r < R; r++) {
      mxGetPr(plhs[10])[2 * r] = it[r];
      mxGetPr(plhs[10])[2 * r + 1] = bad[r];
    }
  }
  vbhmm_batch_free(b);
  vbhmm_free_data(data);
  fail_rc(rc);
}
