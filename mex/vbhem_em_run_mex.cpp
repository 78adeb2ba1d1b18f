// mex/vbhem_em_run_mex.cpp -- MEX gateway of the device-resident EM run.
//
//   [alpha, eta, epsilon, m, W, v, lambda, LogLs, info] = vbhem_em_run_mex(h3m_b.hmm, maxN, tilde_N_k, ...
//        alpha, eta, epsilon, m, W, v, lambda, Sr, prior, m0, W0, tau, max_iter, minDiff [, ngpus])
//
// ngpus (optional 18th argument): shard the base HMMs over that many GPUs (0 = every visible GPU); every device keeps
// the centre hyper-parameters, only the per-centre statistics travel (one NCCL all-reduce per iteration, inside the
// library: vbhem_multi_em_run).  Without it the run uses one GPU.
//
// One call replaces the whole `while 1` loop of /root/reference/src/vbhem/vbhem_h3m_c_step_fc.m:115-433 for one H3M:
// E-step (MEX call :175-194, responsibilities :271-283, statistics vbhem_compute_Statistics.m), lower bound
// (vbhemh3m_lb.m:64-186), convergence test and NaN guard (:311-354), M-step (vbhem_mstep_component.m:42-70, alpha
// :396) -- all on the device, the centre hyper-parameters resident in HBM between iterations.
// Inputs (plain double arrays, column-major; the MATLAB side builds them with cat / cellfun from h3m_r, see
// INTEGRATION.md 3a):
//   h3m_b.hmm  cell {1 x Kb} as for vbhem_hmm_bwd_fwd_mex;  maxN;  tilde_N_k [Kb x 1] (vbhem_h3m_c_step_fc.m:26-30)
//   alpha [1 x K], eta [S x K], epsilon [S x S x K] (row = from), m [d x S x K], W [d x d x S x K] (full covariance)
//   or [d x S x K] (diagonal; decides the covariance mode), v [S x K], lambda [S x K], Sr [1 x K] (states per centre)
//   prior = [alpha0 eta0 epsilon0 lambda0 v0], m0 [d x 1], W0 [1 x 1] or [d x 1], tau, max_iter, minDiff
// Outputs: the hyper-parameters after the run (same shapes), LogLs [1 x passes], info = [iters unstable].
#include "vbhem_mex_common.hpp"

using namespace vbmex;

static MultiCache g_cache;

static const double *dbl_arg(const mxArray *a, size_t n, const char *what) {
  if (!mxIsDouble(a) || mxIsComplex(a) || mxGetNumberOfElements(a) != n)
    mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "%s must be a real double array with %d elements", what, (int)n);
  return mxGetPr(a);
}

enum Arg { A_HB, A_MAXN, A_NTILDE, A_ALPHA, A_ETA, A_EPS, A_M, A_W, A_V, A_LAM, A_SR, A_PRIOR, A_M0, A_W0, A_TAU,
           A_MAXIT, A_MINDIFF, A_COUNT };

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs != A_COUNT && nrhs != A_COUNT + 1)
    mexErrMsgIdAndTxt("MyToolbox:arrayProduct:nrhs", "17 inputs (+ ngpus) required.");
  const int ngpus = nrhs == A_COUNT + 1 ? (int)scalar_arg(prhs[A_COUNT]) : 1;
  if (nlhs != 9) mexErrMsgIdAndTxt("MyToolbox:arrayProduct:nlhs", "9 output required.");
  if (!mxIsCell(prhs[A_HB])) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "1st arg must be cell");
  const int maxN = (int)scalar_arg(prhs[A_MAXN]);
  const int K = (int)mxGetNumberOfElements(prhs[A_ALPHA]);
  if (K < 1) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "alpha is empty");
  const int S = (int)mxGetM(prhs[A_ETA]);
  if (S < 1 || (int)mxGetNumberOfElements(prhs[A_ETA]) != S * K)
    mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "eta must be [S x K]");
  const int d = (int)mxGetNumberOfElements(prhs[A_M0]);
  if (d < 1) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "m0 is empty");
  const size_t nW = mxGetNumberOfElements(prhs[A_W]);
  int covmode = VBHEM_COV_FULL;
  if (nW == (size_t)d * d * S * K) covmode = VBHEM_COV_FULL;
  else if (nW == (size_t)d * S * K) covmode = VBHEM_COV_DIAG;
  else mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "W must be [d x d x S x K] or [d x S x K]");
  const int dd = covmode ? d * d : d;
  const double *alpha = dbl_arg(prhs[A_ALPHA], K, "alpha"), *eta = dbl_arg(prhs[A_ETA], (size_t)S * K, "eta");
  const double *eps = dbl_arg(prhs[A_EPS], (size_t)S * S * K, "epsilon");
  const double *m = dbl_arg(prhs[A_M], (size_t)d * S * K, "m"), *W = dbl_arg(prhs[A_W], (size_t)dd * S * K, "W");
  const double *v = dbl_arg(prhs[A_V], (size_t)S * K, "v"), *lam = dbl_arg(prhs[A_LAM], (size_t)S * K, "lambda");
  const double *SrD = dbl_arg(prhs[A_SR], K, "Sr"), *pr = dbl_arg(prhs[A_PRIOR], 5, "prior");
  const double *m0 = dbl_arg(prhs[A_M0], d, "m0");
  const int nW0 = (int)mxGetNumberOfElements(prhs[A_W0]);
  if (nW0 != 1 && nW0 != d) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "W0 must have 1 or d elements");
  const double *W0 = dbl_arg(prhs[A_W0], nW0, "W0");
  const int tau = (int)scalar_arg(prhs[A_TAU]), max_iter = (int)scalar_arg(prhs[A_MAXIT]);
  const double min_diff = scalar_arg(prhs[A_MINDIFF]);
  if (max_iter < 0) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "max_iter must be >= 0");
  std::vector<int32_t> Sr(K);
  for (int j = 0; j < K; j++) Sr[j] = (int32_t)SrD[j];

  PackedBase b;
  pack_base(prhs[A_HB], maxN, covmode, b);
  if (b.d != d) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "m0 has %d entries, the base emissions %d dimensions", d, b.d);
  if ((int)mxGetNumberOfElements(prhs[A_NTILDE]) != b.N)
    mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "tilde_N_k must have one entry per base HMM");
  vbhem_multi_t *h = acquire_multi(g_cache, b, dbl_arg(prhs[A_NTILDE], b.N, "tilde_N_k"), ngpus);

  fail_rc(vbhem_multi_em_begin(h, K, Sr.data(), S, nullptr, 1, pr[0], pr[1], pr[2], pr[3], pr[4], m0, W0, nW0, alpha, eta, eps,
                         m, W, v, lam));
  std::vector<double> LogLs((size_t)max_iter + 2, 0.0);
  int iters = 0, unstable = 0;
  fail_rc(vbhem_multi_em_run(h, tau, max_iter, min_diff, LogLs.data(), &iters, &unstable));

  plhs[0] = mxCreateDoubleMatrix(1, K, mxREAL);
  plhs[1] = mxCreateDoubleMatrix(S, K, mxREAL);
  const mwSize de[3] = {(mwSize)S, (mwSize)S, (mwSize)K};
  plhs[2] = mxCreateNumericArray(3, de, mxDOUBLE_CLASS, mxREAL);
  const mwSize dm[3] = {(mwSize)d, (mwSize)S, (mwSize)K};
  plhs[3] = mxCreateNumericArray(3, dm, mxDOUBLE_CLASS, mxREAL);
  if (covmode) {
    const mwSize dw[4] = {(mwSize)d, (mwSize)d, (mwSize)S, (mwSize)K};
    plhs[4] = mxCreateNumericArray(4, dw, mxDOUBLE_CLASS, mxREAL);
  } else {
    plhs[4] = mxCreateNumericArray(3, dm, mxDOUBLE_CLASS, mxREAL);
  }
  plhs[5] = mxCreateDoubleMatrix(S, K, mxREAL);
  plhs[6] = mxCreateDoubleMatrix(S, K, mxREAL);
  fail_rc(vbhem_multi_em_fetch(h, mxGetPr(plhs[0]), mxGetPr(plhs[1]), mxGetPr(plhs[2]), mxGetPr(plhs[3]), mxGetPr(plhs[4]),
                         mxGetPr(plhs[5]), mxGetPr(plhs[6])));
  const int n = iters + (unstable ? 1 : 0);
  plhs[7] = mxCreateDoubleMatrix(1, n, mxREAL);
  std::memcpy(mxGetPr(plhs[7]), LogLs.data(), sizeof(double) * n);
  plhs[8] = mxCreateDoubleMatrix(1, 2, mxREAL);
  mxGetPr(plhs[8])[0] = iters;
  mxGetPr(plhs[8])[1] = unstable;
}
