// mex/vbhem_estep_fused_mex.cpp -- MEX gateway of the FUSED E-step (Boundary-2).
//
//   [L_elbo, hat_Z, Nj, Nj_rho1, Nj_rho2rho, Nj_rho, y_bar, S_plus_C, lb] = ...
//       vbhem_estep_fused_mex(h3m_b.hmm, h3m_r.hmm, T, maxN, maxN2, tilde_N_k, logOmegaTilde ...
//                             [, logdetCovPlusDdivlamR, invCovR] [, ngpus])
//
// ngpus (optional last argument, scalar): the base HMMs are sharded over that many GPUs of the box and the
// per-centre statistics all-reduced with NCCL inside the library (vbhem_multi_*); 0 = every visible GPU.  Without
// it the call uses one GPU, as in round 1.
//
// One call replaces, per EM iteration of /root/reference/src/vbhem/vbhem_h3m_c_step_fc.m, the MEX call
// (:175-176 / :193-194), the responsibilities (:271-283) and -- for all centres at once -- the statistics
// reduction of /root/reference/src/vbhem/vbhem_compute_Statistics.m:33-82 that vbhem_mstep_component.m:35
// otherwise runs per centre over N cell entries.  Outputs (all plain double arrays, column-major):
//   L_elbo [Kb x Kr], hat_Z [Kb x Kr] (incl. the +1e-50), Nj [1 x Kr] (incl. the +1e-50),
//   Nj_rho1 [N2 x Kr], Nj_rho2rho [N2 x N2 x Kr] (row = from; 1e-12 where a centre has one state),
//   Nj_rho [N2 x Kr] (incl. the +1e-50), y_bar [N2 x d x Kr], S_plus_C [N2 x d x d x Kr] | [N2 x d x Kr],
//   lb [1 x 2] = [sum(Z .* L_elbo), sum(hat_Z .* log hat_Z)]  (Lt1 and Lt7 of vbhemh3m_lb.m:90,107).
// The first five inputs are those of vbhem_hmm_bwd_fwd_mex; tilde_N_k [Kb x 1] and logOmegaTilde [1 x Kr] are
// the variables of the same names in vbhem_h3m_c_step_fc.m (:26-30, :271-273).
#include "vbhem_mex_common.hpp"

using namespace vbmex;

static MultiCache g_cache;

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs < 7 || nrhs > 10) mexErrMsgIdAndTxt("MyToolbox:arrayProduct:nrhs", "7 or 9 inputs (+ ngpus) required.");
  if (nlhs != 9) mexErrMsgIdAndTxt("MyToolbox:arrayProduct:nlhs", "9 output required.");
  const int covmode = nrhs >= 9 ? VBHEM_COV_FULL : VBHEM_COV_DIAG;
  const bool has_ngpus = nrhs == 8 || nrhs == 10;
  const int ngpus = has_ngpus ? (int)scalar_arg(prhs[nrhs - 1]) : 1;
  if (!mxIsCell(prhs[0])) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "1st arg must be cell");
  if (!mxIsCell(prhs[1])) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "2nd arg must be cell");
  const int T = (int)scalar_arg(prhs[2]);
  const int maxN = (int)scalar_arg(prhs[3]), maxN2 = (int)scalar_arg(prhs[4]);
  if (!mxIsDouble(prhs[5]) || !mxIsDouble(prhs[6]))
    mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "tilde_N_k and logOmegaTilde must be double");

  PackedBase b;
  PackedCentres c;
  pack_base(prhs[0], maxN, covmode, b);
  pack_centres(prhs[1], maxN2, b.d, covmode, covmode ? prhs[7] : nullptr, covmode ? prhs[8] : nullptr, c);
  if ((int)mxGetNumberOfElements(prhs[5]) != b.N)
    mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "tilde_N_k must have one entry per base HMM");
  if ((int)mxGetNumberOfElements(prhs[6]) != c.K)
    mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "logOmegaTilde must have one entry per centre HMM");
  vbhem_multi_t *h = acquire_multi(g_cache, b, mxGetPr(prhs[5]), ngpus);

  const int Kb = b.N, Kr = c.K, d = b.d, dd = b.dd, r = maxN2;
  plhs[0] = mxCreateDoubleMatrix(Kb, Kr, mxREAL);
  plhs[1] = mxCreateDoubleMatrix(Kb, Kr, mxREAL);
  std::vector<double> stats(vbhem_stats_size(Kr, r, d, covmode, 1));
  fail_rc(vbhem_multi_estep(h, Kr, c.Sr.data(), r, nullptr, 1, T, c.logPi.data(), c.logA.data(), c.m.data(),
                            c.vW.data(), c.cterm.data(), mxGetPr(prhs[6]), mxGetPr(plhs[0]), mxGetPr(plhs[1]),
                            stats.data()));
  std::vector<double> lbraw(stats.end() - 2, stats.end());
  fail_rc(vbhem_multi_finalize_stats(h, Kr, c.Sr.data(), r, 1, stats.data()));
  size_t off[7];
  vbhem_stats_offsets(Kr, r, d, covmode, 1, off);

  plhs[2] = mxCreateDoubleMatrix(1, Kr, mxREAL);
  plhs[3] = mxCreateDoubleMatrix(r, Kr, mxREAL);
  const mwSize d3[3] = {(mwSize)r, (mwSize)r, (mwSize)Kr};
  plhs[4] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
  plhs[5] = mxCreateDoubleMatrix(r, Kr, mxREAL);
  const mwSize dy[3] = {(mwSize)r, (mwSize)d, (mwSize)Kr};
  plhs[6] = mxCreateNumericArray(3, dy, mxDOUBLE_CLASS, mxREAL);
  if (covmode) {
    const mwSize ds[4] = {(mwSize)r, (mwSize)d, (mwSize)d, (mwSize)Kr};
    plhs[7] = mxCreateNumericArray(4, ds, mxDOUBLE_CLASS, mxREAL);
  } else {
    plhs[7] = mxCreateNumericArray(3, dy, mxDOUBLE_CLASS, mxREAL);
  }
  plhs[8] = mxCreateDoubleMatrix(1, 2, mxREAL);
  const size_t n[6] = {(size_t)Kr, (size_t)r * Kr, (size_t)r * r * Kr, (size_t)r * Kr, (size_t)r * d * Kr,
                       (size_t)r * dd * Kr};
  for (int k = 0; k < 6; k++) std::memcpy(mxGetPr(plhs[2 + k]), stats.data() + off[k], sizeof(double) * n[k]);
  mxGetPr(plhs[8])[0] = lbraw[0];
  mxGetPr(plhs[8])[1] = lbraw[1];
}
