// mex/vbhem_hmm_bwd_fwd_mex.cpp -- MEX gateway: the reference's own plug-in point, served by the B200 library.
//
// Build with MATLAB:   mex -R2018a mex/vbhem_hmm_bwd_fwd_mex.cpp -Iinclude -L<pkg> -lvbhem_b200
// and put the resulting vbhem_hmm_bwd_fwd_mex.<mexext> ahead of the toolbox's own on the path.  The toolbox then
// picks it up unchanged: /root/reference/src/vbhem/vbhem_h3m_c_step_fc.m:56-83 tests
// exist('vbhem_hmm_bwd_fwd_mex')==3 and calls (:175-176 diag, :193-194 full)
//
//   [L_elbo,nu_1,emit_pr,emit_mu,emit_M,sum_xi] = ...
//       vbhem_hmm_bwd_fwd_mex(h3m_b.hmm,h3m_r.hmm,T,maxN,maxN2[,logdetCovPlusDdivlamR,invCovR])
//
// Same signature, argument meaning, output shapes and error ids as the reference gateway
// (/root/reference/src/vbhem/vbhem_hmm_bwd_fwd_mex.c:288-289 gateway, :335-377 checks, :433-473 argument walk,
// :396-409 / :1109-1120 / :1314-1342 outputs).  The arithmetic runs in libvbhem_b200.so (vbhem_estep_pairs).
//
// In this repository the file is compiled against oracle/mxshim/mex.h and driven by the same flat driver as the
// reference build (oracle/mxshim/mxshim.c), so tests/test_gateway.py runs the identical call through both.
#include "vbhem_mex_common.hpp"

using namespace vbmex;

static HandleCache g_cache;

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs != 5 && nrhs != 7) mexErrMsgIdAndTxt("MyToolbox:arrayProduct:nrhs", "5 or 7 inputs required.");
  if (nlhs != 6) mexErrMsgIdAndTxt("MyToolbox:arrayProduct:nlhs", "6 output required.");
  const int covmode = nrhs == 7 ? VBHEM_COV_FULL : VBHEM_COV_DIAG;
  if (!mxIsCell(prhs[0])) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "1st arg must be cell");
  if (!mxIsCell(prhs[1])) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "2nd arg must be cell");
  const int T = (int)scalar_arg(prhs[2]);
  const int maxN = (int)scalar_arg(prhs[3]), maxN2 = (int)scalar_arg(prhs[4]);

  PackedBase b;
  PackedCentres c;
  pack_base(prhs[0], maxN, covmode, b);
  pack_centres(prhs[1], maxN2, b.d, covmode, covmode ? prhs[5] : nullptr, covmode ? prhs[6] : nullptr, c);
  vbhem_base_t *h = acquire(g_cache, b, nullptr);

  const int Kb = b.N, Kr = c.K, d = b.d, dd = b.dd;
  const size_t P = (size_t)Kb * Kr, r = (size_t)maxN2;
  plhs[0] = mxCreateDoubleMatrix(Kb, Kr, mxREAL);
  // NOTE (cost of the literal signature): the per-pair statistics cross the boundary twice -- device -> these flat
  // staging vectors -> one freshly allocated mxArray per pair and output (5 Kb Kr arrays) -- because the reference's
  // outputs are {Kb x Kr} cells of small matrices (vbhem_hmm_bwd_fwd_mex.c:396-409).  That is inherent to Boundary-1
  // (1 KB and five mxArray headers per pair: 6.5 GB / 32 M arrays at N = 100 k); the fused gateway
  // (vbhem_estep_fused_mex.cpp) returns what the EM loop consumes instead and is the fast path.
  std::vector<double> nu1(r * P), xi(r * r * P), pr(r * P), muo(r * d * P), Muo(r * dd * P);
  fail_rc(vbhem_estep_pairs(h, Kr, c.Sr.data(), maxN2, T, c.logPi.data(), c.logA.data(), c.m.data(), c.vW.data(),
                            c.cterm.data(), mxGetPr(plhs[0]), nu1.data(), xi.data(), pr.data(), muo.data(),
                            Muo.data()));

  // outputs: five {Kb x Kr} cells of small matrices, shaped as the reference shapes them
  for (int k = 1; k < 6; k++) plhs[k] = mxCreateCellMatrix(Kb, Kr);
  for (int j = 0; j < Kr; j++) {
    const int n = c.Sr[j];
    for (int i = 0; i < Kb; i++) {
      const size_t p = (size_t)i + (size_t)Kb * j;
      mxArray *a = mxCreateDoubleMatrix(1, n, mxREAL);                 // sum_nu_1 [1 x N2]
      std::memcpy(mxGetPr(a), &nu1[r * p], sizeof(double) * n);
      mxSetCell(plhs[1], p, a);
      a = mxCreateDoubleMatrix(n, 1, mxREAL);                          // update_emit_pr [N2 x 1]
      std::memcpy(mxGetPr(a), &pr[r * p], sizeof(double) * n);
      mxSetCell(plhs[2], p, a);
      a = mxCreateDoubleMatrix(n, d, mxREAL);                          // update_emit_mu [N2 x d]
      for (int k = 0; k < d; k++)
        std::memcpy(mxGetPr(a) + (size_t)n * k, &muo[r * (k + (size_t)d * p)], sizeof(double) * n);
      mxSetCell(plhs[3], p, a);
      if (covmode) {                                                   // update_emit_Mu [N2 x d x d] | [N2 x d]
        const mwSize dims[3] = {(mwSize)n, (mwSize)d, (mwSize)d};
        a = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
      } else {
        a = mxCreateDoubleMatrix(n, d, mxREAL);
      }
      for (int k = 0; k < dd; k++)
        std::memcpy(mxGetPr(a) + (size_t)n * k, &Muo[r * (k + (size_t)dd * p)], sizeof(double) * n);
      mxSetCell(plhs[4], p, a);
      a = mxCreateDoubleMatrix(n, n, mxREAL);                          // sum_xi [N2 x N2], row = from
      for (int k = 0; k < n; k++) std::memcpy(mxGetPr(a) + (size_t)n * k, &xi[r * (k + r * p)], sizeof(double) * n);
      mxSetCell(plhs[5], p, a);
    }
  }
}
