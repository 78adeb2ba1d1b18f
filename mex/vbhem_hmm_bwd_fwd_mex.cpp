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
#include "mex.h"

#include <cstring>
#include <vector>

#include "../include/vbhem_b200.h"

namespace {

// The base set is packed once and kept across calls of one EM run: the toolbox passes the same h3m_b.hmm cell on
// every iteration.  A call with a different shape re-packs; same shape re-uploads into the existing handle
// (contents may have changed), which costs the copy but no allocation.
struct Cache {
  vbhem_base_t *h = nullptr;
  int N = 0, maxSb = 0, d = 0, covmode = -1;
} g_cache;

void fail_rc(int rc) {
  if (rc != VBHEM_OK) mexErrMsgIdAndTxt("vbhem_mex:cuda", "%s", vbhem_last_error());
}

double scalar_arg(const mxArray *a) {
  if (!mxIsDouble(a) || mxIsComplex(a)) mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "arg must be scalar.");
  return mxGetScalar(a);
}

const mxArray *field(const mxArray *s, const char *name) {
  const mxArray *f = s ? mxGetField(s, 0, name) : nullptr;
  if (!f) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "missing field '%s'", name);
  return f;
}

}  // namespace

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs != 5 && nrhs != 7) mexErrMsgIdAndTxt("MyToolbox:arrayProduct:nrhs", "5 or 7 inputs required.");
  if (nlhs != 6) mexErrMsgIdAndTxt("MyToolbox:arrayProduct:nlhs", "6 output required.");
  const int covmode = nrhs == 7 ? VBHEM_COV_FULL : VBHEM_COV_DIAG;
  if (!mxIsCell(prhs[0])) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "1st arg must be cell");
  if (!mxIsCell(prhs[1])) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "2nd arg must be cell");
  const mxArray *hb = prhs[0], *hr = prhs[1];
  const int Kb = (int)mxGetNumberOfElements(hb), Kr = (int)mxGetNumberOfElements(hr);
  const int T = (int)scalar_arg(prhs[2]);
  const int maxN = (int)scalar_arg(prhs[3]), maxN2 = (int)scalar_arg(prhs[4]);
  if (Kb < 1 || Kr < 1) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "empty HMM list");

  // ---- base HMMs -> packed column-major arrays (zero padded to maxN states)
  const mxArray *e0 = mxGetCell(field(mxGetCell(hb, 0), "emit"), 0);
  const int d = (int)mxGetScalar(field(e0, "nin"));
  const int dd = covmode ? d * d : d;
  std::vector<int32_t> Sb(Kb);
  std::vector<double> prior((size_t)maxN * Kb, 0.0), A((size_t)maxN * maxN * Kb, 0.0), mu((size_t)d * maxN * Kb, 0.0),
      cov((size_t)dd * maxN * Kb, 0.0);
  for (int i = 0; i < Kb; i++) {
    const mxArray *h = mxGetCell(hb, i);
    const mxArray *mA = field(h, "A");
    const int n = (int)mxGetM(mA);
    if (n < 1 || n > maxN) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "base HMM %d has %d states, maxN=%d", i + 1, n, maxN);
    Sb[i] = n;
    const double *pA = mxGetPr(mA), *pp = mxGetPr(field(h, "prior"));
    for (int c = 0; c < n; c++)
      for (int r = 0; r < n; r++) A[r + (size_t)maxN * (c + (size_t)maxN * i)] = pA[r + n * c];
    for (int k = 0; k < n; k++) prior[k + (size_t)maxN * i] = pp[k];
    const mxArray *em = field(h, "emit");
    for (int k = 0; k < n; k++) {
      const mxArray *e = mxGetCell(em, k);
      const double *pc = mxGetPr(field(e, "centres")), *pv = mxGetPr(field(e, "covars"));
      std::memcpy(&mu[(size_t)d * (k + (size_t)maxN * i)], pc, sizeof(double) * d);
      std::memcpy(&cov[(size_t)dd * (k + (size_t)maxN * i)], pv, sizeof(double) * dd);
    }
  }

  // ---- centre HMMs: the pre-pass fields the reference gateway reads
  std::vector<int32_t> Sr(Kr);
  std::vector<double> logPi((size_t)maxN2 * Kr, 0.0), logA((size_t)maxN2 * maxN2 * Kr, 0.0),
      m((size_t)d * maxN2 * Kr, 0.0), vW((size_t)dd * maxN2 * Kr, 0.0), cterm((size_t)maxN2 * Kr, 0.0);
  for (int j = 0; j < Kr; j++) {
    const mxArray *h = mxGetCell(hr, j);
    const mxArray *mA = field(h, "logATilde");
    const int n = (int)mxGetM(mA);
    if (n < 1 || n > maxN2) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "centre HMM %d has %d states, maxN2=%d", j + 1, n, maxN2);
    Sr[j] = n;
    const double *pA = mxGetPr(mA), *pp = mxGetPr(field(h, "logPiTilde"));
    for (int c = 0; c < n; c++)
      for (int r = 0; r < n; r++) logA[r + (size_t)maxN2 * (c + (size_t)maxN2 * j)] = pA[r + n * c];
    for (int k = 0; k < n; k++) logPi[k + (size_t)maxN2 * j] = pp[k];
    const mxArray *em = field(h, "emit");
    const double *ld = covmode ? mxGetPr(mxGetCell(prhs[5], j)) : nullptr;     // [1 x n]
    const double *ic = covmode ? mxGetPr(mxGetCell(prhs[6], j)) : nullptr;     // [d x d x n] = v .* W
    for (int k = 0; k < n; k++) {
      const mxArray *e = mxGetCell(em, k);
      const size_t s = k + (size_t)maxN2 * j;
      std::memcpy(&m[(size_t)d * s], mxGetPr(field(e, "m")), sizeof(double) * d);
      if (covmode) {
        cterm[s] = ld[k];
        std::memcpy(&vW[(size_t)dd * s], ic + (size_t)dd * k, sizeof(double) * dd);
      } else {
        const double v = mxGetScalar(field(e, "v"));
        const double *W = mxGetPr(field(e, "W"));
        cterm[s] = mxGetScalar(field(e, "logLambdaTildePlusDdivlamda"));
        for (int q = 0; q < d; q++) vW[(size_t)dd * s + q] = v * W[q];
      }
    }
  }

  // ---- device: pack (or re-upload) the base set, run the per-pair E-step
  if (g_cache.h && (g_cache.N != Kb || g_cache.maxSb != maxN || g_cache.d != d || g_cache.covmode != covmode)) {
    vbhem_free_base(g_cache.h);
    g_cache.h = nullptr;
  }
  if (!g_cache.h) {
    fail_rc(vbhem_pack_base(0, Kb, Sb.data(), maxN, d, covmode, prior.data(), A.data(), mu.data(), cov.data(), nullptr,
                            &g_cache.h));
    g_cache.N = Kb; g_cache.maxSb = maxN; g_cache.d = d; g_cache.covmode = covmode;
  } else {
    fail_rc(vbhem_repack_base(g_cache.h, Sb.data(), prior.data(), A.data(), mu.data(), cov.data(), nullptr));
  }
  const size_t P = (size_t)Kb * Kr, r = (size_t)maxN2;
  plhs[0] = mxCreateDoubleMatrix(Kb, Kr, mxREAL);
  std::vector<double> nu1(r * P), xi(r * r * P), pr(r * P), muo(r * d * P), Muo(r * dd * P);
  fail_rc(vbhem_estep_pairs(g_cache.h, Kr, Sr.data(), maxN2, T, logPi.data(), logA.data(), m.data(), vW.data(),
                            cterm.data(), mxGetPr(plhs[0]), nu1.data(), xi.data(), pr.data(), muo.data(), Muo.data()));

  // ---- outputs: five {Kb x Kr} cells of small matrices, shaped as the reference shapes them
  for (int c = 1; c < 6; c++) plhs[c] = mxCreateCellMatrix(Kb, Kr);
  for (int j = 0; j < Kr; j++) {
    const int n = Sr[j];
    for (int i = 0; i < Kb; i++) {
      const size_t p = (size_t)i + (size_t)Kb * j;
      mxArray *a = mxCreateDoubleMatrix(1, n, mxREAL);                 // sum_nu_1 [1 x N2]
      std::memcpy(mxGetPr(a), &nu1[r * p], sizeof(double) * n);
      mxSetCell(plhs[1], p, a);
      a = mxCreateDoubleMatrix(n, 1, mxREAL);                          // update_emit_pr [N2 x 1]
      std::memcpy(mxGetPr(a), &pr[r * p], sizeof(double) * n);
      mxSetCell(plhs[2], p, a);
      a = mxCreateDoubleMatrix(n, d, mxREAL);                          // update_emit_mu [N2 x d]
      for (int k = 0; k < d; k++) std::memcpy(mxGetPr(a) + (size_t)n * k, &muo[r * (k + (size_t)d * p)], sizeof(double) * n);
      mxSetCell(plhs[3], p, a);
      if (covmode) {                                                   // update_emit_Mu [N2 x d x d] | [N2 x d]
        const mwSize dims[3] = {(mwSize)n, (mwSize)d, (mwSize)d};
        a = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
      } else {
        a = mxCreateDoubleMatrix(n, d, mxREAL);
      }
      for (int k = 0; k < dd; k++) std::memcpy(mxGetPr(a) + (size_t)n * k, &Muo[r * (k + (size_t)dd * p)], sizeof(double) * n);
      mxSetCell(plhs[4], p, a);
      a = mxCreateDoubleMatrix(n, n, mxREAL);                          // sum_xi [N2 x N2], row = from
      for (int k = 0; k < n; k++) std::memcpy(mxGetPr(a) + (size_t)n * k, &xi[r * (k + r * p)], sizeof(double) * n);
      mxSetCell(plhs[5], p, a);
    }
  }
}
