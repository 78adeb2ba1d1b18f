// mex/vbhem_mex_common.hpp -- argument walk shared by the two MEX gateways.
//
// Turns the toolbox's cell/struct trees into the packed column-major arrays of include/vbhem_b200.h, reading
// exactly the fields the reference gateway reads (/root/reference/src/vbhem/vbhem_hmm_bwd_fwd_mex.c:433-473):
//   h3m_b.hmm{i}: A, prior, emit{k}.{centres, covars, nin}
//   h3m_r.hmm{j}: logATilde, logPiTilde, emit{k}.{m, W, v, logLambdaTildePlusDdivlamda}
//   full covariance only: logdetCovPlusDdivlamR{j} [1 x N2], invCovR{j} [d x d x N2] (= v .* W)
#pragma once
#include "mex.h"

#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/vbhem_b200.h"

namespace vbmex {

inline void fail_rc(int rc) {
  if (rc != VBHEM_OK) mexErrMsgIdAndTxt("vbhem_mex:cuda", "%s", vbhem_last_error());
}

inline double scalar_arg(const mxArray *a) {
  if (!mxIsDouble(a) || mxIsComplex(a)) mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "arg must be scalar.");
  return mxGetScalar(a);
}

inline const mxArray *field(const mxArray *s, const char *name) {
  const mxArray *f = s ? mxGetField(s, 0, name) : nullptr;
  if (!f) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "missing field '%s'", name);
  return f;
}

struct PackedBase {
  int N = 0, maxSb = 0, d = 0, covmode = 0, dd = 0;
  std::vector<int32_t> Sb;
  std::vector<double> prior, A, mu, cov;
};

struct PackedCentres {
  int K = 0, maxSr = 0;
  std::vector<int32_t> Sr;
  std::vector<double> logPi, logA, m, vW, cterm;
};

inline void pack_base(const mxArray *hb, int maxN, int covmode, PackedBase &b) {
  b.N = (int)mxGetNumberOfElements(hb);
  if (b.N < 1) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "empty base HMM list");
  const mxArray *e0 = mxGetCell(field(mxGetCell(hb, 0), "emit"), 0);
  b.d = (int)mxGetScalar(field(e0, "nin"));
  b.covmode = covmode;
  b.dd = covmode ? b.d * b.d : b.d;
  b.maxSb = maxN;
  const int d = b.d, dd = b.dd, N = b.N;
  b.Sb.assign(N, 0);
  b.prior.assign((size_t)maxN * N, 0.0);
  b.A.assign((size_t)maxN * maxN * N, 0.0);
  b.mu.assign((size_t)d * maxN * N, 0.0);
  b.cov.assign((size_t)dd * maxN * N, 0.0);
  for (int i = 0; i < N; i++) {
    const mxArray *h = mxGetCell(hb, i);
    const mxArray *mA = field(h, "A");
    const int n = (int)mxGetM(mA);
    if (n < 1 || n > maxN)
      mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "base HMM %d has %d states, maxN=%d", i + 1, n, maxN);
    b.Sb[i] = n;
    const double *pA = mxGetPr(mA), *pp = mxGetPr(field(h, "prior"));
    for (int c = 0; c < n; c++)
      for (int r = 0; r < n; r++) b.A[r + (size_t)maxN * (c + (size_t)maxN * i)] = pA[r + n * c];
    for (int k = 0; k < n; k++) b.prior[k + (size_t)maxN * i] = pp[k];
    const mxArray *em = field(h, "emit");
    for (int k = 0; k < n; k++) {
      const mxArray *e = mxGetCell(em, k);
      std::memcpy(&b.mu[(size_t)d * (k + (size_t)maxN * i)], mxGetPr(field(e, "centres")), sizeof(double) * d);
      std::memcpy(&b.cov[(size_t)dd * (k + (size_t)maxN * i)], mxGetPr(field(e, "covars")), sizeof(double) * dd);
    }
  }
}

// ldR / icR: the two extra full-covariance arguments (nullptr for diagonal).
inline void pack_centres(const mxArray *hr, int maxN2, int d, int covmode, const mxArray *ldR, const mxArray *icR,
                         PackedCentres &c) {
  const int dd = covmode ? d * d : d;
  c.K = (int)mxGetNumberOfElements(hr);
  if (c.K < 1) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "empty centre HMM list");
  c.maxSr = maxN2;
  const int K = c.K;
  c.Sr.assign(K, 0);
  c.logPi.assign((size_t)maxN2 * K, 0.0);
  c.logA.assign((size_t)maxN2 * maxN2 * K, 0.0);
  c.m.assign((size_t)d * maxN2 * K, 0.0);
  c.vW.assign((size_t)dd * maxN2 * K, 0.0);
  c.cterm.assign((size_t)maxN2 * K, 0.0);
  for (int j = 0; j < K; j++) {
    const mxArray *h = mxGetCell(hr, j);
    const mxArray *mA = field(h, "logATilde");
    const int n = (int)mxGetM(mA);
    if (n < 1 || n > maxN2)
      mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "centre HMM %d has %d states, maxN2=%d", j + 1, n, maxN2);
    c.Sr[j] = n;
    const double *pA = mxGetPr(mA), *pp = mxGetPr(field(h, "logPiTilde"));
    for (int cc = 0; cc < n; cc++)
      for (int r = 0; r < n; r++) c.logA[r + (size_t)maxN2 * (cc + (size_t)maxN2 * j)] = pA[r + n * cc];
    for (int k = 0; k < n; k++) c.logPi[k + (size_t)maxN2 * j] = pp[k];
    const mxArray *em = field(h, "emit");
    const double *ld = covmode ? mxGetPr(mxGetCell(ldR, j)) : nullptr;     // [1 x n]
    const double *ic = covmode ? mxGetPr(mxGetCell(icR, j)) : nullptr;     // [d x d x n] = v .* W
    for (int k = 0; k < n; k++) {
      const mxArray *e = mxGetCell(em, k);
      const size_t s = k + (size_t)maxN2 * j;
      std::memcpy(&c.m[(size_t)d * s], mxGetPr(field(e, "m")), sizeof(double) * d);
      if (covmode) {
        c.cterm[s] = ld[k];
        std::memcpy(&c.vW[(size_t)dd * s], ic + (size_t)dd * k, sizeof(double) * dd);
      } else {
        const double v = mxGetScalar(field(e, "v"));
        const double *W = mxGetPr(field(e, "W"));
        c.cterm[s] = mxGetScalar(field(e, "logLambdaTildePlusDdivlamda"));
        for (int q = 0; q < d; q++) c.vW[(size_t)dd * s + q] = v * W[q];
      }
    }
  }
}

// The base set is packed once and kept across the calls of an EM run: the toolbox passes the same h3m_b.hmm
// on every iteration.  Same shape => re-upload into the existing handle (no allocation); new shape => re-pack.
struct HandleCache {
  vbhem_base_t *h = nullptr;
  int N = 0, maxSb = 0, d = 0, covmode = -1;
};

inline vbhem_base_t *acquire(HandleCache &c, const PackedBase &b, const double *Ntilde) {
  if (c.h && (c.N != b.N || c.maxSb != b.maxSb || c.d != b.d || c.covmode != b.covmode)) {
    vbhem_free_base(c.h);
    c.h = nullptr;
  }
  if (!c.h) {
    fail_rc(vbhem_pack_base(0, b.N, b.Sb.data(), b.maxSb, b.d, b.covmode, b.prior.data(), b.A.data(), b.mu.data(),
                            b.cov.data(), Ntilde, &c.h));
    c.N = b.N; c.maxSb = b.maxSb; c.d = b.d; c.covmode = b.covmode;
  } else {
    fail_rc(vbhem_repack_base(c.h, b.Sb.data(), b.prior.data(), b.A.data(), b.mu.data(), b.cov.data(), Ntilde));
  }
  return c.h;
}

// The same cache for a base set sharded over `ngpus` devices (vbhem_multi_*): ngpus <= 0 = every visible device
// (never more than there are base HMMs).  One MATLAB session thereby drives all GPUs of the box -- what the
// reference gets from `parfor` inside one session (/root/reference/src/vbhem/vbhem_h3m_c.m:28).
struct MultiCache {
  vbhem_multi_t *h = nullptr;
  int N = 0, maxSb = 0, d = 0, covmode = -1, ngpus = 0;
};

inline int resolve_ngpus(int want, int N) {
  int have = 0;
  fail_rc(vbhem_device_count(&have));
  int n = want > 0 ? want : have;
  if (n > have) mexErrMsgIdAndTxt("vbhem_mex:invalidinput", "ngpus = %d, but %d CUDA devices are visible", want, have);
  if (n > N) n = N;
  return n < 1 ? 1 : n;
}

inline vbhem_multi_t *acquire_multi(MultiCache &c, const PackedBase &b, const double *Ntilde, int want_gpus) {
  const int n = resolve_ngpus(want_gpus, b.N);
  if (c.h && (c.N != b.N || c.maxSb != b.maxSb || c.d != b.d || c.covmode != b.covmode || c.ngpus != n)) {
    vbhem_multi_free(c.h);
    c.h = nullptr;
  }
  if (!c.h) {
    fail_rc(vbhem_multi_pack_base(n, nullptr, n, 0, nullptr, b.N, b.Sb.data(), b.maxSb, b.d, b.covmode, b.prior.data(),
                                  b.A.data(), b.mu.data(), b.cov.data(), Ntilde, &c.h));
    c.N = b.N; c.maxSb = b.maxSb; c.d = b.d; c.covmode = b.covmode; c.ngpus = n;
  } else {
    fail_rc(vbhem_multi_repack_base(c.h, b.Sb.data(), b.prior.data(), b.A.data(), b.mu.data(), b.cov.data(), Ntilde));
  }
  return c.h;
}

}  // namespace vbmex
