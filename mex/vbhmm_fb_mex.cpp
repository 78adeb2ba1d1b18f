// mex/vbhmm_fb_mex.cpp -- MEX gateway of the VB-HMM forward-backward pass, served by the B200 library.
//
// Build with MATLAB:   mex -R2018a mex/vbhmm_fb_mex.cpp -Iinclude -L<pkg> -lvbhem_b200
// and put the resulting vbhmm_fb_mex.<mexext> ahead of the toolbox's own on the path.  The toolbox picks it up
// unchanged: /root/reference/src/hmm/vbhmm_fb.m:99-145 tests exist('vbhmm_fb_mex')==3 and calls
//
//   [logrho_Saved, gamma_all, xi_sum, phi_norm] = ...
//       vbhmm_fb_mex(data, K, N, dim, maxT, m, W, v, beta, logLambdaTilde, const_denominator, t_pz1, t_tpztzt1)
//
// Same signature, argument checks, error ids and output shapes as the reference gateway
// (/root/reference/src/hmm/vbhmm_fb_mex.c:185-296; parseMatrix / parseMatrix3 :88-149).  vbhmm_em calls this
// once per EM iteration with the same `data` cell, so the packed sequences are kept in HBM between calls and
// re-used while the cell's data pointers and lengths are unchanged (MATLAB's copy-on-write gives a modified
// array a new data pointer); only the model travels per call.
//
// In this repository the file is compiled against oracle/mxshim/mex.h and driven by the same flat driver as the
// reference build (oracle/mxshim/mxshim.c: shim_vbhmm_fb), so tests/test_fb.py runs the identical call through both.
#include "mex.h"

#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/vbhem_b200.h"

namespace {

double parse_scalar(const mxArray *mx) {
  if (!mxIsDouble(mx) || mxIsComplex(mx)) mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "arg must be scalar.");
  return mxGetScalar(mx);
}

// M x N matrix; pass M = 0 or N = 0 to ignore that dimension (vbhmm_fb_mex.c:88-118)
const double *parse_matrix(const mxArray *mx, int M, int N) {
  if (!mxIsDouble(mx) || mxIsComplex(mx)) mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "arg must be double.");
  if (mxGetNumberOfDimensions(mx) != 2)
    mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "parseMatrix: invalid num dimensions.");
  const size_t nrows = mxGetM(mx), ncols = mxGetN(mx);
  if (((int)nrows != M && M != 0) || ((int)ncols != N && N != 0))
    mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "parseMatrix: invalid size.");
  return mxGetPr(mx);
}

// M x N x D array; MATLAB drops a trailing singleton (vbhmm_fb_mex.c:122-149)
const double *parse_matrix3(const mxArray *mx, int M, int N, int D) {
  if (!mxIsDouble(mx) || mxIsComplex(mx)) mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "arg must be double.");
  const mwSize nd = mxGetNumberOfDimensions(mx);
  if (!((nd == 3 && D > 1) || (nd == 2 && D == 1)))
    mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "parseMatrix3: invalid num dimensions.");
  const mwSize *dims = mxGetDimensions(mx);
  if ((int)dims[0] != M || (int)dims[1] != N || (D > 1 && (int)dims[2] != D))
    mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "parseMatrix3: invalid size.");
  return mxGetPr(mx);
}

struct DataCache {
  vbhmm_data_t *h = nullptr;
  int N = 0, dim = 0;
  uint64_t key = 0;
  ~DataCache() { if (h) vbhmm_free_data(h); }
};
DataCache g_cache;

uint64_t mix(uint64_t k, uint64_t v) { return (k ^ v) * 0x100000001b3ull + 0x9e3779b97f4a7c15ull; }

}  // namespace

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  if (nrhs != 13) mexErrMsgIdAndTxt("MyToolbox:arrayProduct:nrhs", "13 inputs required.");
  if (nlhs != 4) mexErrMsgIdAndTxt("MyToolbox:arrayProduct:nlhs", "One output required.");
  if (!mxIsCell(prhs[0])) mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "1st arg must be cell");
  const int K = (int)parse_scalar(prhs[1]), N = (int)parse_scalar(prhs[2]);
  const int dim = (int)parse_scalar(prhs[3]), maxT = (int)parse_scalar(prhs[4]);
  const double *m = parse_matrix(prhs[5], dim, K);
  const double *W = parse_matrix3(prhs[6], dim, dim, K);
  const double *v = parse_matrix(prhs[7], K, 1);
  const double *beta = parse_matrix(prhs[8], K, 1);
  const double *lLT = parse_matrix(prhs[9], 1, K);
  const double cden = parse_scalar(prhs[10]);
  const double *pz1 = parse_matrix(prhs[11], 1, K);
  const double *A = parse_matrix(prhs[12], K, K);

  // the sequences: data{n} = [T_n x dim]; key = data pointers and lengths
  std::vector<int32_t> T(N);
  std::vector<const double *> ptr(N);
  uint64_t key = mix(mix(0, (uint64_t)N), (uint64_t)dim);
  long long total = 0;
  int seenT = 0;
  for (int n = 0; n < N; n++) {
    const mxArray *c = mxGetCell(prhs[0], n);
    ptr[n] = parse_matrix(c, 0, dim);
    T[n] = (int32_t)mxGetM(c);
    if (T[n] > seenT) seenT = T[n];
    total += T[n];
    key = mix(mix(key, (uint64_t)(uintptr_t)ptr[n]), (uint64_t)T[n]);
  }
  if (seenT > maxT) mexErrMsgIdAndTxt("vbhmm_fb_mex:invalidinput", "a sequence is longer than maxT.");
  if (!g_cache.h || g_cache.key != key || g_cache.N != N || g_cache.dim != dim) {
    if (g_cache.h) { vbhmm_free_data(g_cache.h); g_cache.h = nullptr; }
    std::vector<long long> offs(N);
    std::vector<double> flat((size_t)(total > 0 ? total * dim : 1));
    long long o = 0;
    for (int n = 0; n < N; n++) {
      offs[n] = o;
      std::memcpy(flat.data() + o, ptr[n], sizeof(double) * (size_t)T[n] * dim);
      o += (long long)T[n] * dim;
    }
    if (vbhmm_pack_data(0, N, dim, T.data(), offs.data(), flat.data(), &g_cache.h) != VBHEM_OK)
      mexErrMsgIdAndTxt("vbhmm_fb_mex:cuda", "%s", vbhem_last_error());
    g_cache.key = key; g_cache.N = N; g_cache.dim = dim;
  }

  // outputs [K x N x maxT], [K x N x maxT], [K x K x N], [1 x N]; the library's time extent is max T_n
  const mwSize d3[3] = {(mwSize)K, (mwSize)N, (mwSize)maxT}, dx[3] = {(mwSize)K, (mwSize)K, (mwSize)N};
  plhs[0] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
  plhs[1] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
  plhs[2] = mxCreateNumericArray(3, dx, mxDOUBLE_CLASS, mxREAL);
  plhs[3] = mxCreateDoubleMatrix(1, (mwSize)N, mxREAL);
  // slices t >= max T_n (when the caller's maxT is larger) stay zero, as in the reference
  if (vbhmm_fb(g_cache.h, K, m, W, v, beta, lLT, cden, pz1, A, mxGetPr(plhs[0]), mxGetPr(plhs[1]), mxGetPr(plhs[2]),
               mxGetPr(plhs[3])) != VBHEM_OK)
    mexErrMsgIdAndTxt("vbhmm_fb_mex:cuda", "%s", vbhem_last_error());
}
