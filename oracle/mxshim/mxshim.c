/* oracle/mxshim/mxshim.c -- TEST INFRASTRUCTURE, not product code.
 *
 * (1) An implementation of the mx / mex symbols declared in ./mex.h, backed by
 *     a per-thread bump arena (MATLAB's own allocator is proprietary; the MEX
 *     contract only needs create/get/set semantics).
 * (2) A flat C driver, shim_estep_pairs(), that builds the cell/struct trees
 *     the MEX gateway signature expects
 *       [L_elbo,nu_1,emit_pr,emit_mu,emit_M,sum_xi] =
 *           vbhem_hmm_bwd_fwd_mex(h3m_b.hmm,h3m_r.hmm,T,maxN,maxN2[,logdetCovPlusDdivlamR,invCovR])
 *     (call sites: /root/reference/src/vbhem/vbhem_h3m_c_step_fc.m:175-176,193-194;
 *      argument walk: /root/reference/src/vbhem/vbhem_hmm_bwd_fwd_mex.c:433-473)
 *     from packed column-major arrays, calls whatever mexFunction this file is
 *     linked with, and flattens the six outputs back into packed arrays.
 *
 * Linked with the reference C file it yields oracle/_ref/libvbhem_ref.so (the
 * oracle / CPU baseline).  Linked with mex/vbhem_hmm_bwd_fwd_mex.cpp it
 * exercises this repo's own MEX gateway through the very same driver.
 */
#include "mex.h"

#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ arena */
typedef struct chunk { struct chunk *next; size_t cap, used; } chunk;
static __thread chunk *g_chunks = NULL;

static void *arena_alloc(size_t n) {
  n = (n + 15u) & ~(size_t)15u;
  if (!g_chunks || g_chunks->used + n > g_chunks->cap) {
    size_t cap = (size_t)1 << 24;
    if (cap < n) cap = n;
    chunk *c = (chunk *)malloc(sizeof(chunk) + 16 + cap);
    if (!c) { fprintf(stderr, "mxshim: out of memory\n"); abort(); }
    c->next = g_chunks; c->cap = cap; c->used = 0;
    g_chunks = c;
  }
  char *base = (char *)(g_chunks + 1);
  base = (char *)(((size_t)base + 15u) & ~(size_t)15u);
  void *p = base + g_chunks->used;
  g_chunks->used += n;
  return p;
}
static void arena_reset(void) {
  while (g_chunks) { chunk *n = g_chunks->next; free(g_chunks); g_chunks = n; }
}

/* ---------------------------------------------------------------- mxArray */
#define SHIM_MAXFIELDS 16
struct mxArray_tag {
  mxClassID cls;
  mwSize ndim;
  mwSize dims[4];
  size_t nel;
  double *pr;             /* double data */
  mxArray **cells;        /* cell / struct-field payload */
  int nfields;
  const char *fnames[SHIM_MAXFIELDS];
};

static mxArray *new_array(mxClassID cls, mwSize ndim, const mwSize *dims) {
  mxArray *a = (mxArray *)arena_alloc(sizeof(mxArray));
  memset(a, 0, sizeof(*a));
  a->cls = cls;
  a->ndim = ndim < 2 ? 2 : ndim;
  a->nel = 1;
  for (mwSize i = 0; i < 4; i++) a->dims[i] = 1;
  for (mwSize i = 0; i < ndim && i < 4; i++) { a->dims[i] = dims[i]; a->nel *= dims[i]; }
  return a;
}

int mxIsCell(const mxArray *a) { return a && a->cls == mxCELL_CLASS; }
int mxIsDouble(const mxArray *a) { return a && a->cls == mxDOUBLE_CLASS; }
int mxIsComplex(const mxArray *a) { (void)a; return 0; }
size_t mxGetM(const mxArray *a) { return a->dims[0]; }
size_t mxGetN(const mxArray *a) {          /* MATLAB: product of dims 2..end */
  size_t n = 1;
  for (mwSize i = 1; i < a->ndim; i++) n *= a->dims[i];
  return n;
}
mwSize mxGetNumberOfDimensions(const mxArray *a) { return a->ndim; }
const mwSize *mxGetDimensions(const mxArray *a) { return a->dims; }
size_t mxGetNumberOfElements(const mxArray *a) { return a->nel; }
double *mxGetPr(const mxArray *a) { return a->pr; }
double mxGetScalar(const mxArray *a) { return a->pr[0]; }
mxArray *mxGetCell(const mxArray *a, mwIndex i) { return a->cells[i]; }
mxArray *mxGetField(const mxArray *a, mwIndex i, const char *name) {
  (void)i;
  for (int f = 0; f < a->nfields; f++)
    if (strcmp(a->fnames[f], name) == 0) return a->cells[f];
  return NULL;
}

mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c) {
  (void)c;
  mxArray *a = new_array(cls, ndim, dims);
  a->pr = (double *)arena_alloc(sizeof(double) * (a->nel ? a->nel : 1));
  memset(a->pr, 0, sizeof(double) * a->nel);
  return a;
}
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) {
  mwSize dims[2]; dims[0] = m; dims[1] = n;
  return mxCreateNumericArray(2, dims, mxDOUBLE_CLASS, c);
}
mxArray *mxCreateCellMatrix(mwSize m, mwSize n) {
  mwSize dims[2]; dims[0] = m; dims[1] = n;
  mxArray *a = new_array(mxCELL_CLASS, 2, dims);
  a->cells = (mxArray **)arena_alloc(sizeof(mxArray *) * (a->nel ? a->nel : 1));
  memset(a->cells, 0, sizeof(mxArray *) * a->nel);
  return a;
}
void mxSetCell(mxArray *a, mwIndex i, mxArray *v) { a->cells[i] = v; }

void *mxMalloc(size_t n) { return arena_alloc(n ? n : 1); }
void *mxCalloc(size_t n, size_t sz) { void *p = arena_alloc(n * sz ? n * sz : 1); memset(p, 0, n * sz); return p; }
void mxFree(void *p) { (void)p; }

/* ------------------------------------------------------------------ errors */
static __thread jmp_buf g_jmp;
static __thread int g_jmp_armed = 0;
static __thread char g_errid[128];
static __thread char g_errmsg[512];

void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) {
  va_list ap;
  snprintf(g_errid, sizeof g_errid, "%s", id ? id : "");
  va_start(ap, fmt);
  vsnprintf(g_errmsg, sizeof g_errmsg, fmt, ap);
  va_end(ap);
  if (g_jmp_armed) longjmp(g_jmp, 1);
  fprintf(stderr, "mexErrMsgIdAndTxt(%s): %s\n", g_errid, g_errmsg);
  abort();
}
int mexPrintf(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt);
  int r = vprintf(fmt, ap);
  va_end(ap);
  return r;
}
const char *shim_last_error_id(void) { return g_errid; }
const char *shim_last_error_msg(void) { return g_errmsg; }

/* ------------------------------------------------------------ tree builders */
static mxArray *new_struct(void) {
  mwSize dims[2] = {1, 1};
  mxArray *a = new_array(mxSTRUCT_CLASS, 2, dims);
  a->cells = (mxArray **)arena_alloc(sizeof(mxArray *) * SHIM_MAXFIELDS);
  return a;
}
static void set_field(mxArray *s, const char *name, mxArray *v) {
  s->fnames[s->nfields] = name;
  s->cells[s->nfields] = v;
  s->nfields++;
}
static mxArray *dbl_matrix(mwSize m, mwSize n, const double *src, size_t ld, int transpose_unused) {
  (void)transpose_unused;
  mxArray *a = mxCreateDoubleMatrix(m, n, mxREAL);
  for (mwSize j = 0; j < n; j++)
    for (mwSize i = 0; i < m; i++) a->pr[i + j * m] = src[i + j * ld];
  return a;
}
static mxArray *dbl_scalar(double v) {
  mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL);
  a->pr[0] = v;
  return a;
}

/* ---- cell/struct trees of the two HMM lists ---------------------------------
 * Packed inputs (all column-major, fp64):
 *   base set  : Sb[Kb]; prior[maxSb,Kb]; A[maxSb,maxSb,Kb] (A[from,to]);
 *               mu[d,maxSb,Kb]; cov[d,d,maxSb,Kb] (full) or [d,maxSb,Kb] (diag)
 *   centres   : Sr[Kr]; logPi[maxSr,Kr]; logA[maxSr,maxSr,Kr] (row = from);
 *               m[d,maxSr,Kr]; W[d,d,maxSr,Kr] or [d,maxSr,Kr]; v[maxSr,Kr];
 *               cterm[maxSr,Kr]  (= -logLambdaTilde + d/lambda)
 *   covmode   : 0 diag, 1 full. */

/* h3m_b.hmm : {1 x Kb} of struct(A, prior, emit{1 x Sb} of struct(centres, covars, nin)) */
static mxArray *make_base_tree(int Kb, const int *Sb, int maxSb, int d, int covmode, const double *prior,
                               const double *A, const double *mu, const double *cov) {
  const int dd = covmode ? d * d : d;
  mxArray *hb = mxCreateCellMatrix(1, (mwSize)Kb);
  for (int i = 0; i < Kb; i++) {
    const int S = Sb[i];
    mxArray *h = new_struct();
    set_field(h, "prior", dbl_matrix(S, 1, prior + (size_t)maxSb * i, maxSb, 0));
    set_field(h, "A", dbl_matrix(S, S, A + (size_t)maxSb * maxSb * i, maxSb, 0));
    mxArray *em = mxCreateCellMatrix(1, (mwSize)S);
    for (int k = 0; k < S; k++) {
      mxArray *e = new_struct();
      set_field(e, "centres", dbl_matrix(1, d, mu + (size_t)d * (k + (size_t)maxSb * i), 1, 0));
      if (covmode)
        set_field(e, "covars", dbl_matrix(d, d, cov + (size_t)dd * (k + (size_t)maxSb * i), d, 0));
      else
        set_field(e, "covars", dbl_matrix(1, d, cov + (size_t)dd * (k + (size_t)maxSb * i), 1, 0));
      set_field(e, "nin", dbl_scalar((double)d));
      mxSetCell(em, k, e);
    }
    set_field(h, "emit", em);
    mxSetCell(hb, i, h);
  }
  return hb;
}

/* h3m_r.hmm : {1 x Kr} of struct(logATilde, logPiTilde, emit{1 x Sr} of struct(m, W, v,
 * logLambdaTildePlusDdivlamda)); *c5 = logdetCovPlusDdivlamR {1 x Kr}, *c6 = invCovR {1 x Kr}
 * (invCovR{j}(:,:,n) = v .* W, vbhem_h3m_c_step_fc.m:189). */
static mxArray *make_centre_tree(int Kr, const int *Sr, int maxSr, int d, int covmode, const double *logPi,
                                 const double *logA, const double *m, const double *W, const double *v,
                                 const double *cterm, mxArray **c5out, mxArray **c6out) {
  const int dd = covmode ? d * d : d;
  mxArray *hr = mxCreateCellMatrix(1, (mwSize)Kr);
  mxArray *c5 = mxCreateCellMatrix(1, (mwSize)Kr);
  mxArray *c6 = mxCreateCellMatrix(1, (mwSize)Kr);
  for (int j = 0; j < Kr; j++) {
    const int S = Sr[j];
    mxArray *h = new_struct();
    set_field(h, "logATilde", dbl_matrix(S, S, logA + (size_t)maxSr * maxSr * j, maxSr, 0));
    set_field(h, "logPiTilde", dbl_matrix(S, 1, logPi + (size_t)maxSr * j, maxSr, 0));
    mxArray *em = mxCreateCellMatrix(1, (mwSize)S);
    mwSize d3[3]; d3[0] = d; d3[1] = d; d3[2] = S;
    mxArray *ld = mxCreateDoubleMatrix(1, S, mxREAL);
    mxArray *ic = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);
    for (int k = 0; k < S; k++) {
      const size_t s = k + (size_t)maxSr * j;
      mxArray *e = new_struct();
      set_field(e, "m", dbl_matrix(1, d, m + (size_t)d * s, 1, 0));
      if (covmode) set_field(e, "W", dbl_matrix(d, d, W + (size_t)dd * s, d, 0));
      else         set_field(e, "W", dbl_matrix(1, d, W + (size_t)dd * s, 1, 0));
      set_field(e, "v", dbl_scalar(v[s]));
      set_field(e, "logLambdaTildePlusDdivlamda", dbl_scalar(cterm[s]));
      mxSetCell(em, k, e);
      ld->pr[k] = cterm[s];
      if (covmode)
        for (int q = 0; q < dd; q++) ic->pr[q + (size_t)dd * k] = v[s] * W[(size_t)dd * s + q];
    }
    set_field(h, "emit", em);
    mxSetCell(hr, j, h);
    mxSetCell(c5, j, ld);
    mxSetCell(c6, j, ic);
  }
  *c5out = c5;
  *c6out = c6;
  return hr;
}

/* shim_estep_pairs: the six-output signature of vbhem_hmm_bwd_fwd_mex.
 * Packed outputs (pair index p = i + Kb*j):
 *   L_elbo[Kb,Kr]; nu1[maxSr,p]; xi[maxSr,maxSr,p]; pr[maxSr,p];
 *   mu_out[maxSr,d,p]; Mu_out[maxSr,d,d,p] (full) or [maxSr,d,p] (diag).
 *   Entries beyond a pair's own (Sr_j) extent are left untouched.
 * Returns 0, or 1 if the gateway raised mexErrMsgIdAndTxt (see shim_last_error_*). */
int shim_estep_pairs(int Kb, const int *Sb, int maxSb, int d, int covmode,
                     const double *prior, const double *A, const double *mu, const double *cov,
                     int Kr, const int *Sr, int maxSr,
                     const double *logPi, const double *logA, const double *m, const double *W,
                     const double *v, const double *cterm, int T,
                     double *L_elbo, double *nu1, double *xi, double *pr, double *mu_out, double *Mu_out) {
  const int dd = covmode ? d * d : d;
  g_jmp_armed = 1;
  if (setjmp(g_jmp)) { g_jmp_armed = 0; arena_reset(); return 1; }

  mxArray *c5, *c6;
  mxArray *hb = make_base_tree(Kb, Sb, maxSb, d, covmode, prior, A, mu, cov);
  mxArray *hr = make_centre_tree(Kr, Sr, maxSr, d, covmode, logPi, logA, m, W, v, cterm, &c5, &c6);
  const mxArray *prhs[7];
  mxArray *plhs[6] = {0, 0, 0, 0, 0, 0};
  prhs[0] = hb; prhs[1] = hr;
  prhs[2] = dbl_scalar((double)T);
  prhs[3] = dbl_scalar((double)maxSb);
  prhs[4] = dbl_scalar((double)maxSr);
  prhs[5] = c5; prhs[6] = c6;
  mexFunction(6, plhs, covmode ? 7 : 5, prhs);

  /* flatten outputs */
  memcpy(L_elbo, mxGetPr(plhs[0]), sizeof(double) * (size_t)Kb * Kr);
  for (int j = 0; j < Kr; j++) {
    const int S = Sr[j];
    for (int i = 0; i < Kb; i++) {
      const size_t p = (size_t)i + (size_t)Kb * j;
      const double *s;
      s = mxGetPr(mxGetCell(plhs[1], p));
      for (int a = 0; a < S; a++) nu1[a + (size_t)maxSr * p] = s[a];
      s = mxGetPr(mxGetCell(plhs[2], p));
      for (int a = 0; a < S; a++) pr[a + (size_t)maxSr * p] = s[a];
      s = mxGetPr(mxGetCell(plhs[3], p));
      for (int k = 0; k < d; k++)
        for (int a = 0; a < S; a++) mu_out[a + (size_t)maxSr * (k + (size_t)d * p)] = s[a + S * k];
      s = mxGetPr(mxGetCell(plhs[4], p));
      for (int k = 0; k < dd; k++)
        for (int a = 0; a < S; a++) Mu_out[a + (size_t)maxSr * (k + (size_t)dd * p)] = s[a + S * k];
      s = mxGetPr(mxGetCell(plhs[5], p));
      for (int b = 0; b < S; b++)
        for (int a = 0; a < S; a++) xi[a + (size_t)maxSr * (b + (size_t)maxSr * p)] = s[a + S * b];
    }
  }
  g_jmp_armed = 0;
  arena_reset();
  return 0;
}

/* shim_estep_fused: the nine-output signature of this repo's vbhem_estep_fused_mex
 * (mex/vbhem_estep_fused_mex.cpp).  Extra inputs Ntilde[Kb], logOmega[Kr]; outputs are copied
 * verbatim: L_elbo[Kb,Kr], zhat[Kb,Kr], Nj[Kr], Nj_rho1[maxSr,Kr], Nj_rho2rho[maxSr,maxSr,Kr],
 * Nj_rho[maxSr,Kr], y_bar[maxSr,d,Kr], S_plus_C[maxSr,dd,Kr], lb[2]. */
int shim_estep_fused(int Kb, const int *Sb, int maxSb, int d, int covmode,
                     const double *prior, const double *A, const double *mu, const double *cov,
                     const double *Ntilde, int Kr, const int *Sr, int maxSr,
                     const double *logPi, const double *logA, const double *m, const double *W,
                     const double *v, const double *cterm, const double *logOmega, int T,
                     double *L_elbo, double *zhat, double *Nj, double *Nj_rho1, double *Nj_rho2rho,
                     double *Nj_rho, double *y_bar, double *S_plus_C, double *lb) {
  g_jmp_armed = 1;
  if (setjmp(g_jmp)) { g_jmp_armed = 0; arena_reset(); return 1; }
  mxArray *c5, *c6;
  mxArray *hb = make_base_tree(Kb, Sb, maxSb, d, covmode, prior, A, mu, cov);
  mxArray *hr = make_centre_tree(Kr, Sr, maxSr, d, covmode, logPi, logA, m, W, v, cterm, &c5, &c6);
  const mxArray *prhs[10];
  mxArray *plhs[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  prhs[0] = hb; prhs[1] = hr;
  prhs[2] = dbl_scalar((double)T);
  prhs[3] = dbl_scalar((double)maxSb);
  prhs[4] = dbl_scalar((double)maxSr);
  prhs[5] = dbl_matrix(Kb, 1, Ntilde, Kb, 0);
  prhs[6] = dbl_matrix(1, Kr, logOmega, 1, 0);
  prhs[7] = c5; prhs[8] = c6;
  int nrhs = covmode ? 9 : 7;
  const char *ng = getenv("VBHEM_SHIM_NGPUS");     /* tests: the gateway's optional trailing `ngpus` argument */
  if (ng && *ng) prhs[nrhs++] = dbl_scalar(atof(ng));
  mexFunction(9, plhs, nrhs, prhs);
  double *outs[9] = {L_elbo, zhat, Nj, Nj_rho1, Nj_rho2rho, Nj_rho, y_bar, S_plus_C, lb};
  for (int k = 0; k < 9; k++)
    memcpy(outs[k], mxGetPr(plhs[k]), sizeof(double) * mxGetNumberOfElements(plhs[k]));
  g_jmp_armed = 0;
  arena_reset();
  return 0;
}

/* shim_em_run: the signature of this repo's vbhem_em_run_mex (mex/vbhem_em_run_mex.cpp): base tree + flat
 * hyper-parameter arrays in, the hyper-parameters after the EM run, LogLs[max_iter + 2] and info[2] out. */
int shim_em_run(int Kb, const int *Sb, int maxSb, int d, int covmode,
                const double *prior, const double *A, const double *mu, const double *cov, const double *Ntilde,
                int Kr, const int *Sr, int maxSr, double *alpha, double *eta, double *eps, double *m, double *W,
                double *v, double *lam, const double *hyp0 /* alpha0 eta0 epsilon0 lambda0 v0 */, const double *m0,
                const double *W0, int nW0, int T, int max_iter, double min_diff, double *LogLs, double *info) {
  const int dd = covmode ? d * d : d;
  g_jmp_armed = 1;
  if (setjmp(g_jmp)) { g_jmp_armed = 0; arena_reset(); return 1; }
  mxArray *hb = make_base_tree(Kb, Sb, maxSb, d, covmode, prior, A, mu, cov);
  mxArray *sr = mxCreateDoubleMatrix(1, (mwSize)Kr, mxREAL);
  for (int j = 0; j < Kr; j++) sr->pr[j] = Sr[j];
  const mxArray *prhs[18];
  mxArray *plhs[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  prhs[0] = hb;
  prhs[1] = dbl_scalar((double)maxSb);
  prhs[2] = dbl_matrix(Kb, 1, Ntilde, Kb, 0);
  prhs[3] = dbl_matrix(1, Kr, alpha, 1, 0);
  prhs[4] = dbl_matrix(maxSr, Kr, eta, maxSr, 0);
  prhs[5] = dbl_matrix((mwSize)maxSr * maxSr, Kr, eps, (size_t)maxSr * maxSr, 0);
  prhs[6] = dbl_matrix((mwSize)d * maxSr, Kr, m, (size_t)d * maxSr, 0);
  prhs[7] = dbl_matrix((mwSize)dd * maxSr, Kr, W, (size_t)dd * maxSr, 0);
  prhs[8] = dbl_matrix(maxSr, Kr, v, maxSr, 0);
  prhs[9] = dbl_matrix(maxSr, Kr, lam, maxSr, 0);
  prhs[10] = sr;
  prhs[11] = dbl_matrix(1, 5, hyp0, 1, 0);
  prhs[12] = dbl_matrix(d, 1, m0, d, 0);
  prhs[13] = dbl_matrix(nW0, 1, W0, nW0, 0);
  prhs[14] = dbl_scalar((double)T);
  prhs[15] = dbl_scalar((double)max_iter);
  prhs[16] = dbl_scalar(min_diff);
  int nrhs = 17;
  const char *ng = getenv("VBHEM_SHIM_NGPUS");
  if (ng && *ng) prhs[nrhs++] = dbl_scalar(atof(ng));
  mexFunction(9, plhs, nrhs, prhs);
  double *outs[7] = {alpha, eta, eps, m, W, v, lam};
  for (int k = 0; k < 7; k++) memcpy(outs[k], mxGetPr(plhs[k]), sizeof(double) * mxGetNumberOfElements(plhs[k]));
  const size_t n = mxGetNumberOfElements(plhs[7]);
  memcpy(LogLs, mxGetPr(plhs[7]), sizeof(double) * n);
  info[0] = mxGetPr(plhs[8])[0];
  info[1] = mxGetPr(plhs[8])[1];
  info[2] = (double)n;
  g_jmp_armed = 0;
  arena_reset();
  return 0;
}

/* shim_em_batch: the signature of this repo's vbhmm_em_batch_mex (mex/vbhmm_em_batch_mex.cpp).  The S subjects'
 * sequences arrive concatenated (sequence n = [T_n x dim] column-major at offs[n]; subject s = sequences
 * [sub_start[s], sub_start[s + 1])) and are handed over as the nested cell the toolbox uses. */
int shim_em_batch(int S, const int *sub_start, int dim, const int *T, const long long *offs, const double *seq, int R,
                  const double *run_subject, int K, const double *prior4, const double *m0, const double *W0, int nW0,
                  double *alpha, double *eps, double *beta, double *v, double *m, double *W, int max_iter,
                  double min_diff, double *N1, double *Nk, double *M, double *LLs, double *info) {
  g_jmp_armed = 1;
  if (setjmp(g_jmp)) { g_jmp_armed = 0; arena_reset(); return 1; }
  mxArray *data = mxCreateCellMatrix((mwSize)S, 1);
  for (int s = 0; s < S; s++) {
    const int n0 = sub_start[s], n1 = sub_start[s + 1];
    mxArray *sub = mxCreateCellMatrix((mwSize)(n1 - n0), 1);
    for (int n = n0; n < n1; n++)
      mxSetCell(sub, (mwIndex)(n - n0), dbl_matrix((mwSize)T[n], (mwSize)dim, seq + offs[n], (size_t)T[n], 0));
    mxSetCell(data, (mwIndex)s, sub);
  }
  const mxArray *prhs[14];
  mxArray *plhs[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  prhs[0] = data;
  prhs[1] = dbl_matrix(R, 1, run_subject, R, 0);
  prhs[2] = dbl_scalar((double)K);
  prhs[3] = dbl_matrix(1, 4, prior4, 1, 0);
  prhs[4] = dbl_matrix(dim, 1, m0, dim, 0);
  prhs[5] = dbl_matrix(nW0, 1, W0, nW0, 0);
  prhs[6] = dbl_matrix(K, R, alpha, K, 0);
  prhs[7] = dbl_matrix((mwSize)K * K, R, eps, (size_t)K * K, 0);
  prhs[8] = dbl_matrix(K, R, beta, K, 0);
  prhs[9] = dbl_matrix(K, R, v, K, 0);
  prhs[10] = dbl_matrix((mwSize)dim * K, R, m, (size_t)dim * K, 0);
  prhs[11] = dbl_matrix((mwSize)dim * dim * K, R, W, (size_t)dim * dim * K, 0);
  prhs[12] = dbl_scalar((double)max_iter);
  prhs[13] = dbl_scalar(min_diff);
  mexFunction(11, plhs, 14, prhs);
  double *outs[11] = {alpha, eps, beta, v, m, W, N1, Nk, M, LLs, info};
  for (int k = 0; k < 11; k++) memcpy(outs[k], mxGetPr(plhs[k]), sizeof(double) * mxGetNumberOfElements(plhs[k]));
  g_jmp_armed = 0;
  arena_reset();
  return 0;
}

/* Calls the linked gateway with a deliberately malformed argument list so the
 * tests can check the error ids the boundary raises
 * (vbhem_hmm_bwd_fwd_mex.c:335-366).  mode 0: wrong nrhs; 1: wrong nlhs;
 * 2: first arg not a cell; 3: second arg not a cell. */
int shim_call_malformed(int mode) {
  g_jmp_armed = 1;
  if (setjmp(g_jmp)) { g_jmp_armed = 0; arena_reset(); return 1; }
  const mxArray *prhs[7];
  mxArray *plhs[6] = {0, 0, 0, 0, 0, 0};
  mxArray *cellb = mxCreateCellMatrix(1, 0), *cellr = mxCreateCellMatrix(1, 0);
  for (int i = 0; i < 7; i++) prhs[i] = dbl_scalar(1.0);
  prhs[0] = (mode == 2) ? dbl_scalar(1.0) : cellb;
  prhs[1] = (mode == 3) ? dbl_scalar(1.0) : cellr;
  mexFunction(mode == 1 ? 5 : 6, plhs, mode == 0 ? 4 : 5, prhs);
  g_jmp_armed = 0;
  arena_reset();
  return 0;
}

/* ------------------------------------------------------------------ vbhmm_fb
 * Driver for the reference's VB-HMM forward-backward MEX
 * (/root/reference/src/hmm/vbhmm_fb_mex.c:185-186; call site vbhmm_fb.m:144-145), linked into
 * oracle/_ref/libvbhmm_fb_ref.so only:
 *   [logrho_Saved, gamma_all, xi_sum, phi_norm] =
 *       vbhmm_fb_mex(data, K, N, dim, maxT, m, W, v, beta, logLambdaTilde, const_denominator, t_pz1, t_tpztzt1)
 * seq = concatenated sequences, sequence n = [T_n x dim] column-major at offs[n]. */
int shim_vbhmm_fb(int N, int dim, int K, int maxT, const int *T, const long long *offs, const double *seq,
                  const double *m, const double *W, const double *v, const double *beta,
                  const double *logLambdaTilde, double const_denominator, const double *pz1, const double *A,
                  double *logrho, double *gamma, double *xi, double *phi) {
  g_jmp_armed = 1;
  if (setjmp(g_jmp)) { g_jmp_armed = 0; arena_reset(); return -1; }
  mxArray *data = mxCreateCellMatrix((mwSize)N, 1);
  for (int n = 0; n < N; n++) mxSetCell(data, (mwIndex)n, dbl_matrix((mwSize)T[n], (mwSize)dim, seq + offs[n], (size_t)T[n], 0));
  mwSize wd[3] = {(mwSize)dim, (mwSize)dim, (mwSize)K};
  /* MATLAB drops trailing singleton dimensions: W(:,:,1) alone is 2-D (vbhmm_fb_mex.c:137-139 relies on it) */
  mxArray *Wm = K > 1 ? mxCreateNumericArray(3, wd, mxDOUBLE_CLASS, mxREAL)
                      : mxCreateDoubleMatrix((mwSize)dim, (mwSize)dim, mxREAL);
  memcpy(mxGetPr(Wm), W, sizeof(double) * (size_t)dim * dim * K);
  const mxArray *prhs[13] = {data, dbl_scalar(K), dbl_scalar(N), dbl_scalar(dim), dbl_scalar(maxT),
                             dbl_matrix((mwSize)dim, (mwSize)K, m, (size_t)dim, 0), Wm, dbl_matrix((mwSize)K, 1, v, (size_t)K, 0),
                             dbl_matrix((mwSize)K, 1, beta, (size_t)K, 0), dbl_matrix(1, (mwSize)K, logLambdaTilde, 1, 0),
                             dbl_scalar(const_denominator), dbl_matrix(1, (mwSize)K, pz1, 1, 0),
                             dbl_matrix((mwSize)K, (mwSize)K, A, (size_t)K, 0)};
  mxArray *plhs[4] = {NULL, NULL, NULL, NULL};
  mexFunction(4, plhs, 13, prhs);
  const size_t nkt = (size_t)K * N * maxT;
  memcpy(logrho, mxGetPr(plhs[0]), sizeof(double) * nkt);
  memcpy(gamma, mxGetPr(plhs[1]), sizeof(double) * nkt);
  memcpy(xi, mxGetPr(plhs[2]), sizeof(double) * (size_t)K * K * N);
  memcpy(phi, mxGetPr(plhs[3]), sizeof(double) * (size_t)N);
  g_jmp_armed = 0;
  arena_reset();
  return 0;
}
