/* oracle/fb_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the VB-HMM forward-backward pass of the reference
 * (/root/reference/src/hmm/vbhmm_fb.m:264-380, native twin
 * /root/reference/src/hmm/vbhmm_fb_mex.c:314-950) on flat arrays, used as the parity checker of the
 * CUDA path (vbhmm_fb in libvbhem_b200.so) next to the in-place build of the reference's own C
 * (oracle/_ref/libvbhmm_fb_ref.so).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline may call it.
 *
 * Layouts (all column-major fp64, exactly the MEX arguments):
 *   seq    : concatenated sequences, sequence n = [T_n x dim] (time fastest), start offs[n] (in doubles)
 *   m[dim,K], W[dim,dim,K], v[K], beta[K], logLambdaTilde[K], pz1[K], A[K(from),K(to)]
 *   logrho[K,N,maxT], gamma[K,N,maxT], xi[K,K,N], phi[N]  (zero beyond T_n)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

int fb_oracle(int N, int dim, int K, int maxT, const int *T, const long long *offs, const double *seq,
              const double *m, const double *W, const double *v, const double *beta, const double *logLambdaTilde,
              double const_denominator, const double *pz1, const double *A,
              double *logrho, double *gamma, double *xi, double *phi) {
  const size_t NK = (size_t)N * K;
  double *alpha = (double *)malloc(sizeof(double) * (size_t)maxT * K);
  double *p = (double *)malloc(sizeof(double) * (size_t)maxT * K);
  double *c = (double *)malloc(sizeof(double) * (size_t)maxT);
  double *bt = (double *)malloc(sizeof(double) * 2 * K);
  double *diff = (double *)malloc(sizeof(double) * dim);
  if (!alpha || !p || !c || !bt || !diff) return -1;
  memset(logrho, 0, sizeof(double) * NK * maxT);
  memset(gamma, 0, sizeof(double) * NK * maxT);
  memset(xi, 0, sizeof(double) * (size_t)K * K * N);
  for (int n = 0; n < N; n++) {
    const int tT = T[n];
    const double *x = seq + offs[n];
    double summax = 0.0, sumlogc = 0.0;
    /* emissions: vbhmm_fb.m:279-293 (Bishop 10.64, 10.46), max-shifted :316-317 */
    for (int t = 0; t < tT; t++) {
      double mx = 0.0;
      for (int k = 0; k < K; k++) {
        for (int i = 0; i < dim; i++) diff[i] = x[t + (size_t)i * tT] - m[i + dim * k];
        double mterm = 0.0;
        for (int i = 0; i < dim; i++) {
          double tmp = 0.0;
          /* the MEX walks W sequentially, i.e. uses W(j,i,k): "use the fact that W is symmetric"
             (vbhmm_fb_mex.c:378-391) */
          for (int j = 0; j < dim; j++) tmp += W[j + dim * (i + dim * k)] * diff[j];
          mterm += tmp * diff[i];
        }
        const double delta = dim / beta[k] + v[k] * mterm;
        const double lr = 0.5 * (logLambdaTilde[k] - delta) - const_denominator;
        logrho[k + K * (n + (size_t)N * t)] = lr;
        if (k == 0 || lr > mx) mx = lr;
      }
      summax += mx;
      for (int k = 0; k < K; k++) p[k + K * t] = exp(logrho[k + K * (n + (size_t)N * t)] - mx);
    }
    if (tT >= 1) {
      /* scaled forward: vbhmm_fb.m:330-347 */
      for (int t = 0; t < tT; t++) {
        double ct = 0.0;
        for (int k = 0; k < K; k++) {
          double d;
          if (t == 0) d = pz1[k] * p[k];
          else {
            double tmp = 0.0;
            for (int j = 0; j < K; j++) tmp += alpha[j + K * (t - 1)] * A[j + K * k];
            d = tmp * p[k + K * t];
          }
          alpha[k + K * t] = d;
          ct += d;
        }
        c[t] = ct;
        for (int k = 0; k < K; k++) alpha[k + K * t] /= ct;
        sumlogc += log(ct);
      }
      /* backward, responsibilities, xi: vbhmm_fb.m:349-377 */
      double *b = bt, *bpi = bt + K;
      for (int k = 0; k < K; k++) {
        b[k] = 1.0;
        gamma[k + K * (n + (size_t)N * (tT - 1))] = alpha[k + K * (tT - 1)];
      }
      for (int t = tT - 2; t >= 0; t--) {
        for (int k = 0; k < K; k++) bpi[k] = b[k] * p[k + K * (t + 1)];
        for (int j = 0; j < K; j++) {
          double eta = 0.0;
          for (int k = 0; k < K; k++) eta += bpi[k] * A[j + K * k];
          b[j] = eta / c[t + 1];
        }
        for (int k = 0; k < K; k++) gamma[k + K * (n + (size_t)N * t)] = alpha[k + K * t] * b[k];
        for (int k = 0; k < K; k++)
          for (int j = 0; j < K; j++)
            xi[j + K * (k + (size_t)K * n)] += (A[j + K * k] * (alpha[j + K * t] * bpi[k])) / c[t + 1];
      }
    }
    phi[n] = sumlogc + summax;      /* vbhmm_fb.m:377 */
  }
  free(alpha); free(p); free(c); free(bt); free(diff);
  return 0;
}
