/* oracle/estep_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A plain-C CPU restatement of the VBHEM-H3M E-step on packed column-major
 * arrays.  It is the checker the CUDA path is compared against (tests/,
 * __graft_entry__.smoke(), bench.py's cpu_baseline leg) -- nothing in the
 * product package may call it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below
 * against the reference's own C source compiled in place
 * (oracle/_ref/libvbhem_ref.so, built by oracle/Makefile) and against the
 * committed fixtures in tests/golden/ that were generated from that build.
 * The MATLAB-only stages (responsibilities, statistics reduction) have no
 * native reference; they follow the .m files cited at each function and are
 * cross-checked against oracle/numpy_ref.py (an independent restatement).
 *
 * Each function cites the reference lines it follows
 * (paths relative to /root/reference/src/vbhem/).
 *
 * Packed layouts (fp64, column-major):
 *   base set : Sb[N]; prior[maxSb,N]; A[maxSb,maxSb,N] (A[from,to]);
 *              mu[d,maxSb,N]; cov[d,d,maxSb,N] (full) | [d,maxSb,N] (diag)
 *   centres  : Sr[K]; logPi[maxSr,K]; logA[maxSr,maxSr,K] (row = from);
 *              m[d,maxSr,K]; vW[d,d,maxSr,K] | [d,maxSr,K]  (= v*W);
 *              cterm[maxSr,K] (= -logLambdaTilde + d/lambda)
 *   per pair p = i + N*j : nu1[maxSr,p]; xi[maxSr,maxSr,p] (row=from);
 *              pr[maxSr,p]; mu_out[maxSr,d,p]; Mu_out[maxSr,dd,p]
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* log-sum-exp of one column: max shift, sequential sum, mv + log(sum).
 * Follows logtrick.m:15-18 and vbhem_hmm_bwd_fwd_mex.c:238-266. */
static double lse_col(const double *x, int n) {
  double mv = x[0];
  for (int i = 1; i < n; i++)
    if (x[i] > mv) mv = x[i];
  double s = 0.0;
  for (int i = 0; i < n; i++) s += exp(x[i] - mv);
  return mv + log(s);
}

typedef struct {
  double *E, *L0, *L1, *lt, *ls, *Th, *Th1, *nu, *foo, *x, *n;
} scratch_t;

static void scratch_alloc(scratch_t *s, int b, int r, int T) {
  s->E = malloc(sizeof(double) * b * r);
  s->L0 = malloc(sizeof(double) * b * r);
  s->L1 = malloc(sizeof(double) * b * r);
  s->lt = malloc(sizeof(double) * b * r);
  s->ls = malloc(sizeof(double) * b);
  s->Th = malloc(sizeof(double) * (size_t)r * r * b * (T > 0 ? T : 1));
  s->Th1 = malloc(sizeof(double) * b * r);
  s->nu = malloc(sizeof(double) * b * r);
  s->foo = malloc(sizeof(double) * b * r);
  s->x = malloc(sizeof(double) * b * r);
  s->n = malloc(sizeof(double) * b * r);
}
static void scratch_free(scratch_t *s) {
  free(s->E); free(s->L0); free(s->L1); free(s->lt); free(s->ls); free(s->Th);
  free(s->Th1); free(s->nu); free(s->foo); free(s->x); free(s->n);
}

/* One (base i, centre j) pair.  b = Sb_i, r = Sr_j (true sizes, not padded).
 * Pointers address this pair's own slices; ldb/ldr are the padded leading
 * dimensions of the base / centre arrays; ldo the padded leading dimension of
 * the outputs (maxSr). */
static double pair_estep(int b, int r, int d, int covmode, int T,
                         const double *prior, const double *Ab, int ldb, const double *mu, const double *cov,
                         const double *logPi, const double *logA, int ldr, const double *m, const double *vW,
                         const double *cterm, scratch_t *s,
                         double *nu1, double *xi, double *pr, double *mu_out, double *Mu_out, int ldo) {
  const int dd = covmode ? d * d : d;
  double *E = s->E;

  /* (a5) expected Gaussian log-likelihood E[beta,rho].
   * full: vbhem_hmm_bwd_fwd_mex.c:796-843; diag: :744-759 (v*W pre-multiplied). */
  for (int rho = 0; rho < r; rho++) {
    const double *mr = m + (size_t)d * rho, *P = vW + (size_t)dd * rho;
    for (int be = 0; be < b; be++) {
      const double *mb = mu + (size_t)d * be, *C = cov + (size_t)dd * be;
      double e = d * log(2.0 * M_PI) + cterm[rho];
      if (covmode) {
        for (int q = 0; q < dd; q++) e += P[q] * C[q];
        for (int tj = 0; tj < d; tj++) {
          double acc = 0.0;
          for (int ti = 0; ti < d; ti++) acc += (mb[ti] - mr[ti]) * P[ti + d * tj];
          e += acc * (mb[tj] - mr[tj]);
        }
      } else {
        for (int k = 0; k < d; k++) {
          const double df = mb[k] - mr[k];
          e += P[k] * C[k];
          e += P[k] * (df * df);
        }
      }
      E[be + b * rho] = -0.5 * e;
    }
  }

  /* (a7) backward recursion, vbhem_hmm_bwd_fwd_mex.c:902-1015.
   * Th[rho + r*(sigma + r*(beta + b*t))], t = 1..T-1. */
  double *Lo = s->L0, *Ln = s->L1;
  for (int q = 0; q < b * r; q++) Lo[q] = 0.0;
  for (int t = T - 1; t >= 1; t--) {
    for (int rho = 0; rho < r; rho++) {
      for (int sg = 0; sg < r; sg++)
        for (int be = 0; be < b; be++)
          s->lt[sg + r * be] = logA[rho + ldr * sg] + E[be + b * sg] + Lo[be + b * sg];
      for (int be = 0; be < b; be++) s->ls[be] = lse_col(s->lt + r * be, r);
      for (int bp = 0; bp < b; bp++) {
        double acc = 0.0;
        for (int be = 0; be < b; be++) acc += Ab[bp + ldb * be] * s->ls[be];
        Ln[bp + b * rho] = acc;
      }
      for (int be = 0; be < b; be++)
        for (int sg = 0; sg < r; sg++)
          s->Th[rho + r * (sg + r * (be + (size_t)b * t))] = exp(s->lt[sg + r * be] - s->ls[be]);
    }
    double *tmp = Lo; Lo = Ln; Ln = tmp;
  }

  /* (a8) termination, vbhem_hmm_bwd_fwd_mex.c:1029-1080. */
  for (int sg = 0; sg < r; sg++)
    for (int be = 0; be < b; be++)
      s->lt[sg + r * be] = logPi[sg] + E[be + b * sg] + Lo[be + b * sg];
  double L = 0.0;
  for (int be = 0; be < b; be++) s->ls[be] = lse_col(s->lt + r * be, r);
  for (int be = 0; be < b; be++) L += prior[be] * s->ls[be];
  for (int be = 0; be < b; be++)
    for (int sg = 0; sg < r; sg++) s->Th1[sg + r * be] = exp(s->lt[sg + r * be] - s->ls[be]);

  if (!nu1) return L;   /* caller only wants L_elbo */

  /* (a9) forward recursion, vbhem_hmm_bwd_fwd_mex.c:1100-1298. */
  double *nu = s->nu, *n = s->n;
  for (int ga = 0; ga < b; ga++)
    for (int sg = 0; sg < r; sg++) nu[sg + r * ga] = prior[ga] * s->Th1[sg + r * ga];
  for (int sg = 0; sg < r; sg++) {
    double acc = 0.0;
    for (int ga = 0; ga < b; ga++) acc += nu[sg + r * ga];
    nu1[sg] = acc;
  }
  memcpy(n, nu, sizeof(double) * b * r);
  for (int sg = 0; sg < r; sg++)
    for (int rho = 0; rho < r; rho++) xi[rho + ldo * sg] = 0.0;
  for (int t = 1; t < T; t++) {
    for (int ga = 0; ga < b; ga++)
      for (int rho = 0; rho < r; rho++) {
        double acc = 0.0;
        for (int ka = 0; ka < b; ka++) acc += nu[rho + r * ka] * Ab[ka + ldb * ga];
        s->foo[rho + r * ga] = acc;
      }
    for (int sg = 0; sg < r; sg++) {
      for (int ga = 0; ga < b; ga++)
        for (int rho = 0; rho < r; rho++)
          s->x[rho + r * ga] = s->foo[rho + r * ga] * s->Th[rho + r * (sg + r * (ga + (size_t)b * t))];
      for (int rho = 0; rho < r; rho++) {
        double acc = 0.0;
        for (int ga = 0; ga < b; ga++) acc += s->x[rho + r * ga];
        xi[rho + ldo * sg] += acc;
      }
      for (int ga = 0; ga < b; ga++) {
        double acc = 0.0;
        for (int rho = 0; rho < r; rho++) acc += s->x[rho + r * ga];
        nu[sg + r * ga] = acc;
      }
    }
    for (int q = 0; q < b * r; q++) n[q] += nu[q];
  }

  /* (a10) emission statistics, vbhem_hmm_bwd_fwd_mex.c:1348-1469. */
  for (int sg = 0; sg < r; sg++) {
    double acc = 0.0;
    for (int ga = 0; ga < b; ga++) acc += n[sg + r * ga] * 1.0;
    pr[sg] = acc;
    for (int k = 0; k < d; k++) {
      acc = 0.0;
      for (int ga = 0; ga < b; ga++) acc += n[sg + r * ga] * mu[k + (size_t)d * ga];
      mu_out[sg + ldo * k] = acc;
    }
    if (covmode) {
      for (int k = 0; k < d; k++)
        for (int l = 0; l < d; l++) {
          acc = 0.0;
          for (int ga = 0; ga < b; ga++) {
            const double *mb = mu + (size_t)d * ga;
            acc += n[sg + r * ga] * (mb[k] * mb[l] + cov[k * d + l + (size_t)dd * ga]);
          }
          Mu_out[sg + ldo * (k * d + l)] = acc;
        }
    } else {
      for (int k = 0; k < d; k++) {
        acc = 0.0;
        for (int ga = 0; ga < b; ga++) {
          const double c = mu[k + (size_t)d * ga];
          acc += n[sg + r * ga] * (c * c + cov[k + (size_t)dd * ga]);
        }
        Mu_out[sg + ldo * k] = acc;
      }
    }
  }
  return L;
}

/* All N x K pairs, literal per-pair outputs (the MEX boundary's six outputs,
 * vbhem_hmm_bwd_fwd_mex.c:396-409).  Any of nu1..Mu_out may be NULL together
 * to get L_elbo only.  The centre loop is outermost as in :505,520. */
int oracle_estep_pairs(int N, const int *Sb, int maxSb, int d, int covmode,
                       const double *prior, const double *A, const double *mu, const double *cov,
                       int K, const int *Sr, int maxSr,
                       const double *logPi, const double *logA, const double *m, const double *vW,
                       const double *cterm, int T,
                       double *L_elbo, double *nu1, double *xi, double *pr, double *mu_out, double *Mu_out) {
  const int dd = covmode ? d * d : d;
  scratch_t s;
  scratch_alloc(&s, maxSb, maxSr, T);
  for (int j = 0; j < K; j++)
    for (int i = 0; i < N; i++) {
      const size_t p = (size_t)i + (size_t)N * j;
      const size_t cb = (size_t)maxSb * i, cr = (size_t)maxSr * j;
      L_elbo[p] = pair_estep(
          Sb[i], Sr[j], d, covmode, T, prior + cb, A + (size_t)maxSb * cb, maxSb, mu + (size_t)d * cb,
          cov + (size_t)dd * cb, logPi + cr, logA + (size_t)maxSr * cr, maxSr, m + (size_t)d * cr,
          vW + (size_t)dd * cr, cterm + cr, &s,
          nu1 ? nu1 + (size_t)maxSr * p : NULL, nu1 ? xi + (size_t)maxSr * maxSr * p : NULL,
          nu1 ? pr + (size_t)maxSr * p : NULL, nu1 ? mu_out + (size_t)maxSr * d * p : NULL,
          nu1 ? Mu_out + (size_t)maxSr * dd * p : NULL, maxSr);
    }
  scratch_free(&s);
  return 0;
}

/* Responsibilities (vbhem_h3m_c_step_fc.m:271-283) and the two N-sized
 * lower-bound terms (vbhemh3m_lb.m:90,107).  group[j] gives the softmax
 * segment of centre j (all zeros = one H3M; the (K,S) grid batches several
 * models in one call).  zhat, Z are [N,K].  lb[0]=Lt1=sum Z.*L_elbo,
 * lb[1]=Lt7=sum zhat.*log(zhat), accumulated per group: lb[2*g+{0,1}]. */
int oracle_zhat(int N, int K, const double *L_elbo, const double *logOmega, const double *Ntilde,
                const int *group, int ngroups, double *zhat, double *Z, double *Nj, double *lb) {
  double *lz = malloc(sizeof(double) * K);
  for (int j = 0; j < K; j++) Nj[j] = 0.0;
  for (int g = 0; g < 2 * ngroups; g++) lb[g] = 0.0;
  for (int g = 0; g < ngroups; g++) {
    for (int i = 0; i < N; i++) {
      int first = 1;
      double mv = 0.0, s = 0.0;
      for (int j = 0; j < K; j++) {
        if (group[j] != g) continue;
        lz[j] = Ntilde[i] * (logOmega[j] + L_elbo[i + (size_t)N * j]);
        if (first || lz[j] > mv) { mv = lz[j]; first = 0; }
      }
      for (int j = 0; j < K; j++)
        if (group[j] == g) s += exp(lz[j] - mv);
      const double lse = mv + log(s);
      for (int j = 0; j < K; j++) {
        if (group[j] != g) continue;
        const double z = exp(lz[j] - lse) + 1e-50;
        zhat[i + (size_t)N * j] = z;
        Z[i + (size_t)N * j] = z * Ntilde[i];
      }
    }
    /* column sums in MATLAB's order (down each column) */
    for (int j = 0; j < K; j++) {
      if (group[j] != g) continue;
      double nj = 0.0, t1 = 0.0, t7 = 0.0;
      for (int i = 0; i < N; i++) {
        const size_t p = i + (size_t)N * j;
        nj += Z[p];
        t1 += Z[p] * L_elbo[p];
        t7 += zhat[p] * log(zhat[p]);
      }
      Nj[j] = nj + 1e-50;
      lb[2 * g] += t1;
      lb[2 * g + 1] += t7;
    }
  }
  free(lz);
  return 0;
}

/* Statistics reduction, raw sums only (vbhem_compute_Statistics.m:33-55): for
 * every centre j, sum over bases with Z(i,j) > 1e-8 of Z * per-pair statistic.
 * The normalisation of :57-82 is host logic (numpy_ref.finalize_stats).
 * Outputs: S_nu1[maxSr,K], S_xi[maxSr,maxSr,K], S_pr[maxSr,K],
 *          S_mu[maxSr,d,K], S_Mu[maxSr,dd,K]. */
int oracle_reduce(int N, int K, int maxSr, int d, int covmode, const double *Z,
                  const double *nu1, const double *xi, const double *pr, const double *mu_out,
                  const double *Mu_out, double *S_nu1, double *S_xi, double *S_pr, double *S_mu, double *S_Mu) {
  const int dd = covmode ? d * d : d;
  const int r = maxSr;
  memset(S_nu1, 0, sizeof(double) * r * K);
  memset(S_xi, 0, sizeof(double) * r * r * K);
  memset(S_pr, 0, sizeof(double) * r * K);
  memset(S_mu, 0, sizeof(double) * r * d * K);
  memset(S_Mu, 0, sizeof(double) * r * dd * K);
  for (int j = 0; j < K; j++)
    for (int i = 0; i < N; i++) {
      const size_t p = i + (size_t)N * j;
      const double z = Z[p];
      if (!(z > 1e-8)) continue;
      for (int q = 0; q < r; q++) S_nu1[q + (size_t)r * j] += z * nu1[q + (size_t)r * p];
      for (int q = 0; q < r * r; q++) S_xi[q + (size_t)r * r * j] += z * xi[q + (size_t)r * r * p];
      for (int q = 0; q < r; q++) S_pr[q + (size_t)r * j] += z * pr[q + (size_t)r * p];
      for (int q = 0; q < r * d; q++) S_mu[q + (size_t)r * d * j] += z * mu_out[q + (size_t)r * d * p];
      for (int q = 0; q < r * dd; q++) S_Mu[q + (size_t)r * dd * j] += z * Mu_out[q + (size_t)r * dd * p];
    }
  return 0;
}

/* Fused E-step for sizes where the literal per-pair arrays do not fit: per
 * base i, all K pairs -> softmax -> weighted accumulation.  Same arithmetic as
 * oracle_estep_pairs + oracle_zhat + oracle_reduce, but the sum over i runs
 * base-outer, so results agree with the literal path to rounding
 * (tests/test_oracle.py bounds it).  Outputs as oracle_zhat / oracle_reduce;
 * L_elbo / zhat may be NULL.  Single softmax group. */
int oracle_estep_fused(int N, const int *Sb, int maxSb, int d, int covmode,
                       const double *prior, const double *A, const double *mu, const double *cov,
                       const double *Ntilde,
                       int K, const int *Sr, int maxSr,
                       const double *logPi, const double *logA, const double *m, const double *vW,
                       const double *cterm, const double *logOmega, int T,
                       double *L_elbo, double *zhat, double *Nj, double *lb,
                       double *S_nu1, double *S_xi, double *S_pr, double *S_mu, double *S_Mu) {
  const int dd = covmode ? d * d : d;
  const int r = maxSr;
  const int per = r + r * r + r + r * d + r * dd;
  scratch_t s;
  scratch_alloc(&s, maxSb, maxSr, T);
  double *buf = calloc((size_t)per * K, sizeof(double));
  double *Lrow = malloc(sizeof(double) * K), *lz = malloc(sizeof(double) * K);
  memset(S_nu1, 0, sizeof(double) * r * K);
  memset(S_xi, 0, sizeof(double) * r * r * K);
  memset(S_pr, 0, sizeof(double) * r * K);
  memset(S_mu, 0, sizeof(double) * r * d * K);
  memset(S_Mu, 0, sizeof(double) * r * dd * K);
  for (int j = 0; j < K; j++) Nj[j] = 0.0;
  lb[0] = lb[1] = 0.0;
  for (int i = 0; i < N; i++) {
    const size_t cb = (size_t)maxSb * i;
    for (int j = 0; j < K; j++) {
      const size_t cr = (size_t)maxSr * j;
      double *o = buf + (size_t)per * j;
      memset(o, 0, sizeof(double) * per);
      Lrow[j] = pair_estep(Sb[i], Sr[j], d, covmode, T, prior + cb, A + (size_t)maxSb * cb, maxSb,
                           mu + (size_t)d * cb, cov + (size_t)dd * cb, logPi + cr, logA + (size_t)maxSr * cr,
                           maxSr, m + (size_t)d * cr, vW + (size_t)dd * cr, cterm + cr, &s,
                           o, o + r, o + r + r * r, o + 2 * r + r * r, o + 2 * r + r * r + r * d, maxSr);
    }
    double mv = 0.0, sum = 0.0;
    for (int j = 0; j < K; j++) {
      lz[j] = Ntilde[i] * (logOmega[j] + Lrow[j]);
      if (j == 0 || lz[j] > mv) mv = lz[j];
    }
    for (int j = 0; j < K; j++) sum += exp(lz[j] - mv);
    const double lse = mv + log(sum);
    for (int j = 0; j < K; j++) {
      const double z = exp(lz[j] - lse) + 1e-50, Z = z * Ntilde[i];
      if (L_elbo) L_elbo[i + (size_t)N * j] = Lrow[j];
      if (zhat) zhat[i + (size_t)N * j] = z;
      Nj[j] += Z;
      lb[0] += Z * Lrow[j];
      lb[1] += z * log(z);
      if (!(Z > 1e-8)) continue;
      const double *o = buf + (size_t)per * j;
      for (int q = 0; q < r; q++) S_nu1[q + (size_t)r * j] += Z * o[q];
      for (int q = 0; q < r * r; q++) S_xi[q + (size_t)r * r * j] += Z * o[r + q];
      for (int q = 0; q < r; q++) S_pr[q + (size_t)r * j] += Z * o[r + r * r + q];
      for (int q = 0; q < r * d; q++) S_mu[q + (size_t)r * d * j] += Z * o[2 * r + r * r + q];
      for (int q = 0; q < r * dd; q++) S_Mu[q + (size_t)r * dd * j] += Z * o[2 * r + r * r + r * d + q];
    }
  }
  for (int j = 0; j < K; j++) Nj[j] += 1e-50;
  free(buf); free(Lrow); free(lz);
  scratch_free(&s);
  return 0;
}
