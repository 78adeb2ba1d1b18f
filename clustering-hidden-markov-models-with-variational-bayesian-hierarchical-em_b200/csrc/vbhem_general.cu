// vbhem_general.cu -- the general pair kernel: any state counts, any emission
// dimension, any tau; literal log-space arithmetic.
//
// One warp owns one (base i, centre j) pair and walks the reference's recursion
// in the reference's own order of operations, with the lattice arrays in a
// per-warp slab of global scratch (L1/L2 resident).  It is
//   * the parity anchor on the GPU (same max-shift log-sum-exp as
//     /root/reference/src/vbhem/vbhem_hmm_bwd_fwd_mex.c:211-269),
//   * the path for shapes the tensor-core lattice kernel does not cover
//     (more than 8 states, d > 3, very long tau), and
//   * the in-GPU fall-back for pairs whose linear-domain sums underflow.
// Steps follow vbhem_hmm_bwd_fwd_mex.c: emission expectation :716-865, backward
// :902-1015, termination :1029-1080, forward :1100-1298, statistics :1348-1469.
#include "vbhem_internal.cuh"

namespace vbhem {

namespace {

constexpr int kWarpsPerBlock = 4;

__device__ __forceinline__ double lse_col(const double *x, int n) {
  double mv = x[0];
  for (int i = 1; i < n; i++)
    if (x[i] > mv) mv = x[i];
  double s = 0.0;
  for (int i = 0; i < n; i++) s += exp(x[i] - mv);
  return mv + log(s);
}

struct Slab {
  double *E, *Lo, *Ln, *lt, *ls, *Th1, *nu, *foo, *x, *n, *out, *Th;
};

__device__ __forceinline__ Slab carve(double *p, int b, int r, int T, int per) {
  Slab s;
  const int br = b * r;
  s.E = p; p += br;
  s.Lo = p; p += br;
  s.Ln = p; p += br;
  s.lt = p; p += br;
  s.ls = p; p += b;
  s.Th1 = p; p += br;
  s.nu = p; p += br;
  s.foo = p; p += br;
  s.x = p; p += br;
  s.n = p; p += br;
  s.out = p; p += per;
  s.Th = p;
  return s;
}

// The whole pair.  want_fwd = false stops after L_elbo.  Per-pair statistics land
// in s.out: nu1[r] | xi[r,r] | pr[r] | mu[r,d] | Mu[r,dd]  (leading dim ro).
__device__ double pair_general(const BaseView &B, const CentreView &C, int T, int i, int j, Slab s, int ro,
                               bool want_fwd, int lane) {
  const int b = B.Sb[i], r = C.Sr[j], d = B.d, dd = B.dd;
  const int ldb = B.maxSb, ldr = C.maxSr;
  const double *prior = B.prior + (size_t)ldb * i;
  const double *Ab = B.A + (size_t)ldb * ldb * i;
  const double *mu = B.mu + (size_t)d * ldb * i;
  const double *cov = B.cov + (size_t)dd * ldb * i;
  const double *logPi = C.logPi + (size_t)ldr * j;
  const double *logA = C.logA + (size_t)ldr * ldr * j;
  const double *m = C.m + (size_t)d * ldr * j;
  const double *vW = C.vW + (size_t)dd * ldr * j;
  const double *cterm = C.cterm + (size_t)ldr * j;
  const int br = b * r;

  // emission expectation E[beta + b*rho]
  for (int q = lane; q < br; q += 32) {
    const int be = q % b, rho = q / b;
    const double *mr = m + (size_t)d * rho, *P = vW + (size_t)dd * rho;
    const double *mb = mu + (size_t)d * be, *Cv = cov + (size_t)dd * be;
    double e = d * log(kTwoPi) + cterm[rho];
    if (B.covmode) {
      for (int k = 0; k < dd; k++) e += P[k] * Cv[k];
      for (int tj = 0; tj < d; tj++) {
        double acc = 0.0;
        for (int ti = 0; ti < d; ti++) acc += (mb[ti] - mr[ti]) * P[ti + d * tj];
        e += acc * (mb[tj] - mr[tj]);
      }
    } else {
      for (int k = 0; k < d; k++) {
        const double df = mb[k] - mr[k];
        e += P[k] * Cv[k];
        e += P[k] * (df * df);
      }
    }
    s.E[q] = -0.5 * e;
    s.Lo[q] = 0.0;
  }
  __syncwarp();

  // backward
  double *Lo = s.Lo, *Ln = s.Ln;
  for (int t = T - 1; t >= 1; t--) {
    for (int rho = 0; rho < r; rho++) {
      for (int q = lane; q < br; q += 32) {
        const int be = q % b, sg = q / b;
        s.lt[sg + r * be] = logA[rho + ldr * sg] + s.E[q] + Lo[q];
      }
      __syncwarp();
      for (int be = lane; be < b; be += 32) s.ls[be] = lse_col(s.lt + r * be, r);
      __syncwarp();
      for (int bp = lane; bp < b; bp += 32) {
        double acc = 0.0;
        for (int be = 0; be < b; be++) acc += Ab[bp + ldb * be] * s.ls[be];
        Ln[bp + b * rho] = acc;
      }
      if (want_fwd) {
        double *Th = s.Th + (size_t)r * r * b * t;
        for (int q = lane; q < br; q += 32) {
          const int sg = q % r, be = q / r;
          Th[rho + r * (sg + r * be)] = exp(s.lt[q] - s.ls[be]);
        }
      }
      __syncwarp();
    }
    double *tmp = Lo; Lo = Ln; Ln = tmp;
  }

  // termination
  for (int q = lane; q < br; q += 32) {
    const int be = q % b, sg = q / b;
    s.lt[sg + r * be] = logPi[sg] + s.E[q] + Lo[q];
  }
  __syncwarp();
  for (int be = lane; be < b; be += 32) s.ls[be] = lse_col(s.lt + r * be, r);
  __syncwarp();
  double L = 0.0;
  for (int be = 0; be < b; be++) L += prior[be] * s.ls[be];
  if (!want_fwd) return L;

  // forward
  for (int q = lane; q < br; q += 32) {
    const int be = q / r;
    const double th = exp(s.lt[q] - s.ls[be]);
    s.Th1[q] = th;
    s.nu[q] = prior[be] * th;
    s.n[q] = s.nu[q];
  }
  double *o_nu1 = s.out, *o_xi = s.out + ro, *o_pr = o_xi + ro * ro, *o_mu = o_pr + ro, *o_Mu = o_mu + ro * d;
  for (int q = lane; q < stat_per(ro, d, dd); q += 32) s.out[q] = 0.0;   // rows beyond r stay zero
  __syncwarp();
  for (int sg = lane; sg < r; sg += 32) {
    double acc = 0.0;
    for (int ga = 0; ga < b; ga++) acc += s.nu[sg + r * ga];
    o_nu1[sg] = acc;
  }
  for (int t = 1; t < T; t++) {
    const double *Th = s.Th + (size_t)r * r * b * t;
    for (int q = lane; q < br; q += 32) {
      const int rho = q % r, ga = q / r;
      double acc = 0.0;
      for (int ka = 0; ka < b; ka++) acc += s.nu[rho + r * ka] * Ab[ka + ldb * ga];
      s.foo[q] = acc;
    }
    __syncwarp();
    for (int sg = 0; sg < r; sg++) {
      for (int q = lane; q < br; q += 32) {
        const int rho = q % r, ga = q / r;
        s.x[q] = s.foo[q] * Th[rho + r * (sg + r * ga)];
      }
      __syncwarp();
      for (int rho = lane; rho < r; rho += 32) {
        double acc = 0.0;
        for (int ga = 0; ga < b; ga++) acc += s.x[rho + r * ga];
        o_xi[rho + ro * sg] += acc;
      }
      for (int ga = lane; ga < b; ga += 32) {
        double acc = 0.0;
        for (int rho = 0; rho < r; rho++) acc += s.x[rho + r * ga];
        s.nu[sg + r * ga] = acc;
      }
      __syncwarp();
    }
    for (int q = lane; q < br; q += 32) s.n[q] += s.nu[q];
    __syncwarp();
  }

  // emission statistics
  for (int sg = lane; sg < r; sg += 32) {
    double acc = 0.0;
    for (int ga = 0; ga < b; ga++) acc += s.n[sg + r * ga];
    o_pr[sg] = acc;
  }
  for (int q = lane; q < r * d; q += 32) {
    const int sg = q % r, k = q / r;
    double acc = 0.0;
    for (int ga = 0; ga < b; ga++) acc += s.n[sg + r * ga] * mu[k + (size_t)d * ga];
    o_mu[sg + ro * k] = acc;
  }
  for (int q = lane; q < r * dd; q += 32) {
    const int sg = q % r, kl = q / r;
    double acc = 0.0;
    if (B.covmode) {
      const int k = kl / d, l = kl % d;
      for (int ga = 0; ga < b; ga++) {
        const double *mb = mu + (size_t)d * ga;
        acc += s.n[sg + r * ga] * (mb[k] * mb[l] + cov[kl + (size_t)dd * ga]);
      }
    } else {
      for (int ga = 0; ga < b; ga++) {
        const double c = mu[kl + (size_t)d * ga];
        acc += s.n[sg + r * ga] * (c * c + cov[kl + (size_t)dd * ga]);
      }
    }
    o_Mu[sg + ro * kl] = acc;
  }
  __syncwarp();
  return L;
}

__device__ void store_pair(const BaseView &B, const CentreView &C, const PairOut &O, const Slab &s, int i, int j,
                           int lane) {
  const int ro = C.maxSr, d = B.d, dd = B.dd;
  const size_t p = (size_t)i + (size_t)B.N * j;
  const double *o_nu1 = s.out, *o_xi = s.out + ro, *o_pr = o_xi + ro * ro, *o_mu = o_pr + ro, *o_Mu = o_mu + ro * d;
  for (int q = lane; q < ro; q += 32) {
    O.nu1[q + (size_t)ro * p] = o_nu1[q];
    O.pr[q + (size_t)ro * p] = o_pr[q];
  }
  for (int q = lane; q < ro * ro; q += 32) O.xi[q + (size_t)ro * ro * p] = o_xi[q];
  for (int q = lane; q < ro * d; q += 32) O.mu[q + (size_t)ro * d * p] = o_mu[q];
  for (int q = lane; q < ro * dd; q += 32) O.Mu[q + (size_t)ro * dd * p] = o_Mu[q];
}

// flags: optional per-pair (all-pairs work) or per-item (list work) filter; only
// entries with a non-zero flag are processed.  count: optional device counters; the
// kernel returns at once when count[0] is zero (nothing is flagged).  All-pairs fall-back:
// count[0] = pairs still pending, count[2] = pairs served in this E-step; a served pair's flag is
// cleared, so the flag array is all-zero again when the last pending pair has been served.
// Only bases [i0, i1) are looked at (a slice of a streamed E-step).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_general(BaseView B, CentreView C, int T, PairWork W, PairOut O, unsigned char *flags, int *count,
          double *scratch, size_t per_warp, int i0, int i1) {
  pdl_trigger();
  pdl_wait();
  if (count && *count == 0) return;
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  const int ro = C.maxSr;
  const int per = stat_per(ro, B.d, B.dd);
  Slab s = carve(scratch + per_warp * warp, B.maxSb, C.maxSr, T, per);

  if (W.items == nullptr) {
    const long long ni = i1 - i0, np = ni * C.K;
    for (long long u = warp; u < np; u += nwarps) {
      const int i = i0 + (int)(u % ni), j = (int)(u / ni);
      const long long p = i + (long long)B.N * j;
      VB_ASSERT(i >= 0 && i < B.N && j >= 0 && j < C.K);
      if (flags && !flags[p]) continue;
      const double L = pair_general(B, C, T, i, j, s, ro, O.mode != 0, lane);
      if (lane == 0 && O.L_elbo) O.L_elbo[p] = L;
      if (O.mode == 1) store_pair(B, C, O, s, i, j, lane);
      if (flags && lane == 0) {
        flags[p] = 0;
        if (count) { atomicSub(count, 1); atomicAdd(count + 2, 1); }
      }
      __syncwarp();
    }
  } else {
    // list work (PairWork): one warp per range; slot (range + centre) is owned by that warp alone
    const int nitems = W.totals[0], ipr = W.totals[1];
    for (int w = (int)warp; w < W.nranges; w += (int)nwarps) {
      const int e0 = w * ipr, e1 = min(e0 + ipr, nitems);
      int jcur = -1;
      double *part = nullptr;
      for (int e = e0; e < e1; e++) {
        if (flags && !flags[e]) continue;
        const PairItem it = W.items[e];
        VB_ASSERT((long long)e < W.npairs && it.i >= 0 && it.i < B.N && it.j >= 0 && it.j < C.K &&
                  w + it.j < W.nranges + C.K);
        if (it.j != jcur) {
          jcur = it.j;
          part = O.partial + (size_t)O.per * (w + jcur);
          if (!flags)                                   // primary path: this warp initialises the slot
            for (int q = lane; q < per; q += 32) part[q] = 0.0;
          __syncwarp();
        }
        const double wgt = W.zhat[it.i + (size_t)B.N * it.j] * W.Ntilde[it.i];
        pair_general(B, C, T, it.i, it.j, s, ro, true, lane);
        for (int q = lane; q < per; q += 32) part[q] += wgt * s.out[q];
        __syncwarp();
      }
    }
  }
}

}  // namespace

size_t general_scratch_per_warp(int maxSb, int maxSr, int d, int dd, int T) {
  const size_t br = (size_t)maxSb * maxSr;
  // E Lo Ln lt Th1 nu foo x n (9 x br) + ls + out + Theta
  return 9 * br + maxSb + stat_per(maxSr, d, dd) + (size_t)maxSr * maxSr * maxSb * (T > 0 ? T : 1);
}

int launch_general(const BaseView &B, const CentreView &C, int T, const PairWork &W, const PairOut &O,
                   unsigned char *flags, int *count, double *scratch, size_t scratch_doubles, cudaStream_t st, int i0,
                   int i1) {
  if (i1 < 0) i1 = B.N;
  const size_t per_warp = general_scratch_per_warp(B.maxSb, C.maxSr, B.d, B.dd, T);
  long long units = W.items ? W.nranges : W.npairs;
  if (W.items && W.npairs < units) units = W.npairs;   // never more ranges in use than items
  if (units <= 0) return 0;
  long long warps = (long long)(scratch_doubles / per_warp);
  if (warps > units) warps = units;
  long long blocks = (warps + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (blocks * kWarpsPerBlock * per_warp > scratch_doubles) blocks = warps / kWarpsPerBlock;
  if (blocks < 1) return -1;   // scratch too small for even one block
  launch_chain<1>(k_general, dim3((unsigned)blocks), dim3(kWarpsPerBlock * 32), 0, st, B, C, T, W, O, flags, count, scratch,
               per_warp, i0, i1);
  return 1;
}

}  // namespace vbhem
