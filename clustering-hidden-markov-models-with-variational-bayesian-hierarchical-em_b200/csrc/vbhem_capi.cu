// vbhem_capi.cu -- the C ABI declared in include/vbhem_b200.h.
//
// Host-side orchestration only: argument checks (the reference's gateway checks,
// /root/reference/src/vbhem/vbhem_hmm_bwd_fwd_mex.c:334-377), device buffers, and
// the launch sequence of one E-step.  All arithmetic is in the kernels.
#include "../../include/vbhem_b200.h"
#include "vbhem_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace vbhem;


namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

}  // namespace

// error sink shared with vbhmm_fb.cu (same thread-local message behind vbhem_last_error())
namespace vbhem {
int fail_fb(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
}  // namespace vbhem

namespace {

#define CU(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess)                                                                     \
      return fail(e_ == cudaErrorMemoryAllocation ? VBHEM_ENOMEM : VBHEM_ECUDA, "%s: %s (%s:%d)", #x, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                 \
  } while (0)

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T *as() const { return static_cast<T *>(p); }
};

// Pinned host staging area: small uploads copied here first travel as truly asynchronous copies (a cudaMemcpyAsync
// from pageable memory blocks the host for a staged copy each).  The E-step never synchronises its stream, so the
// handle keeps a ring of these (StageRing) and an event per slot says when the slot's copies have left.
struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

constexpr int kStageSlots = 4;
struct StageRing {
  PinBuf buf[kStageSlots];
  cudaEvent_t ev[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
  bool busy[kStageSlots] = {false, false, false, false};
  int next = 0, cur = 0;
  // a slot whose earlier copies have completed, at least `bytes` large
  cudaError_t acquire(size_t bytes, void **out) {
    cur = next;
    next = (next + 1) % kStageSlots;
    cudaError_t e;
    if (!ev[cur] && (e = cudaEventCreateWithFlags(&ev[cur], cudaEventDisableTiming)) != cudaSuccess) return e;
    if (busy[cur] && (e = cudaEventSynchronize(ev[cur])) != cudaSuccess) return e;
    busy[cur] = false;
    if ((e = buf[cur].ensure(bytes)) != cudaSuccess) return e;
    *out = buf[cur].p;
    return cudaSuccess;
  }
  // call after the copies out of the current slot were enqueued on `st`
  cudaError_t commit(cudaStream_t st) {
    busy[cur] = true;
    return cudaEventRecord(ev[cur], st);
  }
  void release() {
    for (int k = 0; k < kStageSlots; k++) {
      if (ev[k]) { cudaEventSynchronize(ev[k]); cudaEventDestroy(ev[k]); ev[k] = nullptr; }
      buf[k].release();
    }
  }
};

// Ranges of the second-pass list = warps of its grid (one 16-warp CTA per SM): fixed by the device, not by the
// number of active pairs, so nothing about the second pass has to be read back by the host.
int second_pass_ranges() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return 16 * sms;
}

}  // namespace

struct vbhem_base {
  int device = 0;
  cudaStream_t own = nullptr, st = nullptr;
  BaseView B{};
  DevBuf bSb, bprior, bA, bmu, bcov, bNt;
  // centre set (per call)
  DevBuf up, cen_d, rec, tile_i, pair_rec;   // up: the upload arena of stage_centres
  // results and work lists
  DevBuf L, z, flags, fbcount, Njp, lbp, cnt, offs, segs, items, partial, itemflags, scratch, pairs_out, stats, align;
  // streamed E-step (base set in host memory): copy-in / copy-out streams and per-slice events
  cudaStream_t cin = nullptr, cout = nullptr;
  std::vector<cudaEvent_t> slice_ev;
  // device-resident EM run (vbhem_em_*): the hyper-parameters live in em.hyp (alpha | eta | epsilon | W | v | lambda)
  // and em.m; em.prior = m0[d] | W0inv[d]; em.lb = [12 * ngroups] lower-bound terms
  struct Em {
    bool active = false;
    int K = 0, maxSr = 0, ngroups = 1;
    std::vector<int32_t> Sr, group;
    double alpha0 = 0, eta0 = 0, epsilon0 = 0, lambda0 = 0, v0 = 0;
    // The run's own copies of everything it keeps between calls -- hyper-parameters (layout of stage_centres'
    // hyp block), means, state counts | group starts -- so that ordinary E-steps on the same handle (which
    // overwrite the upload arena) cannot corrupt it.  `staged`: cen_d holds the pre-pass of THIS run's last E-step
    // (what vbhem_em_lower_bound reads); any other E-step on the handle clears it.
    DevBuf hyp, ints, m, prior, lb, stats, frozen, draws, gprior, deriv;
    std::vector<double> gprior_host;   // [G0][kGroupPrior] priors of every ORIGINAL group (empty: the handle's scalars)
    bool gprior_on = false, iid_W0 = true;
    bool staged = false;
    double *h_lb = nullptr;   // pinned
  } em;
  std::vector<int> last_lists;   // tile lists as last uploaded to tile_i (skip the copy when unchanged)
  DevBuf fin_i;                  // state counts of vbhem_finalize_stats*
  int *h_totals = nullptr;   // pinned: {#active pairs, items per range, fall-backs first pass, second pass}
  bool lattice_second = false;
  bool totals_pending = false;   // the last E-step's list totals / fall-back counts are still on the device
  const int *d_totals = nullptr;
  StageRing stage;           // pinned staging of the per-call centre uploads (stage_centres)
  bool profile = false;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double timings[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int kernel_choice = 0;
  bool pair_centres = true;   // first pass: two centres per 8-state tile when they fit (vbhem_set_kernel 2 disables)
  bool split_tiles = true;    // ... including 5..7 states next to 3..1 (vbhem_set_kernel 3: halves only)
  long long counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int lastK = 0;
  const int *last_Sr_dev = nullptr;   // state counts of the last E-step's centre set, on the device
};

namespace {

struct CentreArgs {
  int K, maxSr, ngroups, tau;
  const int32_t *Sr, *group;
  // log-domain (vbhem_estep) ...
  const double *logPi, *logA, *m, *vW, *cterm, *logOmega;
  // ... or hyper-parameters (vbhem_estep_hyp)
  const double *alpha, *eta, *epsilon, *W, *v, *lambda;
  bool hyp;
  bool resident = false;   // hyp: the arrays are already in the handle (device-resident EM run)
};

// Host-resident base set and host result buffers of a streamed E-step (vbhem_estep*_stream).
struct StreamIO {
  const int32_t *Sb;
  const double *prior, *A, *mu, *cov, *Ntilde;
  double *L_host, *z_host;   // first row of this handle's bases in the host arrays
  int nslices;
  long long host_ld = 0;     // leading dimension of the host arrays (0 = the handle's N); a shard of a multi-GPU
                             // set writes its rows into the [N_total, K] arrays of the whole set
};

int check_common(vbhem_base *h, const CentreArgs &a, bool need_omega) {
  if (!h) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: base handle is NULL");
  if (a.K < 1) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: K must be >= 1");
  if (a.maxSr < 1) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: maxSr must be >= 1");
  if (a.tau < 1) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: tau must be >= 1");
  if (!a.Sr) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: Sr is NULL");
  for (int j = 0; j < a.K; j++)
    if (a.Sr[j] < 1 || a.Sr[j] > a.maxSr)
      return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: Sr[%d]=%d outside 1..maxSr=%d", j, a.Sr[j], a.maxSr);
  if (a.hyp) {
    if (!a.resident && (!a.alpha || !a.eta || !a.epsilon || !a.m || !a.W || !a.v || !a.lambda))
      return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: a centre hyper-parameter array is NULL");
    if (h->B.d > 8) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: device pre-pass supports d <= 8");
  } else {
    if (!a.logPi || !a.logA || !a.m || !a.vW || !a.cterm || (need_omega && !a.logOmega))
      return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: a centre array is NULL");
  }
  if (a.ngroups < 1) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: ngroups must be >= 1");
  if (a.group) {
    int prev = 0;
    for (int j = 0; j < a.K; j++) {
      if (a.group[j] < prev || a.group[j] > prev + 1 || a.group[j] >= a.ngroups || (j == 0 && a.group[0] != 0))
        return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: group[] must be 0..ngroups-1, contiguous and ascending");
      prev = a.group[j];
    }
    if (prev != a.ngroups - 1) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: group[] does not reach ngroups-1");
  } else if (a.ngroups != 1) {
    return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: ngroups > 1 needs group[]");
  }
  return VBHEM_OK;
}

// Device layout of the per-call centre buffers.
struct CentreDev {
  CentreView C{};
  int *gstart = nullptr;
  double *logLam = nullptr;
  const double *pre_dev = nullptr;   // pre-pass results for the caller's `pre` (hyper-parameter calls)
  size_t pre_doubles = 0;
};

// First-pass tiling (host): a centre of 5..7 states takes, largest first, the largest centre that still fits beside
// it (split tile, class = the smaller one's column count 1..3); the remaining centres of <= 4 states pair up two
// by two (halves, class 0; pairing needs >= 2 of them); everything else has a tile of its own.
// lists = big | class 0 pairs | class 1 | class 2 | class 3 (two entries per paired tile, -1 = no partner).
void tile_lists(const vbhem_base *h, const int32_t *Sr, int K, std::vector<int> &lists, int &nbig, int (&npairs)[kTileClasses]) {
  std::vector<int> big, small[5], cls[kTileClasses];
  for (int j = 0; j < K; j++) {
    if (Sr[j] <= 4 && h->pair_centres) small[Sr[j]].push_back(j);
    else big.push_back(j);
  }
  if (h->pair_centres && h->split_tiles) {
    std::vector<int> order(big);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return Sr[x] > Sr[y]; });
    std::vector<int> alone;
    for (int j : order) {
      int sl = 8 - Sr[j];
      if (sl > 3) sl = 3;
      while (sl >= 1 && small[sl].empty()) sl--;
      if (sl >= 1) {
        cls[sl].push_back(small[sl].back());
        cls[sl].push_back(j);
        small[sl].pop_back();
      } else {
        alone.push_back(j);
      }
    }
    big.swap(alone);
    std::sort(big.begin(), big.end());
  }
  for (int sl = 4; sl >= 1; sl--) cls[0].insert(cls[0].end(), small[sl].begin(), small[sl].end());
  if (cls[0].size() < 2) { big.insert(big.end(), cls[0].begin(), cls[0].end()); cls[0].clear(); std::sort(big.begin(), big.end()); }
  if (cls[0].size() % 2) cls[0].push_back(-1);
  lists = big;
  nbig = (int)big.size();
  for (int c = 0; c < kTileClasses; c++) {
    lists.insert(lists.end(), cls[c].begin(), cls[c].end());
    npairs[c] = (int)cls[c].size() / 2;
  }
}

// Upload the centre set (running the pre-pass on the device if it is given as hyper-parameters) and build the
// lattice records.  Everything the call uploads -- state counts, group starts, tile lists, means, hyper-parameters
// or pre-pass results -- is staged in ONE pinned block and travels as ONE host-to-device copy into the handle's
// upload arena (h->up); a device-resident EM run uploads nothing.  `launches` is incremented.
int stage_centres(vbhem_base *h, const CentreArgs &a, CentreDev &cd, long long &launches) {
  const int K = a.K, r = a.maxSr, d = h->B.d, dd = h->B.dd;
  const size_t nK = K, nr = (size_t)r * K, nrr = (size_t)r * r * K;
  const bool resident = a.hyp && a.resident;
  const bool lattice = h->kernel_choice == 0 && h->B.maxSb <= kS8 && r <= kS8 && d <= 3;
  CU(h->fbcount.ensure(4 * sizeof(int)));
  std::vector<int> lists;
  int nbig = 0, npairs[kTileClasses] = {0, 0, 0, 0};
  if (lattice) tile_lists(h, a.Sr, K, lists, nbig, npairs);

  int *dSr = nullptr, *dlists = nullptr;
  double *dm = nullptr, *dblk = nullptr;     // dblk: hyper-parameter block or log-domain block of the arena
  const size_t th = nK + nr + nrr + dd * nr + nr + nr;        // alpha | eta | epsilon | W | v | lambda
  const size_t tl = nK + nr + nrr + nr + dd * nr;             // logOmega | logPi | logA | cterm | vW
  if (resident) {
    // state counts, group starts, means and hyper-parameters are in the run's own buffers; the tile lists change
    // only with the centre set
    dSr = h->em.ints.as<int>();
    dm = h->em.m.as<double>();
    dblk = h->em.hyp.as<double>();
    if (lattice && !(h->tile_i.p && lists == h->last_lists)) {
      void *sp = nullptr;
      CU(h->stage.acquire(sizeof(int) * (lists.size() + 1), &sp));
      memcpy(sp, lists.data(), sizeof(int) * lists.size());
      CU(h->tile_i.ensure(sizeof(int) * (lists.size() + 1)));
      CU(cudaMemcpyAsync(h->tile_i.p, sp, sizeof(int) * lists.size(), cudaMemcpyHostToDevice, h->st));
      CU(h->stage.commit(h->st));
      h->last_lists = lists;
    }
    dlists = h->tile_i.as<int>();
  } else {
    h->em.staged = false;             // cen_d is about to hold another centre set's pre-pass
    const size_t n_int = (size_t)K + a.ngroups + 1 + lists.size();
    const size_t int_bytes = (n_int * sizeof(int) + 15) / 16 * 16;
    const size_t n_dbl = d * nr + (a.hyp ? th : tl);
    const size_t bytes = int_bytes + sizeof(double) * n_dbl;
    void *sp = nullptr;
    CU(h->stage.acquire(bytes, &sp));
    CU(h->up.ensure(bytes));
    int *s_int = static_cast<int *>(sp);
    double *s_dbl = reinterpret_cast<double *>(static_cast<char *>(sp) + int_bytes);
    memcpy(s_int, a.Sr, sizeof(int) * K);
    int *gs = s_int + K;
    for (int g = 0; g <= a.ngroups; g++) gs[g] = 0;
    if (a.group) {
      for (int j = 0; j < K; j++) gs[a.group[j] + 1] = j + 1;
    } else {
      gs[1] = K;
    }
    if (!lists.empty()) memcpy(gs + a.ngroups + 1, lists.data(), sizeof(int) * lists.size());
    double *w = s_dbl;
    memcpy(w, a.m, sizeof(double) * d * nr); w += d * nr;
    if (a.hyp) {
      memcpy(w, a.alpha, sizeof(double) * nK); w += nK;
      memcpy(w, a.eta, sizeof(double) * nr); w += nr;
      memcpy(w, a.epsilon, sizeof(double) * nrr); w += nrr;
      memcpy(w, a.W, sizeof(double) * dd * nr); w += dd * nr;
      memcpy(w, a.v, sizeof(double) * nr); w += nr;
      memcpy(w, a.lambda, sizeof(double) * nr);
    } else {
      if (a.logOmega) memcpy(w, a.logOmega, sizeof(double) * nK); else memset(w, 0, sizeof(double) * nK);
      w += nK;
      memcpy(w, a.logPi, sizeof(double) * nr); w += nr;
      memcpy(w, a.logA, sizeof(double) * nrr); w += nrr;
      memcpy(w, a.cterm, sizeof(double) * nr); w += nr;
      memcpy(w, a.vW, sizeof(double) * dd * nr);
    }
    CU(cudaMemcpyAsync(h->up.p, sp, bytes, cudaMemcpyHostToDevice, h->st));
    CU(h->stage.commit(h->st));       // the slot is free again once the copy has run
    h->last_lists.clear();            // tile_i is not what the next resident call expects
    dSr = h->up.as<int>();
    dlists = dSr + K + a.ngroups + 1;
    dm = reinterpret_cast<double *>(h->up.as<char>() + int_bytes);
    dblk = dm + d * nr;
  }
  int *dgs = dSr + K;

  cd.C = CentreView{};
  cd.C.K = K; cd.C.maxSr = r; cd.C.Sr = dSr; cd.C.group = nullptr; cd.C.m = dm;
  cd.gstart = dgs;
  cd.logLam = nullptr;
  if (lattice) {
    cd.C.rec_stride = lattice_rec_stride(d, dd);
    CU(h->rec.ensure(sizeof(double) * (size_t)cd.C.rec_stride * K));
    cd.C.rec = h->rec.as<double>();
  }
  int *counters = h->fbcount.as<int>();
  if (a.hyp) {
    // pre-pass results: logOmega[K] | logPi[rK] | logA[rrK] | logLam[rK] | cterm[rK] | vW[dd r K] in cen_d
    CU(h->cen_d.ensure(sizeof(double) * (nK + nr + nrr + nr + nr + dd * nr)));
    double *p = h->cen_d.as<double>();
    double *dOm = p; p += nK;
    double *dPi = p; p += nr;
    double *dA = p; p += nrr;
    double *dLam = p; p += nr;
    double *dCt = p; p += nr;
    double *dvW = p;
    double *q = dblk;
    double *hal = q; q += nK;
    double *het = q; q += nr;
    double *hep = q; q += nrr;
    double *hW = q; q += dd * nr;
    double *hv = q; q += nr;
    double *hl = q;
    cd.C.logPi = dPi; cd.C.logA = dA; cd.C.vW = dvW; cd.C.cterm = dCt; cd.C.logOmega = dOm;
    cd.logLam = dLam;
    // one launch: pre-pass, the centres' lattice records, and the E-step's fall-back counters cleared
    launches += launch_prepass(K, r, d, h->B.covmode, dSr, dgs, a.ngroups, hal, het, hep, hW, hv, hl, dOm, dPi, dA,
                               dLam, dCt, dvW, cd.C, lattice ? h->rec.as<double>() : nullptr, counters, h->st);
    cd.pre_dev = dOm;         // logOmega | logPi | logA | logLambdaTilde are contiguous by construction
    cd.pre_doubles = nK + nr + nrr + nr;
  } else {
    double *q = dblk;
    cd.C.logOmega = q; q += nK;
    cd.C.logPi = q; q += nr;
    cd.C.logA = q; q += nrr;
    cd.C.cterm = q; q += nr;
    cd.C.vW = q;
    if (lattice) launches += launch_build_records(cd.C, d, dd, h->rec.as<double>(), counters, h->st);
    else CU(cudaMemsetAsync(counters, 0, 4 * sizeof(int), h->st));
  }
  if (lattice) {
    int ntile_pairs = 0;
    for (int c = 0; c < kTileClasses; c++) { cd.C.npairs[c] = npairs[c]; ntile_pairs += npairs[c]; }
    cd.C.big_list = dlists;
    cd.C.nbig = nbig;
    cd.C.pair_list = dlists + nbig;
    cd.C.pair_rec = nullptr;
    h->counters[4] = nbig + ntile_pairs;
    h->counters[5] = ntile_pairs - npairs[0];
    if (ntile_pairs > 0) {
      CU(h->pair_rec.ensure(sizeof(double) * (size_t)cd.C.rec_stride * ntile_pairs));
      cd.C.pair_rec = h->pair_rec.as<double>();
      launches += launch_pair_records(cd.C, d, dd, h->pair_rec.as<double>(), h->st);
    }
  }
  return VBHEM_OK;
}

int ensure_scratch(vbhem_base *h, int maxSr, int T, long long want_warps, size_t &doubles) {
  const size_t pw = general_scratch_per_warp(h->B.maxSb, maxSr, h->B.d, h->B.dd, T);
  const size_t cap_bytes = (size_t)2 << 30;
  long long warps = want_warps < 4 ? 4 : want_warps;
  while (warps > 4 && warps * pw * sizeof(double) > cap_bytes) warps /= 2;
  doubles = (size_t)warps * pw;
  CU(h->scratch.ensure(sizeof(double) * doubles));
  return VBHEM_OK;
}

struct StatsPtr { double *Nj, *nu1, *xi, *pr, *mu, *Mu, *lb; };

StatsPtr carve_stats(double *s, int K, int r, int d, int dd, int ng) {
  (void)ng;
  StatsPtr p;
  p.Nj = s; s += K;
  p.nu1 = s; s += (size_t)r * K;
  p.xi = s; s += (size_t)r * r * K;
  p.pr = s; s += (size_t)r * K;
  p.mu = s; s += (size_t)r * d * K;
  p.Mu = s; s += (size_t)r * dd * K;
  p.lb = s;
  return p;
}

// The per-pair fall-back flags: all-zero between E-steps (the general kernel clears what it serves), so they are
// cleared only when the buffer is (re)allocated.
int ensure_flags(vbhem_base *h, long long np) {
  const void *before = h->flags.p;
  CU(h->flags.ensure((size_t)np));
  if (h->flags.p != before) CU(cudaMemsetAsync(h->flags.p, 0, h->flags.cap, h->st));
  return VBHEM_OK;
}

// L_elbo for every pair into h->L (first pass).  The fall-back counters were cleared by stage_centres.
int first_pass(vbhem_base *h, const CentreDev &cd, int T, long long &launches) {
  const long long np = (long long)h->B.N * cd.C.K;
  CU(h->L.ensure(sizeof(double) * np));
  PairWork W{nullptr, nullptr, nullptr, nullptr, 0, np};
  PairOut O{};
  O.mode = 0;
  O.L_elbo = h->L.as<double>();
  if (cd.C.rec && lattice_supported(h->B, cd.C, T, false)) {
    int rf = ensure_flags(h, np);
    if (rf) return rf;
    Fallback fb{h->flags.as<unsigned char>(), h->fbcount.as<int>()};
    launches += launch_lattice_bwd(h->B, cd.C, T, O.L_elbo, fb, h->st);
    size_t sd;
    int rc = ensure_scratch(h, cd.C.maxSr, T, 296, sd);
    if (rc) return rc;
    const int n = launch_general(h->B, cd.C, T, W, O, fb.flags, fb.count, h->scratch.as<double>(), sd, h->st);
    if (n < 0) return fail(VBHEM_ENOMEM, "general kernel scratch too small");
    launches += n;
    h->counters[1] += np;
  } else {
    size_t sd;
    int rc = ensure_scratch(h, cd.C.maxSr, T, np < 2368 ? np : 2368, sd);
    if (rc) return rc;
    const int n = launch_general(h->B, cd.C, T, W, O, nullptr, nullptr, h->scratch.as<double>(), sd, h->st);
    if (n < 0) return fail(VBHEM_ENOMEM, "general kernel scratch too small");
    launches += n;
    h->counters[2] += np;
  }
  return VBHEM_OK;
}

// Front half of a streamed E-step: the base set arrives from host memory in slices (multiples of the
// zhat block, so every fixed-order sum is the one the resident path takes) on the copy-in stream while
// earlier slices run first pass + responsibilities on the handle's stream and their L_elbo / zhat rows
// leave on the copy-out stream.  Afterwards the handle holds the whole set, as after vbhem_repack_base.
int streamed_front(vbhem_base *h, const CentreDev &cd, const CentreArgs &a, const StreamIO &io, const StatsPtr &S,
                   long long &launches) {
  const int N = h->B.N, K = a.K, T = a.tau, maxSb = h->B.maxSb, d = h->B.d, dd = h->B.dd;
  const long long np = (long long)N * K;
  if (!h->cin) CU(cudaStreamCreateWithFlags(&h->cin, cudaStreamNonBlocking));
  if (!h->cout) CU(cudaStreamCreateWithFlags(&h->cout, cudaStreamNonBlocking));
  const int zb = zhat_block_size();
  int ns = io.nslices < 1 ? 1 : (io.nslices > 64 ? 64 : io.nslices);
  // slice length: a multiple of the responsibility block, and of one full wave of the first-pass grid when the
  // set is large enough, so that slicing adds no partially filled wave
  long long unit = zb;
  {
    const long long wave = lattice_bwd_wave();
    long long a = unit, b = wave;
    while (b) { const long long t = a % b; a = b; b = t; }
    const long long l = unit / a * wave;
    if (l * 2 <= N) unit = l;
  }
  long long per64 = (((long long)N + ns - 1) / ns + unit - 1) / unit * unit;
  const int per = (int)(per64 < N ? per64 : ((N + zb - 1) / zb * zb));
  ns = (N + per - 1) / per;
  while ((int)h->slice_ev.size() < 2 * ns + 1) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->slice_ev.push_back(e);
  }
  int rc = ensure_flags(h, np);
  if (rc) return rc;
  Fallback fb{h->flags.as<unsigned char>(), h->fbcount.as<int>()};
  size_t sd;
  rc = ensure_scratch(h, cd.C.maxSr, T, 296, sd);
  if (rc) return rc;
  // earlier work on the handle's stream may still read the base buffers
  CU(cudaEventRecord(h->slice_ev[2 * ns], h->st));
  CU(cudaStreamWaitEvent(h->cin, h->slice_ev[2 * ns], 0));
  PairWork W{nullptr, nullptr, nullptr, nullptr, 0, np};
  PairOut O{};
  O.mode = 0;
  O.L_elbo = h->L.as<double>();
  for (int c = 0; c < ns; c++) {
    const int i0 = c * per, n = (N - i0 < per) ? N - i0 : per;
    const size_t o1 = (size_t)maxSb * i0, n1 = (size_t)maxSb * n;
    CU(cudaMemcpyAsync(h->bSb.as<int>() + i0, io.Sb + i0, sizeof(int) * n, cudaMemcpyHostToDevice, h->cin));
    CU(cudaMemcpyAsync(h->bprior.as<double>() + o1, io.prior + o1, sizeof(double) * n1, cudaMemcpyHostToDevice, h->cin));
    CU(cudaMemcpyAsync(h->bA.as<double>() + o1 * maxSb, io.A + o1 * maxSb, sizeof(double) * n1 * maxSb,
                       cudaMemcpyHostToDevice, h->cin));
    CU(cudaMemcpyAsync(h->bmu.as<double>() + o1 * d, io.mu + o1 * d, sizeof(double) * n1 * d, cudaMemcpyHostToDevice,
                       h->cin));
    CU(cudaMemcpyAsync(h->bcov.as<double>() + o1 * dd, io.cov + o1 * dd, sizeof(double) * n1 * dd,
                       cudaMemcpyHostToDevice, h->cin));
    if (io.Ntilde)
      CU(cudaMemcpyAsync(h->bNt.as<double>() + i0, io.Ntilde + i0, sizeof(double) * n, cudaMemcpyHostToDevice, h->cin));
    CU(cudaEventRecord(h->slice_ev[2 * c], h->cin));
    CU(cudaStreamWaitEvent(h->st, h->slice_ev[2 * c], 0));
    launches += launch_sanitize_base(h->B, h->bprior.as<double>(), h->bA.as<double>(), h->bmu.as<double>(),
                                     h->bcov.as<double>(), h->st, i0, n);
    launches += launch_lattice_bwd(h->B, cd.C, T, O.L_elbo, fb, h->st, i0, n);
    // fall-backs of this slice only (every slice used to rescan all N*K flags)
    const int ng = launch_general(h->B, cd.C, T, W, O, fb.flags, fb.count, h->scratch.as<double>(), sd, h->st, i0,
                                  i0 + n);
    if (ng < 0) return fail(VBHEM_ENOMEM, "general kernel scratch too small");
    launches += ng;
    launches += launch_zhat(N, K, a.ngroups, cd.gstart, h->L.as<double>(), cd.C.logOmega, h->B.Ntilde,
                            h->z.as<double>(), h->Njp.as<double>(), h->cnt.as<int>(), h->lbp.as<double>(), h->st,
                            i0 / zb, (n + zb - 1) / zb);
    CU(cudaEventRecord(h->slice_ev[2 * c + 1], h->st));
    CU(cudaStreamWaitEvent(h->cout, h->slice_ev[2 * c + 1], 0));
    const size_t pitch = sizeof(double) * (size_t)N;
    const size_t hpitch = sizeof(double) * (size_t)(io.host_ld ? io.host_ld : N);
    if (io.L_host)
      CU(cudaMemcpy2DAsync(io.L_host + i0, hpitch, h->L.as<double>() + i0, pitch, sizeof(double) * n, K,
                           cudaMemcpyDeviceToHost, h->cout));
    if (io.z_host)
      CU(cudaMemcpy2DAsync(io.z_host + i0, hpitch, h->z.as<double>() + i0, pitch, sizeof(double) * n, K,
                           cudaMemcpyDeviceToHost, h->cout));
  }
  h->counters[1] += np;
  return VBHEM_OK;
}

// The fused E-step on device buffers; stats_dev receives the raw sums.  With `io` the base set is
// streamed in from host memory (streamed_front) instead of being resident.
int estep_device(vbhem_base *h, const CentreArgs &a, double *stats_dev, double *pre_host,
                 const StreamIO *io = nullptr) {
  const int N = h->B.N, K = a.K, r = a.maxSr, d = h->B.d, dd = h->B.dd, T = a.tau;
  long long launches = 0;
  memset(h->counters, 0, sizeof h->counters);
  h->totals_pending = false;
  auto mark = [&](int k) { if (h->profile) cudaEventRecord(h->ev[k], h->st); };
  mark(0);
  CentreDev cd;
  int rc = stage_centres(h, a, cd, launches);
  if (rc) return rc;
  mark(1);
  const long long np = (long long)N * K;
  const int nb = zhat_blocks(N);
  CU(h->L.ensure(sizeof(double) * np));
  CU(h->z.ensure(sizeof(double) * np));
  CU(h->Njp.ensure(sizeof(double) * (size_t)nb * K));
  CU(h->lbp.ensure(sizeof(double) * (size_t)nb * 2 * a.ngroups));
  CU(h->cnt.ensure(sizeof(int) * (size_t)nb * K));
  CU(h->offs.ensure(sizeof(int) * (size_t)nb * K));
  CU(h->segs.ensure(sizeof(int) * (size_t)(2 * (K + 1) + 4)));
  StatsPtr S = carve_stats(stats_dev, K, r, d, dd, a.ngroups);
  if (io && cd.C.rec && lattice_supported(h->B, cd.C, T, false)) {
    rc = streamed_front(h, cd, a, *io, S, launches);
    if (rc) return rc;
    mark(2);
  } else {
    if (io) {   // shapes the lattice kernel does not take: upload everything, then the resident path
      rc = vbhem_repack_base(h, io->Sb, io->prior, io->A, io->mu, io->cov, io->Ntilde);
      if (rc) return rc;
    }
    rc = first_pass(h, cd, T, launches);
    if (rc) return rc;
    mark(2);
    // responsibilities + deterministic active list
    launches += launch_zhat(N, K, a.ngroups, cd.gstart, h->L.as<double>(), cd.C.logOmega, h->B.Ntilde,
                            h->z.as<double>(), h->Njp.as<double>(), h->cnt.as<int>(), h->lbp.as<double>(), h->st);
    if (io) {
      const size_t pitch = sizeof(double) * (size_t)N;
      const size_t hpitch = sizeof(double) * (size_t)(io->host_ld ? io->host_ld : N);
      if (io->L_host) CU(cudaMemcpy2DAsync(io->L_host, hpitch, h->L.p, pitch, pitch, K, cudaMemcpyDeviceToHost, h->st));
      if (io->z_host) CU(cudaMemcpy2DAsync(io->z_host, hpitch, h->z.p, pitch, pitch, K, cudaMemcpyDeviceToHost, h->st));
    }
  }
  // Deterministic list of the pairs with Z > 1e-8 and the second pass over it.  The list's length stays on the
  // device (totals): buffers are sized by their upper bounds and every grid by N, K and the device alone, so the
  // host enqueues the whole E-step without waiting for anything.
  int *seg_start = h->segs.as<int>(), *totals = seg_start + (K + 1), *segcnt = totals + 4;
  const int nranges = second_pass_ranges();
  const int per = stat_per(r, d, dd);
  if (np >= ((long long)1 << 31)) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: N * K must be below 2^31");
  CU(h->items.ensure(sizeof(PairItem) * (size_t)np));
  CU(h->partial.ensure(sizeof(double) * (size_t)per * ((size_t)nranges + K)));
  launches += launch_list_scan(N, K, a.ngroups, nranges, h->cnt.as<int>(), h->Njp.as<double>(), h->lbp.as<double>(),
                               h->offs.as<int>(), segcnt, seg_start, totals, S.Nj, S.lb, h->st);
  launches += launch_list_fill(N, K, h->z.as<double>(), h->B.Ntilde, h->offs.as<int>(), seg_start,
                               h->items.as<PairItem>(), h->st);
  mark(3);
  {
    PairWork W{h->items.as<PairItem>(), totals, h->z.as<double>(), h->B.Ntilde, nranges, np};
    PairOut O{};
    O.mode = 2;
    O.partial = h->partial.as<double>();
    O.per = per;
    if (cd.C.rec && lattice_supported(h->B, cd.C, T, true)) {
      CU(h->itemflags.ensure((size_t)np));
      Fallback fb{h->itemflags.as<unsigned char>(), h->fbcount.as<int>() + 1};   // count zeroed with the first pass's
      const int n1 = launch_lattice_full(h->B, cd.C, T, W, O, fb, h->st);
      if (n1 < 0) return fail(VBHEM_ECUDA, "lattice kernel launch configuration failed");
      launches += n1;
      size_t sd;
      rc = ensure_scratch(h, r, T, 296, sd);
      if (rc) return rc;
      const int n2 = launch_general(h->B, cd.C, T, W, O, fb.flags, fb.count, h->scratch.as<double>(), sd, h->st);
      if (n2 < 0) return fail(VBHEM_ENOMEM, "general kernel scratch too small");
      launches += n2;
      h->lattice_second = true;
    } else {
      size_t sd;
      rc = ensure_scratch(h, r, T, nranges < 2368 ? nranges : 2368, sd);
      if (rc) return rc;
      const int n2 = launch_general(h->B, cd.C, T, W, O, nullptr, nullptr, h->scratch.as<double>(), sd, h->st);
      if (n2 < 0) return fail(VBHEM_ENOMEM, "general kernel scratch too small");
      launches += n2;
      h->lattice_second = false;
    }
  }
  mark(4);
  // a centre without active pairs has an empty segment, so the reduction writes zeros for it
  launches += launch_partial_reduce(K, per, seg_start, totals, h->partial.as<double>(), S.nu1, S.xi, S.pr, S.mu, S.Mu, r,
                                    d, dd, h->st);
  mark(5);
  // the pre-pass results for the caller leave last: a copy in the middle of the chain would sit on the critical path
  if (pre_host && cd.pre_dev)
    CU(cudaMemcpyAsync(pre_host, cd.pre_dev, sizeof(double) * cd.pre_doubles, cudaMemcpyDeviceToHost, h->st));
  CU(cudaGetLastError());
  h->counters[0] = launches;
  h->lastK = K;
  h->last_Sr_dev = cd.C.Sr;
  h->d_totals = totals;
  h->totals_pending = true;
  return VBHEM_OK;
}

// The counters that live on the device after an E-step (active pairs, fall-backs): read them now.  Synchronises.
int settle_counters(vbhem_base *h) {
  if (!h->totals_pending) return VBHEM_OK;
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(h->h_totals, h->d_totals, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(h->h_totals + 2, h->fbcount.as<int>() + 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  const long long nact = h->h_totals[0];
  h->counters[3] = nact;
  h->counters[2] += h->h_totals[2] + h->h_totals[3];     // fall-backs of the second and of the first pass
  if (h->lattice_second) h->counters[1] += nact; else h->counters[2] += nact;
  h->totals_pending = false;
  return VBHEM_OK;
}

int estep_host(vbhem_base *h, const CentreArgs &a, double *L_elbo, double *zhat, double *stats, double *pre,
               StreamIO *io = nullptr) {
  if (!stats) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: stats is NULL");
  const size_t ns = vbhem_stats_size(a.K, a.maxSr, h->B.d, h->B.covmode, a.ngroups);
  CU(cudaSetDevice(h->device));
  CU(h->stats.ensure(sizeof(double) * ns));
  if (io) { io->L_host = L_elbo; io->z_host = zhat; }
  int rc = estep_device(h, a, h->stats.as<double>(), pre, io);
  if (rc) return rc;
  const size_t np = (size_t)h->B.N * a.K;
  CU(cudaMemcpyAsync(stats, h->stats.p, sizeof(double) * ns, cudaMemcpyDeviceToHost, h->st));
  if (!io) {
    if (L_elbo) CU(cudaMemcpyAsync(L_elbo, h->L.p, sizeof(double) * np, cudaMemcpyDeviceToHost, h->st));
    if (zhat) CU(cudaMemcpyAsync(zhat, h->z.p, sizeof(double) * np, cudaMemcpyDeviceToHost, h->st));
  }
  rc = settle_counters(h);          // synchronises the handle's stream
  if (rc) return rc;
  if (io && h->cout) CU(cudaStreamSynchronize(h->cout));
  return VBHEM_OK;
}

int check_stream_base(vbhem_base *h, const StreamIO &io) {
  if (!h) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: base handle is NULL");
  if (!io.Sb || !io.prior || !io.A || !io.mu || !io.cov)
    return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: a base array is NULL");
  for (int i = 0; i < h->B.N; i++)
    if (io.Sb[i] < 1 || io.Sb[i] > h->B.maxSb)
      return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: Sb[%d]=%d outside 1..maxSb=%d", i, io.Sb[i], h->B.maxSb);
  return VBHEM_OK;
}

}  // namespace

extern "C" {

int vbhem_version(void) { return 100; }
const char *vbhem_last_error(void) { return g_err.c_str(); }

int vbhem_device_count(int *count) {
  if (!count) return fail(VBHEM_EINVAL, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *count = 0; return fail(VBHEM_ENODEVICE, "no CUDA device: %s", cudaGetErrorString(e)); }
  *count = n;
  return VBHEM_OK;
}

int vbhem_pack_base(int device, int N, const int32_t *Sb, int maxSb, int d, int covmode, const double *prior,
                    const double *A, const double *mu, const double *cov, const double *Ntilde, vbhem_base_t **out) {
  if (!out) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: out is NULL");
  *out = nullptr;
  if (N < 1 || maxSb < 1 || d < 1) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: N, maxSb, d must be >= 1");
  if (covmode != VBHEM_COV_DIAG && covmode != VBHEM_COV_FULL)
    return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: covmode must be 0 (diag) or 1 (full)");
  if (!Sb || !prior || !A || !mu || !cov) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: a base array is NULL");
  for (int i = 0; i < N; i++)
    if (Sb[i] < 1 || Sb[i] > maxSb)
      return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: Sb[%d]=%d outside 1..maxSb=%d", i, Sb[i], maxSb);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(VBHEM_ENODEVICE, "no CUDA device (this library has no CPU path)");
  if (device < 0 || device >= ndev) return fail(VBHEM_EINVAL, "device %d out of range (have %d)", device, ndev);
  CU(cudaSetDevice(device));
  vbhem_base *h = new (std::nothrow) vbhem_base;
  if (!h) return fail(VBHEM_ENOMEM, "out of host memory");
  h->device = device;
  const int dd = covmode ? d * d : d;
  auto bail = [&](int code) { vbhem_free_base(h); return code; };
#define CUB(x)                                                                                                   \
  do {                                                                                                           \
    cudaError_t e_ = (x);                                                                                        \
    if (e_ != cudaSuccess)                                                                                       \
      return bail(fail(e_ == cudaErrorMemoryAllocation ? VBHEM_ENOMEM : VBHEM_ECUDA, "%s: %s", #x,               \
                       cudaGetErrorString(e_)));                                                                 \
  } while (0)
  CUB(cudaStreamCreateWithFlags(&h->own, cudaStreamNonBlocking));
  h->st = h->own;
  CUB(cudaMallocHost(&h->h_totals, 4 * sizeof(int)));
  for (int k = 0; k < 6; k++) CUB(cudaEventCreate(&h->ev[k]));
  const size_t nS = (size_t)maxSb * N;
  CUB(h->bSb.ensure(sizeof(int) * N));
  CUB(h->bprior.ensure(sizeof(double) * nS));
  CUB(h->bA.ensure(sizeof(double) * nS * maxSb));
  CUB(h->bmu.ensure(sizeof(double) * nS * d));
  CUB(h->bcov.ensure(sizeof(double) * nS * dd));
  CUB(h->bNt.ensure(sizeof(double) * N));
  CUB(cudaMemcpyAsync(h->bSb.p, Sb, sizeof(int) * N, cudaMemcpyHostToDevice, h->st));
  CUB(cudaMemcpyAsync(h->bprior.p, prior, sizeof(double) * nS, cudaMemcpyHostToDevice, h->st));
  CUB(cudaMemcpyAsync(h->bA.p, A, sizeof(double) * nS * maxSb, cudaMemcpyHostToDevice, h->st));
  CUB(cudaMemcpyAsync(h->bmu.p, mu, sizeof(double) * nS * d, cudaMemcpyHostToDevice, h->st));
  CUB(cudaMemcpyAsync(h->bcov.p, cov, sizeof(double) * nS * dd, cudaMemcpyHostToDevice, h->st));
  if (Ntilde) {
    CUB(cudaMemcpyAsync(h->bNt.p, Ntilde, sizeof(double) * N, cudaMemcpyHostToDevice, h->st));
  } else {
    std::vector<double> ones(N, 1.0);
    CUB(cudaMemcpyAsync(h->bNt.p, ones.data(), sizeof(double) * N, cudaMemcpyHostToDevice, h->st));
    CUB(cudaStreamSynchronize(h->st));
  }
  h->B.N = N; h->B.maxSb = maxSb; h->B.d = d; h->B.covmode = covmode; h->B.dd = dd;
  h->B.Sb = h->bSb.as<int>(); h->B.prior = h->bprior.as<double>(); h->B.A = h->bA.as<double>();
  h->B.mu = h->bmu.as<double>(); h->B.cov = h->bcov.as<double>(); h->B.Ntilde = h->bNt.as<double>();
  launch_sanitize_base(h->B, h->bprior.as<double>(), h->bA.as<double>(), h->bmu.as<double>(), h->bcov.as<double>(),
                       h->st);
  CUB(cudaStreamSynchronize(h->st));
  CUB(cudaGetLastError());
#undef CUB
  *out = h;
  return VBHEM_OK;
}

int vbhem_repack_base(vbhem_base_t *h, const int32_t *Sb, const double *prior, const double *A, const double *mu,
                      const double *cov, const double *Ntilde) {
  if (!h) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: base handle is NULL");
  if (!Sb || !prior || !A || !mu || !cov) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: a base array is NULL");
  const int N = h->B.N, maxSb = h->B.maxSb, d = h->B.d, dd = h->B.dd;
  for (int i = 0; i < N; i++)
    if (Sb[i] < 1 || Sb[i] > maxSb)
      return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: Sb[%d]=%d outside 1..maxSb=%d", i, Sb[i], maxSb);
  CU(cudaSetDevice(h->device));
  const size_t nS = (size_t)maxSb * N;
  CU(cudaMemcpyAsync(h->bSb.p, Sb, sizeof(int) * N, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(h->bprior.p, prior, sizeof(double) * nS, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(h->bA.p, A, sizeof(double) * nS * maxSb, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(h->bmu.p, mu, sizeof(double) * nS * d, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(h->bcov.p, cov, sizeof(double) * nS * dd, cudaMemcpyHostToDevice, h->st));
  if (Ntilde) CU(cudaMemcpyAsync(h->bNt.p, Ntilde, sizeof(double) * N, cudaMemcpyHostToDevice, h->st));
  launch_sanitize_base(h->B, h->bprior.as<double>(), h->bA.as<double>(), h->bmu.as<double>(), h->bcov.as<double>(),
                       h->st);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->st));   // page-locked caller arrays are read by DMA: keep them valid until here
  return VBHEM_OK;
}

int vbhem_free_base(vbhem_base_t *h) {
  if (!h) return VBHEM_OK;
  cudaSetDevice(h->device);
  if (h->own) cudaStreamSynchronize(h->own);
  if (h->cin) { cudaStreamSynchronize(h->cin); cudaStreamDestroy(h->cin); }
  if (h->cout) { cudaStreamSynchronize(h->cout); cudaStreamDestroy(h->cout); }
  for (cudaEvent_t e : h->slice_ev) cudaEventDestroy(e);
  h->em.m.release(); h->em.prior.release(); h->em.lb.release(); h->em.stats.release(); h->em.frozen.release();
  h->em.hyp.release(); h->em.ints.release(); h->em.draws.release(); h->em.gprior.release(); h->em.deriv.release();
  h->fin_i.release();
  if (h->em.h_lb) cudaFreeHost(h->em.h_lb);
  DevBuf *all[] = {&h->bSb, &h->bprior, &h->bA, &h->bmu, &h->bcov, &h->bNt, &h->up, &h->cen_d, &h->rec,
                   &h->L, &h->z, &h->flags, &h->fbcount, &h->Njp, &h->lbp, &h->cnt, &h->offs, &h->segs, &h->items,
                   &h->partial, &h->itemflags, &h->scratch, &h->pairs_out, &h->stats, &h->tile_i, &h->pair_rec, &h->align};
  for (DevBuf *b : all) b->release();
  if (h->h_totals) cudaFreeHost(h->h_totals);
  h->stage.release();
  for (int k = 0; k < 6; k++) if (h->ev[k]) cudaEventDestroy(h->ev[k]);
  if (h->own) cudaStreamDestroy(h->own);
  delete h;
  return VBHEM_OK;
}

int vbhem_set_stream(vbhem_base_t *h, void *cuda_stream) {
  if (!h) return fail(VBHEM_EINVAL, "base handle is NULL");
  h->st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own;
  return VBHEM_OK;
}

int vbhem_synchronize(vbhem_base_t *h) {
  if (!h) return fail(VBHEM_EINVAL, "base handle is NULL");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->st));
  return VBHEM_OK;
}

int vbhem_set_kernel(vbhem_base_t *h, int which) {
  if (!h) return fail(VBHEM_EINVAL, "base handle is NULL");
  if (which < 0 || which > 3)
    return fail(VBHEM_EINVAL, "kernel must be 0 (auto), 1 (general), 2 (auto, one centre per tile) or 3 (auto, halves only)");
  h->kernel_choice = which == 1 ? 1 : 0;
  h->pair_centres = which != 2;
  h->split_tiles = which != 3;
  return VBHEM_OK;
}

size_t vbhem_stats_size(int K, int maxSr, int d, int covmode, int ngroups) {
  const int dd = covmode ? d * d : d;
  return (size_t)K * (1 + stat_per(maxSr, d, dd)) + 2 * (size_t)ngroups;
}

int vbhem_stats_offsets(int K, int maxSr, int d, int covmode, int ngroups, size_t off[7]) {
  (void)ngroups;
  const int dd = covmode ? d * d : d;
  const size_t r = maxSr;
  off[0] = 0;
  off[1] = off[0] + K;
  off[2] = off[1] + r * K;
  off[3] = off[2] + r * r * K;
  off[4] = off[3] + r * K;
  off[5] = off[4] + r * d * K;
  off[6] = off[5] + r * dd * K;
  return VBHEM_OK;
}

int vbhem_estep_pairs(vbhem_base_t *h, int K, const int32_t *Sr, int maxSr, int tau, const double *logPi,
                      const double *logA, const double *m, const double *vW, const double *cterm, double *L_elbo,
                      double *nu1, double *xi, double *pr, double *mu_out, double *Mu_out) {
  CentreArgs a{};
  a.K = K; a.maxSr = maxSr; a.ngroups = 1; a.tau = tau; a.Sr = Sr; a.group = nullptr;
  a.logPi = logPi; a.logA = logA; a.m = m; a.vW = vW; a.cterm = cterm; a.logOmega = nullptr; a.hyp = false;
  int rc = check_common(h, a, false);
  if (rc) return rc;
  if (!L_elbo) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: L_elbo is NULL");
  const bool want = nu1 || xi || pr || mu_out || Mu_out;
  if (want && !(nu1 && xi && pr && mu_out && Mu_out))
    return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: per-pair outputs must be all NULL or all non-NULL");
  CU(cudaSetDevice(h->device));
  long long launches = 0;
  memset(h->counters, 0, sizeof h->counters);
  h->totals_pending = false;
  CentreDev cd;
  rc = stage_centres(h, a, cd, launches);
  if (rc) return rc;
  const int N = h->B.N, d = h->B.d, dd = h->B.dd, r = maxSr;
  const size_t np = (size_t)N * K;
  if (!want) {
    rc = first_pass(h, cd, tau, launches);
    if (rc) return rc;
    CU(cudaMemcpyAsync(L_elbo, h->L.p, sizeof(double) * np, cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    h->counters[0] = launches;
    return VBHEM_OK;
  }
  const int per = stat_per(r, d, dd);
  CU(h->L.ensure(sizeof(double) * np));
  CU(h->pairs_out.ensure(sizeof(double) * np * per));
  CU(cudaMemsetAsync(h->pairs_out.p, 0, sizeof(double) * np * per, h->st));
  double *po = h->pairs_out.as<double>();
  PairOut O{};
  O.mode = 1;
  O.L_elbo = h->L.as<double>();
  O.nu1 = po; po += np * r;
  O.xi = po; po += np * r * r;
  O.pr = po; po += np * r;
  O.mu = po; po += np * r * d;
  O.Mu = po;
  PairWork W{nullptr, nullptr, nullptr, nullptr, 0, (long long)np};
  if (cd.C.rec && lattice_supported(h->B, cd.C, tau, true)) {
    rc = ensure_flags(h, (long long)np);
    if (rc) return rc;
    Fallback fb{h->flags.as<unsigned char>(), h->fbcount.as<int>()};   // counters cleared by stage_centres
    const int n1 = launch_lattice_full(h->B, cd.C, tau, W, O, fb, h->st);
    if (n1 < 0) return fail(VBHEM_ECUDA, "lattice kernel launch configuration failed");
    launches += n1;
    size_t sd;
    rc = ensure_scratch(h, r, tau, 296, sd);
    if (rc) return rc;
    const int n2 = launch_general(h->B, cd.C, tau, W, O, fb.flags, fb.count, h->scratch.as<double>(), sd, h->st);
    if (n2 < 0) return fail(VBHEM_ENOMEM, "general kernel scratch too small");
    launches += n2;
    h->counters[1] = (long long)np;
  } else {
    size_t sd;
    rc = ensure_scratch(h, r, tau, np < 2368 ? (long long)np : 2368, sd);
    if (rc) return rc;
    const int n2 = launch_general(h->B, cd.C, tau, W, O, nullptr, nullptr, h->scratch.as<double>(), sd, h->st);
    if (n2 < 0) return fail(VBHEM_ENOMEM, "general kernel scratch too small");
    launches += n2;
    h->counters[2] = (long long)np;
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(L_elbo, h->L.p, sizeof(double) * np, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(nu1, O.nu1, sizeof(double) * np * r, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(xi, O.xi, sizeof(double) * np * r * r, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(pr, O.pr, sizeof(double) * np * r, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(mu_out, O.mu, sizeof(double) * np * r * d, cudaMemcpyDeviceToHost, h->st));
  CU(cudaMemcpyAsync(Mu_out, O.Mu, sizeof(double) * np * r * dd, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  if (h->fbcount.p) {
    int fbc[3] = {0, 0, 0};                       // {pending, second pass, served}: see k_general
    CU(cudaMemcpy(fbc, h->fbcount.p, sizeof fbc, cudaMemcpyDeviceToHost));
    h->counters[2] += fbc[2];
  }
  h->counters[0] = launches;
  return VBHEM_OK;
}

int vbhem_estep(vbhem_base_t *h, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups, int tau,
                const double *logPi, const double *logA, const double *m, const double *vW, const double *cterm,
                const double *logOmega, double *L_elbo, double *zhat, double *stats) {
  CentreArgs a{};
  a.K = K; a.maxSr = maxSr; a.ngroups = ngroups; a.tau = tau; a.Sr = Sr; a.group = group;
  a.logPi = logPi; a.logA = logA; a.m = m; a.vW = vW; a.cterm = cterm; a.logOmega = logOmega; a.hyp = false;
  int rc = check_common(h, a, true);
  if (rc) return rc;
  return estep_host(h, a, L_elbo, zhat, stats, nullptr);
}

int vbhem_estep_hyp(vbhem_base_t *h, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups, int tau,
                    const double *alpha, const double *eta, const double *epsilon, const double *m, const double *W,
                    const double *v, const double *lambda, double *L_elbo, double *zhat, double *stats, double *pre) {
  CentreArgs a{};
  a.K = K; a.maxSr = maxSr; a.ngroups = ngroups; a.tau = tau; a.Sr = Sr; a.group = group;
  a.alpha = alpha; a.eta = eta; a.epsilon = epsilon; a.m = m; a.W = W; a.v = v; a.lambda = lambda; a.hyp = true;
  int rc = check_common(h, a, true);
  if (rc) return rc;
  return estep_host(h, a, L_elbo, zhat, stats, pre);
}

int vbhem_estep_stream(vbhem_base_t *h, const int32_t *Sb, const double *prior, const double *A, const double *mu,
                       const double *cov, const double *Ntilde, int nslices, int K, const int32_t *Sr, int maxSr,
                       const int32_t *group, int ngroups, int tau, const double *logPi, const double *logA,
                       const double *m, const double *vW, const double *cterm, const double *logOmega, double *L_elbo,
                       double *zhat, double *stats) {
  CentreArgs a{};
  a.K = K; a.maxSr = maxSr; a.ngroups = ngroups; a.tau = tau; a.Sr = Sr; a.group = group;
  a.logPi = logPi; a.logA = logA; a.m = m; a.vW = vW; a.cterm = cterm; a.logOmega = logOmega; a.hyp = false;
  int rc = check_common(h, a, true);
  if (rc) return rc;
  StreamIO io{Sb, prior, A, mu, cov, Ntilde, nullptr, nullptr, nslices};
  rc = check_stream_base(h, io);
  if (rc) return rc;
  return estep_host(h, a, L_elbo, zhat, stats, nullptr, &io);
}

int vbhem_estep_hyp_stream(vbhem_base_t *h, const int32_t *Sb, const double *prior, const double *A, const double *mu,
                           const double *cov, const double *Ntilde, int nslices, int K, const int32_t *Sr, int maxSr,
                           const int32_t *group, int ngroups, int tau, const double *alpha, const double *eta,
                           const double *epsilon, const double *m, const double *W, const double *v,
                           const double *lambda, double *L_elbo, double *zhat, double *stats, double *pre) {
  CentreArgs a{};
  a.K = K; a.maxSr = maxSr; a.ngroups = ngroups; a.tau = tau; a.Sr = Sr; a.group = group;
  a.alpha = alpha; a.eta = eta; a.epsilon = epsilon; a.m = m; a.W = W; a.v = v; a.lambda = lambda; a.hyp = true;
  int rc = check_common(h, a, true);
  if (rc) return rc;
  StreamIO io{Sb, prior, A, mu, cov, Ntilde, nullptr, nullptr, nslices};
  rc = check_stream_base(h, io);
  if (rc) return rc;
  return estep_host(h, a, L_elbo, zhat, stats, pre, &io);
}

int vbhem_estep_dev(vbhem_base_t *h, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups, int tau,
                    const double *logPi, const double *logA, const double *m, const double *vW, const double *cterm,
                    const double *logOmega, double *stats_dev) {
  CentreArgs a{};
  a.K = K; a.maxSr = maxSr; a.ngroups = ngroups; a.tau = tau; a.Sr = Sr; a.group = group;
  a.logPi = logPi; a.logA = logA; a.m = m; a.vW = vW; a.cterm = cterm; a.logOmega = logOmega; a.hyp = false;
  int rc = check_common(h, a, true);
  if (rc) return rc;
  if (!stats_dev) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: stats_dev is NULL");
  CU(cudaSetDevice(h->device));
  return estep_device(h, a, stats_dev, nullptr);
}

int vbhem_estep_hyp_dev(vbhem_base_t *h, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                        int tau, const double *alpha, const double *eta, const double *epsilon, const double *m,
                        const double *W, const double *v, const double *lambda, double *stats_dev, double *pre) {
  CentreArgs a{};
  a.K = K; a.maxSr = maxSr; a.ngroups = ngroups; a.tau = tau; a.Sr = Sr; a.group = group;
  a.alpha = alpha; a.eta = eta; a.epsilon = epsilon; a.m = m; a.W = W; a.v = v; a.lambda = lambda; a.hyp = true;
  int rc = check_common(h, a, true);
  if (rc) return rc;
  if (!stats_dev) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: stats_dev is NULL");
  CU(cudaSetDevice(h->device));
  return estep_device(h, a, stats_dev, pre);
}

// ---- device-resident EM run -------------------------------------------------
namespace {
struct HypPtrs { double *alpha, *eta, *epsilon, *W, *v, *lambda; };
HypPtrs carve_hyp(vbhem_base *h, int K, int r) {
  const size_t nK = K, nr = (size_t)r * K, nrr = (size_t)r * r * K, dd = h->B.dd;
  double *q = h->em.hyp.as<double>();
  HypPtrs p;
  p.alpha = q; q += nK;
  p.eta = q; q += nr;
  p.epsilon = q; q += nrr;
  p.W = q; q += dd * nr;
  p.v = q; q += nr;
  p.lambda = q;
  return p;
}
// The centre set of a device-resident EM run: shapes into h->em, hyper-parameters into HBM.  The prior stays.
int em_upload_state(vbhem_base *h, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                    const double *alpha, const double *eta, const double *epsilon, const double *m, const double *W,
                    const double *v, const double *lambda) {
  auto &em = h->em;
  const int d = h->B.d, dd = h->B.dd;
  em.K = K; em.maxSr = maxSr; em.ngroups = ngroups;
  if (Sr != em.Sr.data()) em.Sr.assign(Sr, Sr + K);
  if (!group) em.group.clear();
  else if (group != em.group.data()) em.group.assign(group, group + K);
  const size_t nK = K, nr = (size_t)maxSr * K, nrr = (size_t)maxSr * maxSr * K;
  CU(em.hyp.ensure(sizeof(double) * (nK + nr + nrr + dd * nr + nr + nr)));
  CU(em.ints.ensure(sizeof(int) * (size_t)(K + ngroups + 1)));
  {
    std::vector<int> iv((size_t)K + ngroups + 1, 0);
    for (int j = 0; j < K; j++) iv[j] = Sr[j];
    if (group) { for (int j = 0; j < K; j++) iv[(size_t)K + group[j] + 1] = j + 1; }
    else iv[(size_t)K + 1] = K;
    CU(cudaMemcpyAsync(em.ints.p, iv.data(), sizeof(int) * iv.size(), cudaMemcpyHostToDevice, h->st));  // pageable: staged before return
  }
  em.staged = false;
  CU(em.m.ensure(sizeof(double) * d * nr));
  CU(em.lb.ensure(sizeof(double) * 12 * ngroups));
  CU(em.stats.ensure(sizeof(double) * vbhem_stats_size(K, maxSr, d, h->B.covmode, ngroups)));
  HypPtrs p = carve_hyp(h, K, maxSr);
  if (!alpha) return VBHEM_OK;                 // shapes only: the caller fills the hyper-parameters on the device
  CU(cudaMemcpyAsync(p.alpha, alpha, sizeof(double) * nK, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(p.eta, eta, sizeof(double) * nr, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(p.epsilon, epsilon, sizeof(double) * nrr, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(p.W, W, sizeof(double) * dd * nr, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(p.v, v, sizeof(double) * nr, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(p.lambda, lambda, sizeof(double) * nr, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(em.m.p, m, sizeof(double) * d * nr, cudaMemcpyHostToDevice, h->st));
  CU(cudaStreamSynchronize(h->st));
  return VBHEM_OK;
}

// Host copy of a centre set's hyper-parameters, C-ABI layouts (one contiguous block per centre in every array).
struct HypHost {
  std::vector<double> alpha, eta, eps, m, W, v, lam;
  void resize(int K, int r, int d, int dd) {
    alpha.resize(K); eta.resize((size_t)r * K); eps.resize((size_t)r * r * K); m.resize((size_t)d * r * K);
    W.resize((size_t)dd * r * K); v.resize((size_t)r * K); lam.resize((size_t)r * K);
  }
  // centre jd of this <- centre js of src
  void copy_centre(int jd, const HypHost &src, int js, int r, int d, int dd) {
    auto cp = [&](std::vector<double> &a, const std::vector<double> &b, size_t w) {
      memcpy(a.data() + w * jd, b.data() + w * js, sizeof(double) * w);
    };
    cp(alpha, src.alpha, 1); cp(eta, src.eta, r); cp(eps, src.eps, (size_t)r * r); cp(m, src.m, (size_t)d * r);
    cp(W, src.W, (size_t)dd * r); cp(v, src.v, r); cp(lam, src.lam, r);
  }
};

CentreArgs em_args(vbhem_base *h, int tau) {
  CentreArgs a{};
  a.K = h->em.K; a.maxSr = h->em.maxSr; a.ngroups = h->em.ngroups; a.tau = tau;
  a.Sr = h->em.Sr.data(); a.group = h->em.group.empty() ? nullptr : h->em.group.data();
  a.hyp = true; a.resident = true;
  return a;
}
}  // namespace

int vbhem_em_begin(vbhem_base_t *h, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                   double alpha0, double eta0, double epsilon0, double lambda0, double v0, const double *m0,
                   const double *W0, int nW0, const double *alpha, const double *eta, const double *epsilon,
                   const double *m, const double *W, const double *v, const double *lambda) {
  CentreArgs a{};
  a.K = K; a.maxSr = maxSr; a.ngroups = ngroups; a.tau = 1; a.Sr = Sr; a.group = group;
  a.alpha = alpha; a.eta = eta; a.epsilon = epsilon; a.m = m; a.W = W; a.v = v; a.lambda = lambda; a.hyp = true;
  int rc = check_common(h, a, true);
  if (rc) return rc;
  const int d = h->B.d;
  if (!m0 || !W0 || (nW0 != 1 && nW0 != d))
    return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: m0[d] and W0[1 or d] are required");
  if (!(v0 > d - 1)) return fail(VBHEM_EINVAL, "v0 not large enough");   // vbhem_mstep_component.m:24-26
  CU(cudaSetDevice(h->device));
  auto &em = h->em;
  em.alpha0 = alpha0; em.eta0 = eta0; em.epsilon0 = epsilon0; em.lambda0 = lambda0; em.v0 = v0;
  CU(em.prior.ensure(sizeof(double) * 2 * d));
  if (!em.h_lb) CU(cudaMallocHost(&em.h_lb, sizeof(double) * 12 * 4096));
  if (ngroups > 4096) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: at most 4096 groups in a device EM run");
  std::vector<double> pr(2 * d);
  for (int q = 0; q < d; q++) { pr[q] = m0[q]; pr[d + q] = 1.0 / W0[nW0 == 1 ? 0 : q]; }
  CU(cudaMemcpyAsync(em.prior.p, pr.data(), sizeof(double) * 2 * d, cudaMemcpyHostToDevice, h->st));
  rc = em_upload_state(h, K, Sr, maxSr, group, ngroups, alpha, eta, epsilon, m, W, v, lambda);   // synchronises
  if (rc) return rc;
  em.active = true;
  em.gprior_on = false;
  em.gprior_host.clear();
  em.iid_W0 = nW0 == 1;
  return VBHEM_OK;
}

// vbhem_em_begin with the `baseem` initialisation (/root/reference/src/vbhem/vbhemhmm_init.m:58-100) done on the
// device from the resident base set: the caller draws the random indices (the reference's vbhemopt.prerand, :66-74)
// and the kernel gathers / inverts.
int vbhem_em_begin_baseem(vbhem_base_t *h, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                          double alpha0, double eta0, double epsilon0, double lambda0, double v0, const double *m0,
                          const double *W0, int nW0, const int32_t *randb, const int32_t *randg, const double *omega,
                          double Nv) {
  CentreArgs a{};
  a.K = K; a.maxSr = maxSr; a.ngroups = ngroups; a.tau = 1; a.Sr = Sr; a.group = group; a.hyp = true; a.resident = true;
  int rc = check_common(h, a, true);
  if (rc) return rc;
  const int d = h->B.d;
  if (!m0 || !W0 || (nW0 != 1 && nW0 != d))
    return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: m0[d] and W0[1 or d] are required");
  if (!(v0 > d - 1)) return fail(VBHEM_EINVAL, "v0 not large enough");
  if (!randb || !randg || !omega) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: randb, randg, omega are required");
  if (d > 8) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: device initialisation supports d <= 8");
  if (ngroups > 4096) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: at most 4096 groups in a device EM run");
  for (int j = 0; j < K; j++)
    for (int k = 0; k < Sr[j]; k++) {
      const int b = randb[k + (size_t)maxSr * j];
      if (b < 0 || b >= h->B.N) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: randb(%d,%d) = %d outside the base set", k, j, b);
      if (randg[k + (size_t)maxSr * j] < 0 || randg[k + (size_t)maxSr * j] >= h->B.maxSb)
        return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: randg(%d,%d) outside 0..maxSb-1", k, j);
    }
  CU(cudaSetDevice(h->device));
  auto &em = h->em;
  em.alpha0 = alpha0; em.eta0 = eta0; em.epsilon0 = epsilon0; em.lambda0 = lambda0; em.v0 = v0;
  CU(em.prior.ensure(sizeof(double) * 2 * d));
  if (!em.h_lb) CU(cudaMallocHost(&em.h_lb, sizeof(double) * 12 * 4096));
  std::vector<double> pr(2 * d);
  for (int q = 0; q < d; q++) { pr[q] = m0[q]; pr[d + q] = 1.0 / W0[nW0 == 1 ? 0 : q]; }
  CU(cudaMemcpyAsync(em.prior.p, pr.data(), sizeof(double) * 2 * d, cudaMemcpyHostToDevice, h->st));
  rc = em_upload_state(h, K, Sr, maxSr, group, ngroups, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc) return rc;
  const size_t nr = (size_t)maxSr * K;
  CU(em.draws.ensure(sizeof(int) * 2 * nr + sizeof(double) * K + 16));
  int *db = em.draws.as<int>(), *dg = db + nr;
  double *dom = reinterpret_cast<double *>(em.draws.as<char>() + ((sizeof(int) * 2 * nr + 15) / 16) * 16);
  CU(cudaMemcpyAsync(db, randb, sizeof(int) * nr, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(dg, randg, sizeof(int) * nr, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(dom, omega, sizeof(double) * K, cudaMemcpyHostToDevice, h->st));
  HypPtrs p = carve_hyp(h, K, maxSr);
  const int *dSr = em.ints.as<int>();
  launch_init_baseem(h->B, K, maxSr, dSr, dSr + K, ngroups, alpha0, eta0, epsilon0, lambda0, v0, Nv, db, dg, dom, p.alpha,
                     p.eta, p.epsilon, em.m.as<double>(), p.W, p.v, p.lambda, h->st);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->st));            // the caller's index arrays may change next
  em.active = true;
  em.gprior_on = false;
  em.gprior_host.clear();
  em.iid_W0 = nW0 == 1;
  return VBHEM_OK;
}

int vbhem_em_estep(vbhem_base_t *h, int tau, double *stats_dev) {
  if (!h || !h->em.active) return fail(VBHEM_EINVAL, "vbhem_em_begin has not been called on this handle");
  CentreArgs a = em_args(h, tau);
  int rc = check_common(h, a, true);
  if (rc) return rc;
  CU(cudaSetDevice(h->device));
  rc = estep_device(h, a, stats_dev ? stats_dev : h->em.stats.as<double>(), nullptr);
  if (rc == VBHEM_OK) h->em.staged = true;
  return rc;
}

namespace {
// Views of the EM state and of a statistics buffer, shared by the lower bound and the M-step.
struct EmCtx { EmView E; StatsPtr S; HypPtrs p; const double *dOm, *dPi, *dA, *dLam; };
EmCtx em_ctx(vbhem_base *h, double *st) {
  auto &em = h->em;
  const int K = em.K, r = em.maxSr, d = h->B.d, dd = h->B.dd, ng = em.ngroups;
  EmCtx c;
  c.S = carve_stats(st, K, r, d, dd, ng);
  const int *dSr = em.ints.as<int>(), *dgs = dSr + K;
  c.p = carve_hyp(h, K, r);
  EmView &E = c.E;
  E = EmView{};
  E.K = K; E.maxSr = r; E.d = d; E.covmode = h->B.covmode; E.ngroups = ng; E.Sr = dSr; E.gstart = dgs;
  E.alpha0 = em.alpha0; E.eta0 = em.eta0; E.epsilon0 = em.epsilon0; E.lambda0 = em.lambda0; E.v0 = em.v0;
  E.m0 = em.prior.as<double>(); E.W0inv = E.m0 + d;
  E.gprior = em.gprior_on ? em.gprior.as<double>() : nullptr;
  E.alpha = c.p.alpha; E.eta = c.p.eta; E.epsilon = c.p.epsilon; E.m = em.m.as<double>(); E.W = c.p.W;
  E.v = c.p.v; E.lambda = c.p.lambda;
  // pre-pass results of the E-step that produced these statistics (layout of stage_centres)
  const size_t nK = K, nr = (size_t)r * K, nrr = (size_t)r * r * K;
  c.dOm = h->cen_d.as<double>(); c.dPi = c.dOm + nK; c.dA = c.dPi + nr; c.dLam = c.dA + nrr;
  return c;
}
}  // namespace

namespace {
// Normalise the statistics (vbhem_compute_Statistics.m:57-82) and evaluate the bound on the device; with `read`
// the 12 * ngroups terms come back to the host (synchronises), without it everything is only enqueued.
int em_lower_bound(vbhem_base *h, double *stats_dev, double *lb_out, bool read) {
  if (!h || !h->em.active) return fail(VBHEM_EINVAL, "vbhem_em_begin has not been called on this handle");
  if (!h->em.staged)
    return fail(VBHEM_EINVAL, "vbhem_em_lower_bound needs the pre-pass of this run's own E-step: call vbhem_em_estep "
                              "first (another E-step ran on the handle since)");
  CU(cudaSetDevice(h->device));
  auto &em = h->em;
  EmCtx c = em_ctx(h, stats_dev ? stats_dev : em.stats.as<double>());
  launch_finalize(em.K, em.maxSr, h->B.d, h->B.covmode, c.E.Sr, c.S.Nj, c.S.xi, c.S.pr, c.S.mu, c.S.Mu, h->st);
  launch_lower_bound(c.E, c.S.Nj, c.S.lb, c.dOm, c.dPi, c.dA, c.dLam, em.lb.as<double>(), h->st);
  CU(cudaGetLastError());
  if (!read) return VBHEM_OK;
  CU(cudaMemcpyAsync(em.h_lb, em.lb.p, sizeof(double) * 12 * em.ngroups, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  if (lb_out) memcpy(lb_out, em.h_lb, sizeof(double) * 12 * em.ngroups);
  return VBHEM_OK;
}
}  // namespace

int vbhem_em_lower_bound(vbhem_base_t *h, double *stats_dev, double *lb_out) {
  return em_lower_bound(h, stats_dev, lb_out, true);
}

namespace {
int em_mstep_masked(vbhem_base *h, double *stats_dev, const unsigned char *frozen_host) {
  if (!h || !h->em.active) return fail(VBHEM_EINVAL, "vbhem_em_begin has not been called on this handle");
  CU(cudaSetDevice(h->device));
  auto &em = h->em;
  EmCtx c = em_ctx(h, stats_dev ? stats_dev : em.stats.as<double>());
  if (frozen_host) {
    CU(em.frozen.ensure((size_t)em.ngroups));
    CU(cudaMemcpyAsync(em.frozen.p, frozen_host, (size_t)em.ngroups, cudaMemcpyHostToDevice, h->st));
    c.E.frozen = em.frozen.as<unsigned char>();
  }
  launch_mstep(c.E, c.S.Nj, c.S.nu1, c.S.xi, c.S.pr, c.S.mu, c.S.Mu, c.p.alpha, c.p.eta, c.p.epsilon,
               em.m.as<double>(), c.p.W, c.p.v, c.p.lambda, h->st);
  if (frozen_host) CU(cudaStreamSynchronize(h->st));     // the caller's mask buffer may change next
  CU(cudaGetLastError());
  return VBHEM_OK;
}
}  // namespace

int vbhem_em_mstep_masked(vbhem_base_t *h, double *stats_dev, const unsigned char *frozen) {
  return em_mstep_masked(h, stats_dev, frozen);
}

int vbhem_em_mstep(vbhem_base_t *h, double *stats_dev) {
  if (!h || !h->em.active) return fail(VBHEM_EINVAL, "vbhem_em_begin has not been called on this handle");
  CU(cudaSetDevice(h->device));
  auto &em = h->em;
  EmCtx c = em_ctx(h, stats_dev ? stats_dev : em.stats.as<double>());
  launch_mstep(c.E, c.S.Nj, c.S.nu1, c.S.xi, c.S.pr, c.S.mu, c.S.Mu, c.p.alpha, c.p.eta, c.p.epsilon,
               em.m.as<double>(), c.p.W, c.p.v, c.p.lambda, h->st);
  CU(cudaGetLastError());
  return VBHEM_OK;
}

namespace {
// rows `groups[0..n)` of the host copy of the per-group priors -> the device (current group order)
int em_upload_gpriors(vbhem_base *h, const int *groups, int n) {
  auto &em = h->em;
  if (em.gprior_host.empty()) { em.gprior_on = false; return VBHEM_OK; }
  std::vector<double> sel((size_t)kGroupPrior * n);
  for (int g = 0; g < n; g++)
    memcpy(sel.data() + (size_t)kGroupPrior * g, em.gprior_host.data() + (size_t)kGroupPrior * groups[g],
           sizeof(double) * kGroupPrior);
  CU(cudaSetDevice(h->device));
  CU(em.gprior.ensure(sizeof(double) * sel.size()));
  CU(cudaMemcpyAsync(em.gprior.p, sel.data(), sizeof(double) * sel.size(), cudaMemcpyHostToDevice, h->st));
  CU(cudaStreamSynchronize(h->st));
  em.gprior_on = true;
  return VBHEM_OK;
}

// raw derivatives of the bound at the current (pre-M-step) state, per current group: out[g][kDerivStride] (HOST)
int em_lb_deriv(vbhem_base *h, double *out) {
  if (!h || !h->em.active) return fail(VBHEM_EINVAL, "vbhem_em_begin has not been called on this handle");
  if (!h->em.staged) return fail(VBHEM_EINVAL, "vbhem_em_lb_deriv needs this run's own E-step: call vbhem_em_estep first");
  CU(cudaSetDevice(h->device));
  auto &em = h->em;
  EmCtx c = em_ctx(h, em.stats.as<double>());
  CU(em.deriv.ensure(sizeof(double) * kDerivStride * em.ngroups));
  launch_lb_deriv(c.E, em.iid_W0, c.dOm, c.dPi, c.dA, c.dLam, em.deriv.as<double>(), h->st);
  CU(cudaMemcpyAsync(out, em.deriv.p, sizeof(double) * kDerivStride * em.ngroups, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  CU(cudaGetLastError());
  return VBHEM_OK;
}
}  // namespace

int vbhem_em_set_group_priors(vbhem_base_t *h, const double *alpha0, const double *eta0, const double *epsilon0,
                              const double *lambda0, const double *v0, const double *m0, const double *W0, int nW0) {
  if (!h || !h->em.active) return fail(VBHEM_EINVAL, "vbhem_em_begin has not been called on this handle");
  if (!alpha0 || !eta0 || !epsilon0 || !lambda0 || !v0 || !m0 || !W0)
    return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: a prior array is NULL");
  const int d = h->B.d, G = h->em.ngroups;
  if (nW0 != 1 && nW0 != d) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: W0 must have 1 or d entries per group");
  if (d > 8) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: per-group priors support d <= 8");
  auto &em = h->em;
  em.gprior_host.assign((size_t)kGroupPrior * G, 0.0);
  std::vector<int> all(G);
  for (int g = 0; g < G; g++) {
    double *p = em.gprior_host.data() + (size_t)kGroupPrior * g;
    if (!(v0[g] > d - 1)) return fail(VBHEM_EINVAL, "v0 not large enough");
    p[0] = alpha0[g]; p[1] = eta0[g]; p[2] = epsilon0[g]; p[3] = lambda0[g]; p[4] = v0[g];
    for (int q = 0; q < d; q++) { p[5 + q] = m0[q + (size_t)d * g]; p[13 + q] = 1.0 / W0[nW0 == 1 ? (size_t)g : q + (size_t)d * g]; }
    all[g] = g;
  }
  em.iid_W0 = nW0 == 1;
  return em_upload_gpriors(h, all.data(), G);
}

int vbhem_em_lb_deriv(vbhem_base_t *h, double *out) {
  if (!out) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: out is NULL");
  return em_lb_deriv(h, out);
}

int vbhem_em_fetch(vbhem_base_t *h, double *alpha, double *eta, double *epsilon, double *m, double *W, double *v,
                   double *lambda) {
  if (!h || !h->em.active) return fail(VBHEM_EINVAL, "vbhem_em_begin has not been called on this handle");
  CU(cudaSetDevice(h->device));
  const int K = h->em.K, r = h->em.maxSr, d = h->B.d, dd = h->B.dd;
  const size_t nK = K, nr = (size_t)r * K, nrr = (size_t)r * r * K;
  HypPtrs p = carve_hyp(h, K, r);
  if (alpha) CU(cudaMemcpyAsync(alpha, p.alpha, sizeof(double) * nK, cudaMemcpyDeviceToHost, h->st));
  if (eta) CU(cudaMemcpyAsync(eta, p.eta, sizeof(double) * nr, cudaMemcpyDeviceToHost, h->st));
  if (epsilon) CU(cudaMemcpyAsync(epsilon, p.epsilon, sizeof(double) * nrr, cudaMemcpyDeviceToHost, h->st));
  if (m) CU(cudaMemcpyAsync(m, h->em.m.p, sizeof(double) * d * nr, cudaMemcpyDeviceToHost, h->st));
  if (W) CU(cudaMemcpyAsync(W, p.W, sizeof(double) * dd * nr, cudaMemcpyDeviceToHost, h->st));
  if (v) CU(cudaMemcpyAsync(v, p.v, sizeof(double) * nr, cudaMemcpyDeviceToHost, h->st));
  if (lambda) CU(cudaMemcpyAsync(lambda, p.lambda, sizeof(double) * nr, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  return VBHEM_OK;
}

// ---- the EM loops, written once over an EmDriver: one handle (vbhem_em_run*) or the shards of a multi-GPU set
// (vbhem_multi_em_run*, vbhem_multi.inl), where every operation fans out to all shards and the E-step is followed
// by the all-reduce of the raw statistics.
namespace {
struct EmDriver {
  virtual ~EmDriver() {}
  virtual vbhem_base *first() = 0;                        // shapes, priors (identical on every shard)
  // E-step from the resident hyper-parameters (+ all-reduce) + lower bound: lb[12 * ngroups] on the host
  virtual int estep_lb(int tau, double *lb) = 0;
  virtual int mstep_masked(const unsigned char *skip) = 0;   // skip == nullptr: every group
  virtual int fetch(double *alpha, double *eta, double *epsilon, double *m, double *W, double *v, double *lambda) = 0;
  virtual int upload_state(int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups, const double *alpha,
                           const double *eta, const double *epsilon, const double *m, const double *W, const double *v,
                           const double *lambda) = 0;
  virtual int sync() = 0;
  virtual int lb_deriv(double *out) = 0;                       // [current groups][kDerivStride]
  virtual int select_priors(const int *groups, int n) = 0;     // per-group priors of these ORIGINAL groups, in order
};

struct SingleDriver : EmDriver {
  vbhem_base *h;
  explicit SingleDriver(vbhem_base *h_) : h(h_) {}
  vbhem_base *first() override { return h; }
  int estep_lb(int tau, double *lb) override {
    int rc = vbhem_em_estep(h, tau, nullptr);
    if (rc) return rc;
    return vbhem_em_lower_bound(h, nullptr, lb);          // of the hyper-parameters this E-step used
  }
  int mstep_masked(const unsigned char *skip) override { return em_mstep_masked(h, nullptr, skip); }
  int fetch(double *alpha, double *eta, double *epsilon, double *m, double *W, double *v, double *lambda) override {
    return vbhem_em_fetch(h, alpha, eta, epsilon, m, W, v, lambda);
  }
  int upload_state(int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups, const double *alpha,
                   const double *eta, const double *epsilon, const double *m, const double *W, const double *v,
                   const double *lambda) override {
    return em_upload_state(h, K, Sr, maxSr, group, ngroups, alpha, eta, epsilon, m, W, v, lambda);
  }
  int sync() override {
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->st));
    return VBHEM_OK;
  }
  int lb_deriv(double *out) override { return em_lb_deriv(h, out); }
  int select_priors(const int *groups, int n) override { return em_upload_gpriors(h, groups, n); }
};

// vbhem_h3m_c_step_fc.m:115-433, one H3M: E-step, lower bound, relative-change test from the third pass on,
// M-step (also after the breaking pass), NaN guard.
int em_run_one(EmDriver &D, int tau, int max_iter, double min_diff, double *LogLs, int *iters, int *unstable) {
  vbhem_base *h = D.first();
  if (!h || !h->em.active) return fail(VBHEM_EINVAL, "vbhem_em_begin has not been called on this handle");
  if (h->em.ngroups != 1) return fail(VBHEM_EINVAL, "vbhem_em_run drives one H3M; batch several with vbhem_em_run_groups");
  if (!LogLs || !iters) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: LogLs / iters is NULL");
  double lastL = -1.79769313486231570e308, lb[12];
  int it = 0, n = 0, bad = 0;
  while (true) {
    int rc = D.estep_lb(tau, lb);
    if (rc) return rc;
    double L = lb[10];
    bool do_break = false;
    if (it > 1 && fabs((L - lastL) / lastL) <= min_diff) do_break = true;
    if (it == max_iter) do_break = true;
    if (L != L) { do_break = true; bad = 1; L = -INFINITY; }     // vbhem_h3m_c_step_fc.m:338-354
    LogLs[n++] = L;
    if (do_break && bad) break;                                  // keep the last stable hyper-parameters
    rc = D.mstep_masked(nullptr);
    if (rc) return rc;
    it++;
    lastL = L;
    if (do_break) break;
  }
  int rc = D.sync();
  if (rc) return rc;
  *iters = it;
  if (unstable) *unstable = bad;
  return VBHEM_OK;
}

// The same loop for several H3Ms batched in one handle (soft-max groups of vbhem_em_begin): every group
// follows the control flow of vbhem_h3m_c_step_fc.m:311-354 on its own; a group that has stopped keeps its
// hyper-parameters (masked M-step) while the others go on.  Once a quarter of the centres in the handle belong
// to stopped groups the centre set is compacted to the running groups (the stopped ones wait in a host copy), so
// the E-step's cost follows the work that is left; at the end the handle holds all groups again, in their
// original order.  LogLs is [ngroups][ld], iters / unstable [ngroups].
// dLL (may be null): [ngroups][kDerivStride], the derivatives of every group's bound with respect to the prior
// hyper-parameters at its breaking pass (vbhem_h3m_c_step_fc.m:357-368: computed before that pass's M-step).
int em_run_groups(EmDriver &D, int tau, int max_iter, double min_diff, double *LogLs, int ld, int *iters,
                  int *unstable, int *ncompact_out, double *dLL = nullptr) {
  vbhem_base *h = D.first();
  if (!h || !h->em.active) return fail(VBHEM_EINVAL, "vbhem_em_begin has not been called on this handle");
  if (!LogLs || !iters || ld < max_iter + 2) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: LogLs[ngroups][>= max_iter + 2], iters required");
  auto &em = h->em;
  const int G = em.ngroups, K0 = em.K, r = em.maxSr, d = h->B.d, dd = h->B.dd;
  const std::vector<int32_t> Sr0 = em.Sr;
  std::vector<int32_t> grp0 = em.group;
  const bool had_groups = G > 1 || !em.group.empty();
  if (grp0.empty()) grp0.assign(K0, 0);
  std::vector<double> lb(12 * (size_t)G), lastL(G, -1.79769313486231570e308);
  std::vector<int> it(G, 0), n(G, 0), bad(G, 0);
  std::vector<unsigned char> done(G, 0), skip(G, 0), stop(G, 0);
  std::vector<int> cur_groups(G), cmap(K0);          // original group / centre of every current one
  for (int g = 0; g < G; g++) cur_groups[g] = g;
  for (int j = 0; j < K0; j++) cmap[j] = j;
  HypHost full, cur;
  std::vector<double> dbuf;
  bool compacted = false;
  int ncompact = 0;
  // current (possibly compacted) device state -> `full`, at the centres' original positions
  auto save_current = [&]() -> int {
    const int Kc = (int)cmap.size();
    cur.resize(Kc, r, d, dd);
    int rc = D.fetch(cur.alpha.data(), cur.eta.data(), cur.eps.data(), cur.m.data(), cur.W.data(), cur.v.data(),
                     cur.lam.data());
    if (rc) return rc;
    if (full.alpha.empty()) full.resize(K0, r, d, dd);
    for (int jc = 0; jc < Kc; jc++) full.copy_centre(cmap[jc], cur, jc, r, d, dd);
    return VBHEM_OK;
  };
  int left = G;
  while (left > 0) {
    const int Gc = (int)cur_groups.size();
    int rc = D.estep_lb(tau, lb.data());
    if (rc) return rc;
    for (int gc = 0; gc < Gc; gc++) {
      const int g = cur_groups[gc];
      stop[gc] = 0;
      skip[gc] = done[g];
      if (done[g]) continue;
      double L = lb[12 * (size_t)gc + 10];
      bool do_break = false;
      if (it[g] > 1 && fabs((L - lastL[g]) / lastL[g]) <= min_diff) do_break = true;
      if (it[g] == max_iter) do_break = true;
      if (L != L) { do_break = true; bad[g] = 1; L = -INFINITY; }
      LogLs[(size_t)ld * g + n[g]++] = L;
      if (do_break && bad[g]) { skip[gc] = 1; stop[gc] = 1; continue; }   // keep the last stable hyper-parameters
      it[g]++;
      lastL[g] = L;
      if (do_break) stop[gc] = 1;                                          // the M-step still runs once more
    }
    if (dLL) {
      bool any = false;
      for (int gc = 0; gc < Gc; gc++) any = any || (stop[gc] && !done[cur_groups[gc]]);
      if (any) {
        dbuf.resize((size_t)kDerivStride * Gc);
        rc = D.lb_deriv(dbuf.data());
        if (rc) return rc;
        for (int gc = 0; gc < Gc; gc++)
          if (stop[gc] && !done[cur_groups[gc]])
            memcpy(dLL + (size_t)kDerivStride * cur_groups[gc], dbuf.data() + (size_t)kDerivStride * gc,
                   sizeof(double) * kDerivStride);
      }
    }
    rc = D.mstep_masked(skip.data());
    if (rc) return rc;
    for (int gc = 0; gc < Gc; gc++)
      if (stop[gc] && !done[cur_groups[gc]]) { done[cur_groups[gc]] = 1; left--; }
    if (left == 0) break;
    // compact when at least a quarter of the current centres belong to stopped groups
    int stopped = 0;
    for (int jc = 0; jc < (int)cmap.size(); jc++) stopped += done[grp0[cmap[jc]]];
    if (4 * stopped < (int)cmap.size()) continue;
    rc = save_current();
    if (rc) return rc;
    std::vector<int> ngroups_list, ncmap;
    std::vector<int32_t> nSr, ngrp;
    for (int gc = 0; gc < Gc; gc++)
      if (!done[cur_groups[gc]]) ngroups_list.push_back(cur_groups[gc]);
    for (int jc = 0; jc < (int)cmap.size(); jc++) {
      const int j = cmap[jc], g = grp0[j];
      if (done[g]) continue;
      const int gi = (int)(std::lower_bound(ngroups_list.begin(), ngroups_list.end(), g) - ngroups_list.begin());
      ncmap.push_back(j);
      nSr.push_back(Sr0[j]);
      ngrp.push_back(gi);
    }
    const int Kc = (int)ncmap.size();
    cur.resize(Kc, r, d, dd);
    for (int jc = 0; jc < Kc; jc++) cur.copy_centre(jc, full, ncmap[jc], r, d, dd);
    rc = D.upload_state(Kc, nSr.data(), r, ngrp.data(), (int)ngroups_list.size(), cur.alpha.data(), cur.eta.data(),
                        cur.eps.data(), cur.m.data(), cur.W.data(), cur.v.data(), cur.lam.data());
    if (rc) return rc;
    cur_groups.swap(ngroups_list);
    cmap.swap(ncmap);
    rc = D.select_priors(cur_groups.data(), (int)cur_groups.size());
    if (rc) return rc;
    compacted = true;
    ncompact++;
  }
  if (compacted) {
    int rc = save_current();
    if (rc) return rc;
    rc = D.upload_state(K0, Sr0.data(), r, had_groups ? grp0.data() : nullptr, G, full.alpha.data(), full.eta.data(),
                        full.eps.data(), full.m.data(), full.W.data(), full.v.data(), full.lam.data());
    if (rc) return rc;
    std::vector<int> all(G);
    for (int g = 0; g < G; g++) all[g] = g;
    rc = D.select_priors(all.data(), G);
    if (rc) return rc;
  }
  int rc = D.sync();
  if (rc) return rc;
  if (ncompact_out) *ncompact_out = compacted ? ncompact : 0;
  for (int g = 0; g < G; g++) {
    iters[g] = it[g];
    if (unstable) unstable[g] = bad[g];
  }
  return VBHEM_OK;
}
}  // namespace

int vbhem_em_run(vbhem_base_t *h, int tau, int max_iter, double min_diff, double *LogLs, int *iters, int *unstable) {
  if (!h) return fail(VBHEM_EINVAL, "base handle is NULL");
  SingleDriver D(h);
  return em_run_one(D, tau, max_iter, min_diff, LogLs, iters, unstable);
}

int vbhem_em_run_groups_deriv(vbhem_base_t *h, int tau, int max_iter, double min_diff, double *LogLs, int ld,
                              int *iters, int *unstable, double *dLL) {
  if (!h) return fail(VBHEM_EINVAL, "base handle is NULL");
  SingleDriver D(h);
  int ncompact = 0;
  int rc = em_run_groups(D, tau, max_iter, min_diff, LogLs, ld, iters, unstable, &ncompact, dLL);
  if (rc) return rc;
  if (ncompact) h->lastK = 0;
  h->counters[6] = ncompact;
  return VBHEM_OK;
}

int vbhem_em_run_groups(vbhem_base_t *h, int tau, int max_iter, double min_diff, double *LogLs, int ld, int *iters,
                        int *unstable) {
  if (!h) return fail(VBHEM_EINVAL, "base handle is NULL");
  SingleDriver D(h);
  int ncompact = 0;
  int rc = em_run_groups(D, tau, max_iter, min_diff, LogLs, ld, iters, unstable, &ncompact);
  if (rc) return rc;
  if (ncompact) h->lastK = 0;   // the pair results in the handle belong to the last compacted E-step: not fetchable
  h->counters[6] = ncompact;
  return VBHEM_OK;
}

int vbhem_fetch(vbhem_base_t *h, double *L_elbo, double *zhat) {
  if (!h) return fail(VBHEM_EINVAL, "base handle is NULL");
  if (h->lastK < 1) return fail(VBHEM_EINVAL, "no E-step has run on this handle");
  CU(cudaSetDevice(h->device));
  const size_t np = (size_t)h->B.N * h->lastK;
  if (L_elbo) CU(cudaMemcpyAsync(L_elbo, h->L.p, sizeof(double) * np, cudaMemcpyDeviceToHost, h->st));
  if (zhat) CU(cudaMemcpyAsync(zhat, h->z.p, sizeof(double) * np, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  return VBHEM_OK;
}

int vbhem_finalize_stats(vbhem_base_t *h, int K, const int32_t *Sr, int maxSr, int ngroups, double *stats) {
  if (!h || !Sr || !stats) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: NULL argument");
  if (K < 1 || maxSr < 1 || ngroups < 1) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: bad sizes");
  CU(cudaSetDevice(h->device));
  const int d = h->B.d, dd = h->B.dd;
  const size_t ns = vbhem_stats_size(K, maxSr, d, h->B.covmode, ngroups);
  CU(h->stats.ensure(sizeof(double) * ns));
  CU(h->fin_i.ensure(sizeof(int) * (size_t)K));
  CU(cudaMemcpyAsync(h->fin_i.p, Sr, sizeof(int) * K, cudaMemcpyHostToDevice, h->st));
  CU(cudaMemcpyAsync(h->stats.p, stats, sizeof(double) * ns, cudaMemcpyHostToDevice, h->st));
  StatsPtr S = carve_stats(h->stats.as<double>(), K, maxSr, d, dd, ngroups);
  launch_finalize(K, maxSr, d, h->B.covmode, h->fin_i.as<int>(), S.Nj, S.xi, S.pr, S.mu, S.Mu, h->st);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(stats, h->stats.p, sizeof(double) * ns, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  return VBHEM_OK;
}

int vbhem_finalize_stats_dev(vbhem_base_t *h, int K, const int32_t *Sr, int maxSr, int ngroups, double *stats_dev) {
  if (!h || !Sr || !stats_dev) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: NULL argument");
  if (K < 1 || maxSr < 1 || ngroups < 1) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: bad sizes");
  CU(cudaSetDevice(h->device));
  // Sr comes from pageable caller memory: the copy is staged by the runtime before the call returns
  CU(h->fin_i.ensure(sizeof(int) * (size_t)K));
  CU(cudaMemcpyAsync(h->fin_i.p, Sr, sizeof(int) * K, cudaMemcpyHostToDevice, h->st));
  StatsPtr S = carve_stats(stats_dev, K, maxSr, h->B.d, h->B.dd, ngroups);
  launch_finalize(K, maxSr, h->B.d, h->B.covmode, h->fin_i.as<int>(), S.Nj, S.xi, S.pr, S.mu, S.Mu, h->st);
  CU(cudaGetLastError());
  return VBHEM_OK;
}

int vbhem_set_profile(vbhem_base_t *h, int on) {
  if (!h) return fail(VBHEM_EINVAL, "base handle is NULL");
  h->profile = on != 0;
  return VBHEM_OK;
}

int vbhem_last_timings(vbhem_base_t *h, double out[8]) {
  if (!h || !out) return fail(VBHEM_EINVAL, "NULL argument");
  if (!h->profile) return fail(VBHEM_EINVAL, "profiling is off (vbhem_set_profile)");
  CU(cudaSetDevice(h->device));
  CU(cudaEventSynchronize(h->ev[5]));
  for (int k = 0; k < 8; k++) out[k] = 0.0;
  for (int k = 0; k < 5; k++) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev[k], h->ev[k + 1]));
    out[k] = ms;
  }
  float tot = 0.f;
  CU(cudaEventElapsedTime(&tot, h->ev[0], h->ev[5]));
  out[5] = tot;
  return VBHEM_OK;
}

int vbhem_measure_fp64_peak(int device, double out[2]) {
  if (!out) return fail(VBHEM_EINVAL, "NULL argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(VBHEM_ENODEVICE, "no CUDA device");
  if (device < 0 || device >= ndev) return fail(VBHEM_EINVAL, "device out of range");
  CU(cudaSetDevice(device));
  if (measure_fp64_peak(out) != 0) return fail(VBHEM_ECUDA, "fp64 peak measurement failed");
  return VBHEM_OK;
}

int vbhem_last_counters(vbhem_base_t *h, long long out[8]) {
  if (!h || !out) return fail(VBHEM_EINVAL, "NULL argument");
  int rc = settle_counters(h);    // active pairs and fall-backs of a non-synchronising E-step are still on the device
  if (rc) return rc;
  memcpy(out, h->counters, sizeof h->counters);
  return VBHEM_OK;
}

}  // extern "C"

#include "vbhem_multi.inl"
