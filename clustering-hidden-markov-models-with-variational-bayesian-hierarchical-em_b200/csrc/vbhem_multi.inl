// vbhem_multi.inl -- the base set sharded over several GPUs behind the C ABI (vbhem_multi_* of
// include/vbhem_b200.h).  Included at the end of vbhem_capi.cu (same translation unit: it drives the handles'
// internals).
//
// The reference's only parallelism is `parfor` inside ONE MATLAB session (/root/reference/src/vbhem/vbhem_h3m_c.m:28),
// so the drop-in must give one host process every GPU of the box: the N base HMMs are split contiguously
// (the pairs (i, j) are independent; the soft-max over j is local to a base), every device holds all centres and
// runs the fused E-step on its shard, and the raw per-centre statistics (vbhem_compute_Statistics.m:33-55: one
// contiguous buffer of vbhem_stats_size() doubles, 66 KB at the north-star shape) are summed over the devices
// with ONE ncclAllReduce(sum, fp64) over NVLink before the normalisation of :57-82.  The same code serves
// (a) one process driving `nlocal` devices (ncclCommInitAll; one persistent host thread per device, so the
//     launch sequences of the devices are enqueued concurrently), and
// (b) one process per device (torchrun / MPI style): nlocal = 1, the ncclUniqueId travels through the caller.
// NCCL is loaded at run time (dlopen of libnccl.so.2 -- the copy torch already loaded when there is one), so the
// library has no link-time dependency on it and single-GPU users never need it.
#include <nccl.h>   // types and enums only; every entry point is resolved with dlsym

#include <dlfcn.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace {

struct NcclApi {
  void *so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
  bool ok = false;
};

NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *env = getenv("VBHEM_NCCL_LIB");
    void *so = nullptr;
    if (env && *env) so = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!so) so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy already in the process (torch's)
    if (!so) so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!so) { api.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    api.so = so;
    bool all = true;
    auto sym = [&](const char *name) { void *p = dlsym(so, name); if (!p) { all = false; api.err = std::string("libnccl lacks ") + name; } return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = all;
  });
  return api;
}

#define NC(x)                                                                                          \
  do {                                                                                                 \
    ncclResult_t r_ = (x);                                                                             \
    if (r_ != ncclSuccess) return fail(VBHEM_ECUDA, "%s: %s", #x, nccl().GetErrorString(r_));          \
  } while (0)

// One persistent host thread per local device: the launch sequence of an E-step is ~15 enqueues, and issuing
// them for eight devices from one thread would start the last device a third of a millisecond after the first.
struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  bool has = false, done = false, quit = false;
  int rc = 0;
  std::string err;
  void start() {
    th = std::thread([this] {
      std::unique_lock<std::mutex> lk(mu);
      while (true) {
        cv.wait(lk, [this] { return has || quit; });
        if (quit) return;
        std::function<int()> f = std::move(job);
        has = false;
        lk.unlock();
        const int r = f();
        std::string e = r ? g_err : std::string();
        lk.lock();
        rc = r;
        err = std::move(e);
        done = true;
        cv.notify_all();
      }
    });
  }
  void post(std::function<int()> f) {
    std::lock_guard<std::mutex> lk(mu);
    job = std::move(f);
    has = true;
    done = false;
    cv.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [this] { return done; });
    if (rc) g_err = err;
    return rc;
  }
  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu);
      quit = true;
      cv.notify_all();
    }
    if (th.joinable()) th.join();
  }
};

}  // namespace

// ---- all-reduce over peer memory -------------------------------------------------------------------------------
// The statistics buffer is 66 KB: an all-reduce of that size is pure latency, and NCCL's costs a kernel launch of
// its own plus its protocol (measured: ~35 us per E-step at 8 GPUs, 2.5 % of the 1.4 ms step).  With NVLink peer
// access every device can simply READ the other devices' buffers, so the exchange is one small kernel per device:
//   1. copy my raw sums into my exchange buffer (double-buffered by call parity), fence system-wide;
//   2. the last CTA to finish writes this call's epoch into MY slot of every peer's flag array (remote stores);
//   3. one thread per CTA waits until every peer's epoch has arrived in the local flag array;
//   4. every thread sums its elements over the ranks 0 .. world-1 IN THAT ORDER (remote loads, uncached) and writes
//      the result in place: identical bits on every device, run to run.
// A peer reads buffer (e & 1) of epoch e only between the flags of epoch e and its own flag of epoch e + 1, and a
// device rewrites that buffer in epoch e + 2, after it has seen every flag of epoch e + 1: two buffers suffice.
// The wait is bounded (a peer that never arrives sets an error word instead of hanging the GPU).
// Exchange area of a device, bytes: [0, 256) flags[world] u64 | [256, 512) arrival counter | [512, 768) error word |
// [768, ...) two buffers of `cap` doubles.
constexpr int kPeerMax = 16;
constexpr size_t kPeerHdr = 768, kPeerCap = (size_t)1 << 18;       // 2 MB per buffer: 262 144 doubles
struct PeerArgs { char *base[kPeerMax]; int world, rank; };

__global__ void __launch_bounds__(256)
k_allreduce_peer(PeerArgs P, double *buf, size_t n, unsigned long long epoch, int slot) {
  vbhem::pdl_trigger();
  vbhem::pdl_wait();
  char *mine = P.base[P.rank];
  volatile unsigned long long *flags = reinterpret_cast<volatile unsigned long long *>(mine);
  unsigned *done = reinterpret_cast<unsigned *>(mine + 256);
  int *err = reinterpret_cast<int *>(mine + 512);
  double *xb = reinterpret_cast<double *>(mine + kPeerHdr) + (size_t)slot * kPeerCap;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  for (size_t x = tid; x < n; x += nth) xb[x] = buf[x];
  __threadfence_system();
  __syncthreads();
  __shared__ int ok;
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(done, 1u);
    if (old == gridDim.x - 1) {                    // every CTA's copy is out: tell the peers
      *done = 0;
      __threadfence_system();
      for (int q = 0; q < P.world; q++)
        reinterpret_cast<volatile unsigned long long *>(P.base[q])[P.rank] = epoch;
    }
    int good = 1;
    const long long t0 = clock64();
    for (int q = 0; q < P.world && good; q++)
      while (flags[q] < epoch)
        if (clock64() - t0 > 120000000000LL) { good = 0; *err = 1; break; }   // ~60 s: a peer never arrived
    ok = good;
  }
  __syncthreads();
  if (!ok) return;
  __threadfence_system();
  for (size_t x = tid; x < n; x += nth) {
    double s = 0.0;
    for (int q = 0; q < P.world; q++)
      s += __ldcv(reinterpret_cast<const double *>(P.base[q] + kPeerHdr) + (size_t)slot * kPeerCap + x);
    buf[x] = s;
  }
}

struct vbhem_multi {
  int nlocal = 0, world = 1, rank_first = 0;
  int N = 0;                               // bases of this process
  std::vector<int> dev, lo, cnt;           // per local shard: device, first base, number of bases
  std::vector<vbhem_base *> h;
  std::vector<ncclComm_t> comm;
  std::vector<Worker *> w;
  std::vector<cudaEvent_t> ev0, ev1;       // around the E-step (+ all-reduce) on every shard's stream
  int lastK = 0;
  size_t last_ns = 0;
  // peer-memory all-reduce: exchange areas of the local shards, the pointer table of every shard (all ranks), and
  // the epoch of the next call per shard
  std::vector<char *> peer_mem;
  std::vector<PeerArgs> peer_args;
  std::vector<unsigned long long> peer_epoch;
  std::vector<void *> peer_imported;       // cudaIpcOpenMemHandle mappings to close
  bool peer_ok = false;
  // f(k) for every local shard, concurrently (one worker thread per shard); first failure wins
  int run_all(const std::function<int(int)> &f) {
    if (nlocal == 1) return f(0);
    for (int k = 0; k < nlocal; k++) w[k]->post([&f, k] { return f(k); });
    int rc = VBHEM_OK;
    std::string e;
    for (int k = 0; k < nlocal; k++) {
      const int r = w[k]->wait();
      if (r && !rc) { rc = r; e = g_err; }
    }
    if (rc) g_err = e;
    return rc;
  }
};

namespace {

// all-reduce (sum) of a shard's raw statistics in place, on the shard's stream: over peer memory when the devices
// can reach each other (k_allreduce_peer), else ncclAllReduce
int allreduce_stats(vbhem_multi *m, int k, double *buf, size_t n) {
  if (m->world == 1) return VBHEM_OK;
  if (m->peer_ok && n <= kPeerCap) {
    const unsigned long long e = ++m->peer_epoch[k];
    int blocks = (int)((n + 255) / 256);
    if (blocks > 128) blocks = 128;              // every CTA must be resident while it waits
    vbhem::launch_chain(k_allreduce_peer, dim3(blocks), dim3(256), 0, m->h[k]->st, m->peer_args[k], buf, n, e,
                        (int)(e & 1));
    CU(cudaGetLastError());
    return VBHEM_OK;
  }
  NC(nccl().AllReduce(buf, buf, n, ncclDouble, ncclSum, m->comm[k], m->h[k]->st));
  return VBHEM_OK;
}

// Exchange areas for the peer all-reduce on every local shard; with one process driving all devices also the
// pointer tables (peer access enabled pairwise).  Returns false (NCCL stays in charge) when a pair cannot map.
bool peer_setup_local(vbhem_multi *m) {
  const char *env = getenv("VBHEM_ALLREDUCE");
  if (env && !strcmp(env, "nccl")) return false;
  if (m->world > kPeerMax) return false;
  const size_t bytes = kPeerHdr + 2 * kPeerCap * sizeof(double);
  m->peer_mem.assign(m->nlocal, nullptr);
  m->peer_args.assign(m->nlocal, PeerArgs{});
  m->peer_epoch.assign(m->nlocal, 0ULL);
  for (int k = 0; k < m->nlocal; k++) {
    if (cudaSetDevice(m->dev[k]) != cudaSuccess) return false;
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return false; }
    m->peer_mem[k] = static_cast<char *>(p);
    if (cudaMemset(p, 0, kPeerHdr) != cudaSuccess) return false;
  }
  if (m->world != m->nlocal) return true;          // several processes: vbhem_multi_peer_export / _import complete it
  for (int k = 0; k < m->nlocal; k++) {
    cudaSetDevice(m->dev[k]);
    for (int j = 0; j < m->nlocal; j++) {
      if (j == k) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, m->dev[k], m->dev[j]) != cudaSuccess || !can) return false;
      const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[j], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return false; }
      cudaGetLastError();
    }
    m->peer_args[k].world = m->world;
    m->peer_args[k].rank = k;
    for (int j = 0; j < m->nlocal; j++) m->peer_args[k].base[j] = m->peer_mem[j];
  }
  for (int k = 0; k < m->nlocal; k++) { cudaSetDevice(m->dev[k]); cudaDeviceSynchronize(); }
  return true;
}

// the error word of the peer all-reduce (a peer that never arrived): checked where the host waits anyway
int peer_check(vbhem_multi *m, int k) {
  if (!m->peer_ok) return VBHEM_OK;
  int e = 0;
  CU(cudaMemcpy(&e, m->peer_mem[k] + 512, sizeof e, cudaMemcpyDeviceToHost));
  if (e) return fail(VBHEM_ECUDA, "peer all-reduce timed out on device %d: a rank did not take part in the call", m->dev[k]);
  return VBHEM_OK;
}

// The fused E-step of every shard + all-reduce; host outputs as for vbhem_estep[_hyp] (stats = RAW sums over ALL
// ranks; L_elbo / zhat = the rows of this process's bases).  With `sio` the base set comes from host memory and
// every shard streams its own rows in (vbhem_estep*_stream).
int multi_estep(vbhem_multi *m, const CentreArgs &a, double *L_elbo, double *zhat, double *stats, double *pre,
                const StreamIO *sio = nullptr) {
  if (!stats) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: stats is NULL");
  const size_t ns = vbhem_stats_size(a.K, a.maxSr, m->h[0]->B.d, m->h[0]->B.covmode, a.ngroups);
  int rc = m->run_all([&](int k) -> int {
    vbhem_base *h = m->h[k];
    CU(cudaSetDevice(h->device));
    CU(h->stats.ensure(sizeof(double) * ns));
    CU(cudaEventRecord(m->ev0[k], h->st));
    const size_t lo = (size_t)m->lo[k];
    int r;
    if (sio) {
      const int ld = h->B.maxSb, d = h->B.d, dd = h->B.dd;
      StreamIO io{sio->Sb + lo, sio->prior + lo * ld, sio->A + lo * ld * ld, sio->mu + lo * ld * d,
                  sio->cov + lo * ld * dd, sio->Ntilde ? sio->Ntilde + lo : nullptr,
                  L_elbo ? L_elbo + lo : nullptr, zhat ? zhat + lo : nullptr, sio->nslices, (long long)m->N};
      r = estep_device(h, a, h->stats.as<double>(), k == 0 ? pre : nullptr, &io);
    } else {
      r = estep_device(h, a, h->stats.as<double>(), k == 0 ? pre : nullptr);
    }
    if (r) return r;
    r = allreduce_stats(m, k, h->stats.as<double>(), ns);
    if (r) return r;
    CU(cudaEventRecord(m->ev1[k], h->st));
    if (k == 0) CU(cudaMemcpyAsync(stats, h->stats.p, sizeof(double) * ns, cudaMemcpyDeviceToHost, h->st));
    if (!sio) {
      const size_t hp = sizeof(double) * (size_t)m->N, dp_ = sizeof(double) * (size_t)m->cnt[k];
      if (L_elbo)
        CU(cudaMemcpy2DAsync(L_elbo + lo, hp, h->L.p, dp_, dp_, a.K, cudaMemcpyDeviceToHost, h->st));
      if (zhat)
        CU(cudaMemcpy2DAsync(zhat + lo, hp, h->z.p, dp_, dp_, a.K, cudaMemcpyDeviceToHost, h->st));
    }
    r = settle_counters(h);      // synchronises the shard's stream
    if (r) return r;
    if (sio && h->cout) CU(cudaStreamSynchronize(h->cout));
    return peer_check(m, k);
  });
  if (rc) return rc;
  m->lastK = a.K;
  return VBHEM_OK;
}

// The same, everything left on the devices and nothing waited for: E-step, all-reduce and (optionally) the
// normalisation of vbhem_compute_Statistics.m:57-82 are enqueued on every shard's stream.
int multi_estep_dev(vbhem_multi *m, const CentreArgs &a, bool finalize) {
  const size_t ns = vbhem_stats_size(a.K, a.maxSr, m->h[0]->B.d, m->h[0]->B.covmode, a.ngroups);
  int rc = m->run_all([&](int k) -> int {
    vbhem_base *h = m->h[k];
    CU(cudaSetDevice(h->device));
    CU(h->stats.ensure(sizeof(double) * ns));
    CU(cudaEventRecord(m->ev0[k], h->st));
    int r = estep_device(h, a, h->stats.as<double>(), nullptr);
    if (r) return r;
    r = allreduce_stats(m, k, h->stats.as<double>(), ns);
    if (r) return r;
    if (finalize) {
      StatsPtr S = carve_stats(h->stats.as<double>(), a.K, a.maxSr, h->B.d, h->B.dd, a.ngroups);
      launch_finalize(a.K, a.maxSr, h->B.d, h->B.covmode, h->last_Sr_dev, S.Nj, S.xi, S.pr, S.mu, S.Mu, h->st);
    }
    CU(cudaEventRecord(m->ev1[k], h->st));
    CU(cudaGetLastError());
    return VBHEM_OK;
  });
  if (rc) return rc;
  m->lastK = a.K;
  m->last_ns = ns;
  return VBHEM_OK;
}

int check_multi(vbhem_multi *m) {
  if (!m || m->h.empty()) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: multi-GPU handle is NULL");
  return VBHEM_OK;
}

// EM loops over the shards: every shard keeps the centre hyper-parameters (identical everywhere: the all-reduced
// statistics are), so nothing but the statistics buffer travels between iterations.
struct MultiDriver : EmDriver {
  vbhem_multi *m;
  explicit MultiDriver(vbhem_multi *m_) : m(m_) {}
  vbhem_base *first() override { return m->h[0]; }
  int estep_lb(int tau, double *lb) override {
    return m->run_all([&](int k) -> int {
      vbhem_base *h = m->h[k];
      CU(cudaSetDevice(h->device));
      CU(cudaEventRecord(m->ev0[k], h->st));
      int r = vbhem_em_estep(h, tau, nullptr);
      if (r) return r;
      const size_t ns = vbhem_stats_size(h->em.K, h->em.maxSr, h->B.d, h->B.covmode, h->em.ngroups);
      r = allreduce_stats(m, k, h->em.stats.as<double>(), ns);
      if (r) return r;
      CU(cudaEventRecord(m->ev1[k], h->st));
      // every shard normalises its copy and evaluates the bound (the M-step needs the normalised statistics);
      // only the first one's value is read back
      return em_lower_bound(h, nullptr, k == 0 ? lb : nullptr, k == 0);
    });
  }
  int mstep_masked(const unsigned char *skip) override {
    return m->run_all([&](int k) { return em_mstep_masked(m->h[k], nullptr, skip); });
  }
  int fetch(double *alpha, double *eta, double *epsilon, double *mm, double *W, double *v, double *lambda) override {
    return vbhem_em_fetch(m->h[0], alpha, eta, epsilon, mm, W, v, lambda);
  }
  int upload_state(int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups, const double *alpha,
                   const double *eta, const double *epsilon, const double *mm, const double *W, const double *v,
                   const double *lambda) override {
    return m->run_all([&](int k) {
      return em_upload_state(m->h[k], K, Sr, maxSr, group, ngroups, alpha, eta, epsilon, mm, W, v, lambda);
    });
  }
  int sync() override {
    return m->run_all([&](int k) -> int {
      CU(cudaSetDevice(m->h[k]->device));
      CU(cudaStreamSynchronize(m->h[k]->st));
      return peer_check(m, k);
    });
  }
  int lb_deriv(double *out) override { return em_lb_deriv(m->h[0], out); }   // identical on every shard
  int select_priors(const int *groups, int n) override {
    return m->run_all([&](int k) { return em_upload_gpriors(m->h[k], groups, n); });
  }
};

}  // namespace

extern "C" {

int vbhem_multi_unique_id(unsigned char id[VBHEM_UNIQUE_ID_BYTES]) {
  if (!id) return fail(VBHEM_EINVAL, "id is NULL");
  static_assert(sizeof(ncclUniqueId) == VBHEM_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
  if (!nccl().ok) return fail(VBHEM_ENODEVICE, "NCCL is not available: %s", nccl().err.c_str());
  ncclUniqueId u;
  NC(nccl().GetUniqueId(&u));
  memcpy(id, &u, sizeof u);
  return VBHEM_OK;
}

int vbhem_multi_free(vbhem_multi_t *m) {
  if (!m) return VBHEM_OK;
  for (Worker *w : m->w) { w->stop(); delete w; }
  for (size_t k = 0; k < m->h.size(); k++) {
    if (m->h[k]) { cudaSetDevice(m->h[k]->device); cudaStreamSynchronize(m->h[k]->st); }
  }
  for (void *p : m->peer_imported) { if (!m->h.empty() && m->h[0]) cudaSetDevice(m->h[0]->device); cudaIpcCloseMemHandle(p); }
  for (size_t k = 0; k < m->peer_mem.size(); k++)
    if (m->peer_mem[k]) { cudaSetDevice(m->dev[k]); cudaFree(m->peer_mem[k]); }
  for (size_t k = 0; k < m->h.size(); k++) {
    if (m->h[k]) cudaSetDevice(m->h[k]->device);
    if (k < m->comm.size() && m->comm[k]) { if (m->h[k]) cudaStreamSynchronize(m->h[k]->st); nccl().CommDestroy(m->comm[k]); }
    if (k < m->ev0.size() && m->ev0[k]) cudaEventDestroy(m->ev0[k]);
    if (k < m->ev1.size() && m->ev1[k]) cudaEventDestroy(m->ev1[k]);
    vbhem_free_base(m->h[k]);
  }
  delete m;
  return VBHEM_OK;
}

int vbhem_multi_pack_base(int nlocal, const int *devices, int world, int rank_first, const unsigned char *id, int N,
                          const int32_t *Sb, int maxSb, int d, int covmode, const double *prior, const double *A,
                          const double *mu, const double *cov, const double *Ntilde, vbhem_multi_t **out) {
  if (!out) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: out is NULL");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(VBHEM_ENODEVICE, "no CUDA device (this library has no CPU path)");
  if (nlocal < 1 || nlocal > ndev) return fail(VBHEM_EINVAL, "nlocal = %d devices requested, %d visible", nlocal, ndev);
  if (world < nlocal || rank_first < 0 || rank_first + nlocal > world)
    return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: ranks [%d, %d) do not fit a world of %d", rank_first,
                rank_first + nlocal, world);
  if (world > nlocal && !id)
    return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: a job of several processes needs the id of vbhem_multi_unique_id");
  if (N < nlocal) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: fewer base HMMs (%d) than devices (%d)", N, nlocal);
  if (maxSb < 1 || d < 1 || (covmode != VBHEM_COV_DIAG && covmode != VBHEM_COV_FULL) || !Sb || !prior || !A || !mu || !cov)
    return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: bad base set arguments");
  if (world > 1 && !nccl().ok) return fail(VBHEM_ENODEVICE, "NCCL is not available: %s", nccl().err.c_str());
  vbhem_multi *m = new (std::nothrow) vbhem_multi;
  if (!m) return fail(VBHEM_ENOMEM, "out of host memory");
  m->nlocal = nlocal; m->world = world; m->rank_first = rank_first; m->N = N;
  const int dd = covmode ? d * d : d;
  auto bail = [&](int code) { vbhem_multi_free(m); return code; };
  for (int k = 0; k < nlocal; k++) {
    const int dv = devices ? devices[k] : k;
    if (dv < 0 || dv >= ndev) return bail(fail(VBHEM_EINVAL, "device %d out of range (have %d)", dv, ndev));
    // contiguous, balanced: the first N % nlocal shards get one base more (dist.shard_bounds)
    const int q = N / nlocal, r = N % nlocal;
    const int lo = k * q + (k < r ? k : r), n = q + (k < r ? 1 : 0);
    m->dev.push_back(dv); m->lo.push_back(lo); m->cnt.push_back(n);
    m->h.push_back(nullptr); m->comm.push_back(nullptr); m->ev0.push_back(nullptr); m->ev1.push_back(nullptr);
  }
  // shards are packed concurrently (the uploads of one device do not wait for the previous device's)
  if (nlocal > 1)
    for (int k = 0; k < nlocal; k++) { m->w.push_back(new Worker); m->w.back()->start(); }
  int rc = m->run_all([&](int k) -> int {
    const size_t lo = (size_t)m->lo[k];
    int r = vbhem_pack_base(m->dev[k], m->cnt[k], Sb + lo, maxSb, d, covmode, prior + lo * maxSb,
                            A + lo * maxSb * maxSb, mu + lo * maxSb * d, cov + lo * maxSb * dd,
                            Ntilde ? Ntilde + lo : nullptr, &m->h[k]);
    if (r) return r;
    CU(cudaEventCreate(&m->ev0[k]));
    CU(cudaEventCreate(&m->ev1[k]));
    return VBHEM_OK;
  });
  if (rc) return bail(rc);
  if (world > 1) {
    if (world == nlocal && !id) {
      ncclResult_t r = nccl().CommInitAll(m->comm.data(), nlocal, m->dev.data());
      if (r != ncclSuccess) return bail(fail(VBHEM_ECUDA, "ncclCommInitAll: %s", nccl().GetErrorString(r)));
    } else {
      ncclUniqueId u;
      memcpy(&u, id, sizeof u);
      ncclResult_t r = nccl().GroupStart();
      for (int k = 0; k < nlocal && r == ncclSuccess; k++) {
        cudaSetDevice(m->dev[k]);
        r = nccl().CommInitRank(&m->comm[k], world, u, rank_first + k);
      }
      if (r == ncclSuccess) r = nccl().GroupEnd();
      if (r != ncclSuccess) return bail(fail(VBHEM_ECUDA, "ncclCommInitRank: %s", nccl().GetErrorString(r)));
    }
  }
  if (world > 1) {
    const bool have = peer_setup_local(m);
    m->peer_ok = have && world == nlocal;          // several processes: after vbhem_multi_peer_import
    if (!have) {
      for (size_t k = 0; k < m->peer_mem.size(); k++)
        if (m->peer_mem[k]) { cudaSetDevice(m->dev[k]); cudaFree(m->peer_mem[k]); }
      m->peer_mem.clear();
    }
  }
  *out = m;
  return VBHEM_OK;
}

int vbhem_multi_peer_export(vbhem_multi_t *m, unsigned char *handles) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (!handles) return fail(VBHEM_EINVAL, "handles is NULL");
  static_assert(sizeof(cudaIpcMemHandle_t) == VBHEM_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  if ((int)m->peer_mem.size() != m->nlocal) return fail(VBHEM_ENODEVICE, "peer all-reduce is not available on this set");
  for (int k = 0; k < m->nlocal; k++) {
    cudaIpcMemHandle_t hd;
    CU(cudaSetDevice(m->dev[k]));
    CU(cudaIpcGetMemHandle(&hd, m->peer_mem[k]));
    memcpy(handles + (size_t)VBHEM_IPC_HANDLE_BYTES * k, &hd, sizeof hd);
  }
  return VBHEM_OK;
}

int vbhem_multi_peer_import(vbhem_multi_t *m, const unsigned char *handles) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (!handles) return fail(VBHEM_EINVAL, "handles is NULL");
  if ((int)m->peer_mem.size() != m->nlocal) return fail(VBHEM_ENODEVICE, "peer all-reduce is not available on this set");
  if (m->nlocal != 1) return fail(VBHEM_EINVAL, "vbhem_multi_peer_import: one device per process (nlocal = 1)");
  if (m->world > kPeerMax) return fail(VBHEM_EINVAL, "peer all-reduce supports at most %d ranks", kPeerMax);
  CU(cudaSetDevice(m->dev[0]));
  PeerArgs &a = m->peer_args[0];
  a.world = m->world;
  a.rank = m->rank_first;
  for (int q = 0; q < m->world; q++) {
    if (q == m->rank_first) { a.base[q] = m->peer_mem[0]; continue; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handles + (size_t)VBHEM_IPC_HANDLE_BYTES * q, sizeof hd);
    void *p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(VBHEM_ECUDA, "cudaIpcOpenMemHandle (rank %d): %s -- NCCL stays in charge of the all-reduce", q,
                  cudaGetErrorString(e));
    }
    m->peer_imported.push_back(p);
    a.base[q] = static_cast<char *>(p);
  }
  m->peer_ok = true;
  return VBHEM_OK;
}

int vbhem_multi_allreduce_mode(vbhem_multi_t *m, int *peer) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (peer) *peer = m->peer_ok ? 1 : 0;
  return VBHEM_OK;
}

int vbhem_multi_repack_base(vbhem_multi_t *m, const int32_t *Sb, const double *prior, const double *A,
                            const double *mu, const double *cov, const double *Ntilde) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (!Sb || !prior || !A || !mu || !cov) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: a base array is NULL");
  return m->run_all([&](int k) {
    vbhem_base *h = m->h[k];
    const size_t lo = (size_t)m->lo[k], ld = h->B.maxSb, d = h->B.d, dd = h->B.dd;
    return vbhem_repack_base(h, Sb + lo, prior + lo * ld, A + lo * ld * ld, mu + lo * ld * d, cov + lo * ld * dd,
                             Ntilde ? Ntilde + lo : nullptr);
  });
}

int vbhem_multi_info(vbhem_multi_t *m, int *nlocal, int *world, int *N) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (nlocal) *nlocal = m->nlocal;
  if (world) *world = m->world;
  if (N) *N = m->N;
  return VBHEM_OK;
}

int vbhem_multi_shard(vbhem_multi_t *m, int k, int *device, int *first, int *count, vbhem_base_t **handle) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (k < 0 || k >= m->nlocal) return fail(VBHEM_EINVAL, "shard %d out of range (have %d)", k, m->nlocal);
  if (device) *device = m->dev[k];
  if (first) *first = m->lo[k];
  if (count) *count = m->cnt[k];
  if (handle) *handle = m->h[k];
  return VBHEM_OK;
}

int vbhem_multi_set_kernel(vbhem_multi_t *m, int which) {
  int rc = check_multi(m);
  if (rc) return rc;
  for (vbhem_base *h : m->h)
    if ((rc = vbhem_set_kernel(h, which)) != VBHEM_OK) return rc;
  return VBHEM_OK;
}

int vbhem_multi_estep(vbhem_multi_t *m, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups, int tau,
                      const double *logPi, const double *logA, const double *mm, const double *vW, const double *cterm,
                      const double *logOmega, double *L_elbo, double *zhat, double *stats) {
  int rc = check_multi(m);
  if (rc) return rc;
  CentreArgs a{};
  a.K = K; a.maxSr = maxSr; a.ngroups = ngroups; a.tau = tau; a.Sr = Sr; a.group = group;
  a.logPi = logPi; a.logA = logA; a.m = mm; a.vW = vW; a.cterm = cterm; a.logOmega = logOmega; a.hyp = false;
  rc = check_common(m->h[0], a, true);
  if (rc) return rc;
  return multi_estep(m, a, L_elbo, zhat, stats, nullptr);
}

int vbhem_multi_estep_hyp(vbhem_multi_t *m, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                          int tau, const double *alpha, const double *eta, const double *epsilon, const double *mm,
                          const double *W, const double *v, const double *lambda, double *L_elbo, double *zhat,
                          double *stats, double *pre) {
  int rc = check_multi(m);
  if (rc) return rc;
  CentreArgs a{};
  a.K = K; a.maxSr = maxSr; a.ngroups = ngroups; a.tau = tau; a.Sr = Sr; a.group = group;
  a.alpha = alpha; a.eta = eta; a.epsilon = epsilon; a.m = mm; a.W = W; a.v = v; a.lambda = lambda; a.hyp = true;
  rc = check_common(m->h[0], a, true);
  if (rc) return rc;
  return multi_estep(m, a, L_elbo, zhat, stats, pre);
}

int vbhem_multi_estep_hyp_stream(vbhem_multi_t *m, const int32_t *Sb, const double *prior, const double *A,
                                 const double *mu, const double *cov, const double *Ntilde, int nslices, int K,
                                 const int32_t *Sr, int maxSr, const int32_t *group, int ngroups, int tau,
                                 const double *alpha, const double *eta, const double *epsilon, const double *mm,
                                 const double *W, const double *v, const double *lambda, double *L_elbo, double *zhat,
                                 double *stats, double *pre) {
  int rc = check_multi(m);
  if (rc) return rc;
  CentreArgs a{};
  a.K = K; a.maxSr = maxSr; a.ngroups = ngroups; a.tau = tau; a.Sr = Sr; a.group = group;
  a.alpha = alpha; a.eta = eta; a.epsilon = epsilon; a.m = mm; a.W = W; a.v = v; a.lambda = lambda; a.hyp = true;
  rc = check_common(m->h[0], a, true);
  if (rc) return rc;
  if (!Sb || !prior || !A || !mu || !cov) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: a base array is NULL");
  for (int i = 0; i < m->N; i++)
    if (Sb[i] < 1 || Sb[i] > m->h[0]->B.maxSb)
      return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: Sb[%d]=%d outside 1..maxSb=%d", i, Sb[i], m->h[0]->B.maxSb);
  StreamIO io{Sb, prior, A, mu, cov, Ntilde, nullptr, nullptr, nslices, 0};
  return multi_estep(m, a, L_elbo, zhat, stats, pre, &io);
}

int vbhem_multi_estep_hyp_dev(vbhem_multi_t *m, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                              int tau, const double *alpha, const double *eta, const double *epsilon, const double *mm,
                              const double *W, const double *v, const double *lambda, int finalize) {
  int rc = check_multi(m);
  if (rc) return rc;
  CentreArgs a{};
  a.K = K; a.maxSr = maxSr; a.ngroups = ngroups; a.tau = tau; a.Sr = Sr; a.group = group;
  a.alpha = alpha; a.eta = eta; a.epsilon = epsilon; a.m = mm; a.W = W; a.v = v; a.lambda = lambda; a.hyp = true;
  rc = check_common(m->h[0], a, true);
  if (rc) return rc;
  return multi_estep_dev(m, a, finalize != 0);
}

int vbhem_multi_synchronize(vbhem_multi_t *m) {
  int rc = check_multi(m);
  if (rc) return rc;
  for (int k = 0; k < m->nlocal; k++) {
    CU(cudaSetDevice(m->h[k]->device));
    CU(cudaStreamSynchronize(m->h[k]->st));
    if ((rc = peer_check(m, k)) != VBHEM_OK) return rc;
  }
  return VBHEM_OK;
}

int vbhem_multi_align(vbhem_multi_t *m) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (m->world == 1) return VBHEM_OK;
  return m->run_all([&](int k) -> int {
    vbhem_base *h = m->h[k];
    CU(cudaSetDevice(h->device));
    CU(h->align.ensure(sizeof(double)));
    CU(cudaMemsetAsync(h->align.p, 0, sizeof(double), h->st));
    return allreduce_stats(m, k, h->align.as<double>(), 1);
  });
}

int vbhem_multi_stats(vbhem_multi_t *m, double *stats) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (!stats || !m->last_ns) return fail(VBHEM_EINVAL, "vbhem_mex:invalidinput: no vbhem_multi_estep_hyp_dev result to fetch");
  vbhem_base *h = m->h[0];
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpyAsync(stats, h->stats.p, sizeof(double) * m->last_ns, cudaMemcpyDeviceToHost, h->st));
  CU(cudaStreamSynchronize(h->st));
  return VBHEM_OK;
}

int vbhem_multi_fetch(vbhem_multi_t *m, double *L_elbo, double *zhat) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (m->h[0]->lastK < 1) return fail(VBHEM_EINVAL, "no E-step has run on this handle");
  const int K = m->h[0]->lastK;
  return m->run_all([&](int k) -> int {
    vbhem_base *h = m->h[k];
    CU(cudaSetDevice(h->device));
    const size_t hp = sizeof(double) * (size_t)m->N, dp_ = sizeof(double) * (size_t)m->cnt[k];
    if (L_elbo) CU(cudaMemcpy2DAsync(L_elbo + m->lo[k], hp, h->L.p, dp_, dp_, K, cudaMemcpyDeviceToHost, h->st));
    if (zhat) CU(cudaMemcpy2DAsync(zhat + m->lo[k], hp, h->z.p, dp_, dp_, K, cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return VBHEM_OK;
  });
}

int vbhem_multi_finalize_stats(vbhem_multi_t *m, int K, const int32_t *Sr, int maxSr, int ngroups, double *stats) {
  int rc = check_multi(m);
  if (rc) return rc;
  return vbhem_finalize_stats(m->h[0], K, Sr, maxSr, ngroups, stats);
}

int vbhem_multi_em_begin(vbhem_multi_t *m, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                         double alpha0, double eta0, double epsilon0, double lambda0, double v0, const double *m0,
                         const double *W0, int nW0, const double *alpha, const double *eta, const double *epsilon,
                         const double *mm, const double *W, const double *v, const double *lambda) {
  int rc = check_multi(m);
  if (rc) return rc;
  return m->run_all([&](int k) {
    return vbhem_em_begin(m->h[k], K, Sr, maxSr, group, ngroups, alpha0, eta0, epsilon0, lambda0, v0, m0, W0, nW0, alpha,
                          eta, epsilon, mm, W, v, lambda);
  });
}

int vbhem_multi_em_run(vbhem_multi_t *m, int tau, int max_iter, double min_diff, double *LogLs, int *iters,
                       int *unstable) {
  int rc = check_multi(m);
  if (rc) return rc;
  MultiDriver D(m);
  return em_run_one(D, tau, max_iter, min_diff, LogLs, iters, unstable);
}

int vbhem_multi_em_run_groups(vbhem_multi_t *m, int tau, int max_iter, double min_diff, double *LogLs, int ld,
                              int *iters, int *unstable) {
  int rc = check_multi(m);
  if (rc) return rc;
  MultiDriver D(m);
  int ncompact = 0;
  rc = em_run_groups(D, tau, max_iter, min_diff, LogLs, ld, iters, unstable, &ncompact);
  if (rc) return rc;
  for (vbhem_base *h : m->h) {
    if (ncompact) h->lastK = 0;
    h->counters[6] = ncompact;
  }
  return VBHEM_OK;
}

int vbhem_multi_em_step(vbhem_multi_t *m, int tau, double *lb_out) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (!m->h[0]->em.active) return fail(VBHEM_EINVAL, "vbhem_multi_em_begin has not been called on this handle");
  MultiDriver D(m);
  std::vector<double> lb(12 * (size_t)m->h[0]->em.ngroups);
  rc = D.estep_lb(tau, lb.data());
  if (rc) return rc;
  if (lb_out) memcpy(lb_out, lb.data(), sizeof(double) * lb.size());
  return D.mstep_masked(nullptr);
}

int vbhem_multi_em_fetch(vbhem_multi_t *m, double *alpha, double *eta, double *epsilon, double *mm, double *W,
                         double *v, double *lambda) {
  int rc = check_multi(m);
  if (rc) return rc;
  return vbhem_em_fetch(m->h[0], alpha, eta, epsilon, mm, W, v, lambda);
}

int vbhem_multi_last_ms(vbhem_multi_t *m, double *ms) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (!ms) return fail(VBHEM_EINVAL, "NULL argument");
  double worst = 0.0;
  for (int k = 0; k < m->nlocal; k++) {
    CU(cudaSetDevice(m->dev[k]));
    CU(cudaEventSynchronize(m->ev1[k]));
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, m->ev0[k], m->ev1[k]));
    if (t > worst) worst = t;
  }
  *ms = worst;
  return VBHEM_OK;
}

int vbhem_multi_counters(vbhem_multi_t *m, long long out[8]) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (!out) return fail(VBHEM_EINVAL, "NULL argument");
  for (int q = 0; q < 8; q++) out[q] = 0;
  for (vbhem_base *h : m->h) {
    long long c[8];
    if ((rc = vbhem_last_counters(h, c)) != VBHEM_OK) return rc;
    for (int q = 0; q < 8; q++) out[q] = (q >= 4) ? c[q] : out[q] + c[q];   // tiles / compactions are per shard
  }
  return VBHEM_OK;
}

}  // extern "C"
