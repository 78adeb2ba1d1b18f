/* vbhem_b200.h -- C ABI of the B200-native VBHEM-H3M E-step (libvbhem_b200.so).
 *
 * This is the drop-in boundary for ONE path of the reference toolbox: the
 * E-step that the reference runs through its MEX plug-in
 *     [L_elbo,nu_1,emit_pr,emit_mu,emit_M,sum_xi] =
 *         vbhem_hmm_bwd_fwd_mex(h3m_b.hmm,h3m_r.hmm,T,maxN,maxN2[,logdetCovPlusDdivlamR,invCovR])
 * (gateway:  /root/reference/src/vbhem/vbhem_hmm_bwd_fwd_mex.c:288-289;
 *  call sites: /root/reference/src/vbhem/vbhem_h3m_c_step_fc.m:175-176,193-194)
 * together with the MATLAB statements that consume its outputs every
 * iteration (pre-pass :118-165,180-191; responsibilities :271-283; statistics
 * reduction /root/reference/src/vbhem/vbhem_compute_Statistics.m:33-82; the
 * N-sized lower-bound terms /root/reference/src/vbhem/vbhemh3m_lb.m:90,93,107).
 *
 * Conventions
 *   - plain pointers and sizes only; every array is fp64 (or int32 where
 *     stated), COLUMN-MAJOR, exactly what mxGetPr() of the corresponding MATLAB
 *     array yields, so a MEX gateway needs no transposes;
 *   - ragged sets are padded: leading dimensions are maxSb / maxSr, entries
 *     beyond an HMM's own state count are ignored on input and zero on output;
 *   - every function returns VBHEM_OK (0) or a negative VBHEM_E* code and never
 *     throws; vbhem_last_error() gives the message of the calling thread's last
 *     failure.  NaN / Inf in the data are data, not errors: they propagate into
 *     L_elbo exactly as in the reference so the caller's NaN guard
 *     (vbhem_h3m_c_step_fc.m:338-354) still fires;
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with VBHEM_ENODEVICE.
 *
 * Index names: N = number of base HMMs (Kb), K = number of centre HMMs (Kr),
 * b = base states, r = centre states, d = emission dimension, dd = d*d (full
 * covariance) or d (diagonal), T = tau, pair index p = i + N*j.
 */
#ifndef VBHEM_B200_H
#define VBHEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VBHEM_OK          0
#define VBHEM_EINVAL     -1   /* bad argument (the reference raises vbhem_mex:invalidinput) */
#define VBHEM_ENODEVICE  -2   /* no usable CUDA device */
#define VBHEM_ECUDA      -3   /* CUDA runtime failure; see vbhem_last_error() */
#define VBHEM_ENOMEM     -4

#define VBHEM_COV_DIAG 0
#define VBHEM_COV_FULL 1

/* Opaque: a base HMM set (reference type h3m_b,
 * /root/reference/src/vbhem/hmms_to_h3m_hem.m:19-140) packed once and resident in
 * HBM, plus the device scratch, stream and per-iteration buffers that go with it.
 * One handle = one device = one stream; calls on one handle must not overlap. */
typedef struct vbhem_base vbhem_base_t;

int         vbhem_version(void);
const char *vbhem_last_error(void);
int         vbhem_device_count(int *count);

/* Pack the base set (replaces the per-call cell/struct walk of
 * vbhem_hmm_bwd_fwd_mex.c:459-473, done once instead of every iteration).
 *   Sb[N] int32          states per base HMM (1..maxSb)
 *   prior[maxSb,N]       hmm{i}.prior
 *   A[maxSb,maxSb,N]     hmm{i}.A, A(from,to)
 *   mu[d,maxSb,N]        hmm{i}.emit{k}.centres
 *   cov[d,d,maxSb,N] (full) | [d,maxSb,N] (diag)   hmm{i}.emit{k}.covars
 *   Ntilde[N]            tilde_N_k = Nv*Kb*omega (vbhem_h3m_c_step_fc.m:26-30); may be NULL (= 1)
 * All pointers are HOST pointers. */
int vbhem_pack_base(int device, int N, const int32_t *Sb, int maxSb, int d, int covmode,
                    const double *prior, const double *A, const double *mu, const double *cov,
                    const double *Ntilde, vbhem_base_t **out);
int vbhem_free_base(vbhem_base_t *base);

/* Re-upload a base set of the SAME shape (N, maxSb, d, covmode) into an existing handle: the
 * host-to-device copies of vbhem_pack_base without its allocations.  Ntilde may be NULL (= keep).
 * Synchronises the handle's stream before returning, so the host arrays may be reused at once. */
int vbhem_repack_base(vbhem_base_t *base, const int32_t *Sb, const double *prior, const double *A,
                      const double *mu, const double *cov, const double *Ntilde);

/* Run subsequent work of this handle on an existing CUDA stream (a cudaStream_t
 * passed as void*), e.g. torch's current stream.  NULL = the handle's own (non-blocking)
 * stream; to name the legacy default stream pass cudaStreamLegacy, not 0. */
int vbhem_set_stream(vbhem_base_t *base, void *cuda_stream);
int vbhem_synchronize(vbhem_base_t *base);

/* Kernel selection, for tests and benchmarks: 0 = automatic (tensor-core
 * lattice kernel when maxSb,maxSr <= 8 and d <= 3, else the general kernel),
 * 1 = force the general (any-size, literal log-space) kernel,
 * 2 = automatic, but without packing two small centres into one 8-state tile,
 * 3 = automatic, packing only centres of <= 4 states two by two (no 5+3 / 6+2 / 7+1 tiles). */
int vbhem_set_kernel(vbhem_base_t *base, int which);

/* ---- Boundary-1: the literal per-pair outputs of the MEX gateway ------------
 * Centre inputs (HOST), the pre-pass results the reference passes to the MEX:
 *   Sr[K] int32; logPi[maxSr,K] = logPiTilde; logA[maxSr,maxSr,K] = logATilde (row = from);
 *   m[d,maxSr,K]; vW[d,d,maxSr,K] | [d,maxSr,K] = v*W (invCovR{j}, vbhem_h3m_c_step_fc.m:189);
 *   cterm[maxSr,K] = logLambdaTildePlusDdivlamda.
 * Outputs (HOST, any of nu1..Mu may be NULL together => L_elbo only):
 *   L_elbo[N,K]; nu1[maxSr,p]; xi[maxSr,maxSr,p] (row = from); pr[maxSr,p];
 *   mu_out[maxSr,d,p]; Mu_out[maxSr,dd,p]
 * (plhs[0..5] of vbhem_hmm_bwd_fwd_mex.c:396-409 flattened over p = i + N*j). */
int vbhem_estep_pairs(vbhem_base_t *base, int K, const int32_t *Sr, int maxSr, int tau,
                      const double *logPi, const double *logA, const double *m, const double *vW,
                      const double *cterm,
                      double *L_elbo, double *nu1, double *xi, double *pr, double *mu_out, double *Mu_out);

/* ---- Boundary-2: the fused E-step ------------------------------------------
 * One call = MEX call + responsibilities + statistics reduction + the N-sized
 * lower-bound terms, without materialising the N*K per-pair statistics.
 * Extra inputs: logOmega[K] = psi(alpha)-psi(sum alpha); group[K] int32 softmax
 * segment of each centre (NULL = all in one H3M; the (K,S) model-selection grid
 * batches several H3Ms in one call), ngroups.
 * Outputs (HOST; L_elbo, zhat may be NULL):
 *   L_elbo[N,K], zhat[N,K]
 *   stats: one contiguous fp64 buffer of vbhem_stats_size() doubles holding the RAW sums
 *     Nj[K] | S_nu1[maxSr,K] | S_xi[maxSr,maxSr,K] | S_pr[maxSr,K] | S_mu[maxSr,d,K] |
 *     S_Mu[maxSr,dd,K] | lb[2,ngroups] (Lt1 = sum Z.*L_elbo, Lt7 = sum zhat.*log zhat)
 *   i.e. vbhem_compute_Statistics.m:33-55 before the normalisation of :57-82, with the
 *   Z_Ni > 1e-8 rule applied; Nj WITHOUT the final +1e-50.  Raw sums are what a
 *   multi-GPU caller all-reduces; vbhem_finalize_stats() then applies :57-82. */
size_t vbhem_stats_size(int K, int maxSr, int d, int covmode, int ngroups);
int    vbhem_stats_offsets(int K, int maxSr, int d, int covmode, int ngroups, size_t off[7]);

int vbhem_estep(vbhem_base_t *base, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                int tau, const double *logPi, const double *logA, const double *m, const double *vW,
                const double *cterm, const double *logOmega,
                double *L_elbo, double *zhat, double *stats);

/* Same, but the centre set is given by its hyper-parameters and the pre-pass
 * (vbhem_h3m_c_step_fc.m:118-165,180-191,271-273) runs on the device:
 *   alpha[K]; eta[maxSr,K]; epsilon[maxSr,maxSr,K] (row = from); m[d,maxSr,K];
 *   W[d,d,maxSr,K] | [d,maxSr,K]; v[maxSr,K]; lambda[maxSr,K].
 * pre (HOST, may be NULL) receives what the host-side lower bound needs:
 *   logOmega[K] | logPi[maxSr,K] | logA[maxSr,maxSr,K] | logLambdaTilde[maxSr,K]
 * contiguous, (K + maxSr*K*(2+maxSr)) doubles. */
int vbhem_estep_hyp(vbhem_base_t *base, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                    int tau, const double *alpha, const double *eta, const double *epsilon, const double *m,
                    const double *W, const double *v, const double *lambda,
                    double *L_elbo, double *zhat, double *stats, double *pre);

/* Streamed variants for a base set that lives in HOST memory (a MEX call that has no resident handle
 * yet, or a base set that changes every call): the base arrays of vbhem_pack_base are uploaded in
 * `nslices` slices on a copy stream while earlier slices already run, and the L_elbo / zhat rows of
 * finished slices are copied back on a third stream, so host<->device transfers hide behind the
 * kernels.  Results are bit-identical to vbhem_repack_base + vbhem_estep[_hyp]; afterwards the handle
 * holds the new base set.  The shape (N, maxSb, d, covmode) is the handle's.  Pinned host buffers are
 * needed for the copies to overlap (pageable memory works, serialised).  Ntilde may be NULL (= keep). */
int vbhem_estep_stream(vbhem_base_t *base, const int32_t *Sb, const double *prior, const double *A, const double *mu,
                       const double *cov, const double *Ntilde, int nslices, int K, const int32_t *Sr, int maxSr,
                       const int32_t *group, int ngroups, int tau, const double *logPi, const double *logA,
                       const double *m, const double *vW, const double *cterm, const double *logOmega, double *L_elbo,
                       double *zhat, double *stats);
int vbhem_estep_hyp_stream(vbhem_base_t *base, const int32_t *Sb, const double *prior, const double *A,
                           const double *mu, const double *cov, const double *Ntilde, int nslices, int K,
                           const int32_t *Sr, int maxSr, const int32_t *group, int ngroups, int tau,
                           const double *alpha, const double *eta, const double *epsilon, const double *m,
                           const double *W, const double *v, const double *lambda, double *L_elbo, double *zhat,
                           double *stats, double *pre);

/* Device-resident variants for multi-GPU hosts: identical to vbhem_estep /
 * vbhem_estep_hyp, but `stats_dev` is a DEVICE pointer (e.g. a torch tensor the
 * caller then all-reduces with NCCL) and nothing is copied back; the WHOLE E-step
 * -- including the second pass, whose size (the pairs with Z_Ni > 1e-8) is known on
 * the device only -- is enqueued on the handle's stream and the call returns without
 * synchronising (tests/test_gpu_parity.py::test_estep_dev_returns_without_synchronising).
 * The centre arrays are copied to pinned staging before the call returns; `pre` is
 * written asynchronously: synchronise before reading it.  L_elbo / zhat stay in the
 * handle (vbhem_fetch); vbhem_last_counters synchronises to read the device-side counts. */
int vbhem_estep_dev(vbhem_base_t *base, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                    int tau, const double *logPi, const double *logA, const double *m, const double *vW,
                    const double *cterm, const double *logOmega, double *stats_dev);
int vbhem_estep_hyp_dev(vbhem_base_t *base, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                        int tau, const double *alpha, const double *eta, const double *epsilon, const double *m,
                        const double *W, const double *v, const double *lambda, double *stats_dev, double *pre);
/* Copy the handle's last L_elbo[N,K] / zhat[N,K] to HOST buffers (either may be NULL). */
int vbhem_fetch(vbhem_base_t *base, double *L_elbo, double *zhat);

/* vbhem_compute_Statistics.m:57-82 applied in place to a stats buffer (HOST
 * pointer; O(K*maxSr*dd) work, run on the device): Nj += 1e-50;
 * S_pr -> Nj_rho = S_pr + 1e-50; S_mu -> y_bar = S_mu / Nj_rho;
 * S_Mu -> S_plus_C = S_Mu / Nj_rho - y_bar*y_bar'; centres with Sr == 1 get
 * Nj_rho2rho = 1e-12. */
int vbhem_finalize_stats(vbhem_base_t *base, int K, const int32_t *Sr, int maxSr, int ngroups, double *stats);

/* As vbhem_finalize_stats, on a DEVICE stats buffer, enqueued on the handle's stream. */
int vbhem_finalize_stats_dev(vbhem_base_t *base, int K, const int32_t *Sr, int maxSr, int ngroups, double *stats_dev);

/* ---- Device-resident EM run --------------------------------
 * The rest of one VBHEM iteration -- lower bound (/root/reference/src/vbhem/vbhemh3m_lb.m:64-186) and M-step
 * (/root/reference/src/vbhem/vbhem_mstep_component.m:42-70; alpha: vbhem_h3m_c_step_fc.m:396) -- on the
 * device, with the centre hyper-parameters resident in the handle between iterations.
 *   vbhem_em_begin       uploads the prior (alpha0, eta0, epsilon0, lambda0, v0, m0[d], W0[nW0 = 1 or d]) and
 *                        the initial hyper-parameters (layouts of vbhem_estep_hyp); several H3Ms may be
 *                        batched through group[] as in vbhem_estep.
 *   vbhem_em_estep       fused E-step from the resident hyper-parameters; raw sums into stats_dev (DEVICE
 *                        pointer, NULL = a buffer of the handle); returns without synchronising, so a
 *                        multi-GPU caller can all-reduce stats_dev next.
 *   vbhem_em_lower_bound normalises the statistics (vbhem_compute_Statistics.m:57-82) and evaluates, per
 *                        group g, lb_out[12 g + 0..9] = Lt1..Lt10, [10] = LB, [11] = sum_j Nj (HOST).
 *   vbhem_em_mstep       replaces the resident hyper-parameters by the M-step of those statistics.
 *   vbhem_em_fetch       copies the resident hyper-parameters to HOST arrays (any may be NULL).
 *   vbhem_em_mstep_masked  the same, leaving the groups with frozen[g] != 0 (HOST bytes, one per group) untouched.
 *   vbhem_em_run         the loop of vbhem_h3m_c_step_fc.m:115-433 for one H3M on one GPU: LogLs[] (room
 *                        for max_iter + 2 values), *iters = M-steps taken, *unstable = NaN guard fired.
 *   vbhem_em_run_groups  that loop for every batched H3M at once (trials, initialisations or (K,S) candidates
 *                        that share the base set: vbhem_h3m_c.m:28-76, vbhem_h3m_cluster.m:261-355): each group
 *                        stops by its own convergence test and keeps its hyper-parameters while the others go
 *                        on; once a quarter of the centres belong to stopped groups the E-step continues on
 *                        the running groups only.  LogLs is [ngroups][ld] (ld >= max_iter + 2), iters /
 *                        unstable are [ngroups].  Afterwards vbhem_em_fetch returns every group in its
 *                        original place; the per-pair results of the last E-step are not fetchable
 *                        (vbhem_fetch) when the set was compacted on the way. */
int vbhem_em_begin(vbhem_base_t *base, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                   double alpha0, double eta0, double epsilon0, double lambda0, double v0, const double *m0,
                   const double *W0, int nW0, const double *alpha, const double *eta, const double *epsilon,
                   const double *m, const double *W, const double *v, const double *lambda);
/* vbhem_em_begin with the toolbox's `baseem` initialisation (/root/reference/src/vbhem/vbhemhmm_init.m:58-100) done on
 * the device from the resident base set, for every batched H3M at once: centre state (k, j) copies the emission of
 * base state randg[k,j] of base HMM randb[k,j] (both [maxSr,K] int32, 0-BASED; the reference's vbhemopt.prerand,
 * :66-74, or the caller's own randi draws), lambda = lambda0 + NLr/Sr, v = v0 + NLr/Sr + 1, m = centres,
 * W = inv((v - d - 1) covars), eta = NLr/Sr + eta0, epsilon = NLr/Sr^2 + epsilon0 (makeAprior's uniform mode),
 * alpha = alpha0 + omega Nv with omega[K] the caller's normalised rand draws, NLr = Nv / Kr of the centre's own H3M,
 * Nv = vbhemopt.Nv * Kb.  Continue with vbhem_em_run[_groups] / vbhem_em_fetch. */
int vbhem_em_begin_baseem(vbhem_base_t *base, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                          double alpha0, double eta0, double epsilon0, double lambda0, double v0, const double *m0,
                          const double *W0, int nW0, const int32_t *randb, const int32_t *randg, const double *omega,
                          double Nv);
int vbhem_em_estep(vbhem_base_t *base, int tau, double *stats_dev);
int vbhem_em_lower_bound(vbhem_base_t *base, double *stats_dev, double *lb_out);
int vbhem_em_mstep(vbhem_base_t *base, double *stats_dev);
int vbhem_em_mstep_masked(vbhem_base_t *base, double *stats_dev, const unsigned char *frozen);
int vbhem_em_fetch(vbhem_base_t *base, double *alpha, double *eta, double *epsilon, double *m, double *W, double *v,
                   double *lambda);
/* Batched hyper-parameter learning (/root/reference/src/vbhem/vbhem_h3m_c_hyp.m:105-129: every evaluation of the
 * optimiser's objective is a whole EM run under candidate prior hyper-parameters; vbhem_h3m_c.m:109-138 runs one
 * optimiser per trial, one after the other).  Here every batched H3M (soft-max group) can run under its OWN prior, so
 * the evaluations of all trials' optimisers are the groups of ONE vbhem_em_run_groups_deriv call:
 *   vbhem_em_set_group_priors  after vbhem_em_begin[_baseem]: alpha0[G], eta0[G], epsilon0[G], lambda0[G], v0[G],
 *                              m0[d,G], W0[nW0,G] (nW0 = 1: W0 = w I, else its diagonal); replaces the common prior
 *                              of vbhem_em_begin in the lower bound and the M-step until the next vbhem_em_begin
 *   vbhem_em_lb_deriv          d LB / d (alpha0, eta0, epsilon0, v0, lambda0, W0[8], m0[8]) of every group at the
 *                              current state (vbhemh3m_lb.m:202-309, before the transforms of :330-339), after
 *                              vbhem_em_estep: out[G][21] (HOST)
 *   vbhem_em_run_groups_deriv  vbhem_em_run_groups that also returns those derivatives taken at every group's
 *                              breaking pass, before that pass's M-step (vbhem_h3m_c_step_fc.m:357-368): dLL[G][21] */
int vbhem_em_set_group_priors(vbhem_base_t *base, const double *alpha0, const double *eta0, const double *epsilon0,
                              const double *lambda0, const double *v0, const double *m0, const double *W0, int nW0);
int vbhem_em_lb_deriv(vbhem_base_t *base, double *out);
int vbhem_em_run_groups_deriv(vbhem_base_t *base, int tau, int max_iter, double min_diff, double *LogLs, int ld,
                              int *iters, int *unstable, double *dLL);
int vbhem_em_run(vbhem_base_t *base, int tau, int max_iter, double min_diff, double *LogLs, int *iters,
                 int *unstable);
int vbhem_em_run_groups(vbhem_base_t *base, int tau, int max_iter, double min_diff, double *LogLs, int ld, int *iters,
                        int *unstable);

/* ---- The base set sharded over several GPUs (BASELINE north_star) -------------------------
 * The reference's only parallelism is `parfor` inside ONE MATLAB session
 * (/root/reference/src/vbhem/vbhem_h3m_c.m:28), so one host process must reach every GPU of the box through this
 * boundary.  A vbhem_multi_t splits the N base HMMs contiguously over the devices (shard k = bases
 * [k*q + min(k, N%n), ...), the first N%n shards one base longer), replicates the centres, runs the fused E-step on
 * every shard concurrently (one host thread per device inside the library) and sums the raw per-centre statistics
 * (vbhem_compute_Statistics.m:33-55, the buffer of vbhem_stats_size() doubles) over all shards with ONE
 * ncclAllReduce(sum, fp64) over NVLink per E-step; the Z_Ni > 1e-8 rule is applied per shard before the sum,
 * the normalisation of vbhem_compute_Statistics.m:57-82 after it.  Results equal the single-GPU ones up to fp64
 * summation order (<= 1e-11 relative) and are bit-identical from run to run.
 *
 * Two arrangements, same entry points:
 *   (a) one process, `nlocal` devices:  world = nlocal, rank_first = 0, id = NULL  (ncclCommInitAll);
 *   (b) several processes (one per device, torchrun / MPI style): every process passes ITS bases, nlocal = 1 (or
 *       more), world = total number of devices, rank_first = global rank of its first device, and the 128-byte id
 *       that ONE process obtained from vbhem_multi_unique_id and the caller broadcast to the others.  Ntilde keeps
 *       its global meaning (Nv * Kb_total * omega_i).
 * NCCL is loaded at run time (libnccl.so.2, or $VBHEM_NCCL_LIB); a set with world = 1 never touches it.
 *
 *   vbhem_multi_pack_base     arrays as for vbhem_pack_base (HOST), N = the bases of THIS process
 *   vbhem_multi_estep[_hyp]   as vbhem_estep[_hyp]: L_elbo / zhat [N, K] = the rows of this process's bases (may be
 *                             NULL), stats = the RAW sums over ALL ranks (finalise with vbhem_multi_finalize_stats),
 *                             pre as for vbhem_estep_hyp
 *   vbhem_multi_em_begin / _em_run / _em_run_groups / _em_fetch   the device-resident EM loops of vbhem_em_* over
 *                             the shards: every device keeps the (identical) centre hyper-parameters, only the
 *                             statistics buffer travels; LogLs etc. as for vbhem_em_run[_groups]
 *   vbhem_multi_em_step       one iteration (E-step + all-reduce + lower bound + M-step); lb_out[12 * ngroups]
 *   vbhem_multi_last_ms       device time of the last E-step + all-reduce: the slowest local shard (CUDA events)
 *   vbhem_multi_shard         shard k of this process: device, first base, number of bases, its single-GPU handle
 *   vbhem_multi_counters      vbhem_last_counters summed over the local shards
 *
 * The all-reduce itself: where the devices can map each other's memory (NVLink / NVSwitch peer access) it is ONE small
 * kernel per device that publishes the device's sums, waits for every peer's epoch flag and adds the peers' buffers
 * in rank order straight out of their memory (deterministic, no NCCL launch; csrc/vbhem_multi.inl); otherwise, for
 * buffers above 2 MB, or with VBHEM_ALLREDUCE=nccl in the environment, ncclAllReduce.  One process driving all
 * devices gets the peer path automatically.  Several processes exchange the IPC handles of their exchange areas once:
 * every process calls vbhem_multi_peer_export (handles[nlocal][64]), the host all-gathers them in rank order and
 * every process calls vbhem_multi_peer_import (handles[world][64]; one device per process).  Without the exchange
 * NCCL is used.  vbhem_multi_allreduce_mode: *peer = 1 when the peer path is active. */
#define VBHEM_UNIQUE_ID_BYTES 128
#define VBHEM_IPC_HANDLE_BYTES 64
typedef struct vbhem_multi vbhem_multi_t;
int vbhem_multi_peer_export(vbhem_multi_t *multi, unsigned char *handles);
int vbhem_multi_peer_import(vbhem_multi_t *multi, const unsigned char *handles);
int vbhem_multi_allreduce_mode(vbhem_multi_t *multi, int *peer);
int vbhem_multi_unique_id(unsigned char id[VBHEM_UNIQUE_ID_BYTES]);
int vbhem_multi_pack_base(int nlocal, const int *devices, int world, int rank_first, const unsigned char *id, int N,
                          const int32_t *Sb, int maxSb, int d, int covmode, const double *prior, const double *A,
                          const double *mu, const double *cov, const double *Ntilde, vbhem_multi_t **out);
int vbhem_multi_free(vbhem_multi_t *multi);
/* Re-upload a base set of the same shape (this process's N bases) into the shards: vbhem_repack_base per shard. */
int vbhem_multi_repack_base(vbhem_multi_t *multi, const int32_t *Sb, const double *prior, const double *A,
                            const double *mu, const double *cov, const double *Ntilde);
int vbhem_multi_info(vbhem_multi_t *multi, int *nlocal, int *world, int *N);
int vbhem_multi_shard(vbhem_multi_t *multi, int k, int *device, int *first, int *count, vbhem_base_t **handle);
int vbhem_multi_set_kernel(vbhem_multi_t *multi, int which);
int vbhem_multi_estep(vbhem_multi_t *multi, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                      int tau, const double *logPi, const double *logA, const double *m, const double *vW,
                      const double *cterm, const double *logOmega, double *L_elbo, double *zhat, double *stats);
int vbhem_multi_estep_hyp(vbhem_multi_t *multi, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                          int tau, const double *alpha, const double *eta, const double *epsilon, const double *m,
                          const double *W, const double *v, const double *lambda, double *L_elbo, double *zhat,
                          double *stats, double *pre);
/* Host-resident base set: every shard streams ITS rows of the arrays (layouts of vbhem_pack_base, N = this process's
 * bases) in slices while earlier slices compute, as vbhem_estep_hyp_stream does on one device. */
int vbhem_multi_estep_hyp_stream(vbhem_multi_t *multi, const int32_t *Sb, const double *prior, const double *A,
                                 const double *mu, const double *cov, const double *Ntilde, int nslices, int K,
                                 const int32_t *Sr, int maxSr, const int32_t *group, int ngroups, int tau,
                                 const double *alpha, const double *eta, const double *epsilon, const double *m,
                                 const double *W, const double *v, const double *lambda, double *L_elbo, double *zhat,
                                 double *stats, double *pre);
/* Nothing copied back, nothing waited for: E-step, all-reduce and (finalize != 0) the normalisation of
 * vbhem_compute_Statistics.m:57-82 are enqueued on every shard's stream.  vbhem_multi_synchronize waits for all
 * shards; vbhem_multi_stats copies the statistics buffer (raw or normalised, as requested) to the HOST;
 * vbhem_multi_fetch the per-pair rows. */
int vbhem_multi_estep_hyp_dev(vbhem_multi_t *multi, int K, const int32_t *Sr, int maxSr, const int32_t *group,
                              int ngroups, int tau, const double *alpha, const double *eta, const double *epsilon,
                              const double *m, const double *W, const double *v, const double *lambda, int finalize);
int vbhem_multi_synchronize(vbhem_multi_t *multi);
/* Device-side rendezvous of all shards of all ranks: enqueues a one-element all-reduce (the same exchange the E-step
 * uses) on every shard's stream and returns.  Whatever is enqueued next starts on every device within a few
 * microseconds of the last device arriving, whatever the hosts' launch skew: an EM loop is aligned like this by the
 * all-reduce of its previous iteration; a caller that issues single E-steps from several processes (bench.py) calls it
 * before the step it wants to time. */
int vbhem_multi_align(vbhem_multi_t *multi);
int vbhem_multi_stats(vbhem_multi_t *multi, double *stats);
int vbhem_multi_fetch(vbhem_multi_t *multi, double *L_elbo, double *zhat);
int vbhem_multi_finalize_stats(vbhem_multi_t *multi, int K, const int32_t *Sr, int maxSr, int ngroups, double *stats);
int vbhem_multi_em_begin(vbhem_multi_t *multi, int K, const int32_t *Sr, int maxSr, const int32_t *group, int ngroups,
                         double alpha0, double eta0, double epsilon0, double lambda0, double v0, const double *m0,
                         const double *W0, int nW0, const double *alpha, const double *eta, const double *epsilon,
                         const double *m, const double *W, const double *v, const double *lambda);
int vbhem_multi_em_run(vbhem_multi_t *multi, int tau, int max_iter, double min_diff, double *LogLs, int *iters,
                       int *unstable);
int vbhem_multi_em_run_groups(vbhem_multi_t *multi, int tau, int max_iter, double min_diff, double *LogLs, int ld,
                              int *iters, int *unstable);
int vbhem_multi_em_step(vbhem_multi_t *multi, int tau, double *lb_out);
int vbhem_multi_em_fetch(vbhem_multi_t *multi, double *alpha, double *eta, double *epsilon, double *m, double *W,
                         double *v, double *lambda);
int vbhem_multi_last_ms(vbhem_multi_t *multi, double *ms);
int vbhem_multi_counters(vbhem_multi_t *multi, long long out[8]);

/* ---- VB-HMM forward-backward (the toolbox's HMM learning) ----
 * Replaces the MEX plug-in of /root/reference/src/hmm/vbhmm_fb.m:99-145,
 *   [logrho_Saved, gamma_all, xi_sum, phi_norm] =
 *       vbhmm_fb_mex(data, K, N, dim, maxT, m, W, v, beta, logLambdaTilde, const_denominator, t_pz1, t_tpztzt1)
 * (gateway /root/reference/src/hmm/vbhmm_fb_mex.c:185-186; body :314-950).  The observation sequences do
 * not change during vbhmm_em, so they are packed once into HBM and every EM iteration passes only the model.
 *   vbhmm_pack_data: N sequences, sequence n = [T_n x dim] COLUMN-MAJOR (time fastest: mxGetPr(data{n})) at
 *     flat[offs[n]] (offs in doubles; the gateway may point offs into separately allocated cells as long as all
 *     lie in one address range, or copy them into one buffer); T_n = 0 is allowed.
 *   vbhmm_fb: model in the MEX layouts -- m[dim,K], W[dim,dim,K], v[K], beta[K], logLambdaTilde[K],
 *     const_denominator = dim*log(2*pi)/2, pz1[K] = exp(logPiTilde), tpztzt1[K,K] = exp(logATilde)(from,to) --
 *     outputs (HOST, any may be NULL): logrho_Saved[K,N,maxT], gamma_all[K,N,maxT], xi_sum[K,K,N], phi_norm[N],
 *     zero beyond a sequence's own length.  K <= 32, dim <= 8 (K <= 8, dim <= 3 take the fast kernel).
 *   vbhmm_fb_dev / vbhmm_fb_fetch: the same split into compute (results stay in the handle, no
 *     synchronisation) and copy-out.  vbhmm_fb_last_ms: device time of the last pass (CUDA events). */
typedef struct vbhmm_data vbhmm_data_t;
int vbhmm_pack_data(int device, int N, int dim, const int32_t *T, const long long *offs, const double *flat,
                    vbhmm_data_t **out);
int vbhmm_free_data(vbhmm_data_t *data);
int vbhmm_data_shape(vbhmm_data_t *data, int *N, int *dim, int *maxT, long long *total);
int vbhmm_fb(vbhmm_data_t *data, int K, const double *m, const double *W, const double *v, const double *beta,
             const double *logLambdaTilde, double const_denominator, const double *pz1, const double *tpztzt1,
             double *logrho_Saved, double *gamma_all, double *xi_sum, double *phi_norm);
int vbhmm_fb_dev(vbhmm_data_t *data, int K, const double *m, const double *W, const double *v, const double *beta,
                 const double *logLambdaTilde, double const_denominator, const double *pz1, const double *tpztzt1);
int vbhmm_fb_fetch(vbhmm_data_t *data, double *logrho_Saved, double *gamma_all, double *xi_sum, double *phi_norm);
/* The same pass fused with what vbhmm_em consumes (/root/reference/src/hmm/vbhmm_em.m:158-246 and the
 * fb-dependent lower-bound terms of vbhmm_em_lb.m:214,217), so that gamma_all / logrho_Saved never leave the device
 * during EM (fetch them once at the end with vbhmm_fb):
 *   Nk1[K] = sum_n gamma(:,n,1) + 1e-50;  Nk[K] = sum_{n,t} gamma + 1e-50;  M[K,K] = sum_n xi_sum (from,to);
 *   xbar[K,dim] = sum gamma x / Nk;  S[dim,dim,K] = sum gamma (x - xbar)(x - xbar)' / Nk;
 *   lb2[0] = Lt63 = sum gamma .* logrho_Saved,  lb2[1] = Lt64 = sum phi_norm.
 * All HOST, column-major, any may be NULL.  K <= 8, dim <= 3. */
int vbhmm_fb_stats(vbhmm_data_t *data, int K, const double *m, const double *W, const double *v, const double *beta,
                   const double *logLambdaTilde, double const_denominator, const double *pz1, const double *tpztzt1,
                   double *Nk1, double *Nk, double *M, double *xbar, double *S, double *lb2);
int vbhmm_fb_last_ms(vbhmm_data_t *data, double *ms);

/* ---- Batched VB-HMM learning: many independent vbhmm_em runs on the device -------------------------------
 * The toolbox learns one HMM per subject by looping vbhmm_learn over the subjects
 * (/root/reference/src/hmm/vbhmm_learn_batch.m), and inside it vbhmm_em over random initialisations and candidate
 * state counts (/root/reference/src/hmm/vbhmm_learn.m:465-557).  Here a RUN is a range [seq_lo, seq_hi) of the packed
 * sequences plus its own K-state model; all R runs of a batch execute the loop of
 * /root/reference/src/hmm/vbhmm_em.m:108-418 on the device -- pre-pass (vbhmm_fb.m:62-80), forward-backward, the sums
 * of vbhmm_em.m:158-246, the lower bound (vbhmm_em_lb.m:74-258), the convergence test |dL / L_last| <= min_diff from
 * the second pass on or max_iter passes (vbhmm_em.m:293-330), the M-step (:353-389, also after the breaking pass) --
 * and the host reads one integer per pass.  Runs may share sequences (several initialisations of one subject).
 *   vbhmm_batch_begin: priors alpha0, epsilon0, beta0, v0, m0[dim], W0[nW0 = 1 | dim] (vbhmm_init.m:128-152) and the
 *     initial variational parameters of every run, MATLAB layouts with the run as the last index:
 *     alpha[K,R], epsilon[K(from),K(to),R], beta[K,R], v[K,R], m[dim,K,R], W[dim,dim,K,R].  K <= 8, dim <= 3.
 *   vbhmm_batch_run: runs every run to its own end; *passes = passes of the slowest run.
 *   vbhmm_batch_fetch (any pointer may be NULL): final parameters (same layouts), the last E-step's N1[K,R] and
 *     N[K,R] (both with the +1e-50 of vbhmm_em.m:163,172) and M[K(from),K(to),R] (vbhmm_em.m:178), LLs[max_iter,R]
 *     (column r = run r's trace, 0 beyond its own passes), iters[R] = passes made,
 *     unstable[R] = the NaN guard of vbhmm_em.m:320-330 fired (then the parameters are those before that pass and the
 *     last LL is -Inf). */
typedef struct vbhmm_batch vbhmm_batch_t;
int vbhmm_batch_begin(vbhmm_data_t *data, int R, const int32_t *seq_lo, const int32_t *seq_hi, int K, double alpha0,
                      double epsilon0, double beta0, double v0, const double *m0, const double *W0, int nW0,
                      const double *alpha, const double *epsilon, const double *beta, const double *v, const double *m,
                      const double *W, int max_iter, vbhmm_batch_t **out);
int vbhmm_batch_run(vbhmm_batch_t *batch, double min_diff, int *passes);
int vbhmm_batch_fetch(vbhmm_batch_t *batch, double *alpha, double *epsilon, double *beta, double *v, double *m,
                      double *W, double *Nk1, double *Nk, double *M, double *LLs, int32_t *iters, int32_t *unstable);
int vbhmm_batch_last_ms(vbhmm_batch_t *batch, double *ms);
int vbhmm_batch_free(vbhmm_batch_t *batch);

/* Batched Gaussian-mixture fits for the random initialisation of the runs above: the K-component full-covariance GMM
 * that /root/reference/src/hmm/vbhmm_init.m:43-61 fits to a run's pooled observations with
 * gmdistribution.fit(data, K, 'Options', struct('TolFun', 1e-5), 'Start', 'randSample') -- EM from K randomly chosen
 * observations -- once per random trial (vbhmm_learn.m:465-557); one block per run on the device.
 *   start_seq[K,R], start_t[K,R]: the observation (sequence index, time step) that starts component k of run r (the
 *     caller draws them); every component starts with the diagonal of the pooled covariance and weight 1/K.
 *   stops when |ll - ll_old| < tol |ll| (TolFun) or after max_iter iterations; reg is added to every covariance diagonal.
 *   outputs (HOST): weights[K,R], means[dim,K,R], covs[dim,dim,K,R], loglik[R] (may be NULL), iters[R] (may be NULL).
 * The toolbox's fit is the Statistics Toolbox's (proprietary); this is the textbook EM, pinned on oracle/numpy_ref.gmm_em. */
int vbhmm_gmm_fit(vbhmm_data_t *data, int R, const int32_t *seq_lo, const int32_t *seq_hi, int K,
                  const int32_t *start_seq, const int32_t *start_t, int max_iter, double tol, double reg,
                  double *weights, double *means, double *covs, double *loglik, int32_t *iters);

/* Phase timing of the handle's E-steps, for bench.py: when enabled, CUDA events are
 * recorded on the handle's stream around each phase and vbhem_last_timings() returns the
 * milliseconds of the last E-step:
 *   out[0] centre staging + pre-pass + records, out[1] first pass (L_elbo of every pair),
 *   out[2] responsibilities + active list, out[3] second pass (forward + statistics),
 *   out[4] reduction, out[5] total, out[6..7] reserved. */
int vbhem_set_profile(vbhem_base_t *base, int on);
int vbhem_last_timings(vbhem_base_t *base, double out[8]);

/* Measures the device's fp64 peaks with dependent-free chains, in TFLOP/s:
 *   out[0] DMMA (mma.sync.m8n8k4.f64), out[1] DFMA.  ~50 ms. */
int vbhem_measure_fp64_peak(int device, double out[2]);

/* Counters of the handle's last E-step, for bench.py / tests:
 *   out[0] kernels launched, out[1] pairs through the tensor-core kernel,
 *   out[2] pairs through the general kernel (incl. numerical fall-backs),
 *   out[3] pairs with Z > 1e-8 (second pass), out[4] 8-state tiles per base in the first pass,
 *   out[5] of which split tiles (5..7 + 3..1 states), out[6] compactions of the last vbhem_em_run_groups,
 *   out[7] reserved. */
int vbhem_last_counters(vbhem_base_t *base, long long out[8]);

#ifdef __cplusplus
}
#endif
#endif /* VBHEM_B200_H */
