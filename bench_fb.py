"""bench.py --workload fb: the VB-HMM forward-backward pass (BASELINE.json configs[4]) on one GPU.

Workload: N = 1 000 000 synthetic 2-D sequences of length 50, S = 5 states (``--bases`` overrides N).  One
"step" is one pass of ``vbhmm_fb`` over every sequence: emission log-densities, scaled forward-backward,
responsibilities, summed pair marginals, phi_norm -- what ``vbhmm_em`` runs once per EM iteration through
the MEX plug-in ``vbhmm_fb_mex`` (/root/reference/src/hmm/vbhmm_fb.m:99-145).

Same JSON contract as bench.py: ``value`` = sequences/s with the packed sequences resident in HBM (CUDA
events around the kernel on its own stream); ``e2e`` = pack (H2D) + pass + every output back to pinned host
memory; ``roofline`` = HBM (this path moves ~3 flop per byte); ``cpu_baseline`` = the reference's own C
(oracle/_ref/libvbhmm_fb_ref.so) on the host cores, bounded sample.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
METRIC = "VB-HMM forward-backward sequences/sec"
UNIT = "sequences/s"


def make_workload(N, T, K, dim, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    means = rng.uniform([0] * dim, [512, 384, 256][:dim], size=(K, dim))
    z = rng.integers(0, K, size=(N, T))
    X = means[z] + rng.normal(size=(N, T, dim)) * 15
    flat = np.ascontiguousarray(np.transpose(X, (0, 2, 1))).ravel()      # each sequence [T x dim] column-major
    Tn = np.full(N, T, np.int32)
    offs = np.arange(N + 1, dtype=np.int64) * T * dim
    vp = dict(v=5 + rng.uniform(size=K) * 10, W=np.stack([np.eye(dim) / (15.0 ** 2 * 8)] * K),
              epsilon=1 + rng.uniform(size=(K, K)) * 5, alpha=1 + rng.uniform(size=K) * 3, m=means,
              beta=np.full(K, 50.0))
    return Tn, offs, flat, vp


def cpu_worker(n, T, K, dim, seed):
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import cref, numpy_ref as nr
    Tn, offs, flat, vp = make_workload(n, T, K, dim, seed)
    par = nr.vbhmm_fb_prepass(vp)
    ps = dict(N=n, dim=dim, T=Tn, offs=offs, flat=flat, maxT=T)
    kind = "reference" if cref.have_fb_ref() else "port"
    fn = cref.ref_vbhmm_fb if kind == "reference" else cref.port_vbhmm_fb
    t0 = time.perf_counter()
    out = fn(ps, par)
    dt = time.perf_counter() - t0
    assert np.isfinite(out["phi_norm"]).all()
    print(json.dumps(dict(seconds=dt, sequences=n, kind=kind)))


def run_cpu_baseline(n, T, K, dim, cores):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "fb", "--cpu-worker", str(n), "--tau", str(T),
           "--states", str(K), "--dim", str(dim)]
    t0 = time.perf_counter()
    procs = [subprocess.Popen(cmd + ["--seed", str(3000 + w)], stdout=subprocess.PIPE, text=True,
                              env=dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES=""))
             for w in range(cores)]
    outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
    wall = time.perf_counter() - t0
    worst = max(o["seconds"] for o in outs)
    tot = sum(o["sequences"] for o in outs)
    return dict(value=tot / worst, unit=UNIT, cores=cores, kind=outs[0]["kind"],
                per_core=statistics.median(o["sequences"] / o["seconds"] for o in outs),
                sample=f"{cores} processes x {n} sequences (T={T}, dim={dim}, S={K}) through the reference MEX gateway "
                       f"vbhmm_fb_mex compiled in place; slowest worker {worst:.2f} s, wall {wall:.1f} s")


def main_fb(a):
    K, dim, T = (a.states if a.states != 8 else 5), a.dim, (a.tau if a.tau != 10 else 50)
    if a.cpu_worker:
        return cpu_worker(a.cpu_worker, T, K, dim, a.seed)
    N = a.bases if a.bases != 100_000 else 1_000_000
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    cores = os.cpu_count() or 1
    config = dict(workload="configs[4] (forward-backward part): synthetic 2-D sequences, one vbhmm_fb pass per step",
                  sequences=N, length=T, states=K, dim=dim,
                  l2="working set 5.8 GB >> L2; no flush needed between steps")
    if a.impl == "reference":
        if rank != 0:
            return
        vals = [run_cpu_baseline(a.cpu_bases or 20000, T, K, dim, cores) for _ in range(a.warmup + a.steps)][a.warmup:]
        v = statistics.median(x["value"] for x in vals)
        line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=1e3 * cores * (a.cpu_bases or 20000) / v, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                    config=dict(config, note="each step = a bounded sample of the workload"),
                    cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind=vals[0]["kind"], sample=vals[0]["sample"]),
                    e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        real_stdout.write(json.dumps(line) + "\n")
        return
    cpu = None if a.no_cpu else run_cpu_baseline(a.cpu_bases or 20000, T, K, dim, cores)

    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import vbhem_b200  # noqa: F401
    from vbhem_b200 import fb
    from vbhem_b200._lib import check, dp, ip, lib
    from bench import ClockSampler
    L = lib()
    W = max(a.warmup, 3)
    Tn, offs, flat, vp = make_workload(N, T, K, dim, a.seed)
    par = fb.prepass(vp)
    c = np.ascontiguousarray
    m, Wm, A = c(par["m"]), c(np.transpose(par["W"], (0, 2, 1))), c(par["A"].T)
    P = lambda x: x.ctypes.data_as(dp)
    model = (K, P(m), P(Wm), P(c(par["v"])), P(c(par["beta"])), P(c(par["logLambdaTilde"])),
             C.c_double(par["const_denominator"]), P(c(par["pz1"])), P(A))
    keep = (m, Wm, A)   # noqa: F841  (the ctypes pointers above borrow these)
    seqs = fb.SequenceSet.from_packed(Tn, offs, flat, dim)
    sampler = ClockSampler(0)
    sampler.start()
    ms = C.c_double()
    for _ in range(W):
        check(L.vbhmm_fb_dev(seqs._h, *model))
    check(L.vbhmm_fb_last_ms(seqs._h, C.byref(ms)))
    sampler.armed = True
    tot = 0.0
    for _ in range(a.steps):
        check(L.vbhmm_fb_dev(seqs._h, *model))
        check(L.vbhmm_fb_last_ms(seqs._h, C.byref(ms)))           # CUDA events on the kernel's own stream
        tot += ms.value
    # the EM-iteration form: the same pass fused with the sums vbhmm_em takes from it (vbhmm_fb_stats); only
    # 2K + K^2 + K dim + K dim^2 + 2 doubles come back, so this is also its end-to-end time per EM iteration
    st = [np.zeros(K), np.zeros(K), np.zeros((K, K)), np.zeros((dim, K)), np.zeros((K, dim, dim)), np.zeros(2)]
    stats_ms, stats_wall = [], []
    for i in range(W + a.steps):
        t0 = time.perf_counter()
        check(L.vbhmm_fb_stats(seqs._h, *model, *[P(x) for x in st]))
        dtw = time.perf_counter() - t0
        check(L.vbhmm_fb_last_ms(seqs._h, C.byref(ms)))
        if i >= W:
            stats_ms.append(ms.value)
            stats_wall.append(1e3 * dtw)
    # real recordings are ragged: the same pass over sequences whose lengths are uniform in 1..T (every warp mixes
    # lengths: its loops run to the longest of its 32 sequences and take the variant with per-lane conditions)
    ragged = None
    if not a.no_extras:
        rng = np.random.default_rng(a.seed + 17)
        Tr = rng.integers(1, T + 1, size=N).astype(np.int32)
        Xr = flat.reshape(N, dim, T)
        flat_r = np.zeros(N * T * dim)
        keepm = np.arange(T)[None, :] < Tr[:, None]
        for dd_ in range(dim):                    # sequence n: dim runs of Tr[n] steps from offs[n] ([Tr x dim] column-major)
            idx = offs[:-1, None] + dd_ * Tr[:, None].astype(np.int64) + np.arange(T, dtype=np.int64)[None, :]
            flat_r[idx[keepm]] = Xr[:, dd_, :][keepm]
        del idx, keepm
        with fb.SequenceSet.from_packed(Tr, offs, flat_r, dim) as sr:
            rms = []
            for i in range(2 + 3):
                check(L.vbhmm_fb_dev(sr._h, *model))
                check(L.vbhmm_fb_last_ms(sr._h, C.byref(ms)))
                if i >= 2:
                    rms.append(ms.value)
        steps_r = int(Tr.sum())
        ragged = dict(ms_per_pass=statistics.median(rms), sequence_steps=steps_r, uniform_sequence_steps=N * T,
                      ns_per_sequence_step=statistics.median(rms) * 1e6 / steps_r,
                      uniform_ns_per_sequence_step=tot / a.steps * 1e6 / (N * T),
                      note="lengths uniform in 1..T: half the sequence steps of the headline workload; a lane idles while "
                           "its warp finishes longer sequences, and ragged warps run the general variant of the loops")
        del flat_r
    # the toolbox's actual use: one HMM per subject (vbhmm_learn_batch.m).  20 000 subjects share the N sequences;
    # every subject's whole EM run (vbhmm_em.m) executes in ONE device-resident loop (vbhmm_batch_*)
    learn = None
    nsub = min(20000, N // 8)
    if nsub >= 1 and K <= 8 and dim <= 3:
        from vbhem_b200 import hmm as Hm
        per = N // nsub
        rng = np.random.default_rng(a.seed + 1)
        runs = [(s * per, (s + 1) * per) for s in range(nsub)]
        inits = [dict(vp, m=vp["m"] + rng.normal(0, 6, vp["m"].shape)) for _ in range(nsub)]
        ini = dict(alpha0=0.1, epsilon0=0.1, beta0=1.0, v0=5.0, mu0=[256.0, 192.0, 128.0][:dim], W0=0.005, maxIter=30,
                   minDiff=1e-5)
        t0 = time.perf_counter()
        hm, info = Hm.vbhmm_em_batch(seqs, runs, K, ini, inits)
        wall = time.perf_counter() - t0
        run_passes = sum(h["iter"] for h in hm)
        # the random-initialisation GMM fits of the same runs (vbhmm_init.m:43-61) on the device, against scikit-learn
        # on the host for a sample of them
        starts = [rng.choice(per * T, size=K, replace=False) for _ in range(nsub)]
        Hm.gmm_fit_batch(seqs, runs[:64], K, starts[:64])                     # warm-up (module load)
        t0 = time.perf_counter()
        gm = Hm.gmm_fit_batch(seqs, runs, K, starts)
        t_gmm = time.perf_counter() - t0
        Xs = np.transpose(flat[:per * T * dim].reshape(per, dim, T), (0, 2, 1)).reshape(-1, dim)
        t0 = time.perf_counter()
        for q in range(5):
            Hm._fit_gmm(Xs, K, q)
        t_sk = (time.perf_counter() - t0) / 5
        learn = dict(subjects=nsub, sequences_per_subject=per, passes=info["passes"], run_passes=run_passes,
                     device_ms=info["device_ms"], wall_ms=1e3 * wall,
                     sequence_passes_per_s=run_passes * per / (info["device_ms"] * 1e-3),
                     unstable=int(sum(h["unstable"] for h in hm)),
                     gmm_init=dict(device_seconds=t_gmm, runs=nsub, mean_iters=float(np.mean([g["iters"] for g in gm])),
                                   sklearn_seconds_per_fit=t_sk, sklearn_seconds_extrapolated=t_sk * nsub),
                     note="hmm.vbhmm_em_batch: every subject's vbhmm_em run (pre-pass, forward-backward, sums, lower "
                          "bound, convergence test |dL/L| <= 1e-5, M-step) in one device-resident loop; device_ms = the "
                          "loop (CUDA events), wall_ms includes stacking the initialisations and fetching 20k models; "
                          "sequence_passes_per_s = (sum over subjects of the passes each needed x its sequences) / device time")
    sampler.armed = False
    sampler.stop_flag = True
    sampler.join()
    value = N * a.steps / (tot * 1e-3)
    alg = N * T * dim * 8 + 2 * N * T * K * 8 + N * K * K * 8 + N * 8
    peak, src = 6552.0, "fallback 6552 GB/s"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy bandwidth, burst)"
    except Exception:  # noqa: BLE001
        pass
    traffic, traffic_src = None, "the committed capture is of the 1 000 000-sequence launch only"
    if (N, T, K, dim) == (1_000_000, 50, 5, 2):
        from bench import ncu_dram_bytes
        traffic, traffic_src = ncu_dram_bytes(os.path.join(ROOT, "profiles", "r02d_k_fb_ncu.json"),
                                              ("vbhmm_fb.cu", "vbhem_math.cuh"))
    ach = alg / (tot / a.steps * 1e-3) / 1e9
    roofline = dict(bound="hbm", kernel="k_fb<5,2> (emissions + scaled forward + backward + gamma + xi, one thread per "
                                        "sequence, persistent grid, alpha_hat parked on chip)", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak,
                    peak_source=src, traffic=traffic, traffic_source=traffic_src, algorithmic_bytes=alg,
                    bytes_per_sequence=alg / N,
                    note="algorithmic bytes = observations in + logrho_Saved, gamma_all, xi_sum, phi_norm out; "
                         "alpha_hat waits for the backward pass in shared memory and in an L2-resident scratch buffer "
                         "(what leaks from L2 is the difference between traffic and the algorithmic bytes)")
    # e2e: pack (H2D of the observations) + pass + all four outputs to pinned host memory, every step
    e2e = None
    if not a.no_e2e:
        pin = lambda x: torch.from_numpy(x).pin_memory().numpy()
        flat_p = pin(flat)
        lr = torch.empty((T, N, K), dtype=torch.float64).pin_memory().numpy()
        gm = torch.empty((T, N, K), dtype=torch.float64).pin_memory().numpy()
        xi = torch.empty((N, K, K), dtype=torch.float64).pin_memory().numpy()
        ph = torch.empty(N, dtype=torch.float64).pin_memory().numpy()
        seqs.close()

        def e2e_step():
            h = C.c_void_p()
            check(L.vbhmm_pack_data(0, N, dim, Tn.ctypes.data_as(ip), offs.ctypes.data_as(C.POINTER(C.c_longlong)),
                                    P(flat_p), C.byref(h)))
            check(L.vbhmm_fb(h, *model, P(lr), P(gm), P(xi), P(ph)))
            L.vbhmm_free_data(h)
            return float(ph[0])

        e2e_step()
        ne = max(2, min(a.steps, 3))
        t0 = time.perf_counter()
        for _ in range(ne):
            e2e_step()
        dt = time.perf_counter() - t0
        e2e = dict(value=N * ne / dt, unit=UNIT, h2d_bytes_per_step=int(flat.nbytes + Tn.nbytes + offs.nbytes),
                   d2h_bytes_per_step=int(lr.nbytes + gm.nbytes + xi.nbytes + ph.nbytes), ms_per_step=1e3 * dt / ne,
                   steps=ne, note="every step: vbhmm_pack_data (observations H2D + repack) + vbhmm_fb (all four outputs "
                                  "D2H into pinned memory) + free; wall clock")
    else:
        seqs.close()
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=1, steps=a.steps, warmup=W, ms_per_step=tot / a.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", config=config,
                clocks=sampler.summary(), e2e=e2e, gpu_launches=a.steps + 3 * a.steps, roofline=roofline,
                em_iteration=dict(device_ms=statistics.median(stats_ms), wall_ms=statistics.median(stats_wall),
                                  sequences_per_s=N / (statistics.median(stats_wall) * 1e-3),
                                  d2h_bytes=int(sum(x.nbytes for x in st)),
                                  note="vbhmm_fb_stats: forward-backward + the M-step sums of vbhmm_em.m:158-246 + Lt63 / "
                                       "Lt64 on the device, observations resident (packed once per vbhmm_em run); wall_ms "
                                       "is host wall clock of the C-ABI call including the result copy"))
    if ragged is not None:
        line["ragged"] = ragged
    if learn is not None:
        line["learn_batch"] = learn
    if cpu is not None:
        line["cpu_baseline"] = cpu
    real_stdout.write(json.dumps(line) + "\n")
    real_stdout.flush()
