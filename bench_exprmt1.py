"""bench.py --workload exprmt1: BASELINE.json configs[0], the reference's own CPU-runnable case, end to end.

/root/reference/Synthetic_experiment/exprmt1_sampledata.m + exprmt1_demo.m: two ground-truth 2-state HMMs that share
their emissions (means [0,0] / [3,3], unit covariances) and differ only in the transition matrix; 20 subjects each,
25 sequences of length 50 per subject.  Pipeline: learn one VB-HMM per subject (S = 2), cluster the 40 HMMs with
VBHEM over K = 1..6 x S = 1..5 (Nv = 100, tau = 50, alpha0 = 1e6, `baseem`, `--trials` random initialisations per
candidate; the reference's default is 100), select (K, S) by the lower bound.  Known answer (syn_evluate.m:19-22): K = 2
and the subjects' group.  On the GPU every subject's HMM learning is one batched device loop (vbhmm_batch_*) and every
(K, S, trial) EM run of the clustering one soft-max group of one batched device loop (vbhem_em_run_groups).

JSON line: `value` = seconds of the clustering stage; `cpu_baseline` = the same EM runs through the oracle loop (the
reference's C E-step + numpy M-step / lower bound) on a bounded sample of them, extrapolated to all runs on the host's
cores (runs are independent, the reference parallelises them with parfor, vbhem_h3m_c.m:28).
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
METRIC = "exprmt1 VBHEM model-selection clustering (K=1..6 x S=1..5) seconds"


def sample_data(seed=1):
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    mus = np.array([[0.0, 0.0], [3.0, 3.0]])
    trans = [np.array([[0.6, 0.4], [0.4, 0.6]]), np.array([[0.4, 0.6], [0.6, 0.4]])]
    subjects, label = [], []
    for g in range(2):
        for _ in range(20):
            seqs = []
            for _ in range(25):
                z = int(rng.integers(2))
                x = np.zeros((50, 2))
                for t in range(50):
                    x[t] = mus[z] + rng.normal(0, 1.0, 2) + rng.normal(0, 0.1)     # exprmt1_sampledata.m:62-63
                    z = int(rng.choice(2, p=trans[g][z]))
                seqs.append(x)
            subjects.append(seqs)
            label.append(g)
    return subjects, np.array(label)


VBOPT = dict(mu0=(1.5, 1.5), W0=1.0)                                               # exprmt1_sampledata.m:92-93
VBHEMOPT = dict(Nv=100.0, tau=50, alpha0=1e6, eta0=1.0, epsilon0=1.0, lambda0=1.0, v0=5.0, W0=1.0, m0=(1.5, 1.5),
                initmode="baseem")                                                 # exprmt1_demo.m:66-79


def cpu_worker(nruns, seed):
    """`nruns` of the clustering stage's EM runs through the oracle loop; prints seconds and EM passes."""
    import numpy as np
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import vbhem_b200  # noqa: F401
    from vbhem_b200 import h3m, synth
    from test_em_loop import oracle_em
    hmms = np.load(os.environ["EXPRMT1_HMMS"], allow_pickle=True)["hmms"].tolist()
    base = h3m.hmms_to_h3m_hem(hmms, Nv=VBHEMOPT["Nv"])
    opt = dict(h3m.CLUSTER_DEFAULTS, **VBHEMOPT)
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    passes = pairs = 0
    for _ in range(nruns):
        K, S = int(rng.integers(1, 7)), int(rng.integers(1, 6))
        hyp = synth.init_centres_baseem(base, K, S, seed=int(rng.integers(1 << 30)), opt=opt)
        _, LL = oracle_em(base, hyp, dict(opt, max_iter=opt["max_iter"]))
        passes += len(LL)
        pairs += len(LL) * K * len(hmms)
    print(json.dumps(dict(seconds=time.perf_counter() - t0, runs=nruns, passes=passes, pair_passes=pairs)))


def main_exprmt1(a):
    if a.cpu_worker:
        return cpu_worker(a.cpu_worker, a.seed)
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import numpy as np
    sys.path.insert(0, ROOT)
    import vbhem_b200  # noqa: F401
    from vbhem_b200 import h3m, hmm
    trials = a.trials
    subjects, label = sample_data(a.seed)
    hmm.vbhmm_learn_batch(subjects[:2], [2], dict(VBOPT, numTrials=1, seed=1, gmm_init="device"))   # warm-up (CUDA context, module load)
    t0 = time.perf_counter()
    hmms = hmm.vbhmm_learn_batch(subjects, [2], dict(VBOPT, numTrials=5, seed=1001, gmm_init="device"))
    t_learn = time.perf_counter() - t0
    keep = [dict(prior=h["prior"], trans=h["trans"], mean=h["mean"], cov=h["cov"],
                 varpar={k: h["varpar"][k] for k in ("alpha", "epsilon", "beta")}) for h in hmms]
    Ks, Ss = list(range(1, 7)), list(range(1, 6))
    h3m.vbhem_h3m_cluster(keep, [1, 2], [1, 2], dict(VBHEMOPT, trials=2, seed=1))          # warm-up (module load)
    t0 = time.perf_counter()
    out = h3m.vbhem_h3m_cluster(keep, Ks, Ss, dict(VBHEMOPT, trials=trials, seed=1001))
    t_cluster = time.perf_counter() - t0
    agree = max(np.mean(out["label"] == label), np.mean(out["label"] == 1 - label)) if out["K"] == 2 else None
    nruns = len(Ks) * len(Ss) * trials
    # every run's E-steps x its K: the (base, centre) pairs the clustering stage processed
    line = dict(metric=METRIC, value=t_cluster, unit="s", n_gpus=1, steps=1, warmup=1, ms_per_step=1e3 * t_cluster,
                higher_is_better=False, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload="configs[0]: Synthetic_experiment/exprmt1_demo (40 subjects x 25 sequences x T=50; "
                                     "VB-HMM learning S=2, then VBHEM clustering K=1..6 x S=1..5)", trials=trials,
                            em_runs=nruns, Nv=100, tau=50),
                learn_seconds=t_learn, cluster_seconds=t_cluster, cluster_breakdown=out["seconds"], best_K=int(out["K"]), best_S=int(out["S"]),
                true_K=2, true_S=2, label_agreement=None if agree is None else float(agree),
                model_LL=[float(x) for x in out["model_LL"]], gpu_launches=None)
    if not a.no_cpu:
        path = os.path.join(ROOT, "gpurun_out", "exprmt1_hmms.npz")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez(path, hmms=np.array(keep, dtype=object))
        cores = os.cpu_count() or 1
        per = a.cpu_bases or 40
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "exprmt1", "--cpu-worker", str(per)]
        t0 = time.perf_counter()
        procs = [subprocess.Popen(cmd + ["--seed", str(3000 + w)], stdout=subprocess.PIPE, text=True,
                                  env=dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="", EXPRMT1_HMMS=path))
                 for w in range(cores)]
        res = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
        wall = time.perf_counter() - t0
        runs = sum(r["runs"] for r in res)
        line["cpu_baseline"] = dict(value=wall * nruns / runs, unit="s", cores=cores, kind="reference",
                                    sample=f"{runs} random (K,S) EM runs ({sum(r['passes'] for r in res)} passes) through the "
                                           f"oracle loop (reference C E-step + numpy M-step / lower bound) on {cores} "
                                           f"processes in {wall:.1f} s, extrapolated to the {nruns} runs of the stage")
    real_stdout.write(json.dumps(line) + "\n")
    real_stdout.flush()
