"""VB-HMM learning loop (vbhmm_em / vbhmm_learn mirrors in vbhem_b200/hmm.py) against the oracle's restatement of
/root/reference/src/hmm/vbhmm_em.m + vbhmm_em_lb.m + vbhmm_init.m.

CPU part: the host formulas (initialisation from a GMM, M-step, lower bound from the fused sums) against the
oracle's literal restatement, which evaluates the lower bound from the full forward-backward outputs.
GPU part: whole EM runs -- same number of iterations, same trace of lower bounds (rtol 1e-9), same final
variational parameters (1e-7) as the oracle loop, whose forward-backward is the reference's own C when built."""
import numpy as np
import pytest

import vbhem_b200  # noqa: F401
from oracle import cref, numpy_ref as nr
from vbhem_b200 import hmm as H

INI = dict(alpha0=0.1, epsilon0=0.1, beta0=1.0, v0=5.0, mu0=[256.0, 192.0], W0=0.005, maxIter=60, minDiff=1e-6)


def make_sequences(n=14, K=3, seed=4, tmin=4, tmax=18, dim=2):
    rng = np.random.Generator(np.random.PCG64(seed))
    mus = rng.uniform([60, 50, 20][:dim], [450, 330, 200][:dim], size=(K, dim))
    A = rng.dirichlet(np.ones(K), size=K) + 1.5 * np.eye(K)
    A /= A.sum(1, keepdims=True)
    data = []
    for _ in range(n):
        T = int(rng.integers(tmin, tmax + 1))
        z = int(rng.integers(K))
        x = np.zeros((T, dim))
        for t in range(T):
            x[t] = mus[z] + rng.normal(0, 14, dim)
            z = int(rng.choice(K, p=A[z]))
        data.append(x)
    gmm = dict(prior=np.full(K, 1.0 / K), mean=mus + rng.normal(0, 8, (K, dim)), cov=np.tile(np.eye(dim) * 300.0, (K, 1, 1)))
    return data, gmm


def oracle_fb(data, par):
    """The reference's own C forward-backward when built here, else its numpy restatement."""
    if cref.have_fb_ref():
        return cref.ref_vbhmm_fb(cref.pack_sequences(data), par)
    return nr.vbhmm_fb(data, par)


@pytest.mark.parametrize("K,dim,W0", [(3, 2, 0.005), (2, 2, [0.004, 0.008]), (4, 3, 0.01)])
def test_host_formulas_match_oracle(K, dim, W0):
    ini = dict(INI, W0=W0, mu0=[256.0, 192.0, 100.0][:dim])
    data, gmm = make_sequences(K=K, dim=dim, seed=11 + K)
    ntot = sum(len(x) for x in data)
    vp_o = nr.vbhmm_init_from_gmm(gmm, ntot, ini)
    vp_h = H.vbhmm_init(data, K, dict(ini, initmode="initgmm", initgmm=gmm))
    vp_t = H.vbhmm_init((ntot, dim), K, dict(ini, initmode="initgmm", initgmm=gmm))
    for k in vp_o:
        np.testing.assert_allclose(vp_h[k], vp_o[k], rtol=1e-12, err_msg=k)
        np.testing.assert_array_equal(vp_t[k], vp_h[k])
    # one E-step by the oracle, then lower bound and M-step both ways
    par = nr.vbhmm_fb_prepass(vp_o)
    fbs = dict(nr.vbhmm_fb(data, par), logLambdaTilde=par["logLambdaTilde"], logPiTilde=par["logPiTilde"],
               logATilde=par["logATilde"])
    st = nr.vbhmm_em_stats(data, fbs)
    want = nr.vbhmm_em_lb(st, vp_o, fbs, ini, len(data))
    fused = dict(st, logLambdaTilde=par["logLambdaTilde"], logPiTilde=par["logPiTilde"], logATilde=par["logATilde"])
    got, terms = H.vbhmm_em_lb(fused, vp_o, ini)
    assert abs(got - want) <= 1e-10 * abs(want)
    assert set(terms) == {f"Lt{i}" for i in range(1, 9)}
    new = H.vbhmm_mstep(st, ini)
    one = nr.vbhmm_em(data, vp_o, dict(ini, maxIter=1))["varpar"]      # maxIter = 1: E-step, break, M-step
    for k in one:
        np.testing.assert_allclose(new[k], one[k], rtol=1e-11, err_msg=k)


def test_oracle_loop_lower_bound_increases():
    """The bound rises from pass to pass; it is not strictly monotone at convergence because the reference's
    Wishart update takes v = v0 + Nk + 1 (vbhmm_em.m:369, "BUG FIX (add 1)" at vbhmm_init.m:194), one more than the
    maximiser of its own bound -- it prints "[LL decreased]" for that case (vbhmm_em.m:299-309).  Parity is with that."""
    data, gmm = make_sequences(K=3, seed=21)
    vp0 = nr.vbhmm_init_from_gmm(gmm, sum(len(x) for x in data), INI)
    out = nr.vbhmm_em(data, vp0, INI)
    assert out["iter"] >= 3 and not out["unstable"]
    assert out["LLs"][1] > out["LLs"][0]
    assert np.all(np.diff(out["LLs"]) > -1e-6 * np.abs(out["LLs"][:-1]))
    np.testing.assert_allclose(out["trans"].sum(1), 1.0, rtol=1e-12)
    np.testing.assert_allclose(out["prior"].sum(), 1.0, rtol=1e-12)


def test_init_errors():
    data, gmm = make_sequences(K=2)
    with pytest.raises(ValueError, match="mu0"):
        H.vbhmm_init(data, 2, dict(INI, mu0=[1.0], initmode="initgmm", initgmm=gmm))
    with pytest.raises(ValueError, match="v0"):
        H.vbhmm_init(data, 2, dict(INI, v0=1.0, initmode="initgmm", initgmm=gmm))
    with pytest.raises(ValueError, match="initgmm"):
        H.vbhmm_init(data, 3, dict(INI, initmode="initgmm", initgmm=gmm))
    with pytest.raises(ValueError, match="initmode"):
        H.vbhmm_init(data, 2, dict(INI, initmode="nope"))


@pytest.mark.gpu
@pytest.mark.parametrize("K,dim,seed", [(3, 2, 4), (2, 2, 9), (4, 3, 5), (1, 2, 6)])
def test_em_run_matches_oracle_loop(K, dim, seed):
    ini = dict(INI, mu0=[256.0, 192.0, 100.0][:dim])
    data, gmm = make_sequences(n=20, K=K, dim=dim, seed=seed)
    vp0 = nr.vbhmm_init_from_gmm(gmm, sum(len(x) for x in data), ini)
    want = nr.vbhmm_em(data, vp0, ini, fb=oracle_fb)
    got = H.vbhmm_em(data, K, dict(ini, initmode="initgmm", initgmm=gmm))
    assert got["iter"] == want["iter"] and got["unstable"] == want["unstable"]
    np.testing.assert_allclose(got["LLs"], want["LLs"], rtol=1e-9)
    for k in want["varpar"]:
        np.testing.assert_allclose(got["varpar"][k], want["varpar"][k], rtol=1e-7, atol=1e-12, err_msg=k)
    for k in ("prior", "trans", "mean", "cov", "M", "N1", "N"):
        np.testing.assert_allclose(got[k], want[k], rtol=1e-7, atol=1e-12, err_msg=k)


@pytest.mark.gpu
def test_learn_then_cluster_recovers_the_two_groups():
    """The toolbox's pipeline at the size of its synthetic experiment (Synthetic_experiment/exprmt1_demo.m: two
    ground-truth HMMs, 20 subjects each): learn one HMM per subject on the GPU (vbhmm_learn), convert
    (hmms_to_h3m_hem), cluster with VBHEM (K = 2) -- the known answer is the subjects' group (syn_evluate.m:19-22)."""
    from vbhem_b200 import BaseSet, h3m, synth
    rng = np.random.Generator(np.random.PCG64(1001))
    truth = [dict(mu=np.array([[120.0, 100.0], [380.0, 280.0]]), A=np.array([[0.8, 0.2], [0.3, 0.7]])),
             dict(mu=np.array([[130.0, 290.0], [370.0, 90.0]]), A=np.array([[0.3, 0.7], [0.7, 0.3]]))]
    hmms, label = [], []
    for s in range(16):
        g = s % 2
        data = []
        for _ in range(12):
            z, x = int(rng.integers(2)), np.zeros((10, 2))
            for t in range(10):
                x[t] = truth[g]["mu"][z] + rng.normal(0, 12, 2)
                z = int(rng.choice(2, p=truth[g]["A"][z]))
            data.append(x)
        hmms.append(H.vbhmm_learn(data, [1, 2, 3], dict(numTrials=3, seed=100 + s)))
        label.append(g)
    assert all(h["K"] == 2 for h in hmms), [h["K"] for h in hmms]
    out = h3m.vbhem_h3m_cluster(hmms, [1, 2, 3], [1, 2, 3], dict(trials=4, seed=1001))
    label = np.array(label)
    assert out["model_bestK"] == 2 and out["K"] == 2, (out["model_LL"], out["model_LL_S"])
    assert (out["label"] == label).all() or (out["label"] == 1 - label).all(), out["label"]
    assert out["group_size"].tolist() == [8, 8]
    np.testing.assert_allclose(out["Z"].sum(1), 1.0, rtol=1e-12)
    np.testing.assert_allclose(out["trans"][:, :out["S"], :out["S"]].sum(2), 1.0, rtol=1e-12)
    # the same selection by hand: every (K, S, trial) as a separate device EM run from the same initialisation
    base = h3m.hmms_to_h3m_hem(hmms)
    opt = out["vbhemopt"]
    with BaseSet(base) as bs:
        for ki, k in enumerate([1, 2, 3]):
            for si, sn in enumerate([1, 2, 3]):
                for t in range(4):
                    hyp = synth.init_centres_baseem(base, k, sn, seed=1001 + t + 1, opt=opt)
                    one = h3m.vbhem_h3m_c_step_fc(hyp, bs, opt, device_loop=True)
                    np.testing.assert_allclose(out["trials_LL"][ki, si, t], one["LL"], rtol=1e-12, err_msg=str((k, sn, t)))


def test_cluster_needs_a_seed():
    from vbhem_b200 import h3m
    with pytest.raises(ValueError, match="seed"):
        h3m.vbhem_h3m_cluster(dict(mu=np.zeros((2, 1, 2))), 2, 2, {})


@pytest.mark.gpu
@pytest.mark.parametrize("K,dim", [(3, 2), (2, 3), (5, 2), (1, 1)])
def test_batched_em_runs_match_oracle_loops(K, dim):
    """vbhmm_em_batch: several subjects x several initialisations as runs of ONE device-resident loop (pre-pass,
    forward-backward, sums, lower bound, convergence test and M-step all on the GPU) -- every run must reproduce the
    oracle's vbhmm_em from the same initialisation: passes, trace of lower bounds (1e-9), parameters (1e-7)."""
    from vbhem_b200 import fb
    ini = dict(INI, mu0=[256.0, 192.0, 100.0][:dim], maxIter=40)
    subjects, gmms = [], []
    for s in range(5):
        data, gmm = make_sequences(n=6 + 3 * s, K=K, dim=dim, seed=40 + s, tmin=1 if s == 1 else 4)
        subjects.append(data)
        gmms.append(gmm)
    flat = [x for sub in subjects for x in sub]
    start = np.concatenate([[0], np.cumsum([len(sub) for sub in subjects])])
    rng = np.random.Generator(np.random.PCG64(77))
    runs, inits, datas = [], [], []
    for s, sub in enumerate(subjects):
        for trial in range(2):                                     # two initialisations share the subject's sequences
            g = dict(gmms[s], mean=gmms[s]["mean"] + (rng.normal(0, 10, gmms[s]["mean"].shape) if trial else 0.0))
            inits.append(nr.vbhmm_init_from_gmm(g, sum(len(x) for x in sub), ini))
            runs.append((int(start[s]), int(start[s + 1])))
            datas.append(sub)
    with fb.SequenceSet(flat) as seqs:
        got, info = H.vbhmm_em_batch(seqs, runs, K, ini, inits)
    assert info["passes"] == max(h["iter"] for h in got)
    assert len(set(h["iter"] for h in got)) > 1 or K == 1        # the runs stop at different passes
    for r, (h, vp0, data) in enumerate(zip(got, inits, datas)):
        want = nr.vbhmm_em(data, vp0, ini, fb=oracle_fb)
        assert h["iter"] == want["iter"] and h["unstable"] == want["unstable"], r
        np.testing.assert_allclose(h["LLs"], want["LLs"], rtol=1e-9, err_msg=f"run {r}")
        for k in want["varpar"]:
            np.testing.assert_allclose(h["varpar"][k], want["varpar"][k], rtol=1e-7, atol=1e-12, err_msg=f"{k} run {r}")
        for k in ("prior", "trans", "mean", "cov", "M", "N1", "N"):
            np.testing.assert_allclose(h[k], want[k], rtol=1e-7, atol=1e-12, err_msg=f"{k} run {r}")


@pytest.mark.gpu
def test_learn_batch_equals_learning_subject_by_subject():
    """vbhmm_learn_batch (all subjects x K x trials in one device loop per K) picks the same models as vbhmm_learn
    subject by subject (same seeds => same initial GMMs)."""
    subjects = [make_sequences(n=10, K=2, seed=60 + s)[0] for s in range(4)]
    opt = dict(numTrials=3, seed=5, maxIter=50)
    got = H.vbhmm_learn_batch(subjects, [1, 2, 3], dict(opt, gmm_init="sklearn"))   # the host GMMs vbhmm_learn takes
    for s, sub in enumerate(subjects):
        one = H.vbhmm_learn(sub, [1, 2, 3], opt)
        assert got[s]["K"] == one["K"]
        np.testing.assert_allclose(got[s]["model_LL"], one["model_LL"], rtol=1e-9)
        # trials that reach the same optimum tie to the last bits, so the winner's states may be permuted
        order = lambda m: m[np.lexsort(m.T[::-1])]
        np.testing.assert_allclose(order(got[s]["mean"]), order(one["mean"]), rtol=1e-6)


def test_weighted_kmeans_and_wtkmeans_init():
    """my_weighted_kmeans.m restated: recovers weighted groups, its energy never rises; init_centres_wtkmeans
    (vbhemhmm_init.m:294-425) returns well-formed centres whose state means lie among the base means' range."""
    from vbhem_b200 import h3m, synth
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.normal(c, 0.3, (40, 2)) for c in ([0, 0], [3, 3], [0, 4])]).T
    w = rng.uniform(0.5, 1.5, 120)
    w /= w.sum()
    cl, cen, en = h3m.my_weighted_kmeans(3, 100, pts, w, pts[:, [0, 50, 100]] + 0.5)
    assert np.bincount(cl).tolist() == [40, 40, 40] and np.all(np.diff(en) <= 1e-12)
    for j in range(3):
        np.testing.assert_allclose(cen[:, j], (pts[:, cl == j] * w[cl == j]).sum(1) / w[cl == j].sum(), rtol=1e-12)
    for covmode in ("full", "diag"):
        base, _ = synth.make_base_set(40, S=3, d=2, C=3, seed=3, covmode=covmode)
        h = h3m.init_centres_wtkmeans(base, 3, 2, seed=5, opt=dict(h3m.CLUSTER_DEFAULTS))
        assert h["m"].shape == (3, 2, 2) and h["W"].shape == ((3, 2, 2, 2) if covmode == "full" else (3, 2, 2))
        assert np.all(h["m"] >= base["mu"].min()) and np.all(h["m"] <= base["mu"].max())
        assert np.all(h["eta"] > 0) and np.all(h["epsilon"] > 0) and np.all(h["v"] > 3)
        again = h3m.init_centres_wtkmeans(base, 3, 2, seed=5, opt=dict(h3m.CLUSTER_DEFAULTS))
        assert all(np.array_equal(h[k], again[k]) for k in h)


@pytest.mark.gpu
def test_cluster_with_wtkmeans_and_auto_init():
    """initmode 'wtkmeans' and 'auto' (trials alternate between baseem and wtkmeans): every trial of 'auto' is the
    same EM run as the trial of that number under its own mode, and the selection takes the best of them."""
    from vbhem_b200 import h3m, synth
    base, lab = synth.make_base_set(90, S=3, d=2, C=3, seed=31, Nv=20.0)
    common = dict(trials=4, seed=3, Nv=20.0, max_iter=60)
    runs = {mode: h3m.vbhem_h3m_cluster(base, [2, 3], 3, dict(common, initmode=mode)) for mode in ("baseem", "wtkmeans", "auto")}
    for out in runs.values():
        assert np.isfinite(out["LL"]) and out["Z"].shape == (90, out["K"]) and out["group_size"].sum() == 90
    # 'baseem' alone is initialised on the device, the baseem trials of 'auto' on the host: equal up to the last bit of W
    np.testing.assert_allclose(runs["auto"]["trials_LL"][:, :, 0::2], runs["baseem"]["trials_LL"][:, :, 0::2], rtol=1e-9)
    np.testing.assert_allclose(runs["auto"]["trials_LL"][:, :, 1::2], runs["wtkmeans"]["trials_LL"][:, :, 1::2], rtol=1e-12)
    assert runs["auto"]["LL"] >= min(runs["baseem"]["LL"], runs["wtkmeans"]["LL"]) - 1e-9 * abs(runs["auto"]["LL"])


def test_oracle_gmm_em_recovers_a_mixture():
    """oracle/numpy_ref.gmm_em (the algorithm of csrc/vbhmm_gmm.inl): likelihood never falls, a well-separated
    two-component mixture is recovered."""
    rng = np.random.default_rng(5)
    X = np.concatenate([rng.normal([0, 0], 1.0, (300, 2)), rng.normal([6, 5], [1.5, 0.5], (200, 2))])
    lls = [nr.gmm_em(X, [3, 400], 2, max_iter=m)["loglik"] for m in (1, 2, 3, 5, 10)]
    assert np.all(np.diff(lls) > -1e-9)
    g = nr.gmm_em(X, [3, 400], 2)
    order = np.argsort(g["mean"][:, 0])
    np.testing.assert_allclose(g["mean"][order], [[0, 0], [6, 5]], atol=0.25)
    np.testing.assert_allclose(g["prior"][order], [0.6, 0.4], atol=0.02)
    assert g["iters"] < 100


@pytest.mark.gpu
@pytest.mark.parametrize("K,dim", [(2, 2), (3, 2), (4, 3), (1, 1)])
def test_device_gmm_fits_match_oracle(K, dim):
    """vbhmm_gmm_fit (one block per run) against oracle gmm_em from the same start points; runs of different sizes,
    two starts per subject.  After 5 EM iterations: log-likelihood 1e-10, parameters 1e-9.  Run on (EM amplifies
    rounding differences where components overlap): iterations equal, log-likelihood 1e-5."""
    from vbhem_b200 import fb
    subjects = [make_sequences(n=6 + 4 * s, K=max(K, 2), dim=dim, seed=90 + s, tmin=1)[0] for s in range(4)]
    flat = [x for sub in subjects for x in sub]
    start = np.concatenate([[0], np.cumsum([len(sub) for sub in subjects])])
    rng = np.random.default_rng(1)
    runs, starts, pts = [], [], []
    for s, sub in enumerate(subjects):
        P = np.concatenate(sub)
        for _ in range(2):
            runs.append((int(start[s]), int(start[s + 1])))
            starts.append(rng.choice(len(P), size=K, replace=False))
            pts.append(P)
    with fb.SequenceSet(flat) as seqs:
        short = H.gmm_fit_batch(seqs, runs, K, starts, max_iter=5, tol=0.0)
        full = H.gmm_fit_batch(seqs, runs, K, starts, max_iter=60, tol=1e-6)
    for g, f, P, st in zip(short, full, pts, starts):
        want = nr.gmm_em(P, st, K, max_iter=5, tol=0.0)
        assert g["iters"] == want["iters"] == 5
        np.testing.assert_allclose(g["loglik"], want["loglik"], rtol=1e-10)
        np.testing.assert_allclose(g["prior"], want["prior"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(g["mean"], want["mean"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(g["cov"], want["cov"], rtol=1e-8, atol=1e-9)
        want = nr.gmm_em(P, st, K, max_iter=60, tol=1e-6)
        assert f["iters"] == want["iters"]
        np.testing.assert_allclose(f["loglik"], want["loglik"], rtol=1e-5)
        np.testing.assert_allclose(f["prior"].sum(), 1.0, rtol=1e-12)


@pytest.mark.gpu
def test_learn_batch_with_device_gmm_init():
    """vbhmm_learn_batch(gmm_init='device'): the random-initialisation GMMs come from the GPU; the learnt models are
    as good as with the host fits (same data, same number of trials) and find the generating state count."""
    subjects = [make_sequences(n=12, K=2, seed=60 + s, tmin=6)[0] for s in range(4)]
    opt = dict(numTrials=4, seed=5, maxIter=60)
    host = H.vbhmm_learn_batch(subjects, [1, 2, 3], opt)
    dev = H.vbhmm_learn_batch(subjects, [1, 2, 3], dict(opt, gmm_init="device"))
    for h, d in zip(host, dev):
        assert d["K"] == h["K"] == 2
        assert d["LL"] >= h["LL"] - 1e-3 * abs(h["LL"])


def test_remove_empty_states_and_groups():
    """vbhmm_remove_empty.m / vbh3m_remove_empty.m mirrors (pure host logic)."""
    from vbhem_b200 import h3m
    rng = np.random.default_rng(2)
    vp = dict(alpha=np.array([3.0, 0.1, 5.0]), epsilon=rng.uniform(1, 4, (3, 3)), beta=np.array([4.0, 1.0, 6.0]),
              v=np.array([8.0, 6.0, 9.0]), m=rng.normal(size=(3, 2)), W=np.stack([np.eye(2)] * 3))
    hm = dict(varpar=vp, M=rng.uniform(size=(3, 3)), N1=np.array([2.0, 1e-9, 4.0]), N=np.array([30.0, 1e-8, 50.0]),
              prior=vp["alpha"] / vp["alpha"].sum(), trans=vp["epsilon"] / vp["epsilon"].sum(1, keepdims=True),
              mean=vp["m"].copy(), cov=np.stack([np.eye(2)] * 3), K=3)
    out, gone = H.vbhmm_remove_empty(hm, 1e-3)
    assert gone.tolist() == [1] and out["K"] == 2 and out["varpar"]["epsilon"].shape == (2, 2)
    np.testing.assert_allclose(out["prior"], [3 / 8, 5 / 8])
    np.testing.assert_allclose(out["trans"].sum(1), 1.0)
    np.testing.assert_array_equal(out["mean"], hm["mean"][[0, 2]])
    same, none = H.vbhmm_remove_empty(out, 1e-3)
    assert same is out and none.size == 0
    res = dict(K=3, Nj=np.array([12.0, 1e-40, 8.0]), omega=np.array([0.6, 0.0, 0.4]), Z=rng.uniform(size=(5, 3)),
               L_elbo=rng.normal(size=(5, 3)), label=np.array([0, 2, 2, 0, 2]), groups=[np.array([0, 3]), np.array([], int),
               np.array([1, 2, 4])], group_size=np.array([2, 0, 3]), prior=np.zeros((3, 2)), trans=np.zeros((3, 2, 2)),
               mean=np.zeros((3, 2, 2)), cov=np.zeros((3, 2, 2, 2)), hyp=dict(alpha=np.ones(3), eta=np.ones((3, 2))))
    o, gone = h3m.vbh3m_remove_empty(res)
    assert gone.tolist() == [1] and o["K"] == 2 and o["Z"].shape == (5, 2) and o["label"].tolist() == [0, 1, 1, 0, 1]
    assert o["group_size"].tolist() == [2, 3] and o["hyp"]["eta"].shape == (2, 2)
    np.testing.assert_allclose(o["omega"], [0.6, 0.4])


def test_batched_init_equals_single_init():
    """hmm.vbhmm_init_from_gmms (all runs at once) gives exactly what one vbhmm_init('initgmm') call per run gives."""
    rng = np.random.default_rng(9)
    gmms, npts = [], []
    for r in range(6):
        A = rng.normal(size=(3, 2, 2))
        gmms.append(dict(prior=rng.dirichlet(np.ones(3)), mean=rng.uniform(0, 400, (3, 2)),
                         cov=A @ np.transpose(A, (0, 2, 1)) * 50 + 20 * np.eye(2)))
        npts.append(int(rng.integers(50, 400)))
    for W0 in (0.005, [0.004, 0.008]):
        ini = dict(INI, W0=W0)
        got = H.vbhmm_init_from_gmms(gmms, npts, 3, ini)
        for g, n, vp in zip(gmms, npts, got):
            want = H.vbhmm_init((n, 2), 3, dict(ini, initmode="initgmm", initgmm=g))
            for k in want:
                np.testing.assert_allclose(vp[k], want[k], rtol=1e-13, err_msg=k)
