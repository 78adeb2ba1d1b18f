"""EM driver (host mirror of vbhem_h3m_c_step_fc) on the GPU against the same loop built from the oracle:
lower-bound trace, iteration count and final hyper-parameters.  The M-step / lower-bound host code in the
package (h3m.py) is checked against oracle/numpy_ref.py on CPU as well."""
import numpy as np
import pytest

from oracle import cref, numpy_ref as nr
from vbhem_b200 import h3m, pack, synth
from util import make_case, prepass_natural

OPT = dict(alpha0=1.0, eta0=1.0, epsilon0=1.0, lambda0=1.0, v0=5.0, W0=0.005, m0=(256.0, 192.0), tau=5,
           max_iter=30, minDiff=1e-6, Nv=20.0)


def oracle_em(base, hyp, opt, covmode="full"):
    """The same control flow with every piece taken from the oracle."""
    pb = pack.pack_base(base)
    K, S = hyp["eta"].shape
    hyp = {k: np.array(v) for k, v in hyp.items()}
    LogLs, lastL, it = [], -np.finfo(float).max, 0
    while True:
        cen = prepass_natural(hyp, covmode)
        St = pack.stats_to_natural(cref.port_estep_fused(pb, pack.pack_centres(cen), opt["tau"], want_pairs=False), 2, 1)
        centres = [dict(eta=hyp["eta"][j], epsilon=hyp["epsilon"][j], m=hyp["m"][j], W=hyp["W"][j], v=hyp["v"][j],
                        lam=hyp["lam"][j]) for j in range(K)]
        pre = [nr.prepass(c, covmode) for c in centres]
        L, _ = nr.lower_bound(St["lb"][0], St["lb"][1], St["Nj"], cen["logOmega"], hyp["alpha"], centres, pre, opt, covmode)
        do_break = it > 1 and abs((L - lastL) / lastL) <= opt["minDiff"]
        do_break = do_break or it == opt["max_iter"]
        new = dict(alpha=opt["alpha0"] + St["Nj"], Sr=hyp["Sr"])
        parts = [nr.mstep_component(*nr.finalize_stats(St["nu1"][j], St["xi"][j], St["pr"][j], St["mu"][j], St["Mu"][j],
                                                       covmode), opt, covmode) for j in range(K)]
        for k in ("eta", "epsilon", "lam", "v", "m", "W"):
            new[k] = np.stack([p[k] for p in parts])
        hyp = new
        it += 1
        LogLs.append(L)
        lastL = L
        if do_break:
            break
    return hyp, np.array(LogLs)


def test_host_mstep_and_lb_match_numpy_ref():
    """CPU: h3m.vbhem_mstep / vbhemh3m_lb (vectorised) against the per-centre numpy_ref restatement."""
    base, hyp, cen = make_case(N=25, K=3, Sb=3, Sr=3, d=2, covmode="full", ragged=False, seed=12, Nv=10.0)
    St = pack.stats_to_natural(cref.port_estep_fused(pack.pack_base(base), pack.pack_centres(cen), 4), 2, 1)
    fins = [nr.finalize_stats(St["nu1"][j], St["xi"][j], St["pr"][j], St["mu"][j], St["Mu"][j]) for j in range(3)]
    fin = dict(Nj=St["Nj"], Nj_rho1=np.stack([f[0] for f in fins]), Nj_rho2rho=np.stack([f[1] for f in fins]),
               Nj_rho=np.stack([f[2] for f in fins]), y_bar=np.stack([f[3] for f in fins]),
               S_plus_C=np.stack([f[4] for f in fins]))
    got = h3m.vbhem_mstep(fin, OPT)
    for j in range(3):
        want = nr.mstep_component(*fins[j], OPT)
        for k in ("eta", "epsilon", "lam", "v", "m", "W"):
            np.testing.assert_allclose(got[k][j], want[k], rtol=1e-12, atol=1e-300, err_msg=k)
    centres = [dict(eta=hyp["eta"][j], epsilon=hyp["epsilon"][j], m=hyp["m"][j], W=hyp["W"][j], v=hyp["v"][j],
                    lam=hyp["lam"][j]) for j in range(3)]
    pre = [nr.prepass(c) for c in centres]
    want_L, want_t = nr.lower_bound(St["lb"][0], St["lb"][1], St["Nj"], cen["logOmega"], hyp["alpha"], centres, pre, OPT)
    pre_s = dict(logOmega=cen["logOmega"], logPi=cen["logPi"], logA=cen["logA"], logLambdaTilde=cen["logLambdaTilde"])
    got_L, got_t = h3m.vbhemh3m_lb(St["lb"][0], St["lb"][1], St["Nj"], pre_s, hyp, OPT)
    for k in want_t:
        np.testing.assert_allclose(got_t[k], want_t[k], rtol=1e-12, atol=1e-9, err_msg=k)
    np.testing.assert_allclose(got_L, want_L, rtol=1e-12)


def test_hmms_to_h3m_hem_layout():
    rng = np.random.default_rng(0)
    hm = []
    for s in (2, 3, None, 1):
        if s is None:
            hm.append(None)
            continue
        cov = np.stack([np.eye(2) * (k + 1) for k in range(s)])
        hm.append(dict(prior=np.full(s, 1 / s), trans=np.full((s, s), 1 / s), mean=rng.normal(size=(s, 2)), cov=cov,
                       varpar=dict(alpha=1 + rng.uniform(size=s), epsilon=1 + rng.uniform(size=(s, s)),
                                   beta=1 + rng.uniform(size=s))))
    b = h3m.hmms_to_h3m_hem(hm, "full", True, Nv=10.0)
    assert list(b["Sb"]) == [2, 3, 1, 1] and b["prior"].shape == (4, 3)
    assert b["Ntilde"][2] == 0.0 and np.isclose(b["Ntilde"].sum(), 40.0)       # dummy HMM has weight 0
    a = hm[0]["varpar"]["alpha"]
    from scipy.special import psi
    np.testing.assert_allclose(b["prior"][0, :2], np.exp(psi(a) - psi(a.sum())))
    assert (b["A"][0, :2, :2].sum(1) < 1).all()                                  # sub-stochastic on purpose
    tb = (hm[1]["varpar"]["beta"] + 1) / hm[1]["varpar"]["beta"]
    np.testing.assert_allclose(b["cov"][1, :3], tb[:, None, None] * hm[1]["cov"])


@pytest.mark.gpu
def test_em_run_matches_oracle_em():
    from vbhem_b200.api import BaseSet
    base, _ = synth.make_base_set(120, S=3, d=2, C=3, covmode="full", seed=77, Nv=OPT["Nv"])
    hyp = synth.init_centres_baseem(base, 3, 3, seed=77, opt=dict(synth.DEFAULT_OPT, Nv=OPT["Nv"]))
    want_hyp, want_L = oracle_em(base, hyp, OPT)
    with BaseSet(base) as bs:
        got = h3m.vbhem_h3m_c_step_fc(hyp, bs, OPT)
    assert got["iters"] == len(want_L) and not got["unstable"]
    np.testing.assert_allclose(got["LogLs"], want_L, rtol=1e-9)
    assert np.all(np.diff(got["LogLs"]) > -1e-6 * np.abs(got["LogLs"][:-1]))     # EM: the bound does not decrease
    for k in ("alpha", "eta", "epsilon", "lam", "v", "m", "W"):
        np.testing.assert_allclose(got["hyp"][k], want_hyp[k], rtol=1e-7, atol=1e-9 * np.abs(want_hyp[k]).max(),
                                   err_msg=k)


@pytest.mark.gpu
def test_device_em_run_matches_oracle_em():
    """vbhem_em_run: E-step, lower bound, M-step and the convergence test through the C ABI with the
    hyper-parameters resident in HBM, against the same loop built from the oracle."""
    from vbhem_b200.api import BaseSet
    base, _ = synth.make_base_set(120, S=3, d=2, C=3, covmode="full", seed=77, Nv=OPT["Nv"])
    hyp = synth.init_centres_baseem(base, 3, 3, seed=77, opt=dict(synth.DEFAULT_OPT, Nv=OPT["Nv"]))
    want_hyp, want_L = oracle_em(base, hyp, OPT)
    with BaseSet(base) as bs:
        bs.em_begin(hyp, OPT)
        got = bs.em_run(OPT["tau"], OPT["max_iter"], OPT["minDiff"])
    assert got["iters"] == len(want_L) and not got["unstable"]
    np.testing.assert_allclose(got["LogLs"], want_L, rtol=1e-9)
    for k in ("alpha", "eta", "epsilon", "lam", "v", "m", "W"):
        np.testing.assert_allclose(got["hyp"][k], want_hyp[k], rtol=1e-7, atol=1e-9 * np.abs(want_hyp[k]).max(),
                                   err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("covmode,d", [("full", 2), ("diag", 3), ("full", 3)])
def test_device_lower_bound_and_mstep_match_host(covmode, d):
    """One iteration, term by term: vbhem_em_lower_bound / vbhem_em_mstep against h3m.vbhemh3m_lb /
    vbhem_mstep (numpy) fed with the same E-step, on ragged centres."""
    from vbhem_b200.api import BaseSet
    base, hyp, cen = make_case(N=90, K=4, Sb=4, Sr=4, d=d, covmode=covmode, ragged=True, seed=31, Nv=15.0)
    opt = dict(OPT, m0=tuple([256.0, 192.0, 128.0][:d]), W0=0.005 if covmode == "full" else [0.005, 0.01, 0.02][:d])
    with BaseSet(base) as bs:
        out = bs.estep_hyp(hyp, 4, want_pairs=False)
        fin = bs.finalize(out["buf"], hyp["Sr"], 4)
        pre = dict(logOmega=out["logOmega"], logPi=out["logPi"], logA=out["logA"], logLambdaTilde=out["logLambdaTilde"])
        bs.em_begin(hyp, opt)
        bs.em_estep(4)
        lb = bs.em_lower_bound()[0]
        bs.em_mstep()
        got = bs.em_fetch()
    # the host formulas assume one state count per H3M; evaluate them centre by centre on the ragged set
    want_terms = np.zeros(10)
    K = len(hyp["Sr"])
    for j in range(K):
        s = int(hyp["Sr"][j])
        hj = {k: np.asarray(hyp[k])[j:j + 1, :s] for k in ("eta", "v", "lam")}
        hj.update(alpha=hyp["alpha"][j:j + 1], epsilon=hyp["epsilon"][j:j + 1, :s, :s], m=hyp["m"][j:j + 1, :s],
                  W=hyp["W"][j:j + 1, :s])
        pj = dict(logOmega=pre["logOmega"][j:j + 1], logPi=pre["logPi"][j:j + 1, :s], logA=pre["logA"][j:j + 1, :s, :s],
                  logLambdaTilde=pre["logLambdaTilde"][j:j + 1, :s])
        _, t = h3m.vbhemh3m_lb(0.0, 0.0, fin["Nj"][j:j + 1], pj, hj, opt, covmode)
        want_terms += np.array([t[f"Lt{i}"] for i in range(1, 11)])
    # terms that couple the centres of one H3M (Dirichlet over omega) are not additive: take them from the whole set
    from scipy.special import gammaln
    a0 = opt["alpha0"]
    want_terms[5] = gammaln(K * a0) - K * gammaln(a0) + (a0 - 1) * pre["logOmega"].sum()
    want_terms[7] = gammaln(hyp["alpha"].sum()) - gammaln(hyp["alpha"]).sum() + np.dot(hyp["alpha"] - 1, pre["logOmega"])
    want_terms[0], want_terms[6] = fin["lb"][0], fin["lb"][1]
    np.testing.assert_allclose(lb[:10], want_terms, rtol=1e-10, atol=1e-9)
    want_LB = want_terms[:6].sum() - want_terms[6:].sum()
    np.testing.assert_allclose(lb[10], want_LB, rtol=1e-11)
    np.testing.assert_allclose(lb[11], fin["Nj"].sum(), rtol=1e-12)
    new = h3m.vbhem_mstep(fin, opt, covmode)
    for j in range(K):
        s = int(hyp["Sr"][j])
        np.testing.assert_allclose(got["alpha"][j], new["alpha"][j], rtol=1e-12)
        for k in ("eta", "lam", "v", "m", "W"):
            np.testing.assert_allclose(got[k][j, :s], new[k][j, :s], rtol=1e-10, atol=1e-14, err_msg=k)
        np.testing.assert_allclose(got["epsilon"][j, :s, :s], new["epsilon"][j, :s, :s], rtol=1e-10, err_msg="epsilon")


@pytest.mark.gpu
def test_sharded_device_em_run_world_1():
    """dist.ShardedEStep.em_run (the multi-GPU loop: E-step on the shard, all-reduce, device lower bound and
    M-step) on one rank equals vbhem_em_run."""
    from vbhem_b200.api import BaseSet
    from vbhem_b200.dist import ShardedEStep
    base, _ = synth.make_base_set(120, S=3, d=2, C=3, covmode="full", seed=77, Nv=OPT["Nv"])
    hyp = synth.init_centres_baseem(base, 3, 3, seed=77, opt=dict(synth.DEFAULT_OPT, Nv=OPT["Nv"]))
    with BaseSet(base) as bs:
        bs.em_begin(hyp, OPT)
        want = bs.em_run(OPT["tau"], OPT["max_iter"], OPT["minDiff"])
    sh = ShardedEStep(base, device=0)
    got = sh.em_run(hyp, OPT)
    sh.close()
    assert got["iters"] == want["iters"]
    assert np.array_equal(got["LogLs"], want["LogLs"])
    for k in ("alpha", "eta", "epsilon", "lam", "v", "m", "W"):
        assert np.array_equal(got["hyp"][k], want["hyp"][k]), k


@pytest.mark.gpu
def test_batched_em_runs_equal_separate_runs():
    """vbhem_em_run_groups: several (K,S) candidates over one base set, each run to its own convergence in one
    batched loop, must reproduce every candidate's own vbhem_em_run (trace, iteration count, hyper-parameters)."""
    from vbhem_b200.api import BaseSet
    opt = dict(OPT, max_iter=25, minDiff=1e-4)
    base, _ = synth.make_base_set(150, S=3, d=2, C=3, covmode="full", seed=61, Nv=opt["Nv"])
    hyp, grp, models = synth.model_selection_grid(base, Ks=range(1, 4), Ss=range(2, 4), seed=61,
                                                  opt=dict(synth.DEFAULT_OPT, Nv=opt["Nv"]))
    with BaseSet(base) as bs:
        bs.em_begin(hyp, opt, group=grp)
        got = bs.em_run_groups(opt["tau"], opt["max_iter"], opt["minDiff"])
        assert len(set(int(i) for i in got["iters"])) > 1          # the candidates stop at different iterations
        assert bs.counters()["compactions"] >= 1                   # ... and the stopped ones left the E-step
        for g, (K, S) in enumerate(models):
            sel = np.where(grp == g)[0]
            bs.em_begin({k: v[sel] for k, v in hyp.items()}, opt)
            one = bs.em_run(opt["tau"], opt["max_iter"], opt["minDiff"])
            assert one["iters"] == got["iters"][g], (K, S)
            np.testing.assert_allclose(got["LogLs"][g], one["LogLs"], rtol=1e-12)
            for k in ("alpha", "eta", "epsilon", "lam", "v", "m", "W"):
                np.testing.assert_allclose(got["hyp"][k][sel], one["hyp"][k], rtol=1e-11, atol=1e-300, err_msg=f"{k} {K},{S}")


def test_model_selection_grid_layout():
    """CPU: synth.model_selection_grid stacks every (K,S) candidate, padded to the largest S, one group per model."""
    base, _ = synth.make_base_set(30, S=3, d=2, C=2, covmode="full", seed=4, Nv=10.0)
    hyp, grp, models = synth.model_selection_grid(base, Ks=range(1, 4), Ss=range(1, 4), seed=4)
    assert len(models) == 9 and hyp["Sr"].shape[0] == 3 * (1 + 2 + 3) == grp.shape[0]
    assert hyp["eta"].shape == (18, 3) and hyp["W"].shape == (18, 3, 2, 2)
    for g, (K, S) in enumerate(models):
        sel = np.where(grp == g)[0]
        assert len(sel) == K and (hyp["Sr"][sel] == S).all()
        assert (hyp["eta"][sel, :S] > 0).all() and not hyp["eta"][sel, S:].any()
    assert (np.diff(grp) >= 0).all() and grp[0] == 0 and grp[-1] == 8


@pytest.mark.gpu
@pytest.mark.parametrize("covmode", ["full", "diag"])
def test_device_baseem_initialisation_equals_host(covmode):
    """vbhem_em_begin_baseem (vbhemhmm_init.m:58-100 on the device, several H3Ms at once) against the numpy
    restatement on the same draws; then the batched EM runs proceed from it."""
    from vbhem_b200.api import BaseSet
    base, _ = synth.make_base_set(90, S=4, d=2, C=3, covmode=covmode, seed=5, Nv=OPT["Nv"], ragged=True)
    shapes = [(1, 1), (2, 3), (3, 4), (2, 2)]
    Sr, group, randb, randg, omega = synth.baseem_draws_batch(base, shapes, seed=17)
    opt = dict(synth.DEFAULT_OPT, Nv=OPT["Nv"], tau=5)
    want = synth.baseem_from_draws(base, Sr, randb, randg, omega, opt, group=group)
    with BaseSet(base) as bs:
        bs.em_begin_baseem(Sr, group, randb, randg, omega, opt)
        got = bs.em_fetch()
        run = bs.em_run_groups(5, 6, 1e-5)
        bs.em_begin(want, opt, group=group)
        ref = bs.em_run_groups(5, 6, 1e-5)
    for k in ("alpha", "eta", "epsilon", "m", "v", "lam"):
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    np.testing.assert_allclose(got["W"], want["W"], rtol=1e-13)
    for a, b in zip(run["LogLs"], ref["LogLs"]):
        np.testing.assert_allclose(a, b, rtol=1e-10)
