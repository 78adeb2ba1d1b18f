"""Hyper-parameter learning around the EM run (h3m.vbhem_h3m_c_hyp; /root/reference/src/vbhem/vbhem_h3m_c_hyp.m,
vbhemh3m_lb.m:202-343, vbhem_get_hypinfo.m, vbhem_clip_hyps.m).

CPU: the analytic derivatives of the lower bound with respect to the prior hyper-parameters against central
differences of the bound itself (h3m.vbhemh3m_lb, which the EM-loop tests pin on the oracle) at fixed variational
posteriors; transforms and clipping.  GPU: the optimiser raises the converged bound."""
import numpy as np
import pytest

import vbhem_b200  # noqa: F401
from vbhem_b200 import h3m, synth


def _case(W0, covmode="full", seed=3):
    rng = np.random.default_rng(seed)
    base, _ = synth.make_base_set(30, S=3, d=2, C=2, seed=seed, covmode=covmode)
    opt = dict(synth.DEFAULT_OPT, W0=W0, alpha0=0.7, eta0=1.3, epsilon0=0.9, lambda0=1.4, v0=6.0)
    hyp = synth.init_centres_baseem(base, 3, 3, seed=1, opt=opt)
    K, S = hyp["eta"].shape
    pre = dict(logOmega=rng.normal(-1, 0.3, K), logPi=rng.normal(-1, 0.3, (K, S)), logA=rng.normal(-1, 0.3, (K, S, S)),
               logLambdaTilde=rng.normal(-5, 0.3, (K, S)))
    return opt, hyp, pre, rng.uniform(1, 5, K)


@pytest.mark.parametrize("W0", [0.005, [0.004, 0.007]])
@pytest.mark.parametrize("covmode", ["full", "diag"])
def test_derivatives_match_central_differences_of_the_bound(W0, covmode):
    opt, hyp, pre, Nj = _case(W0, covmode)
    f = lambda o: h3m.vbhemh3m_lb(-100.0, -3.0, Nj, pre, hyp, o, covmode)[0]
    D = h3m.vbhemh3m_lb_deriv(pre, hyp, opt, covmode)
    for name in ("alpha0", "eta0", "epsilon0", "v0", "lambda0"):
        h = 1e-6 * opt[name]
        num = (f(dict(opt, **{name: opt[name] + h})) - f(dict(opt, **{name: opt[name] - h}))) / (2 * h)
        np.testing.assert_allclose(D["raw"][name][0], num, rtol=2e-6, err_msg=name)
    w = np.atleast_1d(np.asarray(W0, float))
    cv = lambda x: float(x[0]) if w.size == 1 else x
    for i in range(w.size):
        h = 1e-6 * w[i]
        wp, wm = w.copy(), w.copy()
        wp[i] += h
        wm[i] -= h
        num = (f(dict(opt, W0=cv(wp))) - f(dict(opt, W0=cv(wm)))) / (2 * h)
        np.testing.assert_allclose(D["raw"]["W0"][i], num, rtol=2e-6, err_msg="W0")
    for i in range(2):
        mp, mm = np.array(opt["m0"], float), np.array(opt["m0"], float)
        mp[i] += 1e-4
        mm[i] -= 1e-4
        num = (f(dict(opt, m0=mp)) - f(dict(opt, m0=mm))) / 2e-4
        np.testing.assert_allclose(D["raw"]["m0"][i], num, rtol=2e-6, err_msg="m0")
    # chain rule into the optimiser's coordinates (vbhemh3m_lb.m:330-339)
    np.testing.assert_allclose(D["d_logalpha0"], D["raw"]["alpha0"] * opt["alpha0"])
    np.testing.assert_allclose(D["d_logv0D1"], D["raw"]["v0"] * (opt["v0"] - 2 + 1))
    np.testing.assert_allclose(D["d_sqrtW0inv"], D["raw"]["W0"] * w ** 1.5 * -2)
    np.testing.assert_allclose(D["d_logW0"], D["raw"]["W0"] * w)


def test_transforms_round_trip_and_clipping():
    opt = dict(synth.DEFAULT_OPT, W0=[0.004, 0.007], v0=6.0)
    for h in h3m.vbhem_get_hypinfo(["alpha0", "eta0", "epsilon0", "v0", "lambda0", "W0", "W0log", "m0"], opt):
        x = np.atleast_1d(np.asarray(opt[h["optname"]], float))
        np.testing.assert_allclose(h["hyptrans"](h["hypinvtrans"](x)), x, rtol=1e-14)
        assert h["hypdims"] == x.size
    with pytest.raises(ValueError):
        h3m.vbhem_get_hypinfo(["nope"], opt)
    o, c = h3m.vbhem_clip_hyps(dict(opt, alpha0=1e20, lambda0=1e-20, W0=[1e-20, 0.007]))
    assert o["alpha0"] == h3m.HYPS_MAX["alpha0"] and c["alpha0"].tolist() == [1]
    assert o["lambda0"] == h3m.HYPS_MIN["lambda0"] and c["lambda0"].tolist() == [-1]
    assert c["W0"].tolist() == [-1, 0] and o["W0"][0] == h3m.HYPS_MIN["W0"] and c["v0"].tolist() == [0]
    # a derivative that would push a clipped hyper-parameter further out is zeroed (vbhemh3m_lb.m:312-327)
    optc, hyp, pre, _ = _case(0.005)
    g0 = h3m.vbhemh3m_lb_deriv(pre, hyp, optc)["raw"]["alpha0"][0]
    g1 = h3m.vbhemh3m_lb_deriv(pre, hyp, optc, clipped=dict(alpha0=[+1 if g0 > 0 else -1]))["raw"]["alpha0"][0]
    assert g0 != 0.0 and g1 == 0.0


@pytest.mark.gpu
def test_hyper_parameter_learning_raises_the_bound():
    from vbhem_b200 import BaseSet
    base, _ = synth.make_base_set(120, S=3, d=2, C=3, seed=12, Nv=20.0)
    opt = dict(synth.DEFAULT_OPT, Nv=20.0, max_iter=60, minDiff=1e-6)
    init = synth.init_centres_baseem(base, 3, 3, seed=2, opt=opt)
    with BaseSet(base) as bs:
        plain = h3m.vbhem_h3m_c_step_fc(init, bs, opt)
        one = h3m.vbhem_h3m_c_step_fc(init, bs, dict(opt, calc_LLderiv=1))
        learnt = h3m.vbhem_h3m_c_hyp(init, bs, opt, max_evals=25)
    assert one["LL"] == plain["LL"] and set(one["dLL"]) >= {"d_logalpha0", "d_m0", "d_sqrtW0inv"}
    info = learnt["learn_hyps"]
    assert np.isfinite(learnt["LL"]) and learnt["LL"] > plain["LL"]
    assert abs(info["evals"][0] - plain["LL"]) <= 1e-9 * abs(plain["LL"])      # first evaluation = the given hyps
    assert info["hypnames"] == ["alpha0", "eta0", "epsilon0", "v0", "lambda0", "W0", "m0"]
    assert info["vbopt"]["v0"] > 1.0 and np.all(np.asarray(info["vbopt"]["W0"]) > 0)


@pytest.mark.gpu
def test_cluster_with_learn_hyps():
    """vbhem_h3m_cluster(..., learn_hyps=[...]): the chosen model's bound is at least the fixed-hyps one and the
    learnt hyper-parameters come back in vbhemopt."""
    base, lab = synth.make_base_set(60, S=2, d=2, C=2, seed=21, Nv=20.0)
    common = dict(trials=3, seed=7, Nv=20.0, max_iter=50)
    fixed = h3m.vbhem_h3m_cluster(base, [1, 2], 2, common)
    learnt = h3m.vbhem_h3m_cluster(base, [1, 2], 2, dict(common, learn_hyps=["alpha0", "v0", "lambda0"]))
    assert learnt["LL"] >= fixed["LL"] and "learn_hyps" in learnt
    assert learnt["learn_hyps"]["hypnames"] == ["alpha0", "v0", "lambda0"]
    assert learnt["vbhemopt"]["v0"] != fixed["vbhemopt"]["v0"]
    assert learnt["Z"].shape == (60, learnt["K"])


@pytest.mark.gpu
@pytest.mark.parametrize("W0", [0.005, (0.004, 0.007)])
def test_device_derivatives_match_host_formulas(W0):
    """k_lb_deriv (vbhemh3m_lb.m:202-309 on the device, one block per batched H3M, each under its OWN prior) against
    the host restatement ``vbhemh3m_lb_deriv`` evaluated group by group."""
    from vbhem_b200.api import BaseSet
    base, _ = synth.make_base_set(80, S=3, d=2, C=3, seed=4, Nv=15.0, ragged=True)
    opt = dict(synth.DEFAULT_OPT, Nv=15.0, tau=5, W0=W0)
    models = [synth.init_centres_baseem(base, k, s, seed=30 + k, opt=opt) for k, s in ((2, 3), (3, 2), (1, 1))]
    opts = [dict(opt, alpha0=1.0 + 0.5 * g, eta0=0.7 + g, epsilon0=1.3, lambda0=0.5 + g, v0=4.0 + g,
                 m0=(250.0 + 10 * g, 190.0 - 5 * g), W0=np.asarray(W0) * (1 + g)) for g in range(3)]
    hyp, group = h3m.stack_h3ms(models)
    with BaseSet(base) as bs:
        bs.em_begin(hyp, opt, group=group)
        bs.em_set_group_priors(opts)
        bs.em_estep(5)
        lb = bs.em_lower_bound()
        got = bs.em_lb_deriv()
        for g, (mdl, o) in enumerate(zip(models, opts)):
            st = bs.estep_hyp(mdl, 5)
            pre = {k: st[k] for k in ("logOmega", "logPi", "logA", "logLambdaTilde")}
            want = h3m.vbhemh3m_lb_deriv(pre, mdl, o)["raw"]
            row = h3m.device_derivs_to_dict(got[g], 2, np.atleast_1d(W0).size)
            for k in want:
                np.testing.assert_allclose(row[k], want[k], rtol=1e-9, atol=1e-9 * max(1.0, np.abs(want[k]).max()), err_msg=f"{g}:{k}")
            fin = bs.finalize(st["buf"], mdl["Sr"], mdl["eta"].shape[1])
            L, _ = h3m.vbhemh3m_lb(st["lb"][0], st["lb"][1], fin["Nj"], pre, mdl, o)
            np.testing.assert_allclose(lb[g, 10], L, rtol=1e-10)          # the bound under the group's own prior


@pytest.mark.gpu
def test_batched_hyper_parameter_learning():
    """vbhem_h3m_c_hyp_batch: the optimisers of several trials in lock-step, every round of evaluations one batched EM
    loop; each trial's bound rises, and a trial's result does not depend on what it is batched with."""
    from vbhem_b200.api import BaseSet
    base, _ = synth.make_base_set(60, S=2, d=2, C=2, seed=21, Nv=20.0)
    opt = dict(h3m.CLUSTER_DEFAULTS, Nv=20.0, tau=5, max_iter=40, seed=3, learn_hyps=["alpha0", "v0", "lambda0", "m0"])
    inits = [synth.init_centres_baseem(base, k, 2, seed=50 + k, opt=opt) for k in (1, 2, 3)]
    with BaseSet(base) as bs:
        plain = []
        for h0 in inits:
            bs.em_begin(h0, opt)
            plain.append(bs.em_run(5, 40, 1e-5)["LogLs"][-1])
        together = h3m.vbhem_h3m_c_hyp_batch(inits, bs, opt, max_evals=15)
        alone = h3m.vbhem_h3m_c_hyp_batch(inits[1:2], bs, opt, max_evals=15)[0]
    for res, L0 in zip(together, plain):
        assert abs(res["learn_hyps"]["evals"][0] - L0) <= 1e-9 * abs(L0)      # first evaluation = the given hyps
        assert res["LL"] > L0
        assert res["learn_hyps"]["hypnames"] == ["alpha0", "v0", "lambda0", "m0"]
    np.testing.assert_allclose(alone["LL"], together[1]["LL"], rtol=1e-12)
    np.testing.assert_allclose(alone["learn_hyps"]["opt_transhyp"], together[1]["learn_hyps"]["opt_transhyp"], rtol=1e-9)
