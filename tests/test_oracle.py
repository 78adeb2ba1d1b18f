"""CPU tier: pins the oracles.

* the plain-C restatement (oracle/estep_oracle.c) and the numpy restatement
  (oracle/numpy_ref.py) against the reference's own native code compiled in place
  (oracle/_ref, only where /root/reference exists) and against the committed golden
  fixtures generated from it (tests/golden/, everywhere);
* the MATLAB-only stages (responsibilities, reduction, normalisation) C vs numpy;
* the error ids of the reference gateway through the mx shim.
Tolerances: the reference authors' own cross-implementation checks use 1e-10
(/root/reference/src/vbhem/vbhem_hmm_bwd_fwd_fast.m:234-242); we ask 1e-12.
"""
import glob
import os

import numpy as np
import pytest

from oracle import cref, numpy_ref as nr
from vbhem_b200 import pack
from util import assert_pairs_close, make_case, oracle_fused, oracle_pairs

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
              if not os.path.basename(p).startswith(("vbhmm_", "vbhem_em_")))   # E-step fixtures (others: test_fb.py, test_golden_em.py)
HAVE_REF_SRC = os.path.exists("/root/reference/src/vbhem/vbhem_hmm_bwd_fwd_mex.c")


def load_golden(path):
    z = np.load(path)
    grab = lambda pre: {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
    return grab("base_"), grab("hyp_"), grab("cen_"), grab("out_"), int(z["tau"]), str(z["covmode"])


def numpy_pairs(base, cen, T, covmode):
    N, K = base["prior"].shape[0], cen["logPi"].shape[0]
    S = cen["logPi"].shape[1]
    d = base["mu"].shape[2]
    full = covmode == "full"
    out = dict(L_elbo=np.zeros((N, K)), nu1=np.zeros((N, K, S)), xi=np.zeros((N, K, S, S)), pr=np.zeros((N, K, S)),
               mu=np.zeros((N, K, S, d)), Mu=np.zeros((N, K, S, d, d) if full else (N, K, S, d)))
    for i in range(N):
        b = int(base["Sb"][i])
        for j in range(K):
            r = int(cen["Sr"][j])
            L, nu1, pr, mu, Mu, xi = nr.pair_estep(
                base["prior"][i, :b], base["A"][i, :b, :b], base["mu"][i, :b], base["cov"][i, :b], T,
                cen["m"][j, :r], cen["vW"][j, :r], cen["cterm"][j, :r], cen["logA"][j, :r, :r], cen["logPi"][j, :r],
                covmode)
            out["L_elbo"][i, j] = L
            out["nu1"][i, j, :r] = nu1
            out["xi"][i, j, :r, :r] = xi
            out["pr"][i, j, :r] = pr
            out["mu"][i, j, :r] = mu
            out["Mu"][i, j, :r] = Mu
    return out


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_port_matches_golden(path):
    base, hyp, cen, want, tau, covmode = load_golden(path)
    got = oracle_pairs(base, cen, tau, which="port")
    assert_pairs_close(got, want, rtol=1e-12, what="C port vs golden")


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_numpy_matches_golden(path):
    base, hyp, cen, want, tau, covmode = load_golden(path)
    got = numpy_pairs(base, cen, tau, covmode)
    assert_pairs_close(got, want, rtol=1e-11, what="numpy vs golden")


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_numpy_prepass_matches_golden_inputs(path):
    """The stored centre arrays were produced by numpy_ref.prepass; regenerate and compare (guards drift)."""
    from util import prepass_natural
    base, hyp, cen, want, tau, covmode = load_golden(path)
    again = prepass_natural(hyp, covmode)
    for k in ("logPi", "logA", "cterm", "vW", "logOmega"):
        np.testing.assert_allclose(again[k], cen[k], rtol=1e-13, atol=1e-13)


@pytest.mark.skipif(not cref.have_ref() and not HAVE_REF_SRC, reason="reference build not present")
@pytest.mark.parametrize("covmode,d,ragged,T", [("full", 2, True, 5), ("full", 3, False, 3), ("diag", 2, True, 6),
                                                ("diag", 1, False, 2), ("full", 1, True, 1)])
def test_port_matches_reference_build(covmode, d, ragged, T):
    base, hyp, cen = make_case(N=9, K=4, Sb=5, Sr=4, d=d, covmode=covmode, ragged=ragged, seed=17 + d)
    want = oracle_pairs(base, cen, T, which="ref")
    got = oracle_pairs(base, cen, T, which="port")
    assert_pairs_close(got, want, rtol=1e-12, what="C port vs reference build")


@pytest.mark.skipif(not cref.have_ref() and not HAVE_REF_SRC, reason="reference build not present")
def test_reference_gateway_error_ids():
    """Boundary error behaviour, vbhem_hmm_bwd_fwd_mex.c:335-366."""
    assert cref.ref_call_malformed(0) == (1, "MyToolbox:arrayProduct:nrhs")
    assert cref.ref_call_malformed(1) == (1, "MyToolbox:arrayProduct:nlhs")
    assert cref.ref_call_malformed(2) == (1, "vbhem_mex:invalidinput")
    assert cref.ref_call_malformed(3) == (1, "vbhem_mex:invalidinput")


def test_substochastic_mass_is_preserved():
    """sum(nu_1) = sum(prior_b) and sum(xi) = (T-1) * expected mass: a domain invariant of the recursion."""
    base, hyp, cen = make_case(N=5, K=3, Sb=4, Sr=3, d=2, covmode="full", ragged=False, seed=2)
    T = 6
    o = oracle_pairs(base, cen, T)
    np.testing.assert_allclose(o["nu1"].sum(-1), base["prior"].sum(-1)[:, None] * np.ones((1, 3)), rtol=1e-12)
    # pr = sum_t sum_gamma nu_t ; nu_t mass = prior' A^(t-1) 1
    mass = np.zeros(5)
    for i in range(5):
        v = base["prior"][i].copy()
        for t in range(T):
            mass[i] += v.sum()
            v = v @ base["A"][i]
    np.testing.assert_allclose(o["pr"].sum(-1), mass[:, None] * np.ones((1, 3)), rtol=1e-12)


def test_zhat_and_reduce_c_vs_numpy():
    base, hyp, cen = make_case(N=23, K=5, Sb=3, Sr=3, d=2, covmode="full", ragged=True, seed=9, Nv=10.0)
    T = 4
    pb, pc = pack.pack_base(base), pack.pack_centres(cen)
    o = cref.port_estep_pairs(pb, pc, T)
    zz = cref.port_zhat(o["L_elbo"], pc["logOmega"], pb["Ntilde"])
    zhat, Z, Nj = nr.responsibilities(o["L_elbo"].T, cen["logOmega"], base["Ntilde"])
    np.testing.assert_allclose(zz["zhat"].T, zhat, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(zz["Nj"], Nj, rtol=1e-12)
    np.testing.assert_allclose(zz["lb"][0], np.sum(Z * o["L_elbo"].T), rtol=1e-12)
    np.testing.assert_allclose(zz["lb"][1], np.sum(zhat * np.log(zhat)), rtol=1e-12)
    assert (Z <= 1e-8).any() and (Z > 1e-8).any(), "case must exercise the Z > 1e-8 skip rule"
    S = pack.stats_to_natural(cref.port_reduce(zz["Z"], o, pc["maxSr"], pb["d"], pb["covmode"]), 2, 1)
    nat = pack.pairs_to_natural(o, 2, 1)
    for j in range(5):
        a = nr.reduce_raw(Z[:, j], nat["nu1"][:, j], nat["xi"][:, j], nat["pr"][:, j], nat["mu"][:, j],
                          nat["Mu"][:, j])
        for got, want in zip((S["nu1"][j], S["xi"][j], S["pr"][j], S["mu"][j], S["Mu"][j]), a):
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300)


def test_fused_oracle_equals_literal_oracle():
    base, hyp, cen = make_case(N=31, K=4, Sb=4, Sr=4, d=2, covmode="full", ragged=True, seed=21, Nv=20.0)
    pb, pc = pack.pack_base(base), pack.pack_centres(cen)
    lit = oracle_fused(base, cen, 5)
    fus = pack.stats_to_natural(cref.port_estep_fused(pb, pc, 5), 2, 1)
    for k in ("Nj", "nu1", "xi", "pr", "mu", "Mu", "lb", "zhat", "L_elbo"):
        np.testing.assert_allclose(fus[k], lit[k], rtol=1e-11, atol=1e-300, err_msg=k)


def test_grouped_softmax():
    """Two H3Ms batched in one call (the (K,S) grid): each group normalises on its own."""
    base, hyp, cen = make_case(N=11, K=5, Sb=3, Sr=3, d=2, covmode="full", ragged=False, seed=4, Nv=5.0)
    pb, pc = pack.pack_base(base), pack.pack_centres(cen)
    L = cref.port_estep_pairs(pb, pc, 3, lonly=True)["L_elbo"]
    grp = np.array([0, 0, 1, 1, 1], np.int32)
    zz = cref.port_zhat(L, pc["logOmega"], pb["Ntilde"], grp)
    np.testing.assert_allclose(zz["zhat"][:2].sum(0), 1.0, rtol=1e-12)
    np.testing.assert_allclose(zz["zhat"][2:].sum(0), 1.0, rtol=1e-12)
    z0, _, _ = nr.responsibilities(L[:2].T, cen["logOmega"][:2], base["Ntilde"])
    np.testing.assert_allclose(zz["zhat"][:2].T, z0, rtol=1e-12)


def test_finalize_mstep_lb_run():
    """Host-side restatements used as the oracle of the EM iteration: shapes, symmetry, finiteness."""
    base, hyp, cen = make_case(N=15, K=2, Sb=3, Sr=3, d=2, covmode="full", ragged=False, seed=8, Nv=10.0)
    S = oracle_fused(base, cen, 4)
    opt = dict(alpha0=1.0, eta0=1.0, epsilon0=1.0, lambda0=1.0, v0=5.0, W0=0.005, m0=(256.0, 192.0))
    centres, pre = [], []
    for j in range(2):
        f = nr.finalize_stats(S["nu1"][j], S["xi"][j], S["pr"][j], S["mu"][j], S["Mu"][j])
        new = nr.mstep_component(*f, opt)
        assert np.allclose(new["W"], np.transpose(new["W"], (0, 2, 1)))
        assert np.all(np.linalg.eigvalsh(new["W"]) > 0)
        centres.append(dict(eta=hyp["eta"][j], epsilon=hyp["epsilon"][j], m=hyp["m"][j], W=hyp["W"][j],
                            v=hyp["v"][j], lam=hyp["lam"][j]))
        pre.append(nr.prepass(centres[-1]))
    LB, terms = nr.lower_bound(S["lb"][0], S["lb"][1], S["Nj"], cen["logOmega"], hyp["alpha"], centres, pre, opt)
    assert np.isfinite(LB) and all(np.isfinite(v) for v in terms.values())
