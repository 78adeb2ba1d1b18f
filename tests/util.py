"""Shared builders for the parity tests (seeded inputs, oracle front-ends)."""
from __future__ import annotations

import numpy as np

from oracle import cref, numpy_ref as nr
from vbhem_b200 import pack, synth


def prepass_natural(hyp: dict, covmode: str) -> dict:
    """numpy_ref.prepass over a natural hyper-parameter dict -> natural centre dict (padded)."""
    K, S = hyp["eta"].shape
    d = hyp["m"].shape[2]
    full = covmode == "full"
    nat = dict(Sr=np.asarray(hyp["Sr"], np.int32), logPi=np.zeros((K, S)), logA=np.zeros((K, S, S)),
               m=np.array(hyp["m"], float), vW=np.zeros((K, S, d, d) if full else (K, S, d)),
               cterm=np.zeros((K, S)), W=np.array(hyp["W"], float), v=np.array(hyp["v"], float),
               logLambdaTilde=np.zeros((K, S)))
    for j in range(K):
        s = int(hyp["Sr"][j])
        p = nr.prepass(dict(eta=hyp["eta"][j, :s], epsilon=hyp["epsilon"][j, :s, :s], m=hyp["m"][j, :s],
                            W=hyp["W"][j, :s], v=hyp["v"][j, :s], lam=hyp["lam"][j, :s]), covmode)
        nat["logPi"][j, :s] = p["logPi"]
        nat["logA"][j, :s, :s] = p["logA"]
        nat["vW"][j, :s] = p["vW"]
        nat["cterm"][j, :s] = p["cterm"]
        nat["logLambdaTilde"][j, :s] = p["logLambdaTilde"]
    nat["logOmega"] = nr.log_omega_tilde(hyp["alpha"])
    return nat


def make_case(N=7, K=3, Sb=4, Sr=3, d=2, covmode="full", ragged=True, seed=5, Nv=100.0, sr_list=None):
    """(natural base, hyper-parameters, natural centres after the numpy pre-pass)."""
    base, _ = synth.make_base_set(N, S=Sb, d=d, C=max(2, K), covmode=covmode, seed=seed, Nv=Nv, ragged=ragged)
    if sr_list is None:
        rng = np.random.Generator(np.random.PCG64(seed + 99))
        sr_list = rng.integers(1, Sr + 1, size=K) if ragged else np.full(K, Sr)
        sr_list[0] = Sr
    opt = dict(synth.DEFAULT_OPT)
    opt["Nv"] = Nv
    opt["m0"] = tuple([256.0, 192.0, 128.0][:d])
    hyp = synth.init_centres_baseem(base, K, Sr, seed=seed, opt=opt, Sr=sr_list)
    cen = prepass_natural(hyp, covmode)
    return base, hyp, cen


def oracle_pairs(base, cen, T, which="port"):
    pb, pc = pack.pack_base(base), pack.pack_centres(cen)
    fn = cref.ref_estep_pairs if which == "ref" else cref.port_estep_pairs
    return pack.pairs_to_natural(fn(pb, pc, T), pb["d"], pb["covmode"])


def oracle_fused(base, cen, T):
    """Literal path of the oracle: per-pair outputs -> responsibilities -> skip-rule reduction."""
    pb, pc = pack.pack_base(base), pack.pack_centres(cen)
    o = cref.port_estep_pairs(pb, pc, T)
    zz = cref.port_zhat(o["L_elbo"], pc["logOmega"], pb["Ntilde"], pc.get("group"))
    S = cref.port_reduce(zz["Z"], o, pc["maxSr"], pb["d"], pb["covmode"])
    S.update(Nj=zz["Nj"], lb=zz["lb"], L_elbo=o["L_elbo"], zhat=zz["zhat"])
    return pack.stats_to_natural(S, pb["d"], pb["covmode"])


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


def assert_pairs_close(got, want, rtol=1e-9, what=""):
    """L_elbo elementwise relative; the statistics relative to the largest entry of the same pair."""
    np.testing.assert_allclose(got["L_elbo"], want["L_elbo"], rtol=rtol, atol=0, err_msg=f"L_elbo {what}")
    for k in ("nu1", "xi", "pr", "mu", "Mu"):
        g, w = got[k], want[k]
        ax = tuple(range(2, w.ndim))
        scale = np.maximum(np.max(np.abs(w), axis=ax, keepdims=True), 1e-300)
        err = np.max(np.abs(g - w) / scale)
        assert err <= rtol, f"{k} {what}: max scaled error {err:.3e} > {rtol}"
