"""Packed array layouts of the E-step boundary (host side).

The C ABI (include/vbhem_b200.h) takes fp64 **column-major** arrays, the layout a
MATLAB MEX gateway hands over without copies.  Python callers use *natural*
numpy arrays (C order, leading axis = HMM index):

  base set  (reference type ``h3m_b``, /root/reference/src/vbhem/hmms_to_h3m_hem.m:19-140)
      Sb[N] int, prior[N,S], A[N,S,S] (from,to), mu[N,S,d],
      cov[N,S,d,d] (full) or cov[N,S,d] (diag), Ntilde[N]
  centres   (reference type ``h3m_r`` after the pre-pass,
             /root/reference/src/vbhem/vbhem_h3m_c_step_fc.m:118-165,180-191)
      Sr[K] int, logPi[K,S], logA[K,S,S] (from,to), m[K,S,d],
      vW[K,S,d,d] | [K,S,d] (= v*W), cterm[K,S] (= -logLambdaTilde + d/lambda),
      logOmega[K]; optional W, v (only the literal MEX signature needs them)

A numpy C-order array whose axes are the column-major axes reversed has exactly
the column-major memory image, so packing is a transpose of the state axes.
States beyond an HMM's own count (ragged sets) are zero padding.
"""
from __future__ import annotations

import numpy as np


def _c(a, dt=np.float64):
    return np.ascontiguousarray(a, dtype=dt)


def pack_base(nat: dict) -> dict:
    """natural base set -> packed dict (memory = column-major arrays of the C ABI)."""
    prior = _c(nat["prior"])
    N, S = prior.shape
    mu = _c(nat["mu"])
    d = mu.shape[2]
    cov = np.asarray(nat["cov"], dtype=np.float64)
    covmode = 1 if cov.ndim == 4 else 0
    Sb = _c(nat.get("Sb", np.full(N, S)), np.int32)
    Nt = _c(nat.get("Ntilde", np.ones(N)))
    return dict(N=N, maxSb=S, d=d, covmode=covmode, Sb=Sb, prior=prior,
                A=_c(np.transpose(nat["A"], (0, 2, 1))), mu=mu,
                cov=_c(np.transpose(cov, (0, 1, 3, 2))) if covmode else _c(cov), Ntilde=Nt)


def pack_centres(nat: dict) -> dict:
    """natural centre set -> packed dict."""
    logPi = _c(nat["logPi"])
    K, S = logPi.shape
    m = _c(nat["m"])
    vW = np.asarray(nat["vW"], dtype=np.float64)
    covmode = 1 if vW.ndim == 4 else 0
    tr = (lambda a: _c(np.transpose(a, (0, 1, 3, 2)))) if covmode else _c
    out = dict(K=K, maxSr=S, d=m.shape[2], covmode=covmode,
               Sr=_c(nat.get("Sr", np.full(K, S)), np.int32), logPi=logPi,
               logA=_c(np.transpose(nat["logA"], (0, 2, 1))), m=m, vW=tr(vW),
               cterm=_c(nat["cterm"]), logOmega=_c(nat.get("logOmega", np.zeros(K))))
    if "W" in nat:
        out["W"] = tr(np.asarray(nat["W"], dtype=np.float64))
        out["v"] = _c(nat["v"])
    if "group" in nat:
        out["group"] = _c(nat["group"], np.int32)
    return out


def pairs_to_natural(o: dict, d: int, covmode: int) -> dict:
    """packed per-pair outputs -> natural: L_elbo[N,K], nu1[N,K,S], xi[N,K,S,S] (from,to),
    pr[N,K,S], mu[N,K,S,d], Mu[N,K,S,d,d] | [N,K,S,d]."""
    K, N = o["L_elbo"].shape
    S = o["nu1"].shape[2]
    out = dict(L_elbo=o["L_elbo"].T.copy(),
               nu1=np.transpose(o["nu1"], (1, 0, 2)).copy(),
               xi=np.transpose(o["xi"], (1, 0, 3, 2)).copy(),
               pr=np.transpose(o["pr"], (1, 0, 2)).copy(),
               mu=np.transpose(o["mu"], (1, 0, 3, 2)).copy())
    if covmode:
        Mu = o["Mu"].reshape(K, N, d, d, S)            # [K,N,l,k,S] (column-major k + d*l)
        out["Mu"] = np.transpose(Mu, (1, 0, 4, 3, 2)).copy()
    else:
        out["Mu"] = np.transpose(o["Mu"], (1, 0, 3, 2)).copy()
    return out


def stats_to_natural(S: dict, d: int, covmode: int) -> dict:
    """packed per-centre raw sums -> natural: nu1[K,S], xi[K,S,S] (from,to), pr[K,S], mu[K,S,d],
    Mu[K,S,d,d] | [K,S,d] (+ passthrough of Nj, lb, L_elbo[N,K], zhat[N,K])."""
    K, Sr = S["nu1"].shape
    out = dict(nu1=S["nu1"].copy(), xi=np.transpose(S["xi"], (0, 2, 1)).copy(), pr=S["pr"].copy(),
               mu=np.transpose(S["mu"], (0, 2, 1)).copy())
    if covmode:
        out["Mu"] = np.transpose(S["Mu"].reshape(K, d, d, Sr), (0, 3, 2, 1)).copy()
    else:
        out["Mu"] = np.transpose(S["Mu"], (0, 2, 1)).copy()
    for k in ("Nj", "lb"):
        if k in S:
            out[k] = S[k].copy()
    for k in ("L_elbo", "zhat"):
        if S.get(k) is not None:
            out[k] = S[k].T.copy()
    return out
