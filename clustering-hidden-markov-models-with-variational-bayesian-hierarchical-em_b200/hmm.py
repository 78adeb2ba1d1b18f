"""Host mirror of the toolbox's VB-HMM learning entry points around the GPU forward-backward pass.

``vbhmm_em(data, K, ini)`` follows /root/reference/src/hmm/vbhmm_em.m:1-496 and ``vbhmm_learn(data, K, vbopt)``
the selection logic of /root/reference/src/hmm/vbhmm_learn.m:360-620 (candidate state counts, random restarts,
best lower bound), with the reference's option names (``alpha0``, ``epsilon0``, ``mu0``, ``beta0``, ``W0``, ``v0``,
``maxIter``, ``minDiff``, ``initmode``, ``initgmm``, ``inithmm``, ``numTrials``, ``seed``).

Every EM iteration is ONE call of the fused pass ``vbhmm_fb_stats`` (csrc/vbhmm_fb.cu): forward-backward, the sums
vbhmm_em.m:158-246 takes from it and the two sequence-sized lower-bound terms (vbhmm_em_lb.m:214,217) are reduced on
the GPU and 536 bytes come back; what stays on the host is O(K * dim^2): the M-step (vbhmm_em.m:353-414) and the
lower bound (vbhmm_em_lb.m:74-258).  The observations are packed into HBM once per run (``SequenceSet``).

The learned model is a dict with the reference's field names (vbhmm_em.m:426-491), arrays in natural layout:
``prior[K]``, ``trans[K(from),K(to)]``, ``mean[K,dim]``, ``cov[K,dim,dim]`` (the reference's ``pdf{k}.mean/.cov``
stacked), ``LL``, ``M``, ``N1``, ``N``, ``varpar = {alpha, epsilon, beta, v, m, W}``, plus ``iter`` and ``LLs``.  It is
what ``h3m.hmms_to_h3m_hem`` takes, so ``vbhmm_learn`` -> ``hmms_to_h3m_hem`` -> ``vbhem_h3m_c_step_fc`` is the
toolbox's pipeline (learn one HMM per subject, then cluster the HMMs).

Not mirrored: groups, ``fix_cov`` / ``fix_clusters``, hyper-parameter learning (``learn_hyps``, vbhmm_em_hyp.m) and
its derivatives; ``gamma`` (the per-sequence responsibilities, vbhmm_em.m:463-466) is not materialised -- the unfused
pass ``fb.vbhmm_fb(data, hmm['varpar'])`` returns it.  There is no CPU implementation behind this module.
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammaln

import ctypes as C

from . import fb as _fb
from ._lib import check, dp, ip, lib

DEFAULT_VBOPT = dict(alpha0=0.1, epsilon0=0.1, mu0=(256.0, 192.0), beta0=1.0, W0=0.005, v0=5.0, maxIter=100,
                     minDiff=1e-5, initmode="random", numTrials=50, seed=None, verbose=0)   # vbhmm_learn.m:232-300


def _W0inv(ini: dict, dim: int) -> tuple[np.ndarray, bool]:
    w = np.atleast_1d(np.asarray(ini["W0"], float)).ravel()
    if w.size == 1:
        return np.eye(dim) / w[0], True                                    # vbhmm_init.m:136-140,152
    if w.size != dim:
        raise ValueError(f"vbopt.W0 should have dimension D={dim} for diagonal matrix")
    return np.diag(1.0 / w), False                                         # :141-148


def _fit_gmm(X: np.ndarray, K: int, seed) -> dict:
    """Stand-in for ``gmdistribution.fit(data, K, 'Options', struct('TolFun', 1e-5), 'Start', 'randSample')``
    (vbhmm_init.m:60-61; Statistics Toolbox, not reproducible outside MATLAB): scikit-learn's EM from random points."""
    import warnings
    from sklearn.exceptions import ConvergenceWarning
    from sklearn.mixture import GaussianMixture
    with warnings.catch_warnings():   # warning('off', 'stats:gmdistribution:FailedToConverge'), vbhmm_init.m:47
        warnings.simplefilter("ignore", ConvergenceWarning)
        g = GaussianMixture(n_components=K, covariance_type="full", tol=1e-5, init_params="random_from_data",
                            random_state=seed, reg_covar=1e-10).fit(X)
    return dict(prior=g.weights_, mean=g.means_, cov=g.covariances_)


def vbhmm_init(data, K: int, ini: dict) -> dict:
    """vbhmm_init.m:1-216 -> initial variational parameters ``alpha, epsilon, beta, v, m[K,dim], W[K,dim,dim]``.
    initmode ``'initgmm'`` (ini['initgmm'] = dict(prior[K], mean[K,dim], cov[K,dim,dim])), ``'inithmm'``
    (ini['inithmm']['varpar']) or ``'random'`` (a GMM fitted from random points; K = 1 and N <= K as at :25-41).
    ``data``: the sequences, or ``(N, dim)`` = number of observations and their dimension when the mode does not
    look at them ('initgmm', 'inithmm')."""
    mode = ini.get("initmode", "random")
    if isinstance(data, tuple):
        (N, dim), X = data, None
        if mode == "random":
            raise ValueError("initmode 'random' needs the observations")
    else:
        X = np.concatenate([np.asarray(x, float).reshape(len(x), -1) for x in data if len(x)])
        N, dim = X.shape
    if len(np.atleast_1d(ini["mu0"])) != dim:
        raise ValueError(f"vbopt.mu0 should have dimension D={dim}")
    if ini["v0"] <= dim - 1:
        raise ValueError("v0 not large enough")
    if mode == "inithmm":
        vp = ini["inithmm"]["varpar"]
        return {k: np.array(vp[k], float) for k in ("alpha", "epsilon", "beta", "v", "m", "W")}
    if mode == "initgmm":
        g = ini["initgmm"]
        gmm = dict(prior=np.asarray(g["prior"], float).ravel(), mean=np.asarray(g["mean"], float).reshape(-1, dim),
                   cov=np.asarray(g["cov"], float).reshape(-1, dim, dim))
        if gmm["mean"].shape[0] != K:
            raise ValueError("bad initgmm dimensions -- possibly mean is not a row vector")
    elif mode == "random":
        if K == 1:
            gmm = dict(prior=np.ones(1), mean=X.mean(0, keepdims=True), cov=np.atleast_2d(np.cov(X.T))[None])
        elif N <= K:
            t = np.r_[np.ones(N), 1e-6 * np.ones(K - N)]
            gmm = dict(prior=t / t.sum(), mean=np.r_[X, np.zeros((K - N, dim))],
                       cov=np.tile(np.eye(dim) * X.var(0, ddof=1).mean(), (K, 1, 1)))
        else:
            gmm = _fit_gmm(X, K, ini.get("seed"))
    else:
        raise ValueError("bad initmode")
    m0 = np.asarray(ini["mu0"], float).reshape(dim)
    W0inv, _ = _W0inv(ini, dim)
    Nk, Nk2 = N * gmm["prior"], np.full(K, N / K)                          # :166-167
    beta = ini["beta0"] + Nk
    W = np.zeros((K, dim, dim))
    for k in range(K):                                                     # :197-201
        d3 = gmm["mean"][k] - m0
        W[k] = np.linalg.inv(W0inv + Nk[k] * gmm["cov"][k] + ini["beta0"] * Nk[k] / (ini["beta0"] + Nk[k]) * np.outer(d3, d3))
    return dict(alpha=ini["alpha0"] + Nk2, epsilon=np.tile(ini["epsilon0"] + Nk2, (K, 1)), beta=beta,
                v=ini["v0"] + Nk + 1, m=(ini["beta0"] * m0[None] + Nk[:, None] * gmm["mean"]) / beta[:, None], W=W)


def vbhmm_init_from_gmms(gmms: list, npts, K: int, ini: dict) -> list:
    """``vbhmm_init`` (initmode 'initgmm', vbhmm_init.m:166-201) for many runs at once: ``gmms[r]`` = dict(prior, mean,
    cov), ``npts[r]`` = the run's number of observations.  Same numbers as one ``vbhmm_init`` call per run."""
    R = len(gmms)
    pr = np.stack([np.asarray(g["prior"], float) for g in gmms])                        # [R, K]
    mu = np.stack([np.asarray(g["mean"], float) for g in gmms])                         # [R, K, dim]
    cv = np.stack([np.asarray(g["cov"], float) for g in gmms])                          # [R, K, dim, dim]
    dim = mu.shape[2]
    n = np.asarray(npts, float).reshape(R, 1)
    m0 = np.asarray(ini["mu0"], float).reshape(dim)
    W0inv, _ = _W0inv(ini, dim)
    b0 = float(ini["beta0"])
    Nk, Nk2 = n * pr, n / K
    beta = b0 + Nk
    d3 = mu - m0
    W = np.linalg.inv(W0inv + Nk[..., None, None] * cv + (b0 * Nk / (b0 + Nk))[..., None, None] * d3[..., :, None] * d3[..., None, :])
    alpha = ini["alpha0"] + np.broadcast_to(Nk2, (R, K))
    eps = ini["epsilon0"] + np.broadcast_to(Nk2[:, :, None], (R, K, K))
    m = (b0 * m0 + Nk[..., None] * mu) / beta[..., None]
    v = ini["v0"] + Nk + 1
    return [dict(alpha=alpha[r].copy(), epsilon=eps[r].copy(), beta=beta[r], v=v[r], m=m[r], W=W[r]) for r in range(R)]


def vbhmm_mstep(st: dict, ini: dict) -> dict:
    """vbhmm_em.m:353-389 from the sums of one fused pass (``fb.fb_stats``)."""
    K, dim = st["xbar"].shape
    m0 = np.asarray(ini["mu0"], float).reshape(dim)
    W0inv, _ = _W0inv(ini, dim)
    Nk, b0 = st["Nk"], float(ini["beta0"])
    beta = b0 + Nk
    dx = st["xbar"] - m0[None]
    Winv = W0inv[None] + Nk[:, None, None] * st["S"] + (b0 * Nk / (b0 + Nk))[:, None, None] * dx[:, :, None] * dx[:, None, :]
    W = np.linalg.inv(Winv)
    return dict(alpha=ini["alpha0"] + st["Nk1"], epsilon=ini["epsilon0"] + st["M"], beta=beta, v=ini["v0"] + Nk + 1,
                m=(b0 * m0[None] + Nk[:, None] * st["xbar"]) / beta[:, None], W=(W + np.transpose(W, (0, 2, 1))) / 2)


def vbhmm_em_lb(st: dict, vp: dict, ini: dict) -> tuple[float, dict]:
    """vbhmm_em_lb.m:74-258 with the sequence-sized terms taken from the fused pass: Lt2a = sum_k Nk1(k) logPiTilde(k)
    (:143-146), Lt2b = sum M .* logATilde (:162), Lt63, Lt64 (:214,217).  ``vp`` = the parameters the pass ran with."""
    v, W, eps, al, m, beta = (np.asarray(vp[k], float) for k in ("v", "W", "epsilon", "alpha", "m", "beta"))
    K, dim = m.shape
    a0, e0, b0, v0 = (float(ini[k]) for k in ("alpha0", "epsilon0", "beta0", "v0"))
    m0 = np.asarray(ini["mu0"], float).reshape(dim)
    W0inv, iid = _W0inv(ini, dim)
    logdetW0inv = dim * np.log(W0inv[0, 0]) if iid else np.log(np.diag(W0inv)).sum()
    lLT, logPi, logA = st["logLambdaTilde"], st["logPiTilde"], st["logATilde"]
    half = 0.5 * (1 - np.arange(1, dim + 1))
    lnB = lambda vv, logdetW: -(vv / 2) * logdetW - (vv * dim / 2) * np.log(2) - (dim * (dim - 1) / 4) * np.log(np.pi) \
        - gammaln(0.5 * vv + half).sum()
    logdetW = np.linalg.slogdet(W)[1]
    H = sum(-lnB(v[k], logdetW[k]) - 0.5 * (v[k] - dim - 1) * lLT[k] + 0.5 * v[k] * dim for k in range(K))
    trSW = np.einsum("kij,kji->k", st["S"], W)
    dx, dm = st["xbar"] - m, m - m0[None]
    xWx = np.einsum("ki,kij,kj->k", dx, W, dx)
    mWm = np.einsum("ki,kij,kj->k", dm, W, dm)
    trW0W = np.einsum("ij,kji->k", W0inv, W)
    Nk = st["Nk"]
    T = {}
    T["Lt1"] = 0.5 * (Nk * (lLT - dim / beta - v * trSW - v * xWx - dim * np.log(2 * np.pi))).sum()
    Lt2a, Lt2b = (st["Nk1"] * logPi).sum(), (st["M"] * logA).sum()
    T["Lt2"] = Lt2a + Lt2b
    T["Lt3"] = gammaln(K * a0) - K * gammaln(a0) + (a0 - 1) * logPi.sum()
    T["Lt4"] = K * (gammaln(K * e0) - K * gammaln(e0)) + (e0 - 1) * logA.sum()
    T["Lt5"] = 0.5 * (dim * np.log(b0 / (2 * np.pi)) + lLT - dim * b0 / beta - b0 * v * mWm).sum() \
        + K * lnB(v0, -logdetW0inv) + 0.5 * (v0 - dim - 1) * lLT.sum() - 0.5 * (v * trW0W).sum()
    T["Lt6"] = Lt2a + Lt2b + st["Lt63"] - st["Lt64"]
    T["Lt7"] = ((al - 1) * logPi).sum() + gammaln(al.sum()) - gammaln(al).sum() \
        + (((eps - 1) * logA).sum(1) + gammaln(eps.sum(1)) - gammaln(eps).sum(1)).sum()
    T["Lt8"] = 0.5 * (lLT + dim * np.log(beta / (2 * np.pi))).sum() - 0.5 * dim * K - H
    LB = T["Lt1"] + T["Lt2"] + T["Lt3"] + T["Lt4"] + T["Lt5"] - T["Lt6"] - T["Lt7"] - T["Lt8"]
    return float(LB), T


def vbhmm_em(data, K: int, ini: dict) -> dict:
    """[hmm, L] = vbhmm_em(data, K, ini)  (vbhmm_em.m:1).  ``data``: list of [T_n, dim] arrays or a ``SequenceSet``
    (then ``ini['initmode']`` must be 'initgmm' or 'inithmm', which need no look at the observations).
    The per-sequence responsibilities (``hmm.gamma``, vbhmm_em.m:463-466) are not materialised: call
    ``fb.vbhmm_fb(data, hmm['varpar'])`` for them."""
    own = not isinstance(data, _fb.SequenceSet)
    vp = vbhmm_init(data if own else (int(data.T.sum()), data.dim), K, ini)
    seqs = _fb.SequenceSet(data) if own else data
    try:
        max_iter, min_diff = int(ini.get("maxIter", 100)), float(ini.get("minDiff", 1e-5))
        L = lastL = -np.finfo(float).max
        LLs, unstable, it, st = [], False, 0, None
        for it in range(1, max_iter + 1):
            st = _fb.fb_stats(seqs, vp)                                    # E-step + sums (GPU)
            if it > 1:
                lastL = L
            L, _ = vbhmm_em_lb(st, vp, ini)
            do_break = (it > 1 and abs((L - lastL) / lastL) <= min_diff) or it == max_iter       # vbhmm_em.m:293-316
            if np.isnan(L):                                                # :320-330
                do_break, unstable, L = True, True, -np.inf
            LLs.append(L)
            if do_break and unstable:
                break
            vp = vbhmm_mstep(st, ini)                                      # also after the breaking pass (:411-414)
            if do_break:
                break
        dim = vp["m"].shape[1]
        Ci = np.linalg.inv(vp["W"]) / np.where(vp["v"] > dim + 1, vp["v"] - dim - 1, vp["v"])[:, None, None]   # :391-404
        rs = vp["epsilon"].sum(1, keepdims=True)
        hmm = dict(prior=vp["alpha"] / vp["alpha"].sum(), trans=vp["epsilon"] / np.where(rs == 0, 1.0, rs),
                   mean=vp["m"].copy(), cov=(Ci + np.transpose(Ci, (0, 2, 1))) / 2, LL=L, M=st["M"], N1=st["Nk1"],
                   N=st["Nk"], varpar=vp, iter=it, LLs=np.array(LLs), unstable=unstable, K=K)
        return hmm
    finally:
        if own:
            seqs.close()


def vbhmm_learn(data: list, K, vbopt: dict | None = None) -> dict:
    """[hmm, L] = vbhmm_learn(data, K, vbopt)  (vbhmm_learn.m:1): for every candidate K run ``numTrials`` EM runs from
    random initialisations (seed + trial), keep the best lower bound (:465-557); over the candidates keep the best
    ``LL + gammaln(K + 1)`` (:387-405).  The sequences are packed into HBM once for all runs."""
    opt = dict(DEFAULT_VBOPT)
    opt.update(vbopt or {})
    Ks = [int(k) for k in np.atleast_1d(K)]
    X = np.concatenate([np.asarray(x, float).reshape(len(x), -1) for x in data if len(x)])
    best_all, model_LL, per_k = None, [], []
    with _fb.SequenceSet(data) as seqs:
        for k in Ks:
            best = None
            trials = 1 if opt["initmode"] != "random" else int(opt["numTrials"])
            for t in range(trials):
                ini = dict(opt)
                if opt["initmode"] == "random":
                    seed = None if opt.get("seed") is None else int(opt["seed"]) + t
                    ini.update(initmode="inithmm",
                               inithmm=dict(varpar=vbhmm_init([X], k, dict(opt, initmode="random", seed=seed))))
                h = vbhmm_em(seqs, k, ini)
                if best is None or h["LL"] > best["LL"]:
                    best = h
            per_k.append(best)
            model_LL.append(best["LL"] + gammaln(k + 1))
            if best_all is None or model_LL[-1] > best_all[0]:
                best_all = (model_LL[-1], best)
    out = dict(best_all[1])
    out.update(model_LL=np.array(model_LL), model_k=np.array(Ks), model_all=per_k, L=best_all[0])
    return out


def vbhmm_remove_empty(hmm: dict, thresh: float = 1.0) -> tuple[dict, np.ndarray]:
    """[hmmo, emptyinds] = vbhmm_remove_empty(hmm, verbose, thresh)  (vbhmm_remove_empty.m:1-105): drop the states whose
    occupancy ``hmm['N']`` is below ``thresh``; variational parameters, counts, prior / trans (re-normalised from the
    remaining alpha / epsilon) and emissions follow.  Returns the HMM and the indices that were removed."""
    N = np.asarray(hmm["N"], float)
    keep = np.flatnonzero(~(N < thresh))
    gone = np.flatnonzero(N < thresh)
    if gone.size == 0:
        return hmm, gone
    vp = hmm["varpar"]
    nv = dict(alpha=np.asarray(vp["alpha"])[keep], epsilon=np.asarray(vp["epsilon"])[np.ix_(keep, keep)],
              beta=np.asarray(vp["beta"])[keep], v=np.asarray(vp["v"])[keep], m=np.asarray(vp["m"])[keep],
              W=np.asarray(vp["W"])[keep])
    out = dict(hmm)
    out.update(varpar=nv, M=np.asarray(hmm["M"])[np.ix_(keep, keep)], N1=np.asarray(hmm["N1"])[keep], N=N[keep],
               prior=nv["alpha"] / nv["alpha"].sum(), trans=nv["epsilon"] / nv["epsilon"].sum(1, keepdims=True),
               mean=np.asarray(hmm["mean"])[keep], cov=np.asarray(hmm["cov"])[keep], K=int(keep.size))
    return out, gone


# ----------------------------------------------------------------------------- many runs at once, on the device
def _hmm_from_varpar(vp: dict, K: int) -> dict:
    """vbhmm_em.m:391-460: point estimates from the final variational parameters."""
    dim = vp["m"].shape[1]
    Ci = np.linalg.inv(vp["W"]) / np.where(vp["v"] > dim + 1, vp["v"] - dim - 1, vp["v"])[:, None, None]
    rs = vp["epsilon"].sum(1, keepdims=True)
    return dict(prior=vp["alpha"] / vp["alpha"].sum(), trans=vp["epsilon"] / np.where(rs == 0, 1.0, rs),
                mean=vp["m"].copy(), cov=(Ci + np.transpose(Ci, (0, 2, 1))) / 2, varpar=vp, K=K)


def vbhmm_em_batch(seqs: _fb.SequenceSet, runs, K: int, ini: dict, inits: list) -> tuple[list, dict]:
    """Many ``vbhmm_em`` runs over one packed ``SequenceSet`` in a single device-resident loop (``vbhmm_batch_*``).

    ``runs[r] = (lo, hi)``: run r learns from sequences lo..hi-1 (the reference's per-subject loop,
    vbhmm_learn_batch.m, and per-initialisation loop, vbhmm_learn.m:465-557, flattened; runs may share sequences);
    ``inits[r]``: its initial variational parameters (``vbhmm_init``).  All runs have K states.
    Returns (list of hmm dicts as ``vbhmm_em``, info = dict(passes, device_ms))."""
    R = len(runs)
    dim = seqs.dim
    c = np.ascontiguousarray
    lo = c([a for a, _ in runs], dtype=np.int32)
    hi = c([b for _, b in runs], dtype=np.int32)
    stack = lambda k: np.stack([np.asarray(v[k], float) for v in inits])
    al, be, vv = c(stack("alpha")), c(stack("beta")), c(stack("v"))                    # [R, K]     == col-major [K, R]
    ep = c(np.transpose(stack("epsilon"), (0, 2, 1)))                                  # [R, to, from] == [from, to, R]
    mm = c(stack("m"))                                                                 # [R, K, dim]  == [dim, K, R]
    WW = c(np.transpose(stack("W"), (0, 1, 3, 2)))                                     # [R, K, j, i] == [i, j, K, R]
    m0 = c(np.asarray(ini["mu0"], float).reshape(dim))
    W0 = c(np.atleast_1d(np.asarray(ini["W0"], float)).ravel())
    max_iter, min_diff = int(ini.get("maxIter", 100)), float(ini.get("minDiff", 1e-5))
    P = lambda a: a.ctypes.data_as(dp)
    h = C.c_void_p()
    check(lib().vbhmm_batch_begin(seqs._h, R, lo.ctypes.data_as(ip), hi.ctypes.data_as(ip), K, float(ini["alpha0"]),
                                  float(ini["epsilon0"]), float(ini["beta0"]), float(ini["v0"]), P(m0), P(W0),
                                  int(W0.size), P(al), P(ep), P(be), P(vv), P(mm), P(WW), max_iter, C.byref(h)))
    try:
        passes = C.c_int()
        check(lib().vbhmm_batch_run(h, min_diff, C.byref(passes)))
        o = dict(alpha=np.zeros((R, K)), eps=np.zeros((R, K, K)), beta=np.zeros((R, K)), v=np.zeros((R, K)),
                 m=np.zeros((R, K, dim)), W=np.zeros((R, K, dim, dim)), Nk1=np.zeros((R, K)), Nk=np.zeros((R, K)),
                 M=np.zeros((R, K, K)), LLs=np.zeros((R, max_iter)), it=np.zeros(R, np.int32), bad=np.zeros(R, np.int32))
        check(lib().vbhmm_batch_fetch(h, P(o["alpha"]), P(o["eps"]), P(o["beta"]), P(o["v"]), P(o["m"]), P(o["W"]),
                                      P(o["Nk1"]), P(o["Nk"]), P(o["M"]), P(o["LLs"]), o["it"].ctypes.data_as(ip),
                                      o["bad"].ctypes.data_as(ip)))
        ms = C.c_double()
        check(lib().vbhmm_batch_last_ms(h, C.byref(ms)))
    finally:
        check(lib().vbhmm_batch_free(h))
    # point estimates of all runs at once (vbhmm_em.m:391-460; _hmm_from_varpar is the one-run form): the per-run loop
    # below only slices
    eps_all = np.ascontiguousarray(np.transpose(o["eps"], (0, 2, 1)))                  # [R, from, to]
    W_all = np.ascontiguousarray(np.transpose(o["W"], (0, 1, 3, 2)))                   # [R, K, i, j]
    M_all = np.ascontiguousarray(np.transpose(o["M"], (0, 2, 1)))
    vv_all = o["v"]
    Ci = np.linalg.inv(W_all) / np.where(vv_all > dim + 1, vv_all - dim - 1, vv_all)[:, :, None, None]
    cov_all = (Ci + np.transpose(Ci, (0, 1, 3, 2))) / 2
    prior_all = o["alpha"] / o["alpha"].sum(1, keepdims=True)
    rs = eps_all.sum(2, keepdims=True)
    trans_all = eps_all / np.where(rs == 0, 1.0, rs)
    mean_all = o["m"].copy()
    its = o["it"].astype(int)
    LL_last = o["LLs"][np.arange(R), np.maximum(its, 1) - 1]
    out = []
    for r in range(R):
        n = its[r]
        vp = dict(alpha=o["alpha"][r], epsilon=eps_all[r], beta=o["beta"][r], v=vv_all[r], m=o["m"][r], W=W_all[r])
        out.append(dict(prior=prior_all[r], trans=trans_all[r], mean=mean_all[r], cov=cov_all[r], varpar=vp, K=K,
                        LL=float(LL_last[r]), LLs=o["LLs"][r, :n], iter=int(n), unstable=bool(o["bad"][r]),
                        M=M_all[r], N1=o["Nk1"][r], N=o["Nk"][r]))
    return out, dict(passes=passes.value, device_ms=ms.value)


def gmm_fit_batch(seqs: _fb.SequenceSet, runs, K: int, starts, max_iter: int = 100, tol: float = 1e-5,
                  reg: float = 1e-10) -> list:
    """K-component full-covariance GMMs of many runs' pooled observations, one device block per run
    (``vbhmm_gmm_fit``): the random initialisation of vbhmm_init.m:43-61 without the per-run host fit.
    ``runs[r] = (lo, hi)`` sequences; ``starts[r]`` = K indices into the run's pooled observations (sequence after
    sequence, time step after time step) that start its components.  Returns a list of dict(prior, mean, cov, loglik,
    iters) -- the ``initgmm`` argument of ``vbhmm_init``."""
    R, dim = len(runs), seqs.dim
    c = np.ascontiguousarray
    lo = c([a for a, _ in runs], dtype=np.int32)
    hi = c([b for _, b in runs], dtype=np.int32)
    # pooled observation index of a run -> (sequence, time step): offset by the observations before the run's first
    # sequence, then one search in the cumulative lengths of all sequences (vectorised over runs x components)
    cum = np.concatenate([[0], np.cumsum(seqs.T.astype(np.int64))])
    glob_idx = cum[lo.astype(np.int64)][:, None] + np.asarray(starts, np.int64).reshape(R, K)
    if np.any(glob_idx < cum[lo][:, None]) or np.any(glob_idx >= cum[hi][:, None]):
        raise ValueError("a start index lies outside its run's pooled observations")
    seq_idx = np.searchsorted(cum, glob_idx, side="right") - 1
    ss = c(seq_idx, dtype=np.int32)
    st = c(glob_idx - cum[seq_idx], dtype=np.int32)
    w, mu, cv = np.zeros((R, K)), np.zeros((R, K, dim)), np.zeros((R, K, dim, dim))
    ll, it = np.zeros(R), np.zeros(R, np.int32)
    P = lambda a: a.ctypes.data_as(dp)
    I = lambda a: a.ctypes.data_as(ip)
    check(lib().vbhmm_gmm_fit(seqs._h, R, I(lo), I(hi), K, I(ss), I(st), int(max_iter), float(tol), float(reg), P(w), P(mu),
                              P(cv), P(ll), I(it)))
    return [dict(prior=w[r], mean=mu[r], cov=cv[r], loglik=float(ll[r]), iters=int(it[r])) for r in range(R)]


def vbhmm_learn_batch(subjects: list, K, vbopt: dict | None = None) -> list:
    """vbhmm_learn_batch.m: one HMM per subject (``subjects[s]`` = that subject's list of [T, dim] sequences), every
    subject x candidate K x random initialisation as one run of a device-resident batch per K; per subject the best
    trial by LL, then the best K by ``LL + gammaln(K + 1)`` (vbhmm_learn.m:387-405,555).
    The random-initialisation GMMs of all runs are fitted on the GPU (``gmm_fit_batch``, the default);
    ``vbopt['gmm_init'] = 'sklearn'`` takes one scikit-learn fit per run on the host instead (comparison only)."""
    opt = dict(DEFAULT_VBOPT)
    opt.update(vbopt or {})
    Ks = [int(k) for k in np.atleast_1d(K)]
    flat = [x for sub in subjects for x in sub]
    start = np.concatenate([[0], np.cumsum([len(sub) for sub in subjects])]).astype(int)
    trials = 1 if opt["initmode"] != "random" else int(opt["numTrials"])
    best = [None] * len(subjects)
    score = [-np.inf] * len(subjects)
    per_k = [[None] * len(Ks) for _ in subjects]
    device_gmm = opt["initmode"] == "random" and opt.get("gmm_init", "device") == "device"
    with _fb.SequenceSet(flat) as seqs:
        npts = [int(seqs.T[start[s]:start[s + 1]].sum()) for s in range(len(subjects))]
        for ki, k in enumerate(Ks):
            runs, inits, owner = [], [], []
            for s, sub in enumerate(subjects):
                for t in range(trials):
                    runs.append((int(start[s]), int(start[s + 1])))
                    owner.append(s)
            if device_gmm and k > 1 and min(npts) > k:
                # all runs' GMM fits in one launch; the start points are drawn as vbhmm_init's 'randSample' start would
                starts = []
                for s in range(len(subjects)):
                    for t in range(trials):
                        rng = np.random.Generator(np.random.PCG64((0 if opt.get("seed") is None else int(opt["seed"])) + t + 7919 * s))
                        starts.append(rng.choice(npts[s], size=k, replace=False))
                gmms = gmm_fit_batch(seqs, runs, k, starts)
                inits = vbhmm_init_from_gmms(gmms, [npts[o] for o in owner], k, opt)
            else:
                for s, sub in enumerate(subjects):
                    for t in range(trials):
                        seed = None if opt.get("seed") is None else int(opt["seed"]) + t
                        inits.append(vbhmm_init(sub, k, dict(opt, seed=seed)))
            hmms, _ = vbhmm_em_batch(seqs, runs, k, opt, inits)
            for s in range(len(subjects)):
                top = max(hmms[s * trials:(s + 1) * trials], key=lambda h: h["LL"])    # runs are subject-major
                per_k[s][ki] = top
                sc = top["LL"] + gammaln(k + 1)
                if best[s] is None or sc > score[s]:
                    score[s], best[s] = sc, top
    out = []
    for s in range(len(subjects)):
        h = dict(best[s])
        h.update(model_LL=np.array([per_k[s][ki]["LL"] + gammaln(k + 1) for ki, k in enumerate(Ks)]),
                 model_k=np.array(Ks), L=score[s])
        out.append(h)
    return out
