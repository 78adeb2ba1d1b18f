"""Seeded synthetic HMM sets of the BASELINE.json shapes (host side, numpy).

There is no dataset to download; every benchmark and parity case is generated
here from ``numpy.random.Generator(PCG64(seed))`` (1001 is the reference's
customary seed, /root/reference/Synthetic_experiment/exprmt1_demo.m:37,90).

Base HMMs are produced the way the toolbox produces them: a learned VB-HMM has
posterior counts (alpha, epsilon, beta) and point estimates (mean, cov), and
``hmms_to_h3m_hem(hmms, covmode, use_post=1)`` turns those into the base mixture
(/root/reference/src/vbhem/hmms_to_h3m_hem.m:42-58,86-89):
``prior = exp(psi(alpha)-psi(sum alpha))``, ``A = exp(psi(eps)-psi(rowsum eps))``
(sub-stochastic), ``covars = (beta+1)/beta * cov``.

Centre HMMs are initialised with the ``baseem`` rule
(/root/reference/src/vbhem/vbhemhmm_init.m:58-100).
"""
from __future__ import annotations

import numpy as np
from scipy.special import psi

DEFAULT_OPT = dict(alpha0=1.0, eta0=1.0, epsilon0=1.0, lambda0=1.0, v0=5.0, W0=0.005,
                   m0=(256.0, 192.0), Nv=100.0, tau=10)
"""Option defaults that reach the E-step (/root/reference/src/vbhem/vbhem_h3m_cluster.m:141-173)."""


def _rand_cov(rng, n, d, lo=10.0, hi=40.0):
    """n random SPD d x d matrices R diag(s^2) R' with s ~ U(lo,hi)."""
    s = rng.uniform(lo, hi, size=(n, d))
    Q, _ = np.linalg.qr(rng.standard_normal((n, d, d)))
    return np.einsum("nij,nj,nkj->nik", Q, s * s, Q)


def make_truth(rng, C, S, d):
    """C ground-truth cluster HMMs: means ~ U(image box), covariances as above,
    transition rows ~ Dirichlet(1) + 2 I (normalised), priors ~ Dirichlet(1)."""
    box = np.array([512.0, 384.0] + [256.0] * max(0, d - 2))[:d]
    mean = rng.uniform(0.0, 1.0, size=(C, S, d)) * box
    cov = _rand_cov(rng, C * S, d).reshape(C, S, d, d)
    A = rng.dirichlet(np.ones(S), size=(C, S)) + 2.0 * np.eye(S)
    A /= A.sum(axis=2, keepdims=True)
    prior = rng.dirichlet(np.ones(S), size=C)
    return dict(prior=prior, A=A, mean=mean, cov=cov)


def make_base_set(N, S=8, d=2, C=64, covmode="full", seed=1001, Nv=100.0, ragged=False):
    """N base HMMs drawn around C truth HMMs -> natural base dict (see pack.py) + labels.

    ``varpar``-style posterior counts: eps = 1 + 250 A_true * noise, alpha = 1 + 25 pi_true,
    beta = 1 + counts; means jittered by N(0,3^2), covariances scaled by logN(0,0.1^2).
    With ``ragged`` the state counts vary in 1..S (pruned HMMs,
    /root/reference/src/vbhem/vbhem_h3m_cluster.m:116-134) and padding is zero.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    tr = make_truth(rng, C, S, d)
    lab = rng.integers(0, C, size=N)
    eps = 1.0 + 250.0 * tr["A"][lab] * rng.uniform(0.5, 1.5, size=(N, S, S))
    alpha = 1.0 + 25.0 * tr["prior"][lab]
    beta = 1.0 + rng.uniform(5.0, 60.0, size=(N, S))
    mean = tr["mean"][lab] + 3.0 * rng.standard_normal((N, S, d))
    cov = tr["cov"][lab] * np.exp(0.1 * rng.standard_normal((N, S, 1, 1)))
    Sb = np.full(N, S, dtype=np.int32)
    if ragged:
        Sb = rng.integers(1, S + 1, size=N).astype(np.int32)
    mask = np.arange(S)[None, :] < Sb[:, None]                       # [N,S]
    # hmms_to_h3m_hem use_post=1, on each HMM's own (unpadded) states
    alpha_m = np.where(mask, alpha, 0.0)
    prior = np.where(mask, np.exp(psi(np.where(mask, alpha, 1.0)) - psi(alpha_m.sum(1, keepdims=True))), 0.0)
    m2 = mask[:, :, None] & mask[:, None, :]
    eps_m = np.where(m2, eps, 0.0)
    rows = np.where(mask, eps_m.sum(2), 1.0)
    A = np.where(m2, np.exp(psi(np.where(m2, eps, 1.0)) - psi(rows)[:, :, None]), 0.0)
    tb = (beta + 1.0) / beta
    covars = tb[:, :, None, None] * cov
    mean = np.where(mask[:, :, None], mean, 0.0)
    covars = np.where(mask[:, :, None, None], covars, 0.0)
    if covmode == "diag":
        covars = np.diagonal(covars, axis1=2, axis2=3).copy()
    omega = np.full(N, 1.0 / N)
    Ntilde = Nv * N * omega                                          # vbhem_h3m_c_step_fc.m:26-30
    base = dict(Sb=Sb, prior=prior, A=A, mu=mean, cov=covars, Ntilde=Ntilde)
    return base, lab


def baseem_draws(base, K, S, seed=1001, Sr=None):
    """The random draws of one ``baseem`` initialisation (vbhemhmm_init.m:66-74, :91-92), in the order the per-centre
    loop makes them: (Sr[K], randb[K,S] base index, randg[K,S] state index -- both 0-based, -1 beyond a centre's own
    states --, omega[K] normalised)."""
    rng = np.random.Generator(np.random.PCG64(seed + 7))
    N = base["mu"].shape[0]
    Sr = np.full(K, S, dtype=np.int32) if Sr is None else np.asarray(Sr, dtype=np.int32)
    Sb = base["Sb"]
    randb, randg = -np.ones((K, S), np.int32), -np.ones((K, S), np.int32)
    for j in range(K):
        s = int(Sr[j])
        rb = rng.integers(0, N, size=s)
        randb[j, :s] = rb
        randg[j, :s] = rng.integers(0, Sb[rb])          # the stream of one scalar draw per state, vectorised
    omega = rng.uniform(size=K)
    return Sr, randb, randg, omega / omega.sum()


def stack_draws(draws):
    """Draws of several ``baseem`` initialisations (``baseem_draws`` tuples) -> the arguments of
    ``BaseSet.em_begin_baseem``: (Sr[Kt], group[Kt], randb[Kt,maxS], randg[Kt,maxS], omega[Kt])."""
    maxS = max(d[1].shape[1] for d in draws)
    pad = lambda a: np.pad(np.maximum(a, 0), ((0, 0), (0, maxS - a.shape[1])))
    group = np.repeat(np.arange(len(draws), dtype=np.int32), [len(d[0]) for d in draws])
    return (np.concatenate([d[0] for d in draws]).astype(np.int32), group, np.concatenate([pad(d[1]) for d in draws]),
            np.concatenate([pad(d[2]) for d in draws]), np.concatenate([d[3] for d in draws]))


def baseem_draws_batch(base, models, seed):
    """The draws of MANY ``baseem`` initialisations at once, vectorised (one generator for the whole batch):
    ``models`` = [(K, S), ...] -> (Sr[Kt], group[Kt], randb[Kt,maxS], randg[Kt,maxS], omega[Kt] normalised per model),
    the arguments of ``BaseSet.em_begin_baseem``."""
    rng = np.random.Generator(np.random.PCG64(seed + 7))
    N = base["mu"].shape[0]
    Ks = np.array([k for k, _ in models])
    Ss = np.array([s for _, s in models])
    group = np.repeat(np.arange(len(models), dtype=np.int32), Ks)
    Sr = np.repeat(Ss, Ks).astype(np.int32)
    Kt, maxS = int(Ks.sum()), int(Ss.max())
    randb = rng.integers(0, N, size=(Kt, maxS)).astype(np.int32)
    randg = np.floor(rng.uniform(size=(Kt, maxS)) * np.asarray(base["Sb"])[randb]).astype(np.int32)
    live = np.arange(maxS)[None, :] < Sr[:, None]
    randb[~live], randg[~live] = 0, 0
    omega = rng.uniform(size=Kt)
    omega /= np.bincount(group, weights=omega)[group]
    return Sr, group, randb, randg, omega


def baseem_from_draws(base, Sr, randb, randg, omega, opt=DEFAULT_OPT, group=None):
    """``baseem`` initialisation (vbhemhmm_init.m:58-100) from given draws -> centre hyper-parameter dict:
    alpha[K], eta[K,S], epsilon[K,S,S], m[K,S,d], W[K,S,d,d]|[K,S,d], v[K,S], lam[K,S], Sr[K].

    Each centre state copies the emission of a random base state:
    lambda = lambda0 + NLr/S, v = v0 + NLr/S + 1, m = centres, W = inv((v-d-1) covars);
    eta = prior*NLr + eta0, epsilon = A*NLr/S + epsilon0 with uniform prior / A
    (``makeAprior`` in its default mode), alpha = alpha0 + omega*Nv with random omega; NLr = Nv / (centres of the
    same H3M): ``group`` batches several H3Ms.
    """
    N, Sb_max, d = base["mu"].shape
    full = base["cov"].ndim == 4
    Sr = np.asarray(Sr, dtype=np.int32)
    K, S = randb.shape
    group = np.zeros(K, np.int32) if group is None else np.asarray(group)
    Nv = opt["Nv"] * N
    NLr = Nv / np.bincount(group)[group]                                    # [K]
    live = np.arange(S)[None, :] < Sr[:, None]
    per = (NLr / Sr)[:, None]
    lam = np.where(live, opt["lambda0"] + per, 0.0)
    v = np.where(live, opt["v0"] + per + 1.0, 0.0)
    eta = np.where(live, (1.0 / Sr)[:, None] * NLr[:, None] + opt["eta0"], 0.0)
    eps = np.where(live[:, :, None] & live[:, None, :], ((1.0 / Sr) * NLr / Sr)[:, None, None] + opt["epsilon0"], 0.0)
    jj, kk = np.nonzero(live)
    bb, gg = randb[jj, kk], randg[jj, kk]
    m = np.zeros((K, S, d))
    W = np.zeros((K, S, d, d)) if full else np.zeros((K, S, d))
    m[jj, kk] = base["mu"][bb, gg]
    c = base["cov"][bb, gg]
    if full:
        W[jj, kk] = np.linalg.inv((v[jj, kk, None, None] - d - 1.0) * c)
    else:
        W[jj, kk] = 1.0 / ((v[jj, kk, None] - d - 1.0) * c)
    alpha = opt["alpha0"] + omega * Nv
    return dict(Sr=Sr, alpha=alpha, eta=eta, epsilon=eps, m=m, W=W, v=v, lam=lam)


def init_centres_baseem(base, K, S, seed=1001, opt=DEFAULT_OPT, Sr=None):
    """``baseem`` initialisation of one H3M: ``baseem_from_draws(baseem_draws(...))``."""
    Sr, randb, randg, omega = baseem_draws(base, K, S, seed, Sr)
    return baseem_from_draws(base, Sr, np.maximum(randb, 0), np.maximum(randg, 0), omega, opt)


def north_star(N=100_000, K=64, S=8, d=2, seed=1001, covmode="full"):
    """BASELINE.json configs[2]: N base HMMs x K centres, S states both sides, d-D full covariance."""
    base, lab = make_base_set(N, S=S, d=d, C=K, covmode=covmode, seed=seed)
    cen = init_centres_baseem(base, K, S, seed=seed)
    return base, cen, lab


def model_selection_grid(base, Ks=range(1, 17), Ss=range(1, 9), seed=1001, opt=DEFAULT_OPT):
    """BASELINE.json configs[3]: the centres of every (K, S) candidate of the posterior model selection
    (vbhem_h3m_cluster.m:261-355 runs them one after the other) stacked for ONE batched E-step call:
    hyper-parameter dict padded to max(Ss) states, plus ``group`` (one soft-max segment per (K, S) model)
    and ``models`` = [(K, S), ...] in group order."""
    parts, group, models = [], [], []
    maxS = max(Ss)
    for S in Ss:
        for K in Ks:
            h = init_centres_baseem(base, K, S, seed=seed + 131 * K + S, opt=opt)
            pad = maxS - S
            parts.append(dict(Sr=h["Sr"], alpha=h["alpha"], eta=np.pad(h["eta"], ((0, 0), (0, pad))),
                              epsilon=np.pad(h["epsilon"], ((0, 0), (0, pad), (0, pad))),
                              m=np.pad(h["m"], ((0, 0), (0, pad), (0, 0))),
                              W=np.pad(h["W"], ((0, 0), (0, pad)) + ((0, 0),) * (h["W"].ndim - 2)),
                              v=np.pad(h["v"], ((0, 0), (0, pad))), lam=np.pad(h["lam"], ((0, 0), (0, pad)))))
            group += [len(models)] * K
            models.append((K, S))
    hyp = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    return hyp, np.array(group, np.int32), models
