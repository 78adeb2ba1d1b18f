"""oracle/numpy_ref.py -- TEST INFRASTRUCTURE, not product code.

A numpy restatement of the MATLAB side of the VBHEM-H3M E-step and of the host
steps around it, written from the reference's .m files (paths relative to
/root/reference/src/vbhem/).  It is the *second opinion* on the per-pair
arithmetic (the first is the reference's C compiled in place, oracle/_ref) and
the only oracle for the MATLAB-only stages: pre-pass, responsibilities,
statistics normalisation, M-step, lower bound.

Parity status: the per-pair functions are PINNED against oracle/_ref (the
reference's own native code) in tests/test_oracle.py.  The MATLAB-only stages
are "parity unpinned" at the level of MATLAB builtins: `psi`, `gammaln`, `det`,
`inv` are proprietary MathWorks code with no version pin in the reference;
scipy.special.psi/gammaln and numpy.linalg stand in for them (standard
functions, differences are a few ulp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammaln, psi


# --------------------------------------------------------------------------- helpers
def logtrick(lA: np.ndarray, axis: int = 0) -> np.ndarray:
    """log(sum(exp(lA))) over `axis` with a max shift.  logtrick.m:15-18, logtrick2.m:11-16."""
    mv = np.max(lA, axis=axis, keepdims=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        s = mv + np.log(np.sum(np.exp(lA - mv), axis=axis, keepdims=True))
    return np.squeeze(s, axis=axis)


# --------------------------------------------------------------------------- a1: base set
def hmms_to_h3m_base(alpha, epsilon, beta, mean, cov, covmode="full"):
    """Packed rows of `hmms_to_h3m_hem(hmms, covmode, use_post=1)` for ONE learned HMM.

    hmms_to_h3m_hem.m:42-58: prior = exp(psi(alpha) - psi(sum alpha)),
    A = exp(psi(eps) - psi(rowsum eps)) (sub-stochastic on purpose);
    :86-89: covars = (beta+1)/beta * cov (diag: its diagonal).
    alpha[S], epsilon[S,S], beta[S], mean[S,d], cov[S,d,d].
    Returns prior[S], A[S,S], mu[S,d], covars[S,d,d] or [S,d].
    """
    alpha = np.asarray(alpha, float)
    epsilon = np.asarray(epsilon, float)
    prior = np.exp(psi(alpha) - psi(alpha.sum()))
    A = np.exp(psi(epsilon) - psi(epsilon.sum(axis=1, keepdims=True)))
    tb = (np.asarray(beta, float) + 1.0) / np.asarray(beta, float)
    cov = np.asarray(cov, float)
    if covmode == "full":
        covars = tb[:, None, None] * cov
    else:
        covars = tb[:, None] * np.diagonal(cov, axis1=1, axis2=2)
    return prior, A, np.asarray(mean, float), covars


# --------------------------------------------------------------------------- a3: pre-pass
def prepass(centre: dict, covmode="full") -> dict:
    """E-step pre-pass for ONE centre HMM.  vbhem_h3m_c_step_fc.m:118-165 and :180-191.

    centre: eta[S], epsilon[S,S], m[S,d], W[S,d,d] (full) or [S,d] (diag), v[S], lam[S].
    Returns logPi[S], logA[S,S] (row=from), cterm[S] (= -logLambdaTilde + d/lambda),
    logLambdaTilde[S], vW (= v*W, same shape as W).
    """
    eta = np.asarray(centre["eta"], float)
    eps = np.asarray(centre["epsilon"], float)
    W = np.asarray(centre["W"], float)
    v = np.asarray(centre["v"], float)
    lam = np.asarray(centre["lam"], float)
    S = eta.shape[0]
    d = np.asarray(centre["m"]).shape[1]
    idx = np.arange(1, d + 1, dtype=float)
    logLam = np.empty(S)
    for k in range(S):
        t1 = psi(0.5 * (v[k] + 1.0) - 0.5 * idx)
        if covmode == "full":
            logdet = np.log(np.linalg.det(W[k]))
        else:
            logdet = np.sum(np.log(W[k]))
        logLam[k] = np.sum(t1) + d * np.log(2.0) + logdet
    cterm = -logLam + d / lam
    logA = psi(eps) - psi(eps.sum(axis=1, keepdims=True))
    logPi = psi(eta) - psi(eta.sum())
    vW = v[:, None, None] * W if covmode == "full" else v[:, None] * W
    return dict(logPi=logPi, logA=logA, cterm=cterm, logLambdaTilde=logLam, vW=vW)


def log_omega_tilde(alpha) -> np.ndarray:
    """vbhem_h3m_c_step_fc.m:271-273."""
    alpha = np.asarray(alpha, float)
    return psi(alpha) - psi(alpha.sum())


# --------------------------------------------------------------------------- a5..a10: one pair
def pair_estep(prior, A, mu, cov, T, m, vW, cterm, logA, logPi, covmode="full"):
    """One (base, centre) pair, vectorised as in vbhem_hmm_bwd_fwd_fast.m.

    prior[Sb], A[Sb,Sb] (from,to), mu[Sb,d], cov[Sb,d,d]|[Sb,d];
    m[Sr,d], vW[Sr,d,d]|[Sr,d] (= v*W), cterm[Sr], logA[Sr,Sr], logPi[Sr].
    Returns L_elbo, nu1[Sr], pr[Sr], mu_out[Sr,d], Mu_out[Sr,d,d]|[Sr,d], xi[Sr,Sr].
    """
    Sb, d = mu.shape
    Sr = m.shape[0]
    # E3logN [Sb,Sr]  (vbhem_hmm_bwd_fwd_fast.m:80-129; MEX arithmetic :744-759,:796-843)
    diff = mu[:, None, :] - m[None, :, :]                       # [Sb,Sr,d]
    if covmode == "full":
        trcov = np.einsum("rkl,bkl->br", vW, cov)
        dterm = np.einsum("brk,rkl,brl->br", diff, vW, diff)
    else:
        trcov = np.einsum("rk,bk->br", vW, cov)
        dterm = np.einsum("brk,rk->br", diff * diff, vW)
    E = -0.5 * (d * np.log(2 * np.pi) + cterm[None, :] + trcov + dterm)

    # backward (fast.m:171-245).  Theta[rho,sigma,gamma,t]
    Theta = np.zeros((Sr, Sr, Sb, max(T, 1)))
    LL = np.zeros((Sb, Sr))
    for t in range(T - 1, 0, -1):
        # logtheta[sigma, beta, rho] = logA[rho,sigma] + E[beta,sigma] + LL[beta,sigma]
        lt = logA.T[:, None, :] + (E + LL).T[:, :, None]
        ls = logtrick(lt, 0)                                     # [Sb(beta), Sr(rho)]
        LL = A @ ls                                              # LL_new[beta',rho]
        with np.errstate(invalid="ignore"):
            Theta[:, :, :, t] = np.transpose(np.exp(lt - ls[None]), (2, 0, 1))
    # terminate (fast.m:249-257)
    lt = logPi[:, None] + (E + LL).T                             # [Sr, Sb]
    ls = logtrick(lt, 0)
    L_elbo = float(prior @ ls)
    with np.errstate(invalid="ignore"):
        Theta1 = np.exp(lt - ls[None])
    # forward (fast.m:276-341)
    nu = prior[None, :] * Theta1                                 # [Sr(sigma), Sb(gamma)]
    nu1 = nu.sum(axis=1)
    n = nu.copy()
    xi = np.zeros((Sr, Sr))
    for t in range(1, T):
        foo = nu @ A                                             # [rho, gamma]
        x = foo[:, None, :] * Theta[:, :, :, t]                  # [rho, sigma, gamma]
        xi += x.sum(axis=2)
        nu = x.sum(axis=0)
        n += nu
    # emission statistics (fast.m:353-384)
    pr = n.sum(axis=1)
    mu_out = n @ mu
    if covmode == "full":
        Mu_out = np.einsum("sg,gkl->skl", n, mu[:, :, None] * mu[:, None, :] + cov)
    else:
        Mu_out = n @ (mu * mu + cov)
    return L_elbo, nu1, pr, mu_out, Mu_out, xi


# --------------------------------------------------------------------------- a12: responsibilities
def responsibilities(L_elbo, logOmega, Ntilde):
    """vbhem_h3m_c_step_fc.m:271-283.  L_elbo[N,K] -> zhat[N,K], Z[N,K], Nj[K]."""
    Ntilde = np.asarray(Ntilde, float).reshape(-1, 1)
    log_Z = Ntilde * (np.asarray(logOmega, float)[None, :] + L_elbo)
    with np.errstate(invalid="ignore"):
        zhat = np.exp(log_Z - logtrick(log_Z.T, 0)[:, None]) + 1e-50
    Z = zhat * Ntilde
    Nj = Z.sum(axis=0) + 1e-50
    return zhat, Z, Nj


# --------------------------------------------------------------------------- a13: reduction
def reduce_raw(Z, nu1, xi, pr, mu_out, Mu_out):
    """Raw weighted sums of vbhem_compute_Statistics.m:33-55 for ONE centre.

    Z[N]; per-pair arrays with leading axis N.  Bases with Z <= 1e-8 are skipped (:35)."""
    keep = Z > 1e-8
    w = Z[keep]
    f = lambda a: np.tensordot(w, np.asarray(a)[keep], axes=(0, 0))
    return f(nu1), f(xi), f(pr), f(mu_out), f(Mu_out)


def finalize_stats(S_nu1, S_xi, S_pr, S_mu, S_Mu, covmode="full"):
    """vbhem_compute_Statistics.m:57-82 for ONE centre: Nj_rho += 1e-50; y_bar = sum/Nj_rho;
    S_plus_C = sum/Nj_rho - y_bar y_bar'; Sr==1 => Nj_rho2rho = 1e-12."""
    Nj_rho = np.asarray(S_pr, float) + 1e-50
    y_bar = np.asarray(S_mu, float) / Nj_rho[:, None]
    if covmode == "full":
        SC = np.asarray(S_Mu, float) / Nj_rho[:, None, None] - y_bar[:, :, None] * y_bar[:, None, :]
    else:
        SC = np.asarray(S_Mu, float) / Nj_rho[:, None] - y_bar * y_bar
    Nj_rho2rho = np.asarray(S_xi, float)
    if Nj_rho.shape[0] == 1:
        Nj_rho2rho = np.full((1, 1), 1e-12)
    return np.asarray(S_nu1, float), Nj_rho2rho, Nj_rho, y_bar, SC


# --------------------------------------------------------------------------- M-step
def mstep_component(Nj_rho1, Nj_rho2rho, Nj_rho, y_bar, SC, opt, covmode="full"):
    """vbhem_mstep_component.m:42-70.  opt: eta0, epsilon0, m0[d], lambda0, W0 (scalar or [d]), v0."""
    d = y_bar.shape[1]
    W0 = np.atleast_1d(np.asarray(opt["W0"], float))
    W0m = W0[0] * np.eye(d) if W0.size == 1 else np.diag(W0)
    W0inv = np.linalg.inv(W0m)
    m0 = np.asarray(opt["m0"], float).reshape(d)
    lam0, v0 = float(opt["lambda0"]), float(opt["v0"])
    S = Nj_rho.shape[0]
    out = dict(eta=opt["eta0"] + Nj_rho1, epsilon=opt["epsilon0"] + Nj_rho2rho,
               lam=lam0 + Nj_rho, v=v0 + Nj_rho + 1.0, m=np.empty((S, d)),
               W=np.empty((S, d, d)) if covmode == "full" else np.empty((S, d)))
    for k in range(S):
        out["m"][k] = (lam0 * m0 + Nj_rho[k] * y_bar[k]) / (lam0 + Nj_rho[k])
        mult1 = lam0 * Nj_rho[k] / (lam0 + Nj_rho[k])
        diff3 = y_bar[k] - m0
        Sk = SC[k] if covmode == "full" else np.diag(SC[k])
        tW = np.linalg.inv(W0inv + Nj_rho[k] * Sk + mult1 * np.outer(diff3, diff3))
        Wn = (tW + tW.T) / 2
        out["W"][k] = Wn if covmode == "full" else np.diag(Wn)
    return out


# --------------------------------------------------------------------------- lower bound
def lower_bound(Lt1, Lt7, Nj, logOmega, alpha, centres, pre, opt, covmode="full"):
    """vbhemh3m_lb.m:64-186.  Lt1 = sum Z.*L_elbo and Lt7 = sum zhat.*log zhat come from the
    E-step; everything else is O(K*S) host arithmetic.  centres/pre: lists of per-centre dicts
    (hyper-parameters; prepass() results).  Returns (LB, dict of the ten terms)."""
    K = len(centres)
    d = len(np.atleast_1d(opt["m0"]))
    Sr = centres[0]["eta"].shape[0]
    alpha0, eta0, eps0 = float(opt["alpha0"]), float(opt["eta0"]), float(opt["epsilon0"])
    lam0, v0 = float(opt["lambda0"]), float(opt["v0"])
    m0 = np.asarray(opt["m0"], float).reshape(d)
    W0 = np.atleast_1d(np.asarray(opt["W0"], float))
    W0m = W0[0] * np.eye(d) if W0.size == 1 else np.diag(W0)
    W0inv = np.linalg.inv(W0m)
    logdetW0inv = d * np.log(W0inv[0, 0]) if W0.size == 1 else np.sum(np.log(np.diag(W0inv)))
    ii = np.arange(1, d + 1, dtype=float)
    logCalpha0 = gammaln(K * alpha0) - K * gammaln(alpha0)
    logCeta0 = gammaln(Sr * eta0) - Sr * gammaln(eta0)
    logCeps0 = np.full(Sr, gammaln(Sr * eps0) - Sr * gammaln(eps0))
    logB0 = (v0 / 2) * logdetW0inv - (v0 * d / 2) * np.log(2) - (d * (d - 1) / 4) * np.log(np.pi) \
        - np.sum(gammaln(0.5 * (v0 + 1 - ii)))
    const2 = d * np.log(lam0 / (2 * np.pi))
    alpha = np.asarray(alpha, float)
    logCalpha = gammaln(alpha.sum()) - np.sum(gammaln(alpha))
    Lt2 = float(np.asarray(Nj) @ np.asarray(logOmega))
    Lt3 = K * logCeta0 + (eta0 - 1) * sum(np.sum(p["logPi"]) for p in pre)
    Lt4 = K * np.sum(logCeps0) + (eps0 - 1) * sum(np.sum(p["logA"]) for p in pre)
    Lt6 = logCalpha0 + (alpha0 - 1) * np.sum(logOmega)
    Lt8 = logCalpha + float((alpha - 1) @ np.asarray(logOmega))
    Lt5 = Lt9 = Lt10 = 0.0
    for c, p in zip(centres, pre):
        lam, v = np.asarray(c["lam"], float), np.asarray(c["v"], float)
        eta, eps = np.asarray(c["eta"], float), np.asarray(c["epsilon"], float)
        logLam = p["logLambdaTilde"]
        logCeta = gammaln(eta.sum()) - np.sum(gammaln(eta))
        logCeps = gammaln(eps.sum(axis=1)) - np.sum(gammaln(eps), axis=1)
        H = 0.0
        mWm = np.empty(Sr)
        trW = np.empty(Sr)
        for k in range(Sr):
            Wk = c["W"][k] if covmode == "full" else np.diag(c["W"][k])
            logBk = -(v[k] / 2) * np.log(np.linalg.det(Wk)) - (v[k] * d / 2) * np.log(2) \
                - (d * (d - 1) / 4) * np.log(np.pi) - np.sum(gammaln(0.5 * (v[k] + 1 - ii)))
            H += -logBk - 0.5 * (v[k] - d - 1) * logLam[k] + 0.5 * v[k] * d
            df = c["m"][k] - m0
            mWm[k] = df @ Wk @ df
            trW[k] = np.trace(W0inv @ Wk)
        Lt51 = 0.5 * np.sum(const2 + logLam - d * lam0 / lam - lam0 * v * mWm)
        Lt52 = Sr * logB0 + 0.5 * (v0 - d - 1) * np.sum(logLam) - 0.5 * np.sum(v * trW)
        Lt5 += Lt51 + Lt52
        Lt9 += logCeta + float((eta - 1) @ p["logPi"]) + np.sum(logCeps + np.sum((eps - 1) * p["logA"], axis=1))
        Lt10 += 0.5 * np.sum(logLam + d * np.log(lam / (2 * np.pi))) - 0.5 * d * Sr - H
    terms = dict(Lt1=Lt1, Lt2=Lt2, Lt3=Lt3, Lt4=Lt4, Lt5=Lt5, Lt6=Lt6, Lt7=Lt7, Lt8=Lt8, Lt9=Lt9, Lt10=Lt10)
    LB = Lt1 + Lt2 + Lt3 + Lt4 + Lt5 + Lt6 - Lt7 - Lt8 - Lt9 - Lt10
    return float(LB), terms


# ----------------------------------------------------------------------------- VB-HMM forward-backward
def vbhmm_fb_prepass(varpar: dict) -> dict:
    """vbhmm_fb.m:28-80: logLambdaTilde, logATilde, logPiTilde and the MEX arguments derived from them.
    varpar (natural): v[K], W[K,dim,dim], epsilon[K(from),K(to)], alpha[K], m[K,dim], beta[K]."""
    from scipy.special import psi
    v, W, eps, al = (np.asarray(varpar[k], float) for k in ("v", "W", "epsilon", "alpha"))
    K, dim = np.asarray(varpar["m"]).shape
    ii = np.arange(1, dim + 1)
    lLT = np.array([psi(0.5 * (v[k] + 1) - 0.5 * ii).sum() + dim * np.log(2) + np.log(np.linalg.det(W[k]))
                    for k in range(K)])
    logA = psi(eps) - psi(eps.sum(1, keepdims=True))
    logPi = psi(al) - psi(al.sum())
    return dict(m=np.asarray(varpar["m"], float), W=W, v=v, beta=np.asarray(varpar["beta"], float),
                logLambdaTilde=lLT, const_denominator=dim * np.log(2 * np.pi) / 2, pz1=np.exp(logPi), A=np.exp(logA),
                logPiTilde=logPi, logATilde=logA)


def vbhmm_fb(data: list, par: dict) -> dict:
    """vbhmm_fb.m:264-380 (the MATLAB branch), sequence by sequence.  data[n]: [T_n, dim]; par: vbhmm_fb_prepass.
    Returns natural arrays: logrho_Saved / gamma_all [N, maxT, K], xi_sum [N, from, to], phi_norm [N]."""
    N = len(data)
    K, dim = par["m"].shape
    maxT = max((len(x) for x in data), default=0)
    out = dict(logrho_Saved=np.zeros((N, maxT, K)), gamma_all=np.zeros((N, maxT, K)), xi_sum=np.zeros((N, K, K)),
               phi_norm=np.zeros(N))
    A, pz1 = par["A"], par["pz1"]
    for n, x in enumerate(data):
        x = np.asarray(x, float).reshape(len(x), dim)
        T = len(x)
        if T == 0:
            continue
        diff = x[:, None, :] - par["m"][None]                                       # [T, K, dim]
        mterm = np.einsum("tki,kij,tkj->tk", diff, par["W"], diff)
        delta = dim / par["beta"][None] + par["v"][None] * mterm
        logrho = 0.5 * par["logLambdaTilde"][None] - 0.5 * delta - par["const_denominator"]
        out["logrho_Saved"][n, :T] = logrho
        mx = logrho.max(1)
        p = np.exp(logrho - mx[:, None])
        alpha, c = np.zeros((T, K)), np.zeros(T)
        for t in range(T):
            d = pz1 * p[0] if t == 0 else (alpha[t - 1] @ A) * p[t]
            c[t] = d.sum()
            alpha[t] = d / c[t]
        beta = np.ones(K)
        gam = np.zeros((T, K))
        gam[T - 1] = alpha[T - 1]
        xi = np.zeros((K, K))
        for t in range(T - 2, -1, -1):
            bpi = beta * p[t + 1]
            beta = (bpi @ A.T) / c[t + 1]
            gam[t] = alpha[t] * beta
            xi += (A * np.outer(alpha[t], bpi)) / c[t + 1]
        out["gamma_all"][n, :T] = gam
        out["xi_sum"][n] = xi
        out["phi_norm"][n] = np.log(c).sum() + mx.sum()
    return out


def vbhmm_em_stats(data: list, fbstats: dict) -> dict:
    """vbhmm_em.m:158-246 (responsibility sums, transition counts, ROI means and scatter matrices) and the
    fb-dependent lower-bound terms vbhmm_em_lb.m:214,217, from the outputs of vbhmm_fb (natural layout)."""
    gam, xi = fbstats["gamma_all"], fbstats["xi_sum"]                    # [N, maxT, K], [N, from, to]
    N, maxT, K = gam.shape
    dim = np.asarray(data[0]).reshape(len(data[0]), -1).shape[1] if len(data[0]) else 1
    X = np.zeros((N, maxT, dim))
    for n, x in enumerate(data):
        if len(x):
            X[n, :len(x)] = np.asarray(x, float).reshape(len(x), -1)
    Nk1 = gam[:, 0].sum(0) + 1e-50
    Nk = gam.sum((0, 1)) + 1e-50
    M = xi.sum(0)
    xbar = np.einsum("ntk,ntd->kd", gam, X) / Nk[:, None]
    S = np.zeros((K, dim, dim))
    for k in range(K):
        d1 = X - xbar[k]
        S[k] = np.einsum("nt,nti,ntj->ij", gam[:, :, k], d1, d1) / Nk[k]
    return dict(Nk1=Nk1, Nk=Nk, M=M, xbar=xbar, S=S, Lt63=float((gam * fbstats["logrho_Saved"]).sum()),
                Lt64=float(fbstats["phi_norm"].sum()))


# ---------------------------------------------------------------------------- VB-HMM learning loop
# Restatement of /root/reference/src/hmm/vbhmm_init.m, vbhmm_em.m and vbhmm_em_lb.m (no groups, no
# derivatives, no fix_cov / fix_clusters).  Parity status: the forward-backward part is pinned on the
# reference's own C (oracle/_ref/libvbhmm_fb_ref.so); M-step and lower bound exist in MATLAB only, so
# this restatement is unpinned beyond the MATLAB builtins' stand-ins (scipy psi / gammaln, numpy inv / det).

def vbhmm_init_from_gmm(gmm: dict, Ntot: int, ini: dict) -> dict:
    """vbhmm_init.m:128-200 for initmode 'initgmm' (or the GMM any other mode produced).
    gmm: prior[K], mean[K,dim], cov[K,dim,dim]; Ntot = total number of observations (vbhmm_init.m:19-20).
    Returns the variational parameters alpha, epsilon[from,to], beta, v, m[K,dim], W[K,dim,dim]."""
    pc = np.asarray(gmm["prior"], float)
    mu = np.asarray(gmm["mean"], float)
    Sg = np.asarray(gmm["cov"], float)
    K, dim = mu.shape
    m0 = np.asarray(ini["mu0"], float).reshape(dim)
    W0 = np.asarray(ini["W0"], float).ravel()
    W0inv = np.linalg.inv(np.eye(dim) * W0[0] if W0.size == 1 else np.diag(W0))        # :136-152
    Nk = Ntot * pc                                                                      # :166
    Nk2 = Ntot * np.full(K, 1.0 / K)                                                    # :167
    alpha = ini["alpha0"] + Nk2                                                         # :189
    epsilon = np.tile(ini["epsilon0"] + Nk2, (K, 1))                                    # :190-192
    beta = ini["beta0"] + Nk                                                            # :193
    v = ini["v0"] + Nk + 1                                                              # :194
    m = (ini["beta0"] * m0[None] + Nk[:, None] * mu) / beta[:, None]                    # :195
    W = np.zeros((K, dim, dim))
    for k in range(K):                                                                  # :197-201
        mult1 = ini["beta0"] * Nk[k] / (ini["beta0"] + Nk[k])
        d3 = mu[k] - m0
        W[k] = np.linalg.inv(W0inv + Nk[k] * Sg[k] + mult1 * np.outer(d3, d3))
    return dict(alpha=alpha, epsilon=epsilon, beta=beta, v=v, m=m, W=W)


def vbhmm_em_lb(st: dict, varpar: dict, fbstats: dict, ini: dict, N: int) -> float:
    """vbhmm_em_lb.m:74-258 (value only).  st: Nk, M, xbar, S (= t1_S) of vbhmm_em_stats; varpar: the parameters
    the E-step ran with; fbstats: full outputs of vbhmm_fb plus logLambdaTilde / logPiTilde / logATilde."""
    from scipy.special import gammaln
    v, W, eps, al, m, beta = (np.asarray(varpar[k], float) for k in ("v", "W", "epsilon", "alpha", "m", "beta"))
    K, dim = m.shape
    a0, e0, b0, v0 = ini["alpha0"], ini["epsilon0"], ini["beta0"], ini["v0"]
    m0 = np.asarray(ini["mu0"], float).reshape(dim)
    W0 = np.asarray(ini["W0"], float).ravel()
    W0inv = np.linalg.inv(np.eye(dim) * W0[0] if W0.size == 1 else np.diag(W0))
    logdetW0inv = dim * np.log(W0inv[0, 0]) if W0.size == 1 else np.log(np.diag(W0inv)).sum()      # :75-82
    lLT, logPi, logA = fbstats["logLambdaTilde"], fbstats["logPiTilde"], fbstats["logATilde"]
    ii = np.arange(1, dim + 1)
    logCalpha0 = gammaln(K * a0) - K * gammaln(a0)                                                  # :84
    logCeps0 = np.full(K, gammaln(K * e0) - K * gammaln(e0))                                        # :85-87
    logB0 = (v0 / 2) * logdetW0inv - (v0 * dim / 2) * np.log(2) - (dim * (dim - 1) / 4) * np.log(np.pi) \
        - gammaln(0.5 * (v0 + 1 - ii)).sum()                                                        # :88-89
    logCalpha = gammaln(al.sum()) - gammaln(al).sum()                                               # :92
    logCeps = np.array([gammaln(eps[k].sum()) - gammaln(eps[k]).sum() for k in range(K)])           # :93-95
    H = 0.0
    trSW, xWx, mWm, trW0W = np.zeros(K), np.zeros(K), np.zeros(K), np.zeros(K)
    for k in range(K):                                                                              # :107-119
        logBk = -(v[k] / 2) * np.log(np.linalg.det(W[k])) - (v[k] * dim / 2) * np.log(2) \
            - (dim * (dim - 1) / 4) * np.log(np.pi) - gammaln(0.5 * (v[k] + 1 - ii)).sum()
        H = H - logBk - 0.5 * (v[k] - dim - 1) * lLT[k] + 0.5 * v[k] * dim
        trSW[k] = np.trace(st["S"][k] @ W[k])
        d = st["xbar"][k] - m[k]
        xWx[k] = d @ W[k] @ d
        d = m[k] - m0
        mWm[k] = d @ W[k] @ d
        trW0W[k] = np.trace(W0inv @ W[k])
    Nk, M = st["Nk"], st["M"]
    Lt1 = 0.5 * (Nk * (lLT - dim / beta - v * trSW - v * xWx - dim * np.log(2 * np.pi))).sum()      # :124-125
    gamma1 = fbstats["gamma_all"][:, 0, :]                                                          # :128  [N, K]
    Lt2a = (gamma1 * logPi[None]).sum()                                                             # :143-146
    Lt2b = (M * logA).sum()                                                                         # :162
    Lt2 = Lt2a + Lt2b
    Lt3 = logCalpha0 + (a0 - 1) * logPi.sum()                                                       # :176
    Lt4 = (logCeps0 + (e0 - 1) * logA.sum(1)).sum()                                                 # :186-189
    Lt51 = 0.5 * (dim * np.log(b0 / (2 * np.pi)) + lLT - dim * b0 / beta - b0 * v * mWm).sum()      # :201
    Lt52 = K * logB0 + 0.5 * (v0 - dim - 1) * lLT.sum() - 0.5 * (v * trW0W).sum()                   # :202
    Lt5 = Lt51 + Lt52
    Lt63 = (fbstats["gamma_all"] * fbstats["logrho_Saved"]).sum()                                   # :214
    Lt64 = fbstats["phi_norm"].sum()                                                                # :217
    Lt6 = Lt2a + Lt2b + Lt63 - Lt64                                                                 # :221
    Lt71 = ((al - 1) * logPi).sum() + logCalpha                                                     # :225
    Lt72 = (((eps - 1) * logA).sum(1) + logCeps).sum()                                              # :235-238
    Lt7 = Lt71 + Lt72
    Lt8 = 0.5 * (lLT + dim * np.log(beta / (2 * np.pi))).sum() - 0.5 * dim * K - H                  # :253
    return float(Lt1 + Lt2 + Lt3 + Lt4 + Lt5 - Lt6 - Lt7 - Lt8)                                     # :257


def vbhmm_em(data: list, varpar0: dict, ini: dict, fb=None) -> dict:
    """vbhmm_em.m:108-418 from given initial variational parameters (vbhmm_init.m's output): E-step, sums, lower
    bound, convergence test |dL / L_last| <= minDiff from iteration 2 on, M-step (which still runs once after the
    test fires, :411-414).  fb(data, par) -> fbstats defaults to the numpy forward-backward above.
    Returns varpar (final), LL, LLs (per iteration), iter, prior, trans, pdf mean / cov (vbhmm_em.m:429-460)."""
    fb = fb or vbhmm_fb
    vp = {k: np.array(varpar0[k], float) for k in ("alpha", "epsilon", "beta", "v", "m", "W")}
    K, dim = vp["m"].shape
    N = len(data)
    m0 = np.asarray(ini["mu0"], float).reshape(dim)
    W0 = np.asarray(ini["W0"], float).ravel()
    W0inv = np.linalg.inv(np.eye(dim) * W0[0] if W0.size == 1 else np.diag(W0))
    L, lastL, LLs = -np.finfo(float).max, -np.finfo(float).max, []
    it = 0
    unstable = False
    for it in range(1, int(ini["maxIter"]) + 1):
        par = vbhmm_fb_prepass(vp)
        fbs = fb(data, par)
        fbs = dict(fbs, logLambdaTilde=par["logLambdaTilde"], logPiTilde=par["logPiTilde"], logATilde=par["logATilde"])
        st = vbhmm_em_stats(data, fbs)
        if it > 1:
            lastL = L
        L = vbhmm_em_lb(st, vp, fbs, ini, N)
        LLs.append(L)
        do_break = False
        if it > 1 and abs((L - lastL) / lastL) <= ini["minDiff"]:                                   # :293-312
            do_break = True
        if it == int(ini["maxIter"]):
            do_break = True
        if np.isnan(L):                                                                             # :320-330
            do_break, unstable, L = True, True, -np.inf
            LLs[-1] = L
        if do_break and unstable:
            break
        # M-step, vbhmm_em.m:353-400
        vp["alpha"] = ini["alpha0"] + st["Nk1"]
        vp["epsilon"] = ini["epsilon0"] + st["M"]
        Nk = st["Nk"]
        vp["beta"] = ini["beta0"] + Nk
        vp["v"] = ini["v0"] + Nk + 1
        vp["m"] = (ini["beta0"] * m0[None] + Nk[:, None] * st["xbar"]) / vp["beta"][:, None]
        W = np.zeros((K, dim, dim))
        for k in range(K):
            mult1 = ini["beta0"] * Nk[k] / (ini["beta0"] + Nk[k])
            d3 = st["xbar"][k] - m0
            Wk = np.linalg.inv(W0inv + Nk[k] * st["S"][k] + mult1 * np.outer(d3, d3))
            W[k] = (Wk + Wk.T) / 2
        vp["W"] = W
        if do_break:
            break
    C = np.zeros((K, dim, dim))
    for k in range(K):                                                                              # :402-414
        Ck = np.linalg.inv(vp["W"][k]) / (vp["v"][k] - dim - 1 if vp["v"][k] > dim + 1 else vp["v"][k])
        C[k] = (Ck + Ck.T) / 2
    prior = vp["alpha"] / vp["alpha"].sum()                                                         # :430-432
    rs = vp["epsilon"].sum(1, keepdims=True)
    trans = vp["epsilon"] / np.where(rs == 0, 1.0, rs)                                              # :434-439
    return dict(varpar=vp, LL=L, LLs=np.array(LLs), iter=it, prior=prior, trans=trans, mean=vp["m"].copy(), cov=C,
                M=st["M"], N1=st["Nk1"], N=st["Nk"], unstable=unstable)


def gmm_em(points: np.ndarray, start_idx, K: int, max_iter: int = 100, tol: float = 1e-5, reg: float = 1e-10) -> dict:
    """The textbook GMM EM that csrc/vbhmm_gmm.inl runs per block (stand-in for the Statistics Toolbox's
    gmdistribution.fit of /root/reference/src/hmm/vbhmm_init.m:60-61): start from the observations ``start_idx`` as
    means, the diagonal of the pooled covariance, weights 1/K; stop on |dll| < tol |ll| or max_iter."""
    X = np.asarray(points, float)
    n, d = X.shape
    gmean = X.mean(0)
    gvar = ((X - gmean) ** 2).sum(0) / (n - 1 if n > 1 else 1) + reg
    mu = X[np.asarray(start_idx)].copy()
    cov = np.tile(np.diag(gvar), (K, 1, 1))
    w = np.full(K, 1.0 / K)
    ll_old, it = 0.0, 0
    for it in range(1, max_iter + 1):
        P = np.linalg.inv(cov)
        logdet = np.linalg.slogdet(cov)[1]
        df = X[:, None, :] - mu[None]
        q = np.einsum("nki,kij,nkj->nk", df, P, df)
        lp = np.log(w)[None] - 0.5 * (d * np.log(2 * np.pi) + logdet[None] + q)
        mx = lp.max(1, keepdims=True)
        e = np.exp(lp - mx)
        s = e.sum(1, keepdims=True)
        ll = float((mx + np.log(s)).sum())
        g = e / s
        Nk = g.sum(0)
        Xc = X - gmean
        for k in range(K):
            if Nk[k] > 1e-300:
                w[k] = Nk[k] / n
                dk = (g[:, k, None] * Xc).sum(0) / Nk[k]
                mu[k] = gmean + dk
                cov[k] = np.einsum("n,ni,nj->ij", g[:, k], Xc, Xc) / Nk[k] - np.outer(dk, dk) + reg * np.eye(d)
        if it > 1 and abs(ll - ll_old) < tol * abs(ll):
            break
        ll_old = ll
    return dict(prior=w, mean=mu, cov=(cov + np.transpose(cov, (0, 2, 1))) / 2, loglik=ll, iters=it)
