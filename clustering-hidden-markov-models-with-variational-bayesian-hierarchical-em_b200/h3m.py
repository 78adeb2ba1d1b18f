"""Host mirror of the reference's EM driver around the E-step (Python, numpy for the O(K*S) host parts).

``vbhem_h3m_c_step_fc(h3m_r, base, vbhemopt)`` follows the control flow of
/root/reference/src/vbhem/vbhem_h3m_c_step_fc.m:115-433 -- E-step, lower bound, relative-change
convergence test from the third pass on, M-step (also after the breaking pass), NaN guard -- with the same
option names (``tau``, ``max_iter``, ``minDiff``, ``alpha0``, ``eta0``, ``epsilon0``, ``m0``, ``lambda0``, ``W0``,
``v0``).  The E-step (pre-pass, L_elbo, responsibilities, statistics reduction and normalisation) runs on the
GPU through the C ABI; what stays on the host is the part SURVEY.md 8(f) ranks "next": the M-step formulas
(/root/reference/src/vbhem/vbhem_mstep_component.m:42-70) and the lower bound
(/root/reference/src/vbhem/vbhemh3m_lb.m:64-186), both O(K*S*d^2).

Centre hyper-parameters travel as one dict of stacked arrays (natural layout):
    alpha[K], eta[K,S], epsilon[K,S,S] (from,to), m[K,S,d], W[K,S,d,d] | [K,S,d], v[K,S], lam[K,S], Sr[K]
"""
from __future__ import annotations

import time

import numpy as np
from scipy.special import gammaln, psi


def hmms_to_h3m_hem(hmms: list, covmode: str = "full", use_post: bool = True, Nv: float = 100.0) -> dict:
    """List of learned HMMs -> natural base dict (hmms_to_h3m_hem.m:19-140).

    Each HMM is a dict with ``prior[S]``, ``trans[S,S]``, ``mean[S,d]``, ``cov[S,d,d]`` and, for ``use_post``,
    ``varpar = dict(alpha[S], epsilon[S,S], beta[S])``; ``None`` entries become zero-weight dummies (:113-133).
    """
    live = [h for h in hmms if h is not None]
    d = np.asarray(live[0]["mean"]).shape[1]
    S = max(len(h["prior"]) for h in live)
    N = len(hmms)
    full = covmode == "full"
    out = dict(Sb=np.ones(N, np.int32), prior=np.zeros((N, S)), A=np.zeros((N, S, S)), mu=np.zeros((N, S, d)),
               cov=np.zeros((N, S, d, d) if full else (N, S, d)))
    omega = np.ones(N)
    for i, h in enumerate(hmms):
        if h is None:                                   # dummy one-state HMM with weight 0
            out["prior"][i, 0] = 1.0
            out["A"][i, 0, 0] = 1.0
            out["cov"][i, 0] = np.eye(d) if full else np.ones(d)
            omega[i] = 0.0
            continue
        s = len(h["prior"])
        out["Sb"][i] = s
        if use_post:
            al, ep = np.asarray(h["varpar"]["alpha"], float), np.asarray(h["varpar"]["epsilon"], float)
            out["prior"][i, :s] = np.exp(psi(al) - psi(al.sum()))
            out["A"][i, :s, :s] = np.exp(psi(ep) - psi(ep.sum(1, keepdims=True)))
            tb = (np.asarray(h["varpar"]["beta"], float) + 1.0) / np.asarray(h["varpar"]["beta"], float)
        else:
            out["prior"][i, :s] = h["prior"]
            out["A"][i, :s, :s] = h["trans"]
            tb = np.ones(s)
        out["mu"][i, :s] = h["mean"]
        cov = np.asarray(h["cov"], float)
        out["cov"][i, :s] = tb[:, None, None] * cov if full else tb[:, None] * np.diagonal(cov, axis1=1, axis2=2)
    omega /= omega.sum()
    out["Ntilde"] = Nv * N * omega                      # vbhem_h3m_c_step_fc.m:26-30
    return out


def _W0(opt, d):
    w = np.atleast_1d(np.asarray(opt["W0"], float))
    return (w[0] * np.eye(d), True) if w.size == 1 else (np.diag(w), False)


def vbhem_mstep(fin: dict, opt: dict, covmode: str = "full") -> dict:
    """vbhem_mstep_component.m:42-70 for all centres at once.  fin: the normalised statistics of
    BaseSet.finalize (Nj, Nj_rho1[K,S], Nj_rho2rho[K,S,S], Nj_rho[K,S], y_bar[K,S,d], S_plus_C[K,S,d,d]|[K,S,d])."""
    R = fin["Nj_rho"]
    K, S = R.shape
    d = fin["y_bar"].shape[2]
    W0m, _ = _W0(opt, d)
    if opt["v0"] <= d - 1:
        raise ValueError("v0 not large enough")
    W0inv = np.linalg.inv(W0m)
    m0 = np.asarray(opt["m0"], float).reshape(d)
    lam0, v0 = float(opt["lambda0"]), float(opt["v0"])
    diff = fin["y_bar"] - m0
    SC = fin["S_plus_C"] if covmode == "full" else np.einsum("ksi,ij->ksij", fin["S_plus_C"], np.eye(d))
    mult1 = lam0 * R / (lam0 + R)
    tW = np.linalg.inv(W0inv + R[..., None, None] * SC + mult1[..., None, None] * diff[..., :, None] * diff[..., None, :])
    Wn = 0.5 * (tW + np.swapaxes(tW, -1, -2))
    return dict(alpha=opt["alpha0"] + fin["Nj"], eta=opt["eta0"] + fin["Nj_rho1"],
                epsilon=opt["epsilon0"] + fin["Nj_rho2rho"], lam=lam0 + R, v=v0 + R + 1.0,
                m=(lam0 * m0 + R[..., None] * fin["y_bar"]) / (lam0 + R)[..., None],
                W=Wn if covmode == "full" else np.diagonal(Wn, axis1=-2, axis2=-1).copy())


def vbhemh3m_lb(Lt1, Lt7, Nj, pre: dict, hyp: dict, opt: dict, covmode: str = "full"):
    """vbhemh3m_lb.m:64-186.  pre: logOmega[K], logPi[K,S], logA[K,S,S], logLambdaTilde[K,S] (device pre-pass)."""
    K, S = hyp["eta"].shape
    d = hyp["m"].shape[2]
    a0, e0, p0 = float(opt["alpha0"]), float(opt["eta0"]), float(opt["epsilon0"])
    lam0, v0 = float(opt["lambda0"]), float(opt["v0"])
    m0 = np.asarray(opt["m0"], float).reshape(d)
    W0m, iid = _W0(opt, d)
    W0inv = np.linalg.inv(W0m)
    logdetW0inv = d * np.log(W0inv[0, 0]) if iid else np.sum(np.log(np.diag(W0inv)))
    ii = np.arange(1, d + 1, dtype=float)
    logB0 = (v0 / 2) * logdetW0inv - (v0 * d / 2) * np.log(2) - (d * (d - 1) / 4) * np.log(np.pi) \
        - np.sum(gammaln(0.5 * (v0 + 1 - ii)))
    lO, lP, lA, lL = pre["logOmega"], pre["logPi"], pre["logA"], pre["logLambdaTilde"]
    al, eta, eps, lam, v = hyp["alpha"], hyp["eta"], hyp["epsilon"], hyp["lam"], hyp["v"]
    W = hyp["W"] if covmode == "full" else np.einsum("ksi,ij->ksij", hyp["W"], np.eye(d))
    Lt2 = float(np.dot(Nj, lO))
    Lt3 = K * (gammaln(S * e0) - S * gammaln(e0)) + (e0 - 1) * lP.sum()
    Lt4 = K * S * (gammaln(S * p0) - S * gammaln(p0)) + (p0 - 1) * lA.sum()
    Lt6 = gammaln(K * a0) - K * gammaln(a0) + (a0 - 1) * lO.sum()
    Lt8 = gammaln(al.sum()) - gammaln(al).sum() + float(np.dot(al - 1, lO))
    df = hyp["m"] - m0
    mWm = np.einsum("ksi,ksij,ksj->ks", df, W, df)
    trW = np.einsum("ij,ksji->ks", W0inv, W)
    logdetW = np.log(np.linalg.det(W))
    logB = -(v / 2) * logdetW - (v * d / 2) * np.log(2) - (d * (d - 1) / 4) * np.log(np.pi) \
        - gammaln(0.5 * (v[..., None] + 1 - ii)).sum(-1)
    H = (-logB - 0.5 * (v - d - 1) * lL + 0.5 * v * d).sum(1)
    Lt5 = (0.5 * (d * np.log(lam0 / (2 * np.pi)) + lL - d * lam0 / lam - lam0 * v * mWm).sum()
           + K * S * logB0 + 0.5 * (v0 - d - 1) * lL.sum() - 0.5 * (v * trW).sum())
    Lt9 = (gammaln(eta.sum(1)) - gammaln(eta).sum(1) + ((eta - 1) * lP).sum(1)).sum() \
        + (gammaln(eps.sum(2)) - gammaln(eps).sum(2) + ((eps - 1) * lA).sum(2)).sum()
    Lt10 = (0.5 * (lL + d * np.log(lam / (2 * np.pi))).sum(1) - 0.5 * d * S - H).sum()
    LB = Lt1 + Lt2 + Lt3 + Lt4 + Lt5 + Lt6 - Lt7 - Lt8 - Lt9 - Lt10
    return float(LB), dict(Lt1=Lt1, Lt2=Lt2, Lt3=Lt3, Lt4=Lt4, Lt5=Lt5, Lt6=Lt6, Lt7=Lt7, Lt8=Lt8, Lt9=Lt9, Lt10=Lt10)


HYPS_MAX = dict(alpha0=1.0686e+13, eta0=1.0686e+13, epsilon0=1.0686e+13, v0=1e4, lambda0=1.0686e+13, W0=1.0686e+13)
HYPS_MIN = dict(alpha0=1.0686e-13, eta0=1.0686e-13, epsilon0=1.0686e-13, v0=None, lambda0=1.0686e-13, W0=1.0686e-13)
"""vbhem_h3m_cluster.m:182-200 (``v0``'s minimum is exp(-20) + D - 1)."""


def vbhem_clip_hyps(opt: dict) -> tuple[dict, dict]:
    """vbhem_clip_hyps.m:1-86: hyper-parameters clipped to [hyps_min, hyps_max]; clipped[name][i] = +1 / -1 / 0."""
    d = len(np.atleast_1d(opt["m0"]))
    hmax = dict(HYPS_MAX, **opt.get("hyps_max", {}))
    hmin = dict(HYPS_MIN, **opt.get("hyps_min", {}))
    if hmin.get("v0") is None:
        hmin["v0"] = 2.0612e-09 + d - 1
    out, clipped = dict(opt), {}
    for name in hmax:
        val = np.atleast_1d(np.asarray(opt[name], float)).copy()
        c = np.zeros(val.size, int)
        hi, lo = val >= hmax[name], val <= hmin[name]
        val[hi], c[hi] = hmax[name], +1
        val[lo], c[lo] = hmin[name], -1
        out[name] = float(val[0]) if np.ndim(opt[name]) == 0 else val
        clipped[name] = c
    return out, clipped


def vbhemh3m_lb_deriv(pre: dict, hyp: dict, opt: dict, covmode: str = "full", clipped: dict | None = None) -> dict:
    """vbhemh3m_lb.m:202-343: derivatives of the lower bound with respect to the prior hyper-parameters at fixed
    variational posteriors (``hyp`` / ``pre`` = the centre hyper-parameters and pre-pass results of the E-step the
    bound belongs to), then with respect to their transformed versions (the optimiser's coordinates):
    d_logalpha0, d_logeta0, d_logepsilon0, d_logv0D1, d_sqrtv0D1, d_loglambda0, d_sqrtlambda0, d_sqrtW0inv, d_logW0, d_m0."""
    K, S = hyp["eta"].shape
    d = hyp["m"].shape[2]
    a0, e0, p0 = float(opt["alpha0"]), float(opt["eta0"]), float(opt["epsilon0"])
    lam0, v0 = float(opt["lambda0"]), float(opt["v0"])
    m0 = np.asarray(opt["m0"], float).reshape(d)
    W0m, iid = _W0(opt, d)
    W0inv = np.diag(1.0 / np.diag(W0m))
    logdetW0inv = np.log(np.diag(W0inv)).sum()
    ii = np.arange(1, d + 1, dtype=float)
    lam, v = hyp["lam"], hyp["v"]
    W = hyp["W"] if covmode == "full" else np.einsum("ksi,ij->ksij", hyp["W"], np.eye(d))
    df = hyp["m"] - m0
    mWm = np.einsum("ksi,ksij,ksj->ks", df, W, df)
    D = {}
    D["alpha0"] = np.atleast_1d(K * psi(K * a0) - K * psi(a0) + pre["logOmega"].sum())                        # :205-209
    D["eta0"] = np.atleast_1d(K * (S * psi(S * e0) - S * psi(e0)) + pre["logPi"].sum())                        # :213-217
    D["epsilon0"] = np.atleast_1d(K * S * (S * psi(S * p0) - S * psi(p0)) + pre["logA"].sum())                 # :221-225
    d_logB0_v0 = 0.5 * logdetW0inv - (d / 2) * np.log(2) - 0.5 * psi(0.5 * (v0 + 1 - ii)).sum()
    D["v0"] = np.atleast_1d(K * S * d_logB0_v0 + 0.5 * pre["logLambdaTilde"].sum())                           # :229-230
    D["lambda0"] = np.atleast_1d(0.5 * (d / lam0 - d / lam - v * mWm).sum())                                   # :233-241
    diagW = np.diagonal(W, axis1=-2, axis2=-1)                                                                 # [K,S,d]
    if iid:                                                                                                    # :244-267
        w0inv = W0inv[0, 0]
        D["W0"] = np.atleast_1d(K * S * (-0.5 * v0 * d * w0inv) - 0.5 * (-v * w0inv ** 2 * diagW.sum(-1)).sum())
        myW0 = np.atleast_1d(1.0 / w0inv)
    else:                                                                                                      # :270-293
        w0inv = np.diag(W0inv)
        D["W0"] = K * S * (-0.5 * v0 * w0inv) - 0.5 * (-(v[..., None]) * w0inv ** 2 * diagW).sum((0, 1))
        myW0 = 1.0 / w0inv
    D["m0"] = lam0 * np.einsum("ks,ksij,ksj->i", v, W, df)                                                     # :300-309
    return transform_derivs(D, opt, clipped)


def transform_derivs(D: dict, opt: dict, clipped: dict | None = None) -> dict:
    """vbhemh3m_lb.m:312-339: raw derivatives D[name] (alpha0, eta0, epsilon0, v0, lambda0, W0, m0; from the host
    formulas above or from the device, ``BaseSet.em_lb_deriv``) -> zeroed where a clipped hyper-parameter would move
    further out, then chained through the optimiser's transforms."""
    d = len(np.atleast_1d(opt["m0"]))
    a0, e0, p0 = float(opt["alpha0"]), float(opt["eta0"]), float(opt["epsilon0"])
    lam0, v0 = float(opt["lambda0"]), float(opt["v0"])
    myW0 = np.atleast_1d(np.asarray(opt["W0"], float))
    D = {k: np.array(np.atleast_1d(v), float) for k, v in D.items()}
    for name, c in (clipped or {}).items():                                                                    # :312-327
        if name in D:
            g = D[name]
            for i in range(min(len(c), len(g))):
                if (c[i] == +1 and g[i] > 0) or (c[i] == -1 and g[i] < 0):
                    g[i] = 0.0
    return dict(d_logalpha0=D["alpha0"] * a0, d_logeta0=D["eta0"] * e0, d_logepsilon0=D["epsilon0"] * p0,      # :330-339
                d_logv0D1=D["v0"] * (v0 - d + 1), d_sqrtv0D1=D["v0"] * 2 * np.sqrt(v0 - d + 1),
                d_loglambda0=D["lambda0"] * lam0, d_sqrtlambda0=D["lambda0"] * 2 * np.sqrt(lam0),
                d_sqrtW0inv=D["W0"] * myW0 ** 1.5 * (-2), d_logW0=D["W0"] * myW0, d_m0=D["m0"], raw=D)


def device_derivs_to_dict(row: np.ndarray, d: int, nW0: int) -> dict:
    """One row of ``BaseSet.em_lb_deriv`` / ``em_run_groups_deriv`` ([21]: alpha0, eta0, epsilon0, v0, lambda0, W0[8],
    m0[8]) -> the raw-derivative dict of ``transform_derivs``."""
    return dict(alpha0=row[0:1], eta0=row[1:2], epsilon0=row[2:3], v0=row[3:4], lambda0=row[4:5],
                W0=row[5:5 + nW0], m0=row[13:13 + d])


def vbhem_h3m_c_step_fc(h3m_r: dict, base, vbhemopt: dict, covmode: str = "full", estep=None, verbose: int = 0,
                        device_loop: bool = False):
    """One EM run from the centre hyper-parameters ``h3m_r`` (vbhem_h3m_c_step_fc.m:115-433).

    ``device_loop=True`` runs the whole loop through the C ABI (``vbhem_em_run``): lower bound and M-step on
    the device, hyper-parameters resident in HBM; the default keeps them on the host (numpy), which is also
    what a multi-GPU ``estep`` callback uses.

    ``base``: a ``BaseSet`` (device-resident base mixture).  ``estep(hyp, tau)`` may be supplied to run the
    E-step elsewhere (e.g. ``dist.ShardedEStep.estep_hyp`` on several GPUs); it must return the normalised
    statistics plus ``pre`` / ``lb`` as ``BaseSet`` does.
    Returns dict(hyp=final hyper-parameters, LogLs, LL, iters, ms_per_iter, Nj, unstable).
    """
    opt = vbhemopt
    tau, max_iter, min_diff = int(opt["tau"]), int(opt.get("max_iter", 200)), float(opt.get("minDiff", 1e-5))
    hyp = {k: np.array(v, dtype=np.int32 if k == "Sr" else np.float64) for k, v in h3m_r.items()}
    K, S = hyp["eta"].shape
    hyp.setdefault("Sr", np.full(K, S, np.int32))
    if device_loop:
        assert estep is None, "device_loop drives the handle's own E-step"
        t0 = time.perf_counter()
        base.em_begin(hyp, opt)
        out = base.em_run(tau, max_iter, min_diff)
        dt = time.perf_counter() - t0
        return dict(hyp=out["hyp"], LogLs=out["LogLs"], LL=out["LogLs"][-1], iters=out["iters"],
                    unstable=out["unstable"], Nj=out["hyp"]["alpha"] - float(opt["alpha0"]),
                    ms_per_iter=1e3 * dt / max(1, len(out["LogLs"])))

    def default_estep(h, t):
        out = base.estep_hyp(h, t, want_pairs=False)
        fin = base.finalize(out["buf"], h["Sr"], S)
        fin["pre"] = dict(logOmega=out["logOmega"], logPi=out["logPi"], logA=out["logA"],
                          logLambdaTilde=out["logLambdaTilde"])
        return fin

    run = estep or default_estep
    LogLs, lastL, it, unstable, times = [], -np.finfo(float).max, 0, False, []
    fin, dLL = None, None
    while True:
        t0 = time.perf_counter()
        fin = run(hyp, tau)                                               # E-step (GPU)
        L, _ = vbhemh3m_lb(fin["lb"][0], fin["lb"][1], fin["Nj"], fin["pre"], hyp, opt, covmode)
        do_break = False
        if it > 1:
            if abs((L - lastL) / lastL) <= min_diff:
                do_break = True
        if it == max_iter:
            do_break = True
        if np.isnan(L):                                                   # :338-354
            do_break, unstable, L = True, True, -np.inf
        if do_break and opt.get("calc_LLderiv"):                          # before the last M-step (:356-369)
            dLL = vbhemh3m_lb_deriv(fin["pre"], hyp, opt, covmode, opt.get("hyp_clipped"))
            if unstable:
                dLL = {k: (np.full_like(np.asarray(v, float), np.nan) if k != "raw" else v) for k, v in dLL.items()}
        if do_break and unstable:
            LogLs.append(L)
            break
        new = vbhem_mstep(fin, opt, covmode)                              # M-step (host, O(K*S*d^2))
        new["Sr"] = hyp["Sr"]
        hyp = new
        it += 1
        LogLs.append(L)
        lastL = L
        times.append(time.perf_counter() - t0)
        if verbose:
            print(f"iter {it}: L={L:.10g}")
        if do_break:
            break
    return dict(hyp=hyp, LogLs=np.array(LogLs), LL=LogLs[-1], iters=it, unstable=unstable, Nj=fin["Nj"], dLL=dLL,
                ms_per_iter=1e3 * float(np.median(times)) if times else float("nan"))


# ----------------------------------------------------------------------------- user entry: (K, S) model selection
CLUSTER_DEFAULTS = dict(alpha0=1.0, eta0=1.0, epsilon0=1.0, m0=(256.0, 192.0), W0=0.005, lambda0=1.0, v0=5.0,
                        trials=100, max_iter=200, minDiff=1e-5, Nv=100.0, tau=10, use_post=True, seed=None,
                        initmode="baseem", learn_hyps=0, remove_empty=1)
"""vbhem_h3m_cluster.m:141-179 (``initmode`` 'auto' of the reference tries 'baseem', 'gmmNew' and 'wtkmeans';
'baseem' and 'wtkmeans' are mirrored -- the latter with scikit-learn's k-means in place of the Statistics Toolbox's
-- and 'auto' means those two; 'gmmNew' needs ``gmdistribution.fit``).  ``learn_hyps`` defaults to 1 in the
reference (:179); here it is opt-in: 1 or a list of names runs ``vbhem_h3m_c_hyp`` on the best trial of every
(K, S) candidate (the reference optimises every distinct trial solution, vbhem_h3m_c.m:105-150)."""


def my_weighted_kmeans(cluster_num: int, it_max: int, point: np.ndarray, weight: np.ndarray, cluster_center: np.ndarray):
    """my_weighted_kmeans.m:1-100, restated: weighted k-means with the leave-one-out energy adjustment.
    point [dim, n], weight [n], cluster_center [dim, k] (initial).  Returns (cluster[n] 0-based, centres [dim, k],
    energy per iteration)."""
    point, weight = np.asarray(point, float), np.asarray(weight, float)
    k = cluster_num

    def centroids(cluster):
        cw = np.array([weight[cluster == j].sum() for j in range(k)])
        cen = np.stack([(point[:, cluster == j] * weight[cluster == j]).sum(1) for j in range(k)], 1)
        return np.where(cw > 0, cen / np.where(cw > 0, cw, 1.0), cen), cw

    def energy(cluster, cen, cw):
        f = ((point - cen[:, cluster]) ** 2).sum(0)
        ce = np.array([(weight[cluster == j] * f[cluster == j]).sum() for j in range(k)])
        with np.errstate(divide="ignore", invalid="ignore"):
            f = f * cw[cluster] / (cw[cluster] - weight)
        return f, ce

    d2 = ((point[:, None, :] - np.asarray(cluster_center, float)[:, :, None]) ** 2).sum(0)      # [k, n]
    cluster = d2.argmin(0)
    cen, cw = centroids(cluster)
    f, ce = energy(cluster, cen, cw)
    old = ce.sum()
    energies, it = [old], 0
    while it < it_max:
        # energy of moving point i into cluster j (non-members) / of keeping it where it is (members)
        fmat = ((point[:, None, :] - cen[:, :, None]) ** 2).sum(0) * (cw[:, None] / (cw[:, None] + weight[None, :]))
        fmat[cluster, np.arange(point.shape[1])] = f
        fmat = np.where(np.isnan(fmat), np.inf, fmat)          # MATLAB's min skips NaN (a singleton's own 0/0)
        cluster = fmat.argmin(0)
        cen, cw = centroids(cluster)
        f, ce = energy(cluster, cen, cw)
        new = ce.sum()
        if abs(new - old) < 1e-6:
            break
        old = new
        it += 1
        energies.append(new)
    return cluster, cen, np.array(energies)


def init_centres_wtkmeans(base: dict, K: int, S: int, seed: int, opt: dict, mode: str = "r0") -> dict:
    """``wtkmeans`` initialisation (vbhemhmm_init.m:294-425): all base emission means, weighted by their states'
    occupancy (``makeGMMweights``: p A^50 for mode '0', uniform for 'x'), are clustered into K groups by weighted
    k-means started from a plain k-means; each group's means are clustered again into the S state means of a centre.
    Transition / prior pseudo-counts from ``makeAprior`` ('u' uniform, 'r' random).  The reference's plain k-means is
    the Statistics Toolbox's (:331,360; seeds wtseed, wtseed + 1); scikit-learn's stands in.  Full or diagonal."""
    from sklearn.cluster import KMeans
    N, Sb_max, d = base["mu"].shape
    full = base["cov"].ndim == 4
    rng = np.random.Generator(np.random.PCG64(seed + 11))
    pts, wts = [], []
    for i in range(N):
        s = int(base["Sb"][i])
        p, A = base["prior"][i, :s], base["A"][i, :s, :s]
        if "0" in mode:
            for _ in range(50):
                p = p @ A
        elif "x" in mode:
            p = np.full(s, 1.0 / s)
        else:
            raise ValueError("unknown mode")
        pts.append(base["mu"][i, :s])
        wts.append(p)
    pts, wts = np.concatenate(pts).T, np.concatenate(wts)                               # [d, n], [n]
    wts = wts / wts.sum()

    def kmeans(x, k, rs):   # x [n, d] -> centres [<= k, d]
        k = min(k, len(np.unique(x, axis=0)))
        return KMeans(n_clusters=k, n_init=1, random_state=rs).fit(x).cluster_centers_

    init = kmeans(pts.T, K, seed)
    if init.shape[0] < K:
        init = np.concatenate([init, np.repeat(init[:1], K - init.shape[0], 0)])
    idx, _, _ = my_weighted_kmeans(K, 100, pts, wts, init.T)
    centres, empty = [], []
    for j in range(K):
        mean_j = pts[:, idx == j]
        if mean_j.shape[1] == 0:
            centres.append(None)
            empty.append(j)
        elif mean_j.shape[1] <= S:
            centres.append(np.concatenate([mean_j, np.repeat(mean_j[:, :1], S - mean_j.shape[1], 1)], 1))
        else:
            c = kmeans(mean_j.T, S, seed + 1).T
            centres.append(np.concatenate([c, np.repeat(c[:, :1], S - c.shape[1], 1)], 1))
    first = next(c for c in centres if c is not None)
    centres = [first if c is None else c for c in centres]
    NJ = opt["Nv"] * N / K
    Ns = NJ / S
    v = np.full((K, S), opt["v0"] + Ns + 1.0)
    c0 = base["cov"][0, 0]                                                              # h3m_b.hmm{1}.emit{1}.covars
    W1 = np.linalg.inv((v[0, 0] - d - 1.0) * c0) if full else 1.0 / ((v[0, 0] - d - 1.0) * c0)
    eta, eps = np.zeros((K, S)), np.zeros((K, S, S))
    for j in range(K):
        if "u" in mode:
            pr, A = np.full(S, 1.0 / S), np.full((S, S), 1.0 / S)
        elif "r" in mode:
            pr = rng.uniform(size=S)
            pr /= pr.sum()
            A = rng.uniform(size=(S, S))
            A /= A.sum(1, keepdims=True)
        else:
            raise ValueError("unknown mode")
        eta[j], eps[j] = pr * NJ + opt["eta0"], A * NJ + opt["epsilon0"]
    return dict(Sr=np.full(K, S, np.int32), alpha=np.full(K, opt["alpha0"] + NJ), eta=eta, epsilon=eps,
                m=np.stack([c.T for c in centres]), W=np.broadcast_to(W1, (K, S) + W1.shape).copy(), v=v,
                lam=np.full((K, S), opt["lambda0"] + Ns))


def stack_h3ms(models: list) -> tuple[dict, np.ndarray]:
    """Several H3Ms' centre hyper-parameter dicts -> one dict padded to the largest state count, plus ``group``
    (model index of every centre): the argument of ``BaseSet.em_begin(..., group=...)`` / ``estep_hyp``."""
    maxS = max(h["eta"].shape[1] for h in models)
    Kt = sum(h["eta"].shape[0] for h in models)
    first = models[0]
    out = dict(Sr=np.zeros(Kt, np.int32), alpha=np.zeros(Kt), eta=np.zeros((Kt, maxS)), epsilon=np.zeros((Kt, maxS, maxS)),
               m=np.zeros((Kt, maxS) + np.shape(first["m"])[2:]), W=np.zeros((Kt, maxS) + np.shape(first["W"])[2:]),
               v=np.zeros((Kt, maxS)), lam=np.zeros((Kt, maxS)))
    group = np.zeros(Kt, np.int32)
    j = 0
    for g, h in enumerate(models):
        K, S = h["eta"].shape
        sl = slice(j, j + K)
        out["Sr"][sl] = h.get("Sr", S)
        out["alpha"][sl] = h["alpha"]
        out["eta"][sl, :S] = h["eta"]
        out["epsilon"][sl, :S, :S] = h["epsilon"]
        out["m"][sl, :S] = h["m"]
        out["W"][sl, :S] = h["W"]
        out["v"][sl, :S] = h["v"]
        out["lam"][sl, :S] = h["lam"]
        group[sl] = g
        j += K
    return out, group


def _slice_h3m(hyp: dict, sel: np.ndarray, S: int) -> dict:
    out = {k: np.asarray(v)[sel] for k, v in hyp.items()}
    out.update(eta=out["eta"][:, :S], epsilon=out["epsilon"][:, :S, :S], m=out["m"][:, :S], W=out["W"][:, :S],
               v=out["v"][:, :S], lam=out["lam"][:, :S])
    return out


def h3m_parameters(hyp: dict, covmode: str = "full") -> dict:
    """convert_h3mrtoh3mb.m:1-79: point estimates of a centre set from its hyper-parameters -- omega[K],
    prior[K,S], trans[K,S,S], mean[K,S,d], cov[K,S,d,d] | [K,S,d] (zero beyond each centre's own state count)."""
    Sr = np.asarray(hyp["Sr"])
    K, S = hyp["eta"].shape
    d = hyp["m"].shape[2]
    full = covmode == "full"
    out = dict(omega=hyp["alpha"] / hyp["alpha"].sum(), prior=np.zeros((K, S)), trans=np.zeros((K, S, S)),
               mean=np.zeros((K, S, d)), cov=np.zeros((K, S, d, d) if full else (K, S, d)))
    for j in range(K):
        s = int(Sr[j])
        out["prior"][j, :s] = hyp["eta"][j, :s] / hyp["eta"][j, :s].sum()
        rs = hyp["epsilon"][j, :s, :s].sum(1, keepdims=True)
        out["trans"][j, :s, :s] = hyp["epsilon"][j, :s, :s] / np.where(rs == 0, 1.0, rs)
        out["mean"][j, :s] = hyp["m"][j, :s]
        for k in range(s):
            v = hyp["v"][j, k]
            den = v - d - 1 if v > d + 1 else v
            if full:
                tC = np.linalg.inv(hyp["W"][j, k]) / den
                out["cov"][j, k] = (tC + tC.T) / 2
            else:
                out["cov"][j, k] = (1.0 / hyp["W"][j, k]) / den
    return out


def vbh3m_remove_empty(out: dict, thresh: float = 1.0) -> tuple[dict, np.ndarray]:
    """[h3mo, emptyinds] = vbh3m_remove_empty(h3m, thresh)  (vbh3m_remove_empty.m:1-60) on the result of
    ``vbhem_h3m_cluster``: drop the groups whose mass ``Nj`` is below ``thresh``; omega is re-normalised, ``Z``,
    ``L_elbo``, ``label``, ``groups``, ``group_size``, the group HMMs' parameters and ``hyp`` follow."""
    Nj = np.asarray(out["Nj"], float)
    keep = np.flatnonzero(~(Nj < thresh))
    gone = np.flatnonzero(Nj < thresh)
    if gone.size == 0:
        return out, gone
    o = dict(out)
    o["omega"] = out["omega"][keep] / out["omega"][keep].sum()
    for k in ("Z", "L_elbo"):
        o[k] = out[k][:, keep]
    for k in ("Nj", "prior", "trans", "mean", "cov", "group_size"):
        o[k] = np.asarray(out[k])[keep]
    o["hyp"] = {k: np.asarray(v)[keep] for k, v in out["hyp"].items()}
    remap = -np.ones(len(Nj), int)
    remap[keep] = np.arange(keep.size)
    o["label"] = remap[out["label"]]
    o["groups"] = [out["groups"][j] for j in keep]
    o["K"] = int(keep.size)
    return o, gone


def vbhem_h3m_cluster(hmms, K, S, vbhemopt: dict | None = None, covmode: str = "full", device: int = 0) -> dict:
    """[group_hmms] = vbhem_h3m_cluster(hmms, K, S, vbhemopt)  (vbhem_h3m_cluster.m:1): cluster HMMs into K groups
    with S states per group HMM; vectors of K and / or S select the model by the lower bound.

    The reference recurses -- for each K, for each S, ``trials`` EM runs from seeded random initialisations one
    after the other (vbhem_h3m_cluster.m:261-355, vbhem_h3m_c.m:28-76).  Here every candidate (K, S, trial) is a
    soft-max group of ONE batched device-resident EM loop over the shared base set (``vbhem_em_run_groups``), each
    run to its own convergence.  Selection is the reference's: best trial by LL; over S the largest
    ``LL + gammaln(S + 1)`` (:327-331); over K the largest ``LL(best S) + gammaln(K + 1)`` (:279-283).

    ``hmms``: list of learned HMMs (see ``hmms_to_h3m_hem``) or an already converted natural base dict.
    Returns the chosen model with the reference's field names where they exist (form_outputH3M.m:1-61):
    ``K, S, LL, hyp`` (hyper-parameters after the last M-step), ``omega, prior, trans, mean, cov``
    (convert_h3mrtoh3mb.m), ``Z [N,K]``, ``L_elbo [N,K]``, ``label`` (0-based), ``groups``, ``group_size``,
    ``model_LL [nK]``, ``model_k``, ``model_bestK``, ``model_LL_S [nK][nS]``, ``model_S``, ``model_bestS``,
    ``trials_LL [nK,nS,trials]``, ``iters``.  ``Z`` / ``L_elbo`` are evaluated under the returned hyper-parameters --
    one E-step later than the reference's hat_Z (vbhem_h3m_c_step_fc.m:446), equal to it within ``minDiff``."""
    from . import synth
    from .api import BaseSet
    opt = dict(CLUSTER_DEFAULTS)
    opt.update(vbhemopt or {})
    if opt.get("seed") is None:
        raise ValueError("vbhemopt.seed has not been set. This is required to obtain reproducible results.")
    if opt["initmode"] not in ("baseem", "wtkmeans", "auto"):
        raise NotImplementedError(f"initmode {opt['initmode']!r}: 'baseem', 'wtkmeans' and 'auto' (both) are mirrored")
    if not isinstance(hmms, dict) and opt["remove_empty"]:                   # vbhem_h3m_cluster.m:116-131
        from .hmm import vbhmm_remove_empty
        hmms = [h if h is None or "N" not in h else vbhmm_remove_empty(h, 1e-3)[0] for h in hmms]
    base = hmms if isinstance(hmms, dict) else hmms_to_h3m_hem(hmms, covmode, bool(opt["use_post"]), float(opt["Nv"]))
    d = base["mu"].shape[2]
    if opt["v0"] <= d - 1:
        raise ValueError("v0 not large enough...should be > D-1")
    if len(np.atleast_1d(opt["m0"])) != d:
        raise ValueError(f"vbhemopt.m0 should have dimension D={d}")
    Ks, Ss = [int(k) for k in np.atleast_1d(K)], [int(s) for s in np.atleast_1d(S)]
    trials = int(opt["trials"])
    cand, models, shapes = [], [], []
    clock = [time.perf_counter()]
    # 'baseem' (vbhemhmm_init.m:58-100) is built on the device for all candidates at once (vbhem_em_begin_baseem);
    # only the random draws are made here, vectorised
    device_init = opt["initmode"] == "baseem"
    for ki, k in enumerate(Ks):
        for si, s in enumerate(Ss):
            for t in range(trials):
                # 'auto' (vbhem_h3m_cluster.m:360-392) tries the modes one after the other and keeps the best bound;
                # here the trials alternate between them, which the best-trial rule then compares
                mode = opt["initmode"] if opt["initmode"] != "auto" else ("baseem", "wtkmeans")[t % 2]
                cand.append((ki, si, t))
                shapes.append((k, s))
                if device_init:      # the same draws as init_centres_baseem(seed + t + 1); gather / inverse on the device
                    models.append(synth.baseem_draws(base, k, s, seed=int(opt["seed"]) + t + 1))
                    continue
                if mode == "baseem":
                    models.append(synth.init_centres_baseem(base, k, s, seed=int(opt["seed"]) + t + 1, opt=opt))
                else:   # wtseed = seed + it (vbhem_h3m_c.m:50-52)
                    models.append(init_centres_wtkmeans(base, k, s, int(opt["seed"]) + t + 1, opt))
    clock.append(time.perf_counter())
    N = base["mu"].shape[0]
    with BaseSet(base, device=device) as bs:
        # batches of at most 4096 H3Ms (vbhem_em_begin's limit) and 2e8 (base, centre) pairs (3.2 GB of L_elbo + z_hat)
        finals, iters_all = [None] * len(shapes), np.zeros(len(shapes), int)
        LL = np.full((len(Ks), len(Ss), trials), -np.inf)
        lo = 0
        while lo < len(shapes):
            hi, ncen = lo, 0
            while hi < len(shapes) and hi - lo < 4096 and (hi == lo or (ncen + shapes[hi][0]) * N <= 2e8):
                ncen += shapes[hi][0]
                hi += 1
            if device_init:
                bs.em_begin_baseem(*synth.stack_draws(models[lo:hi]), opt)
            else:
                hyp, group = stack_h3ms(models[lo:hi])
                bs.em_begin(hyp, opt, group=group)
            run = bs.em_run_groups(int(opt["tau"]), int(opt["max_iter"]), float(opt["minDiff"]))
            first = np.concatenate([[0], np.cumsum([k for k, _ in shapes[lo:hi]])])
            for g in range(hi - lo):
                ki, si, t = cand[lo + g]
                if len(run["LogLs"][g]) and not run["unstable"][g]:
                    LL[ki, si, t] = run["LogLs"][g][-1]
                finals[lo + g] = _slice_h3m(run["hyp"], slice(int(first[g]), int(first[g + 1])), shapes[lo + g][1])
                iters_all[lo + g] = run["iters"][g]
            lo = hi
        clock.append(time.perf_counter())
        best_t = LL.argmax(2)                                                # vbhem_h3m_c.m: best trial
        LL_ks = np.take_along_axis(LL, best_t[:, :, None], 2)[:, :, 0]
        learnt = {}
        if opt["learn_hyps"]:                                                # vbhem_h3m_c.m:105-150
            # the best trial of every (K, S) candidate; all their optimisers advance together (one batched EM loop
            # per round of evaluations)
            keys = [(ki, si) for ki in range(len(Ks)) for si in range(len(Ss))]
            starts = [finals[cand.index((ki, si, int(best_t[ki, si])))] for ki, si in keys]
            for (ki, si), res in zip(keys, vbhem_h3m_c_hyp_batch(starts, bs, opt, covmode)):
                if np.isfinite(res["LL"]) and res["LL"] > LL_ks[ki, si]:
                    learnt[(ki, si)] = res
                    LL_ks[ki, si] = res["LL"]
        LL_S = LL_ks + gammaln(np.array(Ss) + 1.0)[None]                     # vbhem_h3m_cluster.m:327-331
        best_s = LL_S.argmax(1)
        LL_K = LL_ks[np.arange(len(Ks)), best_s] + gammaln(np.array(Ks) + 1.0)   # :279-283 (the run's own LL)
        bk = int(LL_K.argmax())
        bs_i = int(best_s[bk])
        g = cand.index((bk, bs_i, int(best_t[bk, bs_i])))
        chosen = finals[g]
        if (bk, bs_i) in learnt:
            chosen = learnt[(bk, bs_i)]["hyp"]
            opt = dict(learnt[(bk, bs_i)]["learn_hyps"]["vbopt"])
        last = bs.estep_hyp(chosen, int(opt["tau"]))
    out = dict(K=Ks[bk], S=Ss[bs_i], LL=float(LL_ks[bk, bs_i]), hyp=chosen, Z=last["zhat"], L_elbo=last["L_elbo"],
               Nj=chosen["alpha"] - float(opt["alpha0"]), iters=int(iters_all[g]), model_LL=LL_K,
               model_k=np.array(Ks), model_bestK=Ks[bk], model_LL_S=LL_S, model_S=np.array(Ss), model_bestS=Ss[bs_i],
               trials_LL=LL, vbhemopt=opt,
               seconds=dict(init=clock[1] - clock[0], em_loop=clock[2] - clock[1], passes=int(iters_all.max()) + 1))
    if (bk, bs_i) in learnt:
        out["learn_hyps"] = learnt[(bk, bs_i)]["learn_hyps"]
    out.update(h3m_parameters(chosen, covmode))
    out["label"] = out["Z"].argmax(1)
    out["groups"] = [np.flatnonzero(out["label"] == j) for j in range(out["K"])]
    out["group_size"] = np.array([len(x) for x in out["groups"]])
    return out


# ----------------------------------------------------------------------------- hyper-parameter learning
def vbhem_get_hypinfo(learn_hyps, opt: dict) -> list:
    """vbhem_get_hypinfo.m:1-89: for every learnable hyper-parameter its option name, the name of its derivative in
    ``vbhemh3m_lb_deriv`` and the transform between the optimiser's coordinate and the hyper-parameter."""
    d = len(np.atleast_1d(opt["m0"]))
    nW = np.atleast_1d(opt["W0"]).size
    table = {
        "alpha0": ("alpha0", "d_logalpha0", np.exp, np.log, 1),
        "eta0": ("eta0", "d_logeta0", np.exp, np.log, 1),
        "epsilon0": ("epsilon0", "d_logepsilon0", np.exp, np.log, 1),
        "v0": ("v0", "d_logv0D1", lambda x: np.exp(x) + d - 1, lambda x: np.log(x - d + 1), 1),
        "lambda0": ("lambda0", "d_loglambda0", np.exp, np.log, 1),
        "W0": ("W0", "d_sqrtW0inv", lambda x: x ** -2.0, lambda x: 1.0 / np.sqrt(x), nW),
        "W0isqrt": ("W0", "d_sqrtW0inv", lambda x: x ** -2.0, lambda x: 1.0 / np.sqrt(x), nW),
        "W0log": ("W0", "d_logW0", np.exp, np.log, nW),
        "m0": ("m0", "d_m0", lambda x: x, lambda x: x, d),
    }
    out = []
    for name in learn_hyps:
        if name not in table:
            raise ValueError("bad value of learn_hyp")
        o, dn, tr, inv, dims = table[name]
        out.append(dict(optname=o, derivname=dn, hyptrans=tr, hypinvtrans=inv, hypdims=dims))
    return out


def vbhem_h3m_c_hyp(h3m_init: dict, base, vbhemopt: dict, covmode: str = "full", max_evals: int = 100) -> dict:
    """[h3m, L] = vbhem_h3m_c_hyp(h3m_b, vbopt)  (vbhem_h3m_c_hyp.m:1-210): learn the prior hyper-parameters by
    maximising the converged lower bound.  Every evaluation of the objective is one EM run (``vbhem_h3m_c_step_fc``,
    E-steps on the GPU) from the SAME initial centres ``h3m_init`` (initmode 'inith3m', vbhemhmm_init.m:428-455) under
    the candidate hyper-parameters, its value -LL and its gradient the derivatives of the bound at the breaking pass
    (``vbhemh3m_lb_deriv``) in the transformed coordinates of ``vbhem_get_hypinfo``.

    ``learn_hyps`` = 1 selects {alpha0, eta0, epsilon0, v0, lambda0, W0, m0} (:10-12).  The reference minimises with
    ``minimize_new`` (L-BFGS, 100 function evaluations, :38-53), which is not part of its repository's numerics that
    can be pinned here; scipy's L-BFGS-B with the same budget stands in.  Returns the final run under the optimal
    hyper-parameters plus ``learn_hyps = dict(opt_transhyp, opt_L, vbopt, evals)``."""
    from scipy.optimize import minimize
    opt = dict(vbhemopt)
    lh = opt.get("learn_hyps", 1)
    names = ["alpha0", "eta0", "epsilon0", "v0", "lambda0", "W0", "m0"] if not isinstance(lh, (list, tuple)) else list(lh)
    info = vbhem_get_hypinfo(names, opt)

    def set_vbopt(X):
        o, i = dict(opt), 0
        for h in info:
            val = h["hyptrans"](np.asarray(X[i:i + h["hypdims"]], float))
            o[h["optname"]] = float(val[0]) if h["hypdims"] == 1 and h["optname"] not in ("m0", "W0") else val
            i += h["hypdims"]
        return o

    X0 = np.concatenate([np.atleast_1d(h["hypinvtrans"](np.asarray(opt[h["optname"]], float))).ravel() for h in info])
    evals = []

    def objective(X):
        o, clipped = vbhem_clip_hyps(set_vbopt(X))
        run = vbhem_h3m_c_step_fc(h3m_init, base, dict(o, calc_LLderiv=1, hyp_clipped=clipped), covmode)
        g = -np.concatenate([np.atleast_1d(run["dLL"][h["derivname"]]).ravel() for h in info])
        evals.append(run["LL"])
        if not np.isfinite(run["LL"]):
            return 1e300, np.zeros_like(g)
        return -run["LL"], np.where(np.isfinite(g), g, 0.0)

    res = minimize(objective, X0, jac=True, method="L-BFGS-B", options=dict(maxfun=int(max_evals), maxiter=int(max_evals)))
    opt2, _ = vbhem_clip_hyps(set_vbopt(res.x))
    final = vbhem_h3m_c_step_fc(h3m_init, base, opt2, covmode)
    final["learn_hyps"] = dict(opt_transhyp=res.x, opt_L=float(-res.fun), vbopt=opt2, evals=np.array(evals),
                               hypnames=[h["optname"] for h in info])
    return final


def vbhem_h3m_c_hyp_batch(inits: list, bs, vbhemopt: dict, covmode: str = "full", max_evals: int = 100,
                          history: int = 8) -> list:
    """``vbhem_h3m_c_hyp`` for MANY trials at once (vbhem_h3m_c.m:109-138 runs one optimiser per distinct trial, one
    after the other; every evaluation of an optimiser's objective is a whole EM run, vbhem_h3m_c_hyp.m:105-129).

    All trials' optimisers advance in lock-step: in every round each still-running trial proposes one point (its
    prior hyper-parameters in the transformed coordinates of ``vbhem_get_hypinfo``), the EM runs of all proposals are
    the soft-max groups of ONE batched device loop under per-group priors (``vbhem_em_set_group_priors`` +
    ``vbhem_em_run_groups_deriv``), and value and gradient come back per group -- the gradient from the device
    (``k_lb_deriv``), taken at the breaking pass as in vbhem_h3m_c_step_fc.m:357-368.  The optimiser is L-BFGS with
    Armijo backtracking (own code, vectorised over nothing: a few dozen numbers per trial); the reference's is
    ``minimize_new`` with 100 evaluations.  ``inits``: centre hyper-parameter dicts (their (K, S) may differ);
    ``bs``: the resident ``BaseSet``.  Returns, per trial, the best evaluation's run (hyp, LogLs, LL, iters) plus
    ``learn_hyps = dict(opt_transhyp, opt_L, vbopt, evals, hypnames)`` like ``vbhem_h3m_c_hyp``."""
    opt = dict(vbhemopt)
    lh = opt.get("learn_hyps", 1)
    names = ["alpha0", "eta0", "epsilon0", "v0", "lambda0", "W0", "m0"] if not isinstance(lh, (list, tuple)) else list(lh)
    info = vbhem_get_hypinfo(names, opt)
    d = bs.d
    nW0 = np.atleast_1d(opt["W0"]).size
    tau, max_iter, min_diff = int(opt["tau"]), int(opt.get("max_iter", 200)), float(opt.get("minDiff", 1e-5))

    def set_vbopt(X):
        o, i = dict(opt), 0
        for h in info:
            val = h["hyptrans"](np.asarray(X[i:i + h["hypdims"]], float))
            o[h["optname"]] = float(val[0]) if h["hypdims"] == 1 and h["optname"] not in ("m0", "W0") else val
            i += h["hypdims"]
        return o

    X0 = np.concatenate([np.atleast_1d(h["hypinvtrans"](np.asarray(opt[h["optname"]], float))).ravel() for h in info])
    n = len(inits)
    st = [dict(x=X0.copy(), f=None, g=None, S=[], Y=[], p=None, t=1.0, tries=0, evals=[], done=False, best=None,
               xtry=X0.copy()) for _ in range(n)]

    def evaluate(idx, Xs):
        """objective and gradient of trials idx at points Xs: one batched EM loop"""
        opts, clips = zip(*[vbhem_clip_hyps(set_vbopt(X)) for X in Xs])
        hyp, group = stack_h3ms([inits[i] for i in idx])
        bs.em_begin(hyp, opt, group=group)
        bs.em_set_group_priors(list(opts))
        run = bs.em_run_groups_deriv(tau, max_iter, min_diff)
        first = np.concatenate([[0], np.cumsum([inits[i]["eta"].shape[0] for i in idx])])
        out = []
        for g, i in enumerate(idx):
            LL = run["LogLs"][g][-1] if len(run["LogLs"][g]) and not run["unstable"][g] else -np.inf
            dl = transform_derivs(device_derivs_to_dict(run["dLL"][g], d, nW0), opts[g], clips[g])
            grad = -np.concatenate([np.atleast_1d(dl[h["derivname"]]).ravel() for h in info])
            res = dict(hyp=_slice_h3m(run["hyp"], slice(int(first[g]), int(first[g + 1])), inits[i]["eta"].shape[1]),
                       LogLs=run["LogLs"][g], LL=float(LL), iters=int(run["iters"][g]), unstable=bool(run["unstable"][g]),
                       vbopt=opts[g])
            ok = np.isfinite(LL)
            out.append((-LL if ok else 1e300, np.where(np.isfinite(grad), grad, 0.0) if ok else np.zeros_like(grad), res))
        return out

    def direction(s):
        q, al = s["g"].copy(), []
        for sv, yv in zip(reversed(s["S"]), reversed(s["Y"])):
            a = float(sv @ q) / float(yv @ sv)
            al.append(a)
            q -= a * yv
        if s["S"]:
            q *= float(s["S"][-1] @ s["Y"][-1]) / float(s["Y"][-1] @ s["Y"][-1])
        for (sv, yv), a in zip(zip(s["S"], s["Y"]), reversed(al)):
            q += (a - float(yv @ q) / float(yv @ sv)) * sv
        return -q

    active = list(range(n))
    for s, (f, g, res) in zip(st, evaluate(active, [s["x"] for s in st])):
        s.update(f=f, g=g, best=res)
        s["evals"].append(-f)
        s["p"] = -g
        s["t"] = min(1.0, 1.0 / max(np.abs(g).sum(), 1e-300))
        s["done"] = not np.isfinite(res["LL"]) or np.abs(g).max() < 1e-9
    used = 1
    while used < max_evals:
        active = [i for i in range(n) if not st[i]["done"]]
        if not active:
            break
        for i in active:
            st[i]["xtry"] = st[i]["x"] + st[i]["t"] * st[i]["p"]
        ev = evaluate(active, [st[i]["xtry"] for i in active])
        used += 1
        for i, (f, g, res) in zip(active, ev):
            s = st[i]
            s["evals"].append(-f)
            slope = float(s["g"] @ s["p"])
            if f <= s["f"] + 1e-4 * s["t"] * slope:                       # Armijo: accept
                sv, yv = s["xtry"] - s["x"], g - s["g"]
                rel = abs(s["f"] - f) / max(abs(s["f"]), 1e-300)
                if float(yv @ sv) > 1e-12 * float(yv @ yv):
                    s["S"].append(sv)
                    s["Y"].append(yv)
                    s["S"], s["Y"] = s["S"][-history:], s["Y"][-history:]
                s.update(x=s["xtry"].copy(), f=f, g=g, best=res, t=1.0, tries=0)
                s["p"] = direction(s)
                if float(s["g"] @ s["p"]) >= 0:                            # not a descent direction: restart
                    s["S"], s["Y"], s["p"] = [], [], -g
                s["done"] = rel < 1e-9 or np.abs(g).max() < 1e-9
            else:                                                          # backtrack
                s["t"] *= 0.5
                s["tries"] += 1
                s["done"] = s["tries"] >= 25
    out = []
    for s in st:
        res = dict(s["best"])
        res["learn_hyps"] = dict(opt_transhyp=s["x"], opt_L=float(-s["f"]), vbopt=res.pop("vbopt"),
                                 evals=np.array(s["evals"]), hypnames=[h["optname"] for h in info])
        out.append(res)
    return out
