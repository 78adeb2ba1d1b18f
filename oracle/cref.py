"""oracle/cref.py -- TEST INFRASTRUCTURE, not product code.

ctypes front-ends for the two native oracles:

* ``ref``  -- oracle/_ref/libvbhem_ref.so: the reference's own C E-step
  (/root/reference/src/vbhem/vbhem_hmm_bwd_fwd_mex.c) compiled unmodified and in
  place behind oracle/mxshim (recipe: oracle/Makefile).  ``kind = "reference"``.
* ``port`` -- oracle/_build/libvbhem_oracle.so: this repo's plain-C restatement
  (oracle/estep_oracle.c).  ``kind = "port"``.

Both take the same packed column-major arrays (see estep_oracle.c header).  All
array arguments here are numpy arrays in *numpy C order with the packed axes
reversed*, i.e. ``prior[N, maxSb]``, ``A[N, to, from]`` ... so that their memory
is exactly the column-major layout the C side expects.  Use ``pack.py``-style
helpers in the package (or tests/util.py) to build them.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libvbhem_ref.so")
PORT_SO = os.path.join(HERE, "_build", "libvbhem_oracle.so")
FB_REF_SO = os.path.join(HERE, "_ref", "libvbhmm_fb_ref.so")        # reference's src/hmm/vbhmm_fb_mex.c in place
FB_PORT_SO = os.path.join(HERE, "_build", "libvbhmm_fb_oracle.so")  # oracle/fb_oracle.c
FB_GATEWAY_SO = os.path.join(HERE, "_build", "libvbhmm_fb_gateway.so")  # this repo's mex/vbhmm_fb_mex.cpp behind the shim
GATEWAY_SO = os.path.join(HERE, "_build", "libvbhem_gateway.so")   # this repo's MEX gateway behind the same shim
FUSED_SO = os.path.join(HERE, "_build", "libvbhem_gateway_fused.so")   # ... and its fused-E-step gateway
EM_SO = os.path.join(HERE, "_build", "libvbhem_gateway_em.so")         # ... and its whole-EM-run gateway
EM_BATCH_SO = os.path.join(HERE, "_build", "libvbhmm_em_batch_gateway.so")   # batched VB-HMM learning gateway

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(verbose: bool = False) -> None:
    """Build the port always, and the in-place reference build when /root/reference exists."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def have_ref() -> bool:
    return os.path.exists(REF_SO)


_libs = {}


def _lib(which: str):
    if which not in _libs:
        path = {"ref": REF_SO, "gateway": GATEWAY_SO, "fused": FUSED_SO, "em": EM_SO, "em_batch": EM_BATCH_SO, "fb_ref": FB_REF_SO,
                "fb_port": FB_PORT_SO, "fb_gateway": FB_GATEWAY_SO}.get(which, PORT_SO)
        if not os.path.exists(path):
            build()
        _libs[which] = C.CDLL(path)
    return _libs[which]


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def alloc_pair_outputs(N, K, maxSr, d, covmode, fill=0.0):
    dd = d * d if covmode else d
    P = N * K
    mk = lambda *s: np.full(s, fill, dtype=np.float64)
    return dict(L_elbo=mk(K, N), nu1=mk(K, N, maxSr), xi=mk(K, N, maxSr, maxSr), pr=mk(K, N, maxSr),
                mu=mk(K, N, d, maxSr), Mu=mk(K, N, dd, maxSr))


def ref_estep_pairs(base: dict, cen: dict, T: int, which: str = "ref") -> dict:
    """The MEX gateway signature through the mx shim: the reference's C (`ref`) or this repo's CUDA-backed
    gateway (`gateway`).  cen needs W and v (not vW)."""
    lib = _lib(which)
    N, maxSb = base["prior"].shape
    K, maxSr = cen["logPi"].shape
    d, covmode = base["d"], base["covmode"]
    o = alloc_pair_outputs(N, K, maxSr, d, covmode)
    fn = lib.shim_estep_pairs
    fn.restype = C.c_int
    rc = fn(C.c_int(N), _i(_i32(base["Sb"])), C.c_int(maxSb), C.c_int(d), C.c_int(covmode),
            _d(_f64(base["prior"])), _d(_f64(base["A"])), _d(_f64(base["mu"])), _d(_f64(base["cov"])),
            C.c_int(K), _i(_i32(cen["Sr"])), C.c_int(maxSr),
            _d(_f64(cen["logPi"])), _d(_f64(cen["logA"])), _d(_f64(cen["m"])), _d(_f64(cen["W"])),
            _d(_f64(cen["v"])), _d(_f64(cen["cterm"])), C.c_int(T),
            _d(o["L_elbo"]), _d(o["nu1"]), _d(o["xi"]), _d(o["pr"]), _d(o["mu"]), _d(o["Mu"]))
    if rc != 0:
        lib.shim_last_error_id.restype = C.c_char_p
        lib.shim_last_error_msg.restype = C.c_char_p
        raise RuntimeError(f"{lib.shim_last_error_id().decode()}: {lib.shim_last_error_msg().decode()}")
    return o


def ref_call_malformed(mode: int, which: str = "ref"):
    lib = _lib(which)
    lib.shim_last_error_id.restype = C.c_char_p
    rc = lib.shim_call_malformed(C.c_int(mode))
    return rc, lib.shim_last_error_id().decode()


def gateway_estep_fused(base: dict, cen: dict, T: int) -> dict:
    """mex/vbhem_estep_fused_mex.cpp through the mx shim (nine plain-array outputs, already normalised)."""
    lib = _lib("fused")
    N, maxSb = base["prior"].shape
    K, r = cen["logPi"].shape
    d, covmode = base["d"], base["covmode"]
    dd = d * d if covmode else d
    o = dict(L_elbo=np.zeros((K, N)), zhat=np.zeros((K, N)), Nj=np.zeros(K), Nj_rho1=np.zeros((K, r)),
             Nj_rho2rho=np.zeros((K, r, r)), Nj_rho=np.zeros((K, r)), y_bar=np.zeros((K, d, r)),
             S_plus_C=np.zeros((K, dd, r)), lb=np.zeros(2))
    fn = lib.shim_estep_fused
    fn.restype = C.c_int
    rc = fn(C.c_int(N), _i(_i32(base["Sb"])), C.c_int(maxSb), C.c_int(d), C.c_int(covmode),
            _d(_f64(base["prior"])), _d(_f64(base["A"])), _d(_f64(base["mu"])), _d(_f64(base["cov"])),
            _d(_f64(base["Ntilde"])), C.c_int(K), _i(_i32(cen["Sr"])), C.c_int(r),
            _d(_f64(cen["logPi"])), _d(_f64(cen["logA"])), _d(_f64(cen["m"])), _d(_f64(cen["W"])),
            _d(_f64(cen["v"])), _d(_f64(cen["cterm"])), _d(_f64(cen["logOmega"])), C.c_int(T),
            *[_d(o[k]) for k in ("L_elbo", "zhat", "Nj", "Nj_rho1", "Nj_rho2rho", "Nj_rho", "y_bar", "S_plus_C", "lb")])
    if rc != 0:
        lib.shim_last_error_id.restype = C.c_char_p
        lib.shim_last_error_msg.restype = C.c_char_p
        raise RuntimeError(f"{lib.shim_last_error_id().decode()}: {lib.shim_last_error_msg().decode()}")
    return o


def gateway_em_run(base: dict, hyp: dict, opt: dict) -> dict:
    """mex/vbhem_em_run_mex.cpp through the mx shim.  base: packed base dict; hyp: natural centre hyper-parameters
    (alpha[K], eta[K,S], epsilon[K,S,S] (from,to), m[K,S,d], W[K,S,d,d] | [K,S,d], v, lam, Sr); opt: alpha0, eta0,
    epsilon0, lambda0, v0, m0, W0, tau, max_iter, minDiff.  Returns the hyper-parameters after the run (natural),
    LogLs, iters, unstable."""
    lib = _lib("em")
    N, maxSb = base["prior"].shape
    K, r = hyp["eta"].shape
    d, covmode = base["d"], base["covmode"]
    al, et, v, lam = (_f64(np.array(hyp[k], float)) for k in ("alpha", "eta", "v", "lam"))
    eps = _f64(np.transpose(hyp["epsilon"], (0, 2, 1)))                     # [K, to, from] == column-major [from, to, K]
    m = _f64(np.array(hyp["m"], float))                                      # [K, S, d]     == [d, S, K]
    W = _f64(np.transpose(hyp["W"], (0, 1, 3, 2))) if covmode else _f64(np.array(hyp["W"], float))
    hyp0 = _f64(np.array([opt["alpha0"], opt["eta0"], opt["epsilon0"], opt["lambda0"], opt["v0"]], float))
    m0 = _f64(np.asarray(opt["m0"], float).reshape(d))
    W0 = _f64(np.atleast_1d(np.asarray(opt["W0"], float)))
    max_iter = int(opt.get("max_iter", 200))
    LogLs, info = np.zeros(max_iter + 2), np.zeros(3)
    fn = lib.shim_em_run
    fn.restype = C.c_int
    rc = fn(C.c_int(N), _i(_i32(base["Sb"])), C.c_int(maxSb), C.c_int(d), C.c_int(covmode),
            _d(_f64(base["prior"])), _d(_f64(base["A"])), _d(_f64(base["mu"])), _d(_f64(base["cov"])),
            _d(_f64(base["Ntilde"])), C.c_int(K), _i(_i32(hyp["Sr"])), C.c_int(r), _d(al), _d(et), _d(eps), _d(m), _d(W),
            _d(v), _d(lam), _d(hyp0), _d(m0), _d(W0), C.c_int(W0.size), C.c_int(int(opt["tau"])), C.c_int(max_iter),
            C.c_double(float(opt.get("minDiff", 1e-5))), _d(LogLs), _d(info))
    if rc != 0:
        lib.shim_last_error_id.restype = C.c_char_p
        lib.shim_last_error_msg.restype = C.c_char_p
        raise RuntimeError(f"{lib.shim_last_error_id().decode()}: {lib.shim_last_error_msg().decode()}")
    out = dict(alpha=al, eta=et, epsilon=np.transpose(eps, (0, 2, 1)).copy(), m=m,
               W=np.transpose(W, (0, 1, 3, 2)).copy() if covmode else W, v=v, lam=lam, Sr=np.array(hyp["Sr"]))
    return dict(hyp=out, LogLs=LogLs[:int(info[2])].copy(), iters=int(info[0]), unstable=bool(info[1]))


def port_estep_pairs(base: dict, cen: dict, T: int, lonly: bool = False) -> dict:
    lib = _lib("port")
    N, maxSb = base["prior"].shape
    K, maxSr = cen["logPi"].shape
    d, covmode = base["d"], base["covmode"]
    o = alloc_pair_outputs(N, K, maxSr, d, covmode) if not lonly else dict(L_elbo=np.zeros((K, N)))
    g = (lambda k: None) if lonly else (lambda k: _d(o[k]))
    lib.oracle_estep_pairs(
        C.c_int(N), _i(_i32(base["Sb"])), C.c_int(maxSb), C.c_int(d), C.c_int(covmode),
        _d(_f64(base["prior"])), _d(_f64(base["A"])), _d(_f64(base["mu"])), _d(_f64(base["cov"])),
        C.c_int(K), _i(_i32(cen["Sr"])), C.c_int(maxSr),
        _d(_f64(cen["logPi"])), _d(_f64(cen["logA"])), _d(_f64(cen["m"])), _d(_f64(cen["vW"])),
        _d(_f64(cen["cterm"])), C.c_int(T),
        _d(o["L_elbo"]), g("nu1"), g("xi"), g("pr"), g("mu"), g("Mu"))
    return o


def port_zhat(L_elbo, logOmega, Ntilde, group=None):
    lib = _lib("port")
    K, N = L_elbo.shape
    group = np.zeros(K, np.int32) if group is None else _i32(group)
    ng = int(group.max()) + 1
    zhat, Z, Nj, lb = np.zeros((K, N)), np.zeros((K, N)), np.zeros(K), np.zeros(2 * ng)
    lib.oracle_zhat(C.c_int(N), C.c_int(K), _d(_f64(L_elbo)), _d(_f64(logOmega)), _d(_f64(Ntilde)),
                    _i(group), C.c_int(ng), _d(zhat), _d(Z), _d(Nj), _d(lb))
    return dict(zhat=zhat, Z=Z, Nj=Nj, lb=lb)


def port_reduce(Z, o, maxSr, d, covmode):
    lib = _lib("port")
    K, N = Z.shape
    dd = d * d if covmode else d
    S = dict(nu1=np.zeros((K, maxSr)), xi=np.zeros((K, maxSr, maxSr)), pr=np.zeros((K, maxSr)),
             mu=np.zeros((K, d, maxSr)), Mu=np.zeros((K, dd, maxSr)))
    lib.oracle_reduce(C.c_int(N), C.c_int(K), C.c_int(maxSr), C.c_int(d), C.c_int(covmode), _d(_f64(Z)),
                      _d(o["nu1"]), _d(o["xi"]), _d(o["pr"]), _d(o["mu"]), _d(o["Mu"]),
                      _d(S["nu1"]), _d(S["xi"]), _d(S["pr"]), _d(S["mu"]), _d(S["Mu"]))
    return S


def port_estep_fused(base: dict, cen: dict, T: int, want_pairs: bool = True) -> dict:
    lib = _lib("port")
    N, maxSb = base["prior"].shape
    K, maxSr = cen["logPi"].shape
    d, covmode = base["d"], base["covmode"]
    dd = d * d if covmode else d
    L = np.zeros((K, N)) if want_pairs else None
    z = np.zeros((K, N)) if want_pairs else None
    S = dict(Nj=np.zeros(K), lb=np.zeros(2), nu1=np.zeros((K, maxSr)), xi=np.zeros((K, maxSr, maxSr)),
             pr=np.zeros((K, maxSr)), mu=np.zeros((K, d, maxSr)), Mu=np.zeros((K, dd, maxSr)))
    lib.oracle_estep_fused(
        C.c_int(N), _i(_i32(base["Sb"])), C.c_int(maxSb), C.c_int(d), C.c_int(covmode),
        _d(_f64(base["prior"])), _d(_f64(base["A"])), _d(_f64(base["mu"])), _d(_f64(base["cov"])),
        _d(_f64(base["Ntilde"])),
        C.c_int(K), _i(_i32(cen["Sr"])), C.c_int(maxSr),
        _d(_f64(cen["logPi"])), _d(_f64(cen["logA"])), _d(_f64(cen["m"])), _d(_f64(cen["vW"])),
        _d(_f64(cen["cterm"])), _d(_f64(cen["logOmega"])), C.c_int(T),
        _d(L), _d(z), _d(S["Nj"]), _d(S["lb"]),
        _d(S["nu1"]), _d(S["xi"]), _d(S["pr"]), _d(S["mu"]), _d(S["Mu"]))
    S["L_elbo"], S["zhat"] = L, z
    return S


# ----------------------------------------------------------------------------- VB-HMM forward-backward
def have_fb_ref() -> bool:
    return os.path.exists(FB_REF_SO)


def pack_sequences(data: list) -> dict:
    """data[n]: array [T_n, dim] (time, coordinate) -> flat buffer with sequence n stored [T_n x dim] column-major
    (exactly mxGetPr(data{n}) of the reference), offsets in doubles."""
    dim = int(np.asarray(data[0]).reshape(len(data[0]), -1).shape[1]) if len(data) else 1
    T = np.array([len(x) for x in data], np.int32)
    offs = np.zeros(len(data) + 1, np.int64)
    offs[1:] = np.cumsum(T.astype(np.int64) * dim)
    flat = np.zeros(max(1, int(offs[-1])))
    for n, x in enumerate(data):
        x = np.asarray(x, float).reshape(len(x), dim)
        flat[offs[n]:offs[n + 1]] = x.T.ravel()          # column-major [T x dim]
    return dict(N=len(data), dim=dim, T=T, offs=offs, flat=flat, maxT=int(T.max()) if len(T) else 0)


def _fb(which: str, ps: dict, par: dict) -> dict:
    """par: m[K,dim], W[K,dim,dim], v[K], beta[K], logLambdaTilde[K], const_denominator, pz1[K], A[K(from),K(to)]."""
    L = _lib(which)
    fn = L.fb_oracle if which == "fb_port" else L.shim_vbhmm_fb
    N, dim, maxT = ps["N"], ps["dim"], ps["maxT"]
    K = len(par["v"])
    c = np.ascontiguousarray
    m = c(par["m"], dtype=np.float64)                                     # [K, dim]  == column-major [dim, K]
    W = c(np.transpose(par["W"], (0, 2, 1)), dtype=np.float64)            # [K, j, i] == column-major [i, j, K]
    A = c(np.asarray(par["A"], float).T)                                  # [to, from] == column-major [from, to]
    out = dict(logrho=np.zeros((maxT, N, K)), gamma=np.zeros((maxT, N, K)), xi=np.zeros((N, K, K)), phi=np.zeros(N))
    fn.restype = C.c_int
    rc = fn(N, dim, K, maxT, ps["T"].ctypes.data_as(_ip), ps["offs"].ctypes.data_as(C.POINTER(C.c_longlong)),
            _d(ps["flat"]), _d(m), _d(W), _d(c(par["v"], dtype=np.float64)), _d(c(par["beta"], dtype=np.float64)),
            _d(c(par["logLambdaTilde"], dtype=np.float64)), C.c_double(par["const_denominator"]),
            _d(c(par["pz1"], dtype=np.float64)), _d(A), _d(out["logrho"]), _d(out["gamma"]), _d(out["xi"]),
            _d(out["phi"]))
    if rc != 0:
        raise RuntimeError(f"{which} vbhmm_fb failed ({rc})")
    # natural views: logrho / gamma [N, maxT, K]; xi [N, from, to]
    return dict(logrho_Saved=np.transpose(out["logrho"], (1, 0, 2)).copy(),
                gamma_all=np.transpose(out["gamma"], (1, 0, 2)).copy(),
                xi_sum=np.transpose(out["xi"], (0, 2, 1)).copy(), phi_norm=out["phi"])


def ref_vbhmm_fb(ps: dict, par: dict) -> dict:
    return _fb("fb_ref", ps, par)


def port_vbhmm_fb(ps: dict, par: dict) -> dict:
    return _fb("fb_port", ps, par)


def gateway_em_batch(subjects: list, run_subject, K: int, ini: dict, inits: list) -> list:
    """mex/vbhmm_em_batch_mex.cpp through the mx shim.  subjects[s]: list of [T, dim] sequences; run_subject[r]: 0-based
    subject of run r; inits[r]: natural variational parameters (alpha[K], epsilon[K,K] (from,to), beta, v, m[K,dim],
    W[K,dim,dim]); ini: alpha0, epsilon0, beta0, v0, mu0, W0, maxIter, minDiff.  Returns per run a dict with the
    learnt varpar (natural), N1, N, M (from,to), LLs, iter, unstable."""
    lib = _lib("em_batch")
    flat_seqs = [x for sub in subjects for x in sub]
    ps = pack_sequences(flat_seqs)
    start = np.concatenate([[0], np.cumsum([len(sub) for sub in subjects])]).astype(np.int32)
    R, dim = len(inits), ps["dim"]
    st = lambda k: np.stack([np.asarray(v[k], float) for v in inits])
    al, be, vv = _f64(st("alpha")), _f64(st("beta")), _f64(st("v"))
    ep = _f64(np.transpose(st("epsilon"), (0, 2, 1)))
    mm = _f64(st("m"))
    WW = _f64(np.transpose(st("W"), (0, 1, 3, 2)))
    prior = _f64(np.array([ini["alpha0"], ini["epsilon0"], ini["beta0"], ini["v0"]], float))
    m0 = _f64(np.asarray(ini["mu0"], float).reshape(dim))
    W0 = _f64(np.atleast_1d(np.asarray(ini["W0"], float)))
    max_iter = int(ini["maxIter"])
    N1, Nk, M = np.zeros((R, K)), np.zeros((R, K)), np.zeros((R, K, K))
    LLs, info = np.zeros((R, max_iter)), np.zeros((R, 2))
    rs = _f64(np.asarray(run_subject, float) + 1.0)
    fn = lib.shim_em_batch
    fn.restype = C.c_int
    rc = fn(C.c_int(len(subjects)), _i(start), C.c_int(dim), _i(_i32(ps["T"])), ps["offs"].ctypes.data_as(C.POINTER(C.c_longlong)),
            _d(ps["flat"]), C.c_int(R), _d(rs), C.c_int(K), _d(prior), _d(m0), _d(W0), C.c_int(W0.size), _d(al), _d(ep),
            _d(be), _d(vv), _d(mm), _d(WW), C.c_int(max_iter), C.c_double(float(ini["minDiff"])), _d(N1), _d(Nk), _d(M),
            _d(LLs), _d(info))
    if rc != 0:
        lib.shim_last_error_id.restype = C.c_char_p
        lib.shim_last_error_msg.restype = C.c_char_p
        raise RuntimeError(f"{lib.shim_last_error_id().decode()}: {lib.shim_last_error_msg().decode()}")
    out = []
    for r in range(R):
        n = int(info[r, 0])
        out.append(dict(varpar=dict(alpha=al[r], epsilon=ep[r].T.copy(), beta=be[r], v=vv[r], m=mm[r],
                                    W=np.transpose(WW[r], (0, 2, 1)).copy()),
                        N1=N1[r], N=Nk[r], M=M[r].T.copy(), LLs=LLs[r, :n].copy(), iter=n, unstable=bool(info[r, 1])))
    return out


def gateway_vbhmm_fb(ps: dict, par: dict) -> dict:
    """mex/vbhmm_fb_mex.cpp (CUDA library behind it) through the same mx shim and driver as the reference's MEX."""
    return _fb("fb_gateway", ps, par)
