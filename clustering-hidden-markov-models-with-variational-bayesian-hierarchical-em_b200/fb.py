"""Python host mirror of the toolbox's VB-HMM forward-backward entry (``vbhmm_fb``).

``vbhmm_fb(data, hmm_varpar, opt)`` follows /root/reference/src/hmm/vbhmm_fb.m:1-393: the O(K) pre-pass
(``logLambdaTilde``, ``logATilde``, ``logPiTilde``; :62-80) on the host, then the per-sequence emission
densities, scaled forward-backward recursion, responsibilities and summed pair marginals on the GPU through
the C ABI (``vbhmm_fb`` in include/vbhem_b200.h, which replaces the MEX plug-in ``vbhmm_fb_mex`` selected at
vbhmm_fb.m:99-145).  Returned ``fbstats`` has the reference's field names; arrays are in natural numpy
layout -- ``logrho_Saved`` / ``gamma_all`` [N, maxT, K], ``xi_sum`` [N, K(from), K(to)], ``phi_norm`` [N].

The observation sequences do not change over the EM iterations of ``vbhmm_em``, so they are packed into HBM
once (``SequenceSet``) and each call passes only the model.  There is no CPU implementation behind this.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from scipy.special import psi

from ._lib import check, dp, ip, lib


class SequenceSet:
    """N observation sequences resident in HBM (``vbhmm_pack_data``).  data[n]: array [T_n, dim]."""

    def __init__(self, data: list, device: int = 0):
        dim = None
        for x in data:
            if len(x):
                dim = int(np.asarray(x).reshape(len(x), -1).shape[1])
                break
        if dim is None:
            raise ValueError("every sequence is empty")
        T = np.array([len(x) for x in data], np.int32)
        offs = np.zeros(len(data) + 1, np.int64)
        offs[1:] = np.cumsum(T.astype(np.int64) * dim)
        flat = np.zeros(max(1, int(offs[-1])))
        for n, x in enumerate(data):
            if len(x):
                flat[offs[n]:offs[n + 1]] = np.asarray(x, float).reshape(len(x), dim).T.ravel()   # [T x dim] col-major
        self.N, self.dim, self.maxT, self.T = len(data), dim, int(T.max()), T
        self._h = C.c_void_p()
        check(lib().vbhmm_pack_data(device, self.N, dim, T.ctypes.data_as(ip),
                                    offs.ctypes.data_as(C.POINTER(C.c_longlong)), flat.ctypes.data_as(dp),
                                    C.byref(self._h)))

    @classmethod
    def from_packed(cls, T: np.ndarray, offs: np.ndarray, flat: np.ndarray, dim: int, device: int = 0):
        """Already concatenated sequences (each [T_n x dim] column-major at flat[offs[n]])."""
        self = cls.__new__(cls)
        T = np.ascontiguousarray(T, np.int32)
        offs = np.ascontiguousarray(offs, np.int64)
        self.N, self.dim, self.maxT, self.T = len(T), dim, int(T.max()), T
        self._h = C.c_void_p()
        check(lib().vbhmm_pack_data(device, self.N, dim, T.ctypes.data_as(ip),
                                    offs.ctypes.data_as(C.POINTER(C.c_longlong)), flat.ctypes.data_as(dp),
                                    C.byref(self._h)))
        return self

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().vbhmm_free_data(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def prepass(hmm_varpar: dict) -> dict:
    """vbhmm_fb.m:28-80.  hmm_varpar (natural): v[K], W[K,dim,dim], epsilon[K(from),K(to)], alpha[K], m[K,dim], beta[K]."""
    v, W, eps, al = (np.asarray(hmm_varpar[k], float) for k in ("v", "W", "epsilon", "alpha"))
    m = np.asarray(hmm_varpar["m"], float)
    K, dim = m.shape
    ii = np.arange(1, dim + 1)
    lLT = np.array([psi(0.5 * (v[k] + 1) - 0.5 * ii).sum() + dim * np.log(2) + np.log(np.linalg.det(W[k]))
                    for k in range(K)])
    logA = psi(eps) - psi(eps.sum(1, keepdims=True))
    logPi = psi(al) - psi(al.sum())
    return dict(m=m, W=W, v=v, beta=np.asarray(hmm_varpar["beta"], float), logLambdaTilde=lLT,
                const_denominator=dim * np.log(2 * np.pi) / 2, pz1=np.exp(logPi), A=np.exp(logA),
                logPiTilde=logPi, logATilde=logA)


def fb_call(seqs: SequenceSet, par: dict, want=("logrho_Saved", "gamma_all", "xi_sum", "phi_norm")) -> dict:
    """One pass with an explicit model (the MEX argument list).  ``want`` selects what is copied back."""
    c = np.ascontiguousarray
    K = len(par["v"])
    N, maxT = seqs.N, seqs.maxT
    m = c(par["m"], dtype=np.float64)                                  # [K, dim]  == column-major [dim, K]
    W = c(np.transpose(par["W"], (0, 2, 1)), dtype=np.float64)         # [K, j, i] == column-major [i, j, K]
    A = c(np.asarray(par["A"], float).T)                               # [to, from] == column-major [from, to]
    bufs = dict(logrho_Saved=np.zeros((maxT, N, K)) if "logrho_Saved" in want else None,
                gamma_all=np.zeros((maxT, N, K)) if "gamma_all" in want else None,
                xi_sum=np.zeros((N, K, K)) if "xi_sum" in want else None,
                phi_norm=np.zeros(N) if "phi_norm" in want else None)
    P = lambda a: None if a is None else a.ctypes.data_as(dp)
    check(lib().vbhmm_fb(seqs._h, K, P(m), P(W), P(c(par["v"], dtype=np.float64)),
                         P(c(par["beta"], dtype=np.float64)), P(c(par["logLambdaTilde"], dtype=np.float64)),
                         C.c_double(float(par["const_denominator"])), P(c(par["pz1"], dtype=np.float64)), P(A),
                         P(bufs["logrho_Saved"]), P(bufs["gamma_all"]), P(bufs["xi_sum"]), P(bufs["phi_norm"])))
    out = {}
    if bufs["logrho_Saved"] is not None:
        out["logrho_Saved"] = np.transpose(bufs["logrho_Saved"], (1, 0, 2))
    if bufs["gamma_all"] is not None:
        out["gamma_all"] = np.transpose(bufs["gamma_all"], (1, 0, 2))
    if bufs["xi_sum"] is not None:
        out["xi_sum"] = np.transpose(bufs["xi_sum"], (0, 2, 1))
    if bufs["phi_norm"] is not None:
        out["phi_norm"] = bufs["phi_norm"]
    return out


def fb_stats(seqs: SequenceSet, hmm_varpar: dict) -> dict:
    """One forward-backward pass fused with the sums ``vbhmm_em`` takes from it (vbhmm_em.m:158-246,
    vbhmm_em_lb.m:214,217) -- ``vbhmm_fb_stats``: nothing of size N leaves the device.
    Returns Nk1[K], Nk[K], M[K(from),K(to)], xbar[K,dim], S[K,dim,dim], Lt63, Lt64 and the pre-pass results."""
    c = np.ascontiguousarray
    par = prepass(hmm_varpar)
    K, dim = par["m"].shape
    m = c(par["m"], dtype=np.float64)
    W = c(np.transpose(par["W"], (0, 2, 1)), dtype=np.float64)
    A = c(np.asarray(par["A"], float).T)
    Nk1, Nk, M, xb, S, lb = np.zeros(K), np.zeros(K), np.zeros((K, K)), np.zeros((dim, K)), np.zeros((K, dim, dim)), np.zeros(2)
    P = lambda a: a.ctypes.data_as(dp)
    check(lib().vbhmm_fb_stats(seqs._h, K, P(m), P(W), P(c(par["v"], dtype=np.float64)),
                               P(c(par["beta"], dtype=np.float64)), P(c(par["logLambdaTilde"], dtype=np.float64)),
                               C.c_double(float(par["const_denominator"])), P(c(par["pz1"], dtype=np.float64)), P(A),
                               P(Nk1), P(Nk), P(M), P(xb), P(S), P(lb)))
    return dict(Nk1=Nk1, Nk=Nk, M=M.T.copy(), xbar=xb.T.copy(), S=np.transpose(S, (0, 2, 1)).copy(), Lt63=lb[0],
                Lt64=lb[1], logLambdaTilde=par["logLambdaTilde"], logPiTilde=par["logPiTilde"],
                logATilde=par["logATilde"])


def vbhmm_fb(data, hmm_varpar: dict, opt: dict | None = None) -> dict:
    """[fbstats] = vbhmm_fb(data, hmm_varpar, opt)  (vbhmm_fb.m:1).

    ``data``: a list of [T_n, dim] arrays or a ``SequenceSet``.  Options honoured: ``usegroups`` must be 0 / absent
    (call once per group, as vbhmm_fb.m:150-177 does); ``savexi`` is not supported, as in the reference's MEX path
    (vbhmm_fb.m:104-107)."""
    opt = opt or {}
    if opt.get("usegroups", 0):
        raise NotImplementedError("usegroups: call vbhmm_fb once per group (vbhmm_fb.m:150-177)")
    if opt.get("savexi", 0):
        raise NotImplementedError("vbhmm_fb: savexi not supported in the GPU implementation (vbhmm_fb.m:104-107)")
    own = not isinstance(data, SequenceSet)
    seqs = SequenceSet(data) if own else data
    try:
        par = prepass(hmm_varpar)
        out = fb_call(seqs, par)
    finally:
        if own:
            seqs.close()
    out.update(logLambdaTilde=par["logLambdaTilde"], logPiTilde=par["logPiTilde"], logATilde=par["logATilde"])
    return out
