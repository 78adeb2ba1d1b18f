"""Python host mirror of the E-step boundary (thin wrapper over the C ABI).

``BaseSet`` is the resident base mixture (reference type ``h3m_b``); its methods
mirror the two boundaries of include/vbhem_b200.h:

* ``estep_pairs``  -- the six outputs of the reference's MEX plug-in
  ``vbhem_hmm_bwd_fwd_mex`` (/root/reference/src/vbhem/vbhem_hmm_bwd_fwd_mex.c:288-289,
  call sites /root/reference/src/vbhem/vbhem_h3m_c_step_fc.m:175-176,193-194);
* ``estep`` / ``estep_hyp`` -- the fused E-step: MEX call + responsibilities
  (vbhem_h3m_c_step_fc.m:271-283) + statistics reduction
  (/root/reference/src/vbhem/vbhem_compute_Statistics.m:33-82).

Arrays are *natural* numpy arrays (see pack.py); all compute happens in the CUDA
library -- there is no numpy implementation behind these calls.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import pack
from ._lib import check, dp, ip, lib


def _d(a):
    return None if a is None else a.ctypes.data_as(dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(ip)


class BaseSet:
    """A base HMM set packed once into HBM (``vbhem_pack_base``)."""

    def __init__(self, base: dict, device: int = 0, packed: bool = False):
        pb = base if packed else pack.pack_base(base)
        self.N, self.maxSb, self.d, self.covmode = pb["N"], pb["maxSb"], pb["d"], pb["covmode"]
        self.dd = self.d * self.d if self.covmode else self.d
        self._h = C.c_void_p()
        self.device = device
        check(lib().vbhem_pack_base(device, self.N, _i(pb["Sb"]), self.maxSb, self.d, self.covmode,
                                    _d(pb["prior"]), _d(pb["A"]), _d(pb["mu"]), _d(pb["cov"]), _d(pb["Ntilde"]),
                                    C.byref(self._h)))

    def repack(self, base: dict, packed: bool = False):
        """Re-upload a base set of the same shape into this handle (``vbhem_repack_base``)."""
        pb = base if packed else pack.pack_base(base)
        assert (pb["N"], pb["maxSb"], pb["d"], pb["covmode"]) == (self.N, self.maxSb, self.d, self.covmode)
        check(lib().vbhem_repack_base(self._h, _i(pb["Sb"]), _d(pb["prior"]), _d(pb["A"]), _d(pb["mu"]),
                                      _d(pb["cov"]), _d(pb["Ntilde"])))

    # -- lifetime -----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().vbhem_free_base(self._h)
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

    # -- knobs --------------------------------------------------------------
    def set_kernel(self, which: str = "auto"):
        check(lib().vbhem_set_kernel(self._h, {"auto": 0, "general": 1, "unpaired": 2, "halves": 3}[which]))

    def set_stream(self, cuda_stream: int | None):
        """Run this handle's work on an existing CUDA stream (``torch.cuda.current_stream().cuda_stream``).
        ``None`` = the handle's own stream.  ``0`` is torch's default stream, i.e. the legacy default stream,
        which the C ABI takes as ``cudaStreamLegacy`` (NULL there means "own stream")."""
        if cuda_stream is None:
            handle = 0
        else:
            handle = cuda_stream if cuda_stream else 1          # (cudaStream_t)0x1 == cudaStreamLegacy
        check(lib().vbhem_set_stream(self._h, C.c_void_p(handle)))

    def synchronize(self):
        check(lib().vbhem_synchronize(self._h))

    def counters(self) -> dict:
        out = (C.c_longlong * 8)()
        check(lib().vbhem_last_counters(self._h, out))
        return dict(launches=out[0], lattice_pairs=out[1], general_pairs=out[2], active_pairs=out[3], tiles=out[4],
                    split_tiles=out[5], compactions=out[6])

    # -- Boundary-1 ---------------------------------------------------------
    def estep_pairs(self, cen: dict, tau: int, L_only: bool = False, packed: bool = False) -> dict:
        """Literal per-pair outputs (natural layout, see pack.pairs_to_natural)."""
        pc = cen if packed else pack.pack_centres(cen)
        K, r, N, d, dd = pc["K"], pc["maxSr"], self.N, self.d, self.dd
        L = np.zeros((K, N))
        if L_only:
            check(lib().vbhem_estep_pairs(self._h, K, _i(pc["Sr"]), r, tau, _d(pc["logPi"]), _d(pc["logA"]),
                                          _d(pc["m"]), _d(pc["vW"]), _d(pc["cterm"]), _d(L),
                                          None, None, None, None, None))
            return dict(L_elbo=L.T.copy())
        o = dict(L_elbo=L, nu1=np.zeros((K, N, r)), xi=np.zeros((K, N, r, r)), pr=np.zeros((K, N, r)),
                 mu=np.zeros((K, N, d, r)), Mu=np.zeros((K, N, dd, r)))
        check(lib().vbhem_estep_pairs(self._h, K, _i(pc["Sr"]), r, tau, _d(pc["logPi"]), _d(pc["logA"]),
                                      _d(pc["m"]), _d(pc["vW"]), _d(pc["cterm"]), _d(o["L_elbo"]), _d(o["nu1"]),
                                      _d(o["xi"]), _d(o["pr"]), _d(o["mu"]), _d(o["Mu"])))
        return pack.pairs_to_natural(o, d, self.covmode)

    # -- Boundary-2 ---------------------------------------------------------
    def _split_stats(self, buf, K, r, ng):
        off = (C.c_size_t * 7)()
        lib().vbhem_stats_offsets(K, r, self.d, self.covmode, ng, off)
        d, dd = self.d, self.dd
        S = dict(Nj=buf[off[0]:off[1]], nu1=buf[off[1]:off[2]].reshape(K, r),
                 xi=buf[off[2]:off[3]].reshape(K, r, r), pr=buf[off[3]:off[4]].reshape(K, r),
                 mu=buf[off[4]:off[5]].reshape(K, d, r), Mu=buf[off[5]:off[6]].reshape(K, dd, r),
                 lb=buf[off[6]:off[6] + 2 * ng])
        return S

    def stats_size(self, K, r, ng=1):
        return int(lib().vbhem_stats_size(K, r, self.d, self.covmode, ng))

    def estep(self, cen: dict, tau: int, want_pairs: bool = True, packed: bool = False, raw: bool = False) -> dict:
        """Fused E-step from pre-pass results (cen: logPi, logA, m, vW, cterm, logOmega[, group])."""
        pc = cen if packed else pack.pack_centres(cen)
        K, r = pc["K"], pc["maxSr"]
        grp = pc.get("group")
        ng = int(grp.max()) + 1 if grp is not None else 1
        buf = np.zeros(self.stats_size(K, r, ng))
        L = np.zeros((K, self.N)) if want_pairs else None
        z = np.zeros((K, self.N)) if want_pairs else None
        check(lib().vbhem_estep(self._h, K, _i(pc["Sr"]), r, _i(grp), ng, tau, _d(pc["logPi"]), _d(pc["logA"]),
                                _d(pc["m"]), _d(pc["vW"]), _d(pc["cterm"]), _d(pc["logOmega"]),
                                _d(L), _d(z), _d(buf)))
        if raw:
            return dict(buf=buf, L_elbo=L, zhat=z)
        S = self._split_stats(buf, K, r, ng)
        S["L_elbo"], S["zhat"] = L, z
        out = pack.stats_to_natural(S, self.d, self.covmode)
        out["buf"] = buf
        return out

    def estep_hyp(self, hyp: dict, tau: int, want_pairs: bool = True, group=None, stream_base: dict | None = None,
                  nslices: int = 8) -> dict:
        """Fused E-step from centre hyper-parameters; the pre-pass runs on the device.

        With ``stream_base`` (a base set of this handle's shape, natural or packed) the base set is streamed
        in from host memory slice by slice while earlier slices compute (``vbhem_estep_hyp_stream``).

        hyp (natural): Sr[K], alpha[K], eta[K,S], epsilon[K,S,S], m[K,S,d], W[K,S,d,d]|[K,S,d], v[K,S], lam[K,S].
        Also returns the pre-pass results the host-side lower bound needs
        (logOmega, logPi, logA (from,to), logLambdaTilde)."""
        c = np.ascontiguousarray
        K, r = hyp["eta"].shape
        Sr = c(hyp.get("Sr", np.full(K, r)), dtype=np.int32)
        full = self.covmode == 1
        W = np.asarray(hyp["W"], dtype=np.float64)
        Wp = c(np.transpose(W, (0, 1, 3, 2))) if full else c(W)
        eps = c(np.transpose(hyp["epsilon"], (0, 2, 1)), dtype=np.float64)
        grp = None if group is None else c(group, dtype=np.int32)
        ng = int(grp.max()) + 1 if grp is not None else 1
        buf = np.zeros(self.stats_size(K, r, ng))
        pre = np.zeros(K + r * K * (2 + r))
        L = np.zeros((K, self.N)) if want_pairs else None
        z = np.zeros((K, self.N)) if want_pairs else None
        f64 = lambda a: c(a, dtype=np.float64)
        al, et, m, v, lam = f64(hyp["alpha"]), f64(hyp["eta"]), f64(hyp["m"]), f64(hyp["v"]), f64(hyp["lam"])
        if stream_base is not None:
            pb = stream_base if "Sb" in stream_base and "maxSb" in stream_base else pack.pack_base(stream_base)
            assert (pb["N"], pb["maxSb"], pb["d"], pb["covmode"]) == (self.N, self.maxSb, self.d, self.covmode)
            check(lib().vbhem_estep_hyp_stream(self._h, _i(pb["Sb"]), _d(pb["prior"]), _d(pb["A"]), _d(pb["mu"]),
                                               _d(pb["cov"]), _d(pb["Ntilde"]), nslices, K, _i(Sr), r, _i(grp), ng,
                                               tau, _d(al), _d(et), _d(eps), _d(m), _d(Wp), _d(v), _d(lam), _d(L),
                                               _d(z), _d(buf), _d(pre)))
        else:
            check(lib().vbhem_estep_hyp(self._h, K, _i(Sr), r, _i(grp), ng, tau, _d(al), _d(et), _d(eps), _d(m),
                                        _d(Wp), _d(v), _d(lam), _d(L), _d(z), _d(buf), _d(pre)))
        S = self._split_stats(buf, K, r, ng)
        S["L_elbo"], S["zhat"] = L, z
        out = pack.stats_to_natural(S, self.d, self.covmode)
        out["buf"] = buf
        o = 0
        out["logOmega"] = pre[o:o + K].copy(); o += K
        out["logPi"] = pre[o:o + r * K].reshape(K, r).copy(); o += r * K
        out["logA"] = np.transpose(pre[o:o + r * r * K].reshape(K, r, r), (0, 2, 1)).copy(); o += r * r * K
        out["logLambdaTilde"] = pre[o:o + r * K].reshape(K, r).copy()
        return out

    # -- device-resident EM run (vbhem_em_*) ---------------------------------
    def em_begin(self, hyp: dict, opt: dict, group=None):
        """Upload the prior (opt: alpha0, eta0, epsilon0, lambda0, v0, m0, W0) and the initial centre
        hyper-parameters (natural layout, as for ``estep_hyp``); they then stay in HBM between iterations."""
        c = np.ascontiguousarray
        f64 = lambda a: c(a, dtype=np.float64)
        K, r = hyp["eta"].shape
        self._em = dict(K=K, r=r, Sr=c(hyp.get("Sr", np.full(K, r)), dtype=np.int32),
                        grp=None if group is None else c(group, dtype=np.int32))
        ng = int(self._em["grp"].max()) + 1 if group is not None else 1
        self._em["ng"] = ng
        full = self.covmode == 1
        W = np.asarray(hyp["W"], dtype=np.float64)
        Wp = c(np.transpose(W, (0, 1, 3, 2))) if full else c(W)
        eps = c(np.transpose(hyp["epsilon"], (0, 2, 1)), dtype=np.float64)
        m0 = f64(np.asarray(opt["m0"], float).reshape(self.d))
        W0 = f64(np.atleast_1d(np.asarray(opt["W0"], float)))
        al, et, m, v, lam = f64(hyp["alpha"]), f64(hyp["eta"]), f64(hyp["m"]), f64(hyp["v"]), f64(hyp["lam"])
        check(lib().vbhem_em_begin(self._h, K, _i(self._em["Sr"]), r, _i(self._em["grp"]), ng,
                                   float(opt["alpha0"]), float(opt["eta0"]), float(opt["epsilon0"]),
                                   float(opt["lambda0"]), float(opt["v0"]), _d(m0), _d(W0), int(W0.size),
                                   _d(al), _d(et), _d(eps), _d(m), _d(Wp), _d(v), _d(lam)))

    def em_begin_baseem(self, Sr, group, randb, randg, omega, opt: dict):
        """``em_begin`` with the ``baseem`` initialisation built on the device (``vbhem_em_begin_baseem``) from draws:
        Sr[K], group[K] (or None), randb / randg [K, maxS] 0-based, omega[K] normalised per H3M."""
        c = np.ascontiguousarray
        Sr = c(Sr, dtype=np.int32)
        K, r = np.shape(randb)
        grp = None if group is None else c(group, dtype=np.int32)
        ng = int(grp.max()) + 1 if grp is not None else 1
        self._em = dict(K=K, r=r, Sr=Sr, grp=grp, ng=ng)
        m0 = c(np.asarray(opt["m0"], float).reshape(self.d))
        W0 = c(np.atleast_1d(np.asarray(opt["W0"], float)))
        rb, rg = c(randb, dtype=np.int32), c(randg, dtype=np.int32)
        check(lib().vbhem_em_begin_baseem(self._h, K, _i(Sr), r, _i(grp), ng, float(opt["alpha0"]), float(opt["eta0"]),
                                          float(opt["epsilon0"]), float(opt["lambda0"]), float(opt["v0"]), _d(m0),
                                          _d(W0), int(W0.size), _i(rb), _i(rg), _d(c(omega, dtype=np.float64)),
                                          float(opt["Nv"]) * self.N))

    def em_estep(self, tau: int, stats_dev: int | None = None):
        check(lib().vbhem_em_estep(self._h, tau, C.c_void_p(stats_dev or 0)))

    def em_lower_bound(self, stats_dev: int | None = None) -> np.ndarray:
        """[ngroups, 12]: Lt1..Lt10, LB, sum Nj."""
        out = np.zeros((self._em["ng"], 12))
        check(lib().vbhem_em_lower_bound(self._h, C.c_void_p(stats_dev or 0), _d(out)))
        return out

    def em_mstep(self, stats_dev: int | None = None):
        check(lib().vbhem_em_mstep(self._h, C.c_void_p(stats_dev or 0)))

    def em_fetch(self) -> dict:
        """The resident hyper-parameters, natural layout."""
        K, r, d, dd = self._em["K"], self._em["r"], self.d, self.dd
        o = dict(alpha=np.zeros(K), eta=np.zeros((K, r)), eps=np.zeros((K, r, r)), m=np.zeros((K, r, d)),
                 W=np.zeros((K, r, d, d) if self.covmode else (K, r, d)), v=np.zeros((K, r)), lam=np.zeros((K, r)))
        check(lib().vbhem_em_fetch(self._h, _d(o["alpha"]), _d(o["eta"]), _d(o["eps"]), _d(o["m"]), _d(o["W"]),
                                   _d(o["v"]), _d(o["lam"])))
        return dict(alpha=o["alpha"], eta=o["eta"], epsilon=np.transpose(o["eps"], (0, 2, 1)).copy(), m=o["m"],
                    W=np.transpose(o["W"], (0, 1, 3, 2)).copy() if self.covmode else o["W"], v=o["v"], lam=o["lam"],
                    Sr=self._em["Sr"].copy())

    def em_run(self, tau: int, max_iter: int = 200, min_diff: float = 1e-5) -> dict:
        """vbhem_h3m_c_step_fc.m:115-433 entirely through the C ABI (one H3M, one GPU)."""
        LogLs = np.zeros(max_iter + 2)
        it, bad = C.c_int(), C.c_int()
        check(lib().vbhem_em_run(self._h, tau, max_iter, min_diff, _d(LogLs), C.byref(it), C.byref(bad)))
        n = it.value + (1 if bad.value else 0)
        return dict(LogLs=LogLs[:n].copy(), iters=it.value, unstable=bool(bad.value), hyp=self.em_fetch())

    def em_run_groups(self, tau: int, max_iter: int = 200, min_diff: float = 1e-5) -> dict:
        """Every batched H3M of ``em_begin(..., group=...)`` run to its own convergence in one loop
        (``vbhem_em_run_groups``): LogLs is a list of per-group traces."""
        G, ld = self._em["ng"], max_iter + 2
        LogLs = np.zeros((G, ld))
        it, bad = np.zeros(G, np.int32), np.zeros(G, np.int32)
        check(lib().vbhem_em_run_groups(self._h, tau, max_iter, min_diff, _d(LogLs), ld,
                                        it.ctypes.data_as(C.POINTER(C.c_int)), bad.ctypes.data_as(C.POINTER(C.c_int))))
        return dict(LogLs=[LogLs[g, :it[g] + (1 if bad[g] else 0)].copy() for g in range(G)], iters=it,
                    unstable=bad.astype(bool), hyp=self.em_fetch())

    def em_set_group_priors(self, opts: list):
        """Every batched H3M under its own prior (``vbhem_em_set_group_priors``): one option dict per soft-max group
        (alpha0, eta0, epsilon0, lambda0, v0, m0, W0 -- all W0 scalars or all d-vectors)."""
        c = np.ascontiguousarray
        G = self._em["ng"]
        assert len(opts) == G
        col = lambda k: c([float(o[k]) for o in opts])
        m0 = c(np.stack([np.asarray(o["m0"], float).reshape(self.d) for o in opts]))
        W0 = c(np.stack([np.atleast_1d(np.asarray(o["W0"], float)) for o in opts]))
        check(lib().vbhem_em_set_group_priors(self._h, _d(col("alpha0")), _d(col("eta0")), _d(col("epsilon0")),
                                              _d(col("lambda0")), _d(col("v0")), _d(m0), _d(W0), int(W0.shape[1])))

    def em_lb_deriv(self) -> np.ndarray:
        """[ngroups, 21] raw derivatives of the bound (``vbhem_em_lb_deriv``): alpha0, eta0, epsilon0, v0, lambda0,
        W0[8], m0[8]."""
        out = np.zeros((self._em["ng"], 21))
        check(lib().vbhem_em_lb_deriv(self._h, _d(out)))
        return out

    def em_run_groups_deriv(self, tau: int, max_iter: int = 200, min_diff: float = 1e-5) -> dict:
        """``em_run_groups`` + the derivatives of every group's bound at its breaking pass (``dLL`` [G, 21])."""
        G, ld = self._em["ng"], max_iter + 2
        LogLs, dLL = np.zeros((G, ld)), np.zeros((G, 21))
        it, bad = np.zeros(G, np.int32), np.zeros(G, np.int32)
        check(lib().vbhem_em_run_groups_deriv(self._h, tau, max_iter, min_diff, _d(LogLs), ld,
                                              it.ctypes.data_as(C.POINTER(C.c_int)),
                                              bad.ctypes.data_as(C.POINTER(C.c_int)), _d(dLL)))
        return dict(LogLs=[LogLs[g, :it[g] + (1 if bad[g] else 0)].copy() for g in range(G)], iters=it,
                    unstable=bad.astype(bool), dLL=dLL, hyp=self.em_fetch())

    def finalize(self, buf: np.ndarray, Sr, r: int, ng: int = 1) -> dict:
        """vbhem_compute_Statistics.m:57-82 on a raw stats buffer (e.g. after an all-reduce)."""
        Sr = np.ascontiguousarray(Sr, dtype=np.int32)
        K = Sr.shape[0]
        buf = np.ascontiguousarray(buf, dtype=np.float64).copy()
        check(lib().vbhem_finalize_stats(self._h, K, _i(Sr), r, ng, _d(buf)))
        S = self._split_stats(buf, K, r, ng)
        out = pack.stats_to_natural(S, self.d, self.covmode)
        return dict(Nj=out["Nj"], Nj_rho1=out["nu1"], Nj_rho2rho=out["xi"], Nj_rho=out["pr"], y_bar=out["mu"],
                    S_plus_C=out["Mu"], lb=out["lb"])


def _hyp_arrays(hyp: dict, covmode: int):
    """Natural centre hyper-parameters -> the C-ABI arrays of vbhem_estep_hyp (contiguous, column-major meaning)."""
    c = np.ascontiguousarray
    f64 = lambda a: c(a, dtype=np.float64)
    K, r = hyp["eta"].shape
    Sr = c(hyp.get("Sr", np.full(K, r)), dtype=np.int32)
    W = np.asarray(hyp["W"], dtype=np.float64)
    Wp = c(np.transpose(W, (0, 1, 3, 2))) if covmode == 1 else c(W)
    eps = c(np.transpose(hyp["epsilon"], (0, 2, 1)), dtype=np.float64)
    return K, r, Sr, [f64(hyp["alpha"]), f64(hyp["eta"]), eps, f64(hyp["m"]), Wp, f64(hyp["v"]), f64(hyp["lam"])]


class MultiBaseSet:
    """A base set sharded over several GPUs behind the C ABI (``vbhem_multi_*``): one host process drives every
    device (the arrangement of the reference's single-session ``parfor``, vbhem_h3m_c.m:28), or -- with ``world`` /
    ``rank_first`` / ``unique_id`` -- one process per device.  Same results as :class:`BaseSet` up to fp64
    summation order; the raw statistics are all-reduced with NCCL inside the library."""

    def __init__(self, base: dict, devices=None, packed: bool = False, world: int | None = None, rank_first: int = 0,
                 unique_id: bytes | None = None):
        pb = base if packed else pack.pack_base(base)
        self.N, self.maxSb, self.d, self.covmode = pb["N"], pb["maxSb"], pb["d"], pb["covmode"]
        self.dd = self.d * self.d if self.covmode else self.d
        if devices is None:
            n = C.c_int()
            check(lib().vbhem_device_count(C.byref(n)))
            devices = list(range(n.value))
        self.devices = [int(x) for x in devices]
        dv = np.ascontiguousarray(self.devices, dtype=np.int32)
        self.world = len(self.devices) if world is None else int(world)
        idp = None
        if unique_id is not None:
            self._id = (C.c_ubyte * 128).from_buffer_copy(unique_id)
            idp = self._id
        self._h = C.c_void_p()
        check(lib().vbhem_multi_pack_base(len(self.devices), _i(dv), self.world, int(rank_first), idp, self.N,
                                          _i(pb["Sb"]), self.maxSb, self.d, self.covmode, _d(pb["prior"]),
                                          _d(pb["A"]), _d(pb["mu"]), _d(pb["cov"]), _d(pb["Ntilde"]),
                                          C.byref(self._h)))

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_ubyte * 128)()
        check(lib().vbhem_multi_unique_id(buf))
        return bytes(buf)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().vbhem_multi_free(self._h)
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

    def set_kernel(self, which: str = "auto"):
        check(lib().vbhem_multi_set_kernel(self._h, {"auto": 0, "general": 1, "unpaired": 2, "halves": 3}[which]))

    # -- peer-memory all-reduce between processes: exchange the IPC handles once (else NCCL is used) -----------
    def peer_export(self) -> bytes:
        buf = (C.c_ubyte * (64 * len(self.devices)))()
        check(lib().vbhem_multi_peer_export(self._h, buf))
        return bytes(buf)

    def peer_import(self, handles_in_rank_order: list):
        blob = b"".join(handles_in_rank_order)
        assert len(blob) == 64 * self.world
        check(lib().vbhem_multi_peer_import(self._h, (C.c_ubyte * len(blob)).from_buffer_copy(blob)))

    def allreduce_mode(self) -> str:
        p = C.c_int()
        check(lib().vbhem_multi_allreduce_mode(self._h, C.byref(p)))
        return "peer" if p.value else ("nccl" if self.world > 1 else "none")

    def shards(self):
        out = []
        for k in range(len(self.devices)):
            dev, lo, n = C.c_int(), C.c_int(), C.c_int()
            check(lib().vbhem_multi_shard(self._h, k, C.byref(dev), C.byref(lo), C.byref(n), None))
            out.append((dev.value, lo.value, n.value))
        return out

    def counters(self) -> dict:
        out = (C.c_longlong * 8)()
        check(lib().vbhem_multi_counters(self._h, out))
        return dict(launches=out[0], lattice_pairs=out[1], general_pairs=out[2], active_pairs=out[3], tiles=out[4],
                    split_tiles=out[5], compactions=out[6])

    def last_ms(self) -> float:
        ms = C.c_double()
        check(lib().vbhem_multi_last_ms(self._h, C.byref(ms)))
        return ms.value

    _split_stats = BaseSet._split_stats
    stats_size = BaseSet.stats_size

    def estep_hyp(self, hyp: dict, tau: int, want_pairs: bool = True, group=None) -> dict:
        """As :meth:`BaseSet.estep_hyp`; the statistics are the sums over ALL shards (and ranks)."""
        K, r, Sr, H = _hyp_arrays(hyp, self.covmode)
        grp = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
        ng = int(grp.max()) + 1 if grp is not None else 1
        buf = np.zeros(self.stats_size(K, r, ng))
        pre = np.zeros(K + r * K * (2 + r))
        L = np.zeros((K, self.N)) if want_pairs else None
        z = np.zeros((K, self.N)) if want_pairs else None
        check(lib().vbhem_multi_estep_hyp(self._h, K, _i(Sr), r, _i(grp), ng, tau, *[_d(h) for h in H], _d(L), _d(z),
                                          _d(buf), _d(pre)))
        S = self._split_stats(buf, K, r, ng)
        S["L_elbo"], S["zhat"] = L, z
        out = pack.stats_to_natural(S, self.d, self.covmode)
        out["buf"] = buf
        return out

    def estep(self, cen: dict, tau: int, want_pairs: bool = True) -> dict:
        pc = pack.pack_centres(cen)
        K, r = pc["K"], pc["maxSr"]
        grp = pc.get("group")
        ng = int(grp.max()) + 1 if grp is not None else 1
        buf = np.zeros(self.stats_size(K, r, ng))
        L = np.zeros((K, self.N)) if want_pairs else None
        z = np.zeros((K, self.N)) if want_pairs else None
        check(lib().vbhem_multi_estep(self._h, K, _i(pc["Sr"]), r, _i(grp), ng, tau, _d(pc["logPi"]), _d(pc["logA"]),
                                      _d(pc["m"]), _d(pc["vW"]), _d(pc["cterm"]), _d(pc["logOmega"]), _d(L), _d(z),
                                      _d(buf)))
        S = self._split_stats(buf, K, r, ng)
        S["L_elbo"], S["zhat"] = L, z
        out = pack.stats_to_natural(S, self.d, self.covmode)
        out["buf"] = buf
        return out

    # -- non-synchronising / host-streamed variants (bench.py) ---------------
    def prepare_hyp(self, hyp: dict, group=None) -> dict:
        """The C-ABI arrays of a centre set, converted once (the timed loops of bench.py pass them as they are)."""
        K, r, Sr, H = _hyp_arrays(hyp, self.covmode)
        grp = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
        return dict(K=K, r=r, Sr=Sr, H=H, grp=grp, ng=int(grp.max()) + 1 if grp is not None else 1)

    def estep_hyp_dev(self, P: dict, tau: int, finalize: bool = True):
        """Enqueue E-step + all-reduce (+ normalisation) on every shard and return (``vbhem_multi_estep_hyp_dev``)."""
        check(lib().vbhem_multi_estep_hyp_dev(self._h, P["K"], _i(P["Sr"]), P["r"], _i(P["grp"]), P["ng"], tau,
                                              *[_d(h) for h in P["H"]], 1 if finalize else 0))

    def synchronize(self):
        check(lib().vbhem_multi_synchronize(self._h))

    def align(self):
        """Device-side rendezvous of every shard of every rank (``vbhem_multi_align``): what is enqueued next starts on
        all devices together, whatever the launch skew of the host processes."""
        check(lib().vbhem_multi_align(self._h))

    def stats(self, P: dict) -> np.ndarray:
        buf = np.zeros(self.stats_size(P["K"], P["r"], P["ng"]))
        check(lib().vbhem_multi_stats(self._h, _d(buf)))
        return buf

    def fetch(self, K: int):
        L, z = np.zeros((K, self.N)), np.zeros((K, self.N))
        check(lib().vbhem_multi_fetch(self._h, _d(L), _d(z)))
        return L, z

    def estep_hyp_stream(self, pb: dict, P: dict, tau: int, nslices: int, L, z, stats, pre):
        """Host-resident (packed, ideally pinned) base set streamed in by every shard; host outputs as given."""
        check(lib().vbhem_multi_estep_hyp_stream(self._h, _i(pb["Sb"]), _d(pb["prior"]), _d(pb["A"]), _d(pb["mu"]),
                                                 _d(pb["cov"]), _d(pb["Ntilde"]), nslices, P["K"], _i(P["Sr"]), P["r"],
                                                 _i(P["grp"]), P["ng"], tau, *[_d(h) for h in P["H"]], _d(L), _d(z),
                                                 _d(stats), _d(pre)))

    def shard_handle(self, k: int = 0):
        h = C.c_void_p()
        check(lib().vbhem_multi_shard(self._h, k, None, None, None, C.byref(h)))
        return h

    def finalize(self, buf: np.ndarray, Sr, r: int, ng: int = 1) -> dict:
        Sr = np.ascontiguousarray(Sr, dtype=np.int32)
        K = Sr.shape[0]
        buf = np.ascontiguousarray(buf, dtype=np.float64).copy()
        check(lib().vbhem_multi_finalize_stats(self._h, K, _i(Sr), r, ng, _d(buf)))
        S = self._split_stats(buf, K, r, ng)
        out = pack.stats_to_natural(S, self.d, self.covmode)
        return dict(Nj=out["Nj"], Nj_rho1=out["nu1"], Nj_rho2rho=out["xi"], Nj_rho=out["pr"], y_bar=out["mu"],
                    S_plus_C=out["Mu"], lb=out["lb"])

    # -- device-resident EM over the shards ----------------------------------
    def em_begin(self, hyp: dict, opt: dict, group=None):
        K, r, Sr, H = _hyp_arrays(hyp, self.covmode)
        grp = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
        ng = int(grp.max()) + 1 if grp is not None else 1
        self._em = dict(K=K, r=r, Sr=Sr, grp=grp, ng=ng)
        m0 = np.ascontiguousarray(np.asarray(opt["m0"], float).reshape(self.d))
        W0 = np.ascontiguousarray(np.atleast_1d(np.asarray(opt["W0"], float)))
        check(lib().vbhem_multi_em_begin(self._h, K, _i(Sr), r, _i(grp), ng, float(opt["alpha0"]), float(opt["eta0"]),
                                         float(opt["epsilon0"]), float(opt["lambda0"]), float(opt["v0"]), _d(m0),
                                         _d(W0), int(W0.size), *[_d(h) for h in H]))

    def em_step(self, tau: int) -> np.ndarray:
        out = np.zeros((self._em["ng"], 12))
        check(lib().vbhem_multi_em_step(self._h, tau, _d(out)))
        return out

    def em_fetch(self) -> dict:
        K, r, d = self._em["K"], self._em["r"], self.d
        o = dict(alpha=np.zeros(K), eta=np.zeros((K, r)), eps=np.zeros((K, r, r)), m=np.zeros((K, r, d)),
                 W=np.zeros((K, r, d, d) if self.covmode else (K, r, d)), v=np.zeros((K, r)), lam=np.zeros((K, r)))
        check(lib().vbhem_multi_em_fetch(self._h, _d(o["alpha"]), _d(o["eta"]), _d(o["eps"]), _d(o["m"]), _d(o["W"]),
                                         _d(o["v"]), _d(o["lam"])))
        return dict(alpha=o["alpha"], eta=o["eta"], epsilon=np.transpose(o["eps"], (0, 2, 1)).copy(), m=o["m"],
                    W=np.transpose(o["W"], (0, 1, 3, 2)).copy() if self.covmode else o["W"], v=o["v"], lam=o["lam"],
                    Sr=self._em["Sr"].copy())

    def em_run(self, tau: int, max_iter: int = 200, min_diff: float = 1e-5) -> dict:
        LogLs = np.zeros(max_iter + 2)
        it, bad = C.c_int(), C.c_int()
        check(lib().vbhem_multi_em_run(self._h, tau, max_iter, min_diff, _d(LogLs), C.byref(it), C.byref(bad)))
        n = it.value + (1 if bad.value else 0)
        return dict(LogLs=LogLs[:n].copy(), iters=it.value, unstable=bool(bad.value), hyp=self.em_fetch())

    def em_run_groups(self, tau: int, max_iter: int = 200, min_diff: float = 1e-5) -> dict:
        G, ld = self._em["ng"], max_iter + 2
        LogLs = np.zeros((G, ld))
        it, bad = np.zeros(G, np.int32), np.zeros(G, np.int32)
        check(lib().vbhem_multi_em_run_groups(self._h, tau, max_iter, min_diff, _d(LogLs), ld,
                                              it.ctypes.data_as(C.POINTER(C.c_int)),
                                              bad.ctypes.data_as(C.POINTER(C.c_int))))
        return dict(LogLs=[LogLs[g, :it[g] + (1 if bad[g] else 0)].copy() for g in range(G)], iters=it,
                    unstable=bad.astype(bool), hyp=self.em_fetch())
