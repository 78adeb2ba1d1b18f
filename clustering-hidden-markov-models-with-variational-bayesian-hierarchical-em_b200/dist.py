"""Multi-GPU E-step: base HMMs sharded over the ranks of one box, centres replicated, ONE all-reduce.

The pairs (i, j) are independent; the only couplings are the soft-max over the centres j of one base i
(local, because a base's K pairs live on one rank) and the sums over bases i of the weighted statistics
(/root/reference/src/vbhem/vbhem_compute_Statistics.m:33-55), which are what is exchanged: one contiguous
fp64 buffer of ``vbhem_stats_size()`` doubles -- Nj | S_nu1 | S_xi | S_pr | S_mu | S_Mu | (Lt1, Lt7) --
64 x 129 + 2 doubles = 66 KB at the north-star shape, all-reduced (sum) over NCCL / NVLink.  The Z > 1e-8
skip rule is applied per shard *before* accumulation and the normalisation of vbhem_compute_Statistics.m:57-82
*after* the all-reduce, so the result does not depend on the number of ranks beyond fp64 summation order.

One process per GPU (torchrun); torch.distributed is plumbing only.  The compute callable is injectable so
that the sharding / exchange logic can be exercised with the gloo backend on CPU-only machines (tests use the
oracle there; the default is the CUDA library and has no CPU fallback).
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_bounds(N: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split of N bases: the first N % world ranks get one extra."""
    q, r = divmod(N, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_base(base: dict, world: int, rank: int) -> dict:
    """Rank's slice of a natural base dict (Ntilde keeps its GLOBAL meaning, Nv * Kb_total * omega_i)."""
    N = np.asarray(base["prior"]).shape[0]
    lo, hi = shard_bounds(N, world, rank)
    return {k: (np.asarray(v)[lo:hi] if np.ndim(v) >= 1 and np.asarray(v).shape[0] == N else v)
            for k, v in base.items()}


class ShardedEStep:
    """E-step over this rank's shard + all-reduce of the raw statistics + normalisation.

    compute(hyp, tau) -> (raw stats buffer as a torch tensor on the exchange device, extras dict)
    finalize(buffer)  -> dict of normalised statistics (in place on the same buffer)
    Defaults: the CUDA library on ``device`` (vbhem_estep_hyp_dev / vbhem_finalize_stats_dev).
    """

    def __init__(self, base_shard: dict, device: int | None = None, group=None, compute=None, finalize=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._compute, self._finalize = compute, finalize
        self.bs = None
        if compute is None:
            import torch
            from .api import BaseSet
            self.torch = torch
            self.device = torch.cuda.current_device() if device is None else device
            self.bs = BaseSet(base_shard, device=self.device)
            self.bs.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
            self._stats = None

    def close(self):
        if self.bs is not None:
            self.bs.close()

    # -- CUDA path ----------------------------------------------------------
    def _cuda_compute(self, hyp, tau):
        from ._lib import check, dp, ip, lib
        torch, bs = self.torch, self.bs
        c = np.ascontiguousarray
        K, r = hyp["eta"].shape
        Sr = c(hyp.get("Sr", np.full(K, r)), dtype=np.int32)
        W = np.asarray(hyp["W"], dtype=np.float64)
        Wp = c(np.transpose(W, (0, 1, 3, 2))) if bs.covmode == 1 else c(W)
        eps = c(np.transpose(hyp["epsilon"], (0, 2, 1)), dtype=np.float64)
        f64 = lambda a: c(a, dtype=np.float64)
        al, et, m, v, lam = f64(hyp["alpha"]), f64(hyp["eta"]), f64(hyp["m"]), f64(hyp["v"]), f64(hyp["lam"])
        ns = bs.stats_size(K, r)
        if self._stats is None or self._stats.numel() != ns:
            self._stats = torch.zeros(ns, dtype=torch.float64, device=f"cuda:{self.device}")
        pre = np.zeros(K + r * K * (2 + r))
        P = lambda a: a.ctypes.data_as(dp)
        check(lib().vbhem_estep_hyp_dev(bs._h, K, Sr.ctypes.data_as(ip), r, None, 1, tau, P(al), P(et), P(eps), P(m),
                                        P(Wp), P(v), P(lam), C.c_void_p(self._stats.data_ptr()), P(pre)))
        return self._stats, dict(Sr=Sr, r=r, K=K, pre=pre)

    def _cuda_finalize(self, buf, extra):
        from ._lib import check, ip, lib
        check(lib().vbhem_finalize_stats_dev(self.bs._h, extra["K"], extra["Sr"].ctypes.data_as(ip), extra["r"], 1,
                                             C.c_void_p(buf.data_ptr())))
        self.bs.synchronize()
        host = buf.cpu().numpy()
        S = self.bs._split_stats(host, extra["K"], extra["r"], 1)
        from . import pack
        out = pack.stats_to_natural(S, self.bs.d, self.bs.covmode)
        return dict(Nj=out["Nj"], Nj_rho1=out["nu1"], Nj_rho2rho=out["xi"], Nj_rho=out["pr"], y_bar=out["mu"],
                    S_plus_C=out["Mu"], lb=out["lb"], pre=extra["pre"])

    # -- one E-step ---------------------------------------------------------
    def estep_hyp(self, hyp: dict, tau: int) -> dict:
        compute = self._compute or self._cuda_compute
        finalize = self._finalize or self._cuda_finalize
        buf, extra = compute(hyp, tau)
        if self.world > 1:
            self.dist.all_reduce(buf, op=self.dist.ReduceOp.SUM, group=self.group)
        return finalize(buf, extra)

    # -- a whole EM run, hyper-parameters resident on every rank -------------
    def _cuda_em_ops(self, hyp, opt):
        """The device-resident EM state as five callables (vbhem_em_*); the statistics buffer is a CUDA tensor."""
        torch, bs = self.torch, self.bs
        K, r = hyp["eta"].shape
        bs.em_begin(hyp, opt)
        stats = torch.zeros(bs.stats_size(K, r), dtype=torch.float64, device=f"cuda:{self.device}")

        def estep(tau):
            bs.em_estep(tau, stats.data_ptr())
            return stats

        return dict(estep=estep, lower_bound=lambda buf: float(bs.em_lower_bound(buf.data_ptr())[0, 10]),
                    mstep=lambda buf: bs.em_mstep(buf.data_ptr()), fetch=bs.em_fetch, sync=bs.synchronize)

    def em_run(self, hyp: dict, opt: dict, covmode: str = "full", ops=None) -> dict:
        """vbhem_h3m_c_step_fc.m:115-433 over the sharded base set: every rank keeps the centre hyper-parameters
        (``vbhem_em_begin``: in HBM), runs the fused E-step on its shard, all-reduces the raw statistics (the one
        collective), and applies the identical lower bound + M-step, so no parameters travel between iterations.

        ``ops`` (estep(tau) -> raw statistics tensor, lower_bound(buf) -> L after normalising buf in place,
        mstep(buf), fetch() -> hyper-parameters, sync()) defaults to the CUDA library; the gloo test injects the
        oracle so that the loop and the exchange are exercised on CPU-only machines."""
        import time
        ops = ops or self._cuda_em_ops(hyp, opt)
        tau, max_iter, min_diff = int(opt["tau"]), int(opt.get("max_iter", 200)), float(opt.get("minDiff", 1e-5))
        LogLs, lastL, it, unstable, times = [], -np.finfo(float).max, 0, False, []
        while True:
            t0 = time.perf_counter()
            stats = ops["estep"](tau)
            if self.world > 1:
                self.dist.all_reduce(stats, op=self.dist.ReduceOp.SUM, group=self.group)
            L = ops["lower_bound"](stats)
            do_break = it > 1 and abs((L - lastL) / lastL) <= min_diff
            do_break = do_break or it == max_iter
            if np.isnan(L):
                do_break, unstable, L = True, True, -np.inf
            LogLs.append(L)
            if do_break and unstable:
                break
            ops["mstep"](stats)
            it += 1
            lastL = L
            times.append(time.perf_counter() - t0)
            if do_break:
                break
        ops["sync"]()
        return dict(hyp=ops["fetch"](), LogLs=np.array(LogLs), LL=LogLs[-1], iters=it, unstable=unstable,
                    ms_per_iter=1e3 * float(np.median(times)) if times else float("nan"))


def run_sharded(items: list, fn, group=None) -> list:
    """Independent work units over the ranks of one box, no data-path collective.

    The batched workloads next to the E-step partition instead of sharding: the subjects of ``hmm.vbhmm_learn_batch``
    (the reference's vbhmm_learn_batch.m loop) and the (K, S, trial) candidates of a model selection are independent
    runs.  ``items`` is split contiguously (``shard_bounds``); every rank calls ``fn(its slice)`` -- which must return
    one result per item -- on its own GPU, and the complete list comes back on every rank through one
    ``all_gather_object`` of the (small) results.  Without an initialised process group it is ``fn(items)``."""
    import torch.distributed as dist
    if not dist.is_initialized():
        return list(fn(items))
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(len(items), world, rank)
    mine = list(fn(items[lo:hi])) if hi > lo else []
    if len(mine) != hi - lo:
        raise ValueError("fn must return one result per work item")
    parts = [None] * world
    dist.all_gather_object(parts, mine, group=group)
    return [x for p in parts for x in p]
