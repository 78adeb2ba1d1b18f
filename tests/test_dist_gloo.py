"""world_size-2 gloo test (CPU): the sharding + single all-reduce of vbhem_b200.dist against the un-sharded
oracle.  The compute callable is the oracle here (no GPU in this tier); on the GPU box the same class runs the
CUDA library over NCCL (bench.py --gpus N, tests/test_gpu_parity.py::test_sharded_matches_single)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import vbhem_b200  # noqa: F401
    from vbhem_b200 import pack
    from vbhem_b200.dist import ShardedEStep, shard_base, shard_bounds
    from oracle import cref, numpy_ref as nr
    from util import make_case
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, hyp, cen = make_case(N=N, K=4, Sb=4, Sr=3, d=2, covmode="full", ragged=True, seed=91, Nv=6.0)
    shard = shard_base(base, world, rank)
    lo, hi = shard_bounds(N, world, rank)
    assert shard["prior"].shape[0] == hi - lo and np.array_equal(shard["Ntilde"], base["Ntilde"][lo:hi])
    pc = pack.pack_centres(cen)

    def compute(hyp_, tau):
        S = cref.port_estep_fused(pack.pack_base(shard), pc, tau, want_pairs=False)
        S["Nj"] = S["Nj"] - 1e-50                 # raw sums travel; the +1e-50 belongs to the normalisation
        buf = np.concatenate([S[k].ravel() for k in ("Nj", "nu1", "xi", "pr", "mu", "Mu", "lb")])
        return torch.from_numpy(buf), None

    def finalize(buf, _):
        return buf.numpy()

    got = ShardedEStep(shard, compute=compute, finalize=finalize).estep_hyp(hyp, 5)
    if rank == 0:
        np.save(os.path.join(out_dir, "got.npy"), got)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [37, 38])
def test_two_rank_allreduce_equals_unsharded(tmp_path, N):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import cref
    from vbhem_b200 import pack
    from util import make_case
    port = _free_port()
    mp.spawn(_worker, args=(2, port, N, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(tmp_path, "got.npy"))
    base, hyp, cen = make_case(N=N, K=4, Sb=4, Sr=3, d=2, covmode="full", ragged=True, seed=91, Nv=6.0)
    S = cref.port_estep_fused(pack.pack_base(base), pack.pack_centres(cen), 5, want_pairs=False)
    S["Nj"] = S["Nj"] - 1e-50
    want = np.concatenate([S[k].ravel() for k in ("Nj", "nu1", "xi", "pr", "mu", "Mu", "lb")])
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12 * np.abs(want).max())


def test_shard_bounds_cover_everything():
    from vbhem_b200.dist import shard_bounds
    for N in (1, 7, 8, 100_000):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(N, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == N
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _em_worker(rank, world, port, out_dir):
    """Sharded EM loop (dist.ShardedEStep.em_run) with every device op replaced by the oracle / numpy formulas."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import vbhem_b200  # noqa: F401
    from vbhem_b200 import h3m, pack, synth
    from vbhem_b200.dist import ShardedEStep, shard_base
    from oracle import cref, numpy_ref as nr
    from util import prepass_natural
    from test_em_loop import OPT
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, _ = synth.make_base_set(60, S=3, d=2, C=3, covmode="full", seed=77, Nv=OPT["Nv"])
    hyp0 = synth.init_centres_baseem(base, 3, 3, seed=77, opt=dict(synth.DEFAULT_OPT, Nv=OPT["Nv"]))
    shard = pack.pack_base(shard_base(base, world, rank))
    state = dict(hyp={k: np.array(v) for k, v in hyp0.items()}, pre=None, fin=None)
    keys = ("Nj", "nu1", "xi", "pr", "mu", "Mu", "lb")

    def estep(tau):
        cen = prepass_natural(state["hyp"], "full")
        state["pre"] = dict(logOmega=cen["logOmega"], logPi=cen["logPi"], logA=cen["logA"],
                            logLambdaTilde=cen["logLambdaTilde"])
        S = cref.port_estep_fused(shard, pack.pack_centres(cen), tau, want_pairs=False)
        S["Nj"] = S["Nj"] - 1e-50                     # raw sums travel
        state["shapes"] = [S[k].shape for k in keys]
        return torch.from_numpy(np.concatenate([S[k].ravel() for k in keys]))

    def lower_bound(buf):
        parts, o = {}, 0
        for k, sh in zip(keys, state["shapes"]):
            n = int(np.prod(sh))
            parts[k] = buf.numpy()[o:o + n].reshape(sh)
            o += n
        St = pack.stats_to_natural(parts, 2, 1)
        K = St["Nj"].shape[0]
        fins = [nr.finalize_stats(St["nu1"][j], St["xi"][j], St["pr"][j], St["mu"][j], St["Mu"][j]) for j in range(K)]
        state["fin"] = dict(Nj=St["Nj"] + 1e-50, Nj_rho1=np.stack([f[0] for f in fins]),
                            Nj_rho2rho=np.stack([f[1] for f in fins]), Nj_rho=np.stack([f[2] for f in fins]),
                            y_bar=np.stack([f[3] for f in fins]), S_plus_C=np.stack([f[4] for f in fins]))
        L, _ = h3m.vbhemh3m_lb(St["lb"][0], St["lb"][1], state["fin"]["Nj"], state["pre"], state["hyp"], OPT)
        return L

    def mstep(buf):
        new = h3m.vbhem_mstep(state["fin"], OPT)
        new["Sr"] = state["hyp"]["Sr"]
        state["hyp"] = new

    ops = dict(estep=estep, lower_bound=lower_bound, mstep=mstep, fetch=lambda: state["hyp"], sync=lambda: None)
    got = ShardedEStep(None, compute=lambda *a: None).em_run(hyp0, OPT, ops=ops)
    if rank == 0:
        np.savez(os.path.join(out_dir, "em.npz"), LogLs=got["LogLs"], iters=got["iters"],
                 **{f"hyp_{k}": v for k, v in got["hyp"].items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_em_loop_equals_single_process_oracle_em(tmp_path):
    """The multi-GPU EM loop (E-step per shard, ONE all-reduce, identical lower bound / M-step on every rank,
    convergence rule of vbhem_h3m_c_step_fc.m:311-354) on two gloo ranks against the un-sharded oracle loop."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from vbhem_b200 import synth
    from test_em_loop import OPT, oracle_em
    port = _free_port()
    mp.spawn(_em_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(tmp_path, "em.npz"))
    base, _ = synth.make_base_set(60, S=3, d=2, C=3, covmode="full", seed=77, Nv=OPT["Nv"])
    hyp0 = synth.init_centres_baseem(base, 3, 3, seed=77, opt=dict(synth.DEFAULT_OPT, Nv=OPT["Nv"]))
    want_hyp, want_L = oracle_em(base, hyp0, OPT)
    assert int(got["iters"]) == len(want_L)
    np.testing.assert_allclose(got["LogLs"], want_L, rtol=1e-10)
    for k in ("alpha", "eta", "epsilon", "lam", "v", "m", "W"):
        np.testing.assert_allclose(got[f"hyp_{k}"], want_hyp[k], rtol=1e-8, atol=1e-10 * np.abs(want_hyp[k]).max(), err_msg=k)


def _worker_runs(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    import vbhem_b200  # noqa: F401
    from vbhem_b200.dist import run_sharded, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    items = [dict(subject=i, data=np.arange(i + 1.0)) for i in range(n)]
    seen = []

    def fn(chunk):                       # stands for hmm.vbhmm_learn_batch on this rank's GPU
        seen.extend(c["subject"] for c in chunk)
        return [dict(subject=c["subject"], total=float(c["data"].sum()), rank=rank) for c in chunk]

    out = run_sharded(items, fn)
    lo, hi = shard_bounds(n, world, rank)
    assert seen == list(range(lo, hi))
    if rank == 1:
        np.save(os.path.join(out_dir, "runs.npy"), np.array([[o["subject"], o["total"], o["rank"]] for o in out]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 1])
def test_independent_runs_partition_over_two_ranks(tmp_path, n):
    """dist.run_sharded: the batched workloads (subjects of vbhmm_learn_batch, candidates of a model selection)
    partition over the ranks with no data-path collective; every rank ends up with every result, in order."""
    mp.spawn(_worker_runs, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "runs.npy"))
    assert got[:, 0].tolist() == list(range(n))
    assert got[:, 1].tolist() == [i * (i + 1) / 2 for i in range(n)]
    assert got[:, 2].tolist() == [0] * ((n + 1) // 2) + [1] * (n // 2)
