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
