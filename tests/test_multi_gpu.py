"""GPU tier: the base set sharded over several GPUs behind the C ABI (vbhem_multi_*, SURVEY.md 8e / VERDICT r1 row e').

One host process drives every device -- the arrangement of the reference's single-session parfor
(/root/reference/src/vbhem/vbhem_h3m_c.m:28) -- or one process per device with the NCCL id passed by the caller.
The multi-device tests need >= 2 GPUs (`gpurun --gpus 2`); the single-device ones run everywhere a GPU is.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import make_case, oracle_fused

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ndev():
    import ctypes as C
    from vbhem_b200._lib import lib
    n = C.c_int()
    lib().vbhem_device_count(C.byref(n))
    return n.value


def two():
    if ndev() < 2:
        pytest.skip("needs >= 2 GPUs")
    return list(range(min(ndev(), 8)))


def api():
    from vbhem_b200 import api as a
    return a


def close(got, want, rtol=1e-11):
    for k in ("Nj", "nu1", "xi", "pr", "mu", "Mu", "lb"):
        np.testing.assert_allclose(got[k], want[k], rtol=rtol, atol=rtol * np.abs(want[k]).max(), err_msg=k)


def test_multi_with_one_device_is_the_single_gpu_path():
    base, hyp, cen = make_case(N=500, K=6, Sb=5, Sr=4, d=2, covmode="full", ragged=True, seed=31, Nv=20.0)
    with api().BaseSet(base) as bs:
        want = bs.estep_hyp(hyp, 6)
    with api().MultiBaseSet(base, devices=[0]) as ms:
        got = ms.estep_hyp(hyp, 6)
        got2 = ms.estep(cen, 6)
        assert ms.shards() == [(0, 0, 500)]
    for k in ("L_elbo", "zhat", "buf"):
        assert np.array_equal(got[k], want[k]), k
    close(got2, want)


@pytest.mark.parametrize("N", [2003, 517])
def test_multi_estep_equals_single_gpu(N):
    """the multi entry equals the single-GPU result to 1e-11 and is bit-identical run to run;
    the per-pair rows (L_elbo, z_hat) do not depend on the sharding at all."""
    devs = two()
    base, hyp, cen = make_case(N=N, K=7, Sb=6, Sr=5, d=2, covmode="full", ragged=True, seed=11, Nv=15.0)
    with api().BaseSet(base) as bs:
        want = bs.estep_hyp(hyp, 8)
        fin_want = bs.finalize(want["buf"], hyp["Sr"], 5)
    with api().MultiBaseSet(base, devices=devs) as ms:
        got = ms.estep_hyp(hyp, 8)
        again = ms.estep_hyp(hyp, 8)
        fin = ms.finalize(got["buf"], hyp["Sr"], 5)
        sh = ms.shards()
        c = ms.counters()
    assert [s[2] for s in sh] == [N // len(devs) + (1 if k < N % len(devs) else 0) for k in range(len(devs))]
    assert np.array_equal(got["L_elbo"], want["L_elbo"]) and np.array_equal(got["zhat"], want["zhat"])
    assert np.array_equal(got["buf"], again["buf"]), "run-to-run"
    close(got, want)
    for k in fin_want:
        np.testing.assert_allclose(fin[k], fin_want[k], rtol=1e-11, atol=1e-11 * np.abs(fin_want[k]).max(), err_msg=k)
    Z = want["zhat"] * base["Ntilde"][:, None]
    assert c["active_pairs"] == int((Z > 1e-8).sum())
    want_or = oracle_fused(base, cen, 8)
    np.testing.assert_allclose(got["L_elbo"], want_or["L_elbo"], rtol=1e-9)
    np.testing.assert_allclose(got["zhat"], want_or["zhat"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_align_is_a_rendezvous_and_leaves_results_alone(mode, monkeypatch):
    """vbhem_multi_align: a one-element all-reduce on every shard's stream (peer epochs advance with it); the E-steps
    around it give the bits they give without it."""
    devs = two()
    if mode == "nccl":
        monkeypatch.setenv("VBHEM_ALLREDUCE", "nccl")
    base, hyp, cen = make_case(N=901, K=5, Sb=6, Sr=5, d=2, covmode="full", ragged=True, seed=17, Nv=15.0)
    with api().MultiBaseSet(base, devices=devs) as ms:
        assert ms.allreduce_mode() == mode
        want = ms.estep_hyp(hyp, 8, want_pairs=False)["buf"]
        for _ in range(3):                 # odd count: the next E-step uses the other exchange buffer
            ms.align()
        ms.synchronize()
        got = ms.estep_hyp(hyp, 8, want_pairs=False)["buf"]
        ms.align()
        got2 = ms.estep_hyp(hyp, 8, want_pairs=False)["buf"]
    assert np.array_equal(got, want) and np.array_equal(got2, want)
    with api().MultiBaseSet(base, devices=[0]) as one:
        one.align()                        # one device: nothing to meet
        one.synchronize()


def test_peer_and_nccl_all_reduce_agree(monkeypatch):
    """The all-reduce over peer memory (k_allreduce_peer, the default where the devices can map each other) against
    ncclAllReduce (VBHEM_ALLREDUCE=nccl): same sums up to the order of the additions."""
    devs = two()
    base, hyp, cen = make_case(N=1203, K=7, Sb=6, Sr=5, d=2, covmode="full", ragged=True, seed=13, Nv=15.0)
    with api().MultiBaseSet(base, devices=devs) as ms:
        assert ms.allreduce_mode() == "peer"
        got = [ms.estep_hyp(hyp, 8, want_pairs=False)["buf"] for _ in range(3)]
    monkeypatch.setenv("VBHEM_ALLREDUCE", "nccl")
    with api().MultiBaseSet(base, devices=devs) as ms:
        assert ms.allreduce_mode() == "nccl"
        ref = ms.estep_hyp(hyp, 8, want_pairs=False)["buf"]
    assert np.array_equal(got[0], got[1]) and np.array_equal(got[0], got[2])
    np.testing.assert_allclose(got[0], ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())


def test_multi_em_run_equals_single_gpu():
    devs = two()
    from vbhem_b200 import synth
    base, hyp, cen = make_case(N=1200, K=5, Sb=4, Sr=4, d=2, covmode="full", ragged=True, seed=5, Nv=10.0)
    opt = dict(synth.DEFAULT_OPT, Nv=10.0, tau=6)
    with api().BaseSet(base) as bs:
        bs.em_begin(hyp, opt)
        want = bs.em_run(6, 12, 1e-6)
    with api().MultiBaseSet(base, devices=devs) as ms:
        ms.em_begin(hyp, opt)
        got = ms.em_run(6, 12, 1e-6)
        ms.em_begin(hyp, opt)
        again = ms.em_run(6, 12, 1e-6)
    assert got["iters"] == want["iters"] and not got["unstable"]
    np.testing.assert_allclose(got["LogLs"], want["LogLs"], rtol=1e-10)
    assert np.array_equal(got["LogLs"], again["LogLs"])
    for k in ("alpha", "eta", "epsilon", "m", "W", "v", "lam"):
        np.testing.assert_allclose(got["hyp"][k], want["hyp"][k], rtol=1e-8, atol=1e-10, err_msg=k)


def test_multi_em_run_groups_equals_single_gpu():
    """Batched candidates (soft-max groups) with compaction, over the shards."""
    devs = two()
    from vbhem_b200 import synth
    base, _, _ = make_case(N=900, K=3, Sb=4, Sr=3, d=2, covmode="full", ragged=True, seed=8, Nv=10.0)
    opt = dict(synth.DEFAULT_OPT, Nv=10.0, tau=5)
    hyp, grp, models = synth.model_selection_grid(base, Ks=range(1, 5), Ss=range(1, 4), seed=3, opt=opt)
    with api().BaseSet(base) as bs:
        bs.em_begin(hyp, opt, group=grp)
        want = bs.em_run_groups(5, 15, 1e-4)
    with api().MultiBaseSet(base, devices=devs) as ms:
        ms.em_begin(hyp, opt, group=grp)
        got = ms.em_run_groups(5, 15, 1e-4)
    assert list(got["iters"]) == list(want["iters"])
    for a, b in zip(got["LogLs"], want["LogLs"]):
        np.testing.assert_allclose(a, b, rtol=1e-10)
    np.testing.assert_allclose(got["hyp"]["m"], want["hyp"]["m"], rtol=1e-8, atol=1e-10)


WORKER = r"""
import sys, os, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import vbhem_b200
from vbhem_b200 import api
from vbhem_b200.dist import shard_base
from util import make_case
rank, world, idfile, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
base, hyp, cen = make_case(N=1501, K=6, Sb=5, Sr=4, d=2, covmode="full", ragged=True, seed=77, Nv=12.0)
uid = open(idfile, "rb").read()
import time
with api.MultiBaseSet(shard_base(base, world, rank), devices=[rank], world=world, rank_first=rank, unique_id=uid) as ms:
    nccl = ms.estep_hyp(hyp, 7)
    assert ms.allreduce_mode() == "nccl"
    # the peer-memory all-reduce: exchange the IPC handles of the exchange areas through files
    open(idfile + ".h" + str(rank) + ".tmp", "wb").write(ms.peer_export())
    os.rename(idfile + ".h" + str(rank) + ".tmp", idfile + ".h" + str(rank))
    hs = []
    for r in range(world):
        while not os.path.exists(idfile + ".h" + str(r)):
            time.sleep(0.01)
        hs.append(open(idfile + ".h" + str(r), "rb").read())
    ms.peer_import(hs)
    assert ms.allreduce_mode() == "peer"
    got = ms.estep_hyp(hyp, 7)
    again = ms.estep_hyp(hyp, 7)
assert np.array_equal(got["buf"], again["buf"])
np.testing.assert_allclose(got["buf"], nccl["buf"], rtol=1e-12, atol=1e-12 * np.abs(nccl["buf"]).max())
np.savez(out, buf=got["buf"], L=got["L_elbo"])
"""


def test_one_process_per_device_with_caller_passed_id(tmp_path):
    """Arrangement (b) of include/vbhem_b200.h: every process owns one device and its shard; the 128-byte NCCL id
    travels through the caller (here: a file).  Every rank ends up with the same all-reduced statistics, equal to
    the single-GPU ones."""
    two()
    from vbhem_b200.dist import shard_bounds
    world = 2
    base, hyp, cen = make_case(N=1501, K=6, Sb=5, Sr=4, d=2, covmode="full", ragged=True, seed=77, Nv=12.0)
    with api().BaseSet(base) as bs:
        want = bs.estep_hyp(hyp, 7)
    idfile = tmp_path / "id.bin"
    idfile.write_bytes(api().MultiBaseSet.unique_id())
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), str(idfile), str(tmp_path / f"o{r}.npz")])
             for r in range(world)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    outs = [np.load(tmp_path / f"o{r}.npz") for r in range(world)]
    assert np.array_equal(outs[0]["buf"], outs[1]["buf"])
    np.testing.assert_allclose(outs[0]["buf"], want["buf"], rtol=1e-11, atol=1e-11 * np.abs(want["buf"]).max())
    for r in range(world):
        lo, hi = shard_bounds(1501, world, r)
        assert np.array_equal(outs[r]["L"], want["L_elbo"][lo:hi])


def test_mex_gateways_take_ngpus(monkeypatch):
    """The MEX gateways' optional trailing `ngpus` argument (mex/vbhem_estep_fused_mex.cpp, vbhem_em_run_mex.cpp;
    matlab/*: vbhemopt.ngpus) shards the base set inside the library: same results as the one-GPU call."""
    devs = two()
    from oracle import cref
    from vbhem_b200 import pack, synth
    base, hyp, cen = make_case(N=700, K=5, Sb=4, Sr=4, d=2, covmode="full", ragged=True, seed=21, Nv=10.0)
    pb, pc = pack.pack_base(base), pack.pack_centres(cen)
    opt = dict(synth.DEFAULT_OPT, Nv=10.0, tau=5, max_iter=8, minDiff=1e-6)
    monkeypatch.delenv("VBHEM_SHIM_NGPUS", raising=False)
    one = cref.gateway_estep_fused(pb, pc, 5)
    em1 = cref.gateway_em_run(pb, hyp, opt)
    for ng in (str(len(devs)), "0"):                 # 0 = every visible GPU
        monkeypatch.setenv("VBHEM_SHIM_NGPUS", ng)
        two_ = cref.gateway_estep_fused(pb, pc, 5)
        em2 = cref.gateway_em_run(pb, hyp, opt)
        assert np.array_equal(two_["L_elbo"], one["L_elbo"]) and np.array_equal(two_["zhat"], one["zhat"])
        for k in ("Nj", "Nj_rho1", "Nj_rho2rho", "Nj_rho", "y_bar", "S_plus_C", "lb"):
            np.testing.assert_allclose(two_[k], one[k], rtol=1e-11, atol=1e-11 * np.abs(one[k]).max(), err_msg=k)
        assert em2["iters"] == em1["iters"]
        np.testing.assert_allclose(em2["LogLs"], em1["LogLs"], rtol=1e-10)


@pytest.mark.parametrize("covmode,d,kernel", [("diag", 3, "auto"), ("full", 2, "general")])
def test_multi_other_shapes_and_groups(covmode, d, kernel):
    """Diagonal covariances, the general kernel, several soft-max groups (a (K,S) grid) and a base count that leaves
    the shards uneven and smaller than one responsibility block -- sharded against single GPU."""
    devs = two()
    from vbhem_b200 import synth
    N = 37 * len(devs) + 3
    base, _ = synth.make_base_set(N, S=4, d=d, C=3, covmode=covmode, seed=9, Nv=8.0, ragged=True)
    opt = dict(synth.DEFAULT_OPT, Nv=8.0, tau=4, m0=tuple([256.0, 192.0, 128.0][:d]))
    hyp, grp, models = synth.model_selection_grid(base, Ks=range(1, 4), Ss=range(1, 4), seed=2, opt=opt)
    with api().BaseSet(base) as bs:
        bs.set_kernel(kernel)
        want = bs.estep_hyp(hyp, 4, group=grp)
    with api().MultiBaseSet(base, devices=devs) as ms:
        ms.set_kernel(kernel)
        got = ms.estep_hyp(hyp, 4, group=grp)
    assert np.array_equal(got["L_elbo"], want["L_elbo"]) and np.array_equal(got["zhat"], want["zhat"])
    close(got, want)
