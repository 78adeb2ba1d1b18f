"""GPU tier: the CUDA path through the C ABI against the oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): L_elbo rtol 1e-9; z_hat rtol 1e-9 with
atol 1e-12; per-pair statistics and reduced sums 1e-9 relative to the largest entry of the same
block.  Both kernels are exercised: the tensor-core lattice kernel ("auto") and the general
log-space kernel ("general").
"""
import glob
import os

import numpy as np
import pytest

from test_oracle import load_golden
from util import assert_pairs_close, make_case, oracle_fused, oracle_pairs

pytestmark = pytest.mark.gpu

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
              if not os.path.basename(p).startswith(("vbhmm_", "vbhem_em_")))   # E-step fixtures (others: test_fb.py, test_golden_em.py)
RTOL = 1e-9


def BaseSet(*a, **k):
    from vbhem_b200.api import BaseSet as B
    return B(*a, **k)


@pytest.mark.parametrize("kernel", ["auto", "general"])
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_pairs_match_reference_golden(path, kernel):
    base, hyp, cen, want, tau, covmode = load_golden(path)
    with BaseSet(base) as bs:
        bs.set_kernel(kernel)
        got = bs.estep_pairs(cen, tau)
        c = bs.counters()
    assert_pairs_close(got, want, rtol=RTOL, what=f"cuda({kernel}) vs reference golden")
    assert c["launches"] > 0
    if kernel == "general":
        assert c["lattice_pairs"] == 0


SHAPES = [
    # N, K, Sb, Sr, d, covmode, ragged, T
    (13, 5, 8, 8, 2, "full", False, 10),
    (13, 5, 8, 8, 2, "full", True, 10),
    (9, 4, 4, 3, 2, "diag", True, 5),
    (9, 4, 3, 4, 3, "full", True, 5),
    (9, 4, 3, 4, 3, "diag", False, 2),
    (7, 3, 1, 1, 2, "full", False, 4),
    (7, 3, 2, 1, 1, "full", True, 3),
    (6, 2, 8, 8, 1, "diag", True, 1),
    (5, 3, 6, 7, 2, "full", True, 50),
]


@pytest.mark.parametrize("kernel", ["auto", "general"])
@pytest.mark.parametrize("shape", SHAPES, ids=[f"N{s[0]}K{s[1]}b{s[2]}r{s[3]}d{s[4]}{s[5]}{'R' if s[6] else ''}T{s[7]}"
                                                for s in SHAPES])
def test_pairs_match_oracle(shape, kernel):
    N, K, Sb, Sr, d, covmode, ragged, T = shape
    base, hyp, cen = make_case(N=N, K=K, Sb=Sb, Sr=Sr, d=d, covmode=covmode, ragged=ragged, seed=100 + N + d)
    want = oracle_pairs(base, cen, T)
    with BaseSet(base) as bs:
        bs.set_kernel(kernel)
        got = bs.estep_pairs(cen, T)
        L = bs.estep_pairs(cen, T, L_only=True)["L_elbo"]
    assert_pairs_close(got, want, rtol=RTOL, what=f"cuda({kernel}) vs oracle")
    np.testing.assert_allclose(L, want["L_elbo"], rtol=RTOL)


def test_general_kernel_for_large_state_counts():
    """More than 8 states and d = 4: only the general kernel applies (automatic selection)."""
    base, hyp, cen = make_case(N=5, K=2, Sb=11, Sr=9, d=2, covmode="full", ragged=True, seed=77)
    want = oracle_pairs(base, cen, 4)
    with BaseSet(base) as bs:
        got = bs.estep_pairs(cen, 4)
        c = bs.counters()
    assert_pairs_close(got, want, rtol=RTOL)
    assert c["lattice_pairs"] == 0 and c["general_pairs"] == 10


def check_fused(got, want, N):
    np.testing.assert_allclose(got["L_elbo"], want["L_elbo"], rtol=RTOL)
    np.testing.assert_allclose(got["zhat"], want["zhat"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(got["Nj"] + 1e-50, want["Nj"], rtol=RTOL)
    # Lt1 relative; Lt7 = sum z log z is a sum of O(eps * |log_Z|) rounding residues when z_hat is
    # sharp, so it is held to the scale it enters the lower bound with (LB = Lt1 - Lt7 + ...)
    np.testing.assert_allclose(got["lb"][0::2], want["lb"][0::2], rtol=RTOL)
    np.testing.assert_allclose(got["lb"][1::2], want["lb"][1::2], rtol=RTOL, atol=1e-12 * np.abs(want["lb"][0::2]).max())
    for k in ("nu1", "xi", "pr", "mu", "Mu"):
        g, w = got[k], want[k]
        scale = np.maximum(np.max(np.abs(w), axis=tuple(range(1, w.ndim)), keepdims=True), 1e-300)
        err = np.max(np.abs(g - w) / scale)
        assert err <= RTOL, f"{k}: {err:.3e}"


@pytest.mark.parametrize("kernel", ["auto", "general"])
@pytest.mark.parametrize("Nv", [100.0, 3.0])
def test_fused_estep_matches_oracle(kernel, Nv):
    """Boundary-2: responsibilities, Z > 1e-8 skip rule, reduced sums, Lt1 / Lt7.  Nv = 100 gives the
    reference's sharp z_hat (most pairs skipped), Nv = 3 soft ones (most pairs kept)."""
    base, hyp, cen = make_case(N=300, K=6, Sb=5, Sr=4, d=2, covmode="full", ragged=True, seed=31, Nv=Nv)
    want = oracle_fused(base, cen, 6)
    with BaseSet(base) as bs:
        bs.set_kernel(kernel)
        got = bs.estep(cen, 6)
        c = bs.counters()
    check_fused(got, want, 300)
    Z = want["zhat"] * base["Ntilde"][:, None]
    assert c["active_pairs"] == int((Z > 1e-8).sum())


def test_fused_estep_from_hyperparameters():
    """Device pre-pass (digamma, log det) + fused E-step against numpy pre-pass + oracle."""
    base, hyp, cen = make_case(N=120, K=5, Sb=4, Sr=4, d=2, covmode="full", ragged=True, seed=41, Nv=30.0)
    want = oracle_fused(base, cen, 5)
    with BaseSet(base) as bs:
        got = bs.estep_hyp(hyp, 5)
    np.testing.assert_allclose(got["logOmega"], cen["logOmega"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(got["logPi"], cen["logPi"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(got["logA"], cen["logA"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(got["logLambdaTilde"], cen["logLambdaTilde"], rtol=1e-12, atol=1e-13)
    check_fused(got, want, 120)


def test_prepass_diag_d3():
    base, hyp, cen = make_case(N=20, K=3, Sb=3, Sr=3, d=3, covmode="diag", ragged=False, seed=43, Nv=30.0)
    want = oracle_fused(base, cen, 3)
    with BaseSet(base) as bs:
        got = bs.estep_hyp(hyp, 3)
    np.testing.assert_allclose(got["logLambdaTilde"], cen["logLambdaTilde"], rtol=1e-12, atol=1e-13)
    check_fused(got, want, 20)


def test_grouped_centres_one_call():
    """The (K,S) grid batches several H3Ms: per-group softmax and per-group Lt1 / Lt7."""
    base, hyp, cen = make_case(N=90, K=5, Sb=3, Sr=3, d=2, covmode="full", ragged=False, seed=47, Nv=10.0)
    cen = dict(cen)
    cen["group"] = np.array([0, 0, 1, 1, 1], np.int32)
    # per-group logOmega, as the reference would compute for two separate H3Ms
    from oracle import numpy_ref as nr
    cen["logOmega"] = np.concatenate([nr.log_omega_tilde(hyp["alpha"][:2]), nr.log_omega_tilde(hyp["alpha"][2:])])
    want = oracle_fused(base, cen, 4)
    with BaseSet(base) as bs:
        got = bs.estep(cen, 4)
    assert got["lb"].shape == (4,)
    check_fused(got, want, 90)


def test_finalize_matches_numpy():
    from oracle import numpy_ref as nr
    base, hyp, cen = make_case(N=60, K=4, Sb=3, Sr=3, d=2, covmode="full", ragged=True, seed=53, Nv=10.0,
                               sr_list=[3, 1, 2, 3])
    with BaseSet(base) as bs:
        got = bs.estep(cen, 4)
        fin = bs.finalize(got["buf"], cen["Sr"], 3)
    for j in range(4):
        r = int(cen["Sr"][j])
        a = nr.finalize_stats(got["nu1"][j, :r], got["xi"][j, :r, :r], got["pr"][j, :r], got["mu"][j, :r],
                              got["Mu"][j, :r])
        np.testing.assert_allclose(fin["Nj_rho1"][j, :r], a[0], rtol=1e-13)
        np.testing.assert_allclose(fin["Nj_rho2rho"][j, :r, :r], a[1], rtol=1e-13)
        np.testing.assert_allclose(fin["Nj_rho"][j, :r], a[2], rtol=1e-13)
        np.testing.assert_allclose(fin["y_bar"][j, :r], a[3], rtol=1e-13)
        np.testing.assert_allclose(fin["S_plus_C"][j, :r], a[4], rtol=1e-9, atol=1e-9 * np.abs(a[4]).max())
    np.testing.assert_allclose(fin["Nj"], got["Nj"] + 1e-50, rtol=1e-15)


def test_nan_and_inf_propagate_like_the_reference():
    """NaN / Inf are data (SURVEY 8b 'Errors'): they must reach L_elbo exactly where the reference puts them,
    through the in-GPU fall-back of the lattice kernel."""
    base, hyp, cen = make_case(N=6, K=3, Sb=4, Sr=4, d=2, covmode="full", ragged=False, seed=61)
    base = {k: np.array(v) for k, v in base.items()}
    cen = {k: np.array(v) for k, v in cen.items()}
    base["mu"][1, 2, 0] = np.nan            # NaN in one base emission
    cen["logA"][2, 1, :] = -np.inf          # a centre state with no way out
    cen["cterm"][0, 3] = np.inf             # a centre state nobody can emit from
    want = oracle_pairs(base, cen, 4)
    for kernel in ("auto", "general"):
        with BaseSet(base) as bs:
            bs.set_kernel(kernel)
            got = bs.estep_pairs(cen, 4)
            c = bs.counters()
        assert np.array_equal(np.isnan(got["L_elbo"]), np.isnan(want["L_elbo"])), kernel
        assert np.isnan(want["L_elbo"]).any()
        ok = ~np.isnan(want["L_elbo"])
        np.testing.assert_allclose(got["L_elbo"][ok], want["L_elbo"][ok], rtol=RTOL)
        if kernel == "auto":
            assert c["general_pairs"] > 0, "the poisoned pairs must have gone through the fall-back"


def test_underflow_falls_back_in_gpu():
    """A transition row whose only large entries point at states the emissions rule out: the
    linear-domain sum underflows, the pair is redone in log space, results still match."""
    base, hyp, cen = make_case(N=5, K=2, Sb=3, Sr=3, d=2, covmode="full", ragged=False, seed=67)
    cen = {k: np.array(v) for k, v in cen.items()}
    cen["logA"][0, 0, :] = [-0.001, -900.0, -900.0]
    cen["cterm"][0, 0] += 3000.0            # state 0 of centre 0 is ~e^-1500 less likely than the others
    want = oracle_pairs(base, cen, 5)
    with BaseSet(base) as bs:
        got = bs.estep_pairs(cen, 5)
        c = bs.counters()
    assert np.isfinite(want["L_elbo"]).all()
    assert_pairs_close(got, want, rtol=RTOL)
    assert c["general_pairs"] > 0


def test_bitwise_deterministic():
    base, hyp, cen = make_case(N=500, K=8, Sb=6, Sr=5, d=2, covmode="full", ragged=True, seed=71, Nv=8.0)
    with BaseSet(base) as bs:
        a = bs.estep(cen, 7)["buf"].copy()
        b = bs.estep(cen, 7)["buf"].copy()
    assert np.array_equal(a, b)


def test_error_codes():
    from vbhem_b200._lib import VbhemError
    base, hyp, cen = make_case(N=4, K=2, Sb=2, Sr=2, d=2, covmode="full", ragged=False, seed=73)
    with BaseSet(base) as bs:
        with pytest.raises(VbhemError) as ei:
            bs.estep_pairs(cen, 0)
        assert ei.value.code == -1 and "tau" in ei.value.msg
        bad = dict(cen)
        bad["Sr"] = np.array([2, 5], np.int32)
        with pytest.raises(VbhemError) as ei:
            bs.estep_pairs(bad, 3)
        assert ei.value.code == -1 and "vbhem_mex:invalidinput" in ei.value.msg


def test_north_star_shape_properties():
    """BASELINE configs[2] shape at a size the GPU finishes in a blink but the oracle would not:
    size-independent properties instead of element-wise comparison, plus an oracle spot check."""
    from vbhem_b200 import synth
    from util import prepass_natural
    N, K = 20000, 64
    base, hyp, _ = synth.north_star(N=N, K=K)
    cen = prepass_natural(hyp, "full")
    with BaseSet(base) as bs:
        got = bs.estep(cen, 10)
        c = bs.counters()
    z = got["zhat"]
    np.testing.assert_allclose(z.sum(1), 1.0, rtol=1e-12)                     # softmax rows
    np.testing.assert_allclose(got["Nj"].sum(), base["Ntilde"].sum(), rtol=1e-12)
    Z = z * base["Ntilde"][:, None]
    act = Z > 1e-8
    assert c["active_pairs"] == int(act.sum())
    # mass conservation through the skip rule: sum_sigma S_nu1[j] = sum_{i active} Z_ij * sum(prior_i)
    np.testing.assert_allclose(got["nu1"].sum(1), (np.where(act, Z, 0.0) * base["prior"].sum(1)[:, None]).sum(0),
                               rtol=1e-10)
    np.testing.assert_allclose(got["xi"].sum((1, 2)) + got["nu1"].sum(1), got["pr"].sum(1), rtol=1e-10)
    np.testing.assert_allclose(got["lb"][0], np.sum(Z * got["L_elbo"]), rtol=1e-12)
    # spot check 40 bases against the oracle
    idx = np.arange(0, N, N // 40)[:40]
    sub = {k: v[idx] for k, v in base.items()}
    want = oracle_pairs(sub, cen, 10)
    np.testing.assert_allclose(got["L_elbo"][idx], want["L_elbo"], rtol=RTOL)


def test_sharded_matches_single():
    """Two shards on one GPU, raw buffers summed on the host (what NCCL does across GPUs), then normalised:
    equal to the un-sharded E-step up to fp64 summation order."""
    from vbhem_b200.dist import shard_base
    base, hyp, cen = make_case(N=401, K=6, Sb=5, Sr=4, d=2, covmode="full", ragged=True, seed=83, Nv=9.0)
    with BaseSet(base) as bs:
        one = bs.estep(cen, 6)
        fin_one = bs.finalize(one["buf"], cen["Sr"], 4)
        bufs = []
        for r in range(2):
            with BaseSet(shard_base(base, 2, r)) as sh:
                bufs.append(sh.estep(cen, 6, want_pairs=False)["buf"])
        fin_two = bs.finalize(bufs[0] + bufs[1], cen["Sr"], 4)
    for k in fin_one:
        np.testing.assert_allclose(fin_two[k], fin_one[k], rtol=1e-11, atol=1e-11 * np.abs(fin_one[k]).max(), err_msg=k)


def test_repack_same_shape():
    base, hyp, cen = make_case(N=50, K=3, Sb=4, Sr=3, d=2, covmode="full", ragged=True, seed=85)
    base2, _, _ = make_case(N=50, K=3, Sb=4, Sr=3, d=2, covmode="full", ragged=True, seed=86)
    with BaseSet(base) as bs:
        bs.estep_pairs(cen, 4)
        bs.repack(base2)
        got = bs.estep_pairs(cen, 4)
    assert_pairs_close(got, oracle_pairs(base2, cen, 4), rtol=RTOL)


@pytest.mark.parametrize("nslices", [1, 3, 8])
def test_streamed_estep_is_bit_identical_to_resident(nslices):
    """vbhem_estep_hyp_stream: the base set arrives from host memory in slices (multiples of the
    256-base responsibility block) while earlier slices compute; every result must be bit-identical
    to vbhem_repack_base + vbhem_estep_hyp, and match the oracle."""
    base, hyp, cen = make_case(N=1500, K=5, Sb=4, Sr=4, d=2, covmode="full", ragged=True, seed=91, Nv=30.0)
    other, _, _ = make_case(N=1500, K=5, Sb=4, Sr=4, d=2, covmode="full", ragged=True, seed=92, Nv=30.0)
    with BaseSet(other) as bs:                      # the handle starts with a DIFFERENT base set of the same shape
        bs.estep_hyp(hyp, 5)
        got = bs.estep_hyp(hyp, 5, stream_base=base, nslices=nslices)
        again = bs.estep_hyp(hyp, 5)                # the streamed set is now resident
    with BaseSet(base) as bs:
        ref = bs.estep_hyp(hyp, 5)
    for k in ("L_elbo", "zhat", "buf"):
        assert np.array_equal(got[k], ref[k]), k
        assert np.array_equal(again[k], ref[k]), k
    check_fused(got, oracle_fused(base, cen, 5), 1500)


def test_streamed_estep_general_kernel_shapes():
    """Shapes the lattice kernel does not take (here 10 base states) go through upload + resident path."""
    base, hyp, cen = make_case(N=40, K=3, Sb=10, Sr=3, d=2, covmode="full", ragged=True, seed=93, Nv=20.0)
    with BaseSet(base) as bs:
        ref = bs.estep_hyp(hyp, 4)
        got = bs.estep_hyp(hyp, 4, stream_base=base, nslices=4)
    for k in ("L_elbo", "zhat", "buf"):
        assert np.array_equal(got[k], ref[k]), k


def test_set_stream_orders_with_torch_default_stream():
    """A handle bound to torch's current (default) stream must be ordered with torch work on that stream:
    the statistics buffer is consumed by a torch op right after vbhem_estep_hyp_dev without a synchronise."""
    import ctypes as C
    import torch
    from vbhem_b200._lib import check, dp, ip, lib
    base, hyp, cen = make_case(N=4000, K=6, Sb=4, Sr=4, d=2, covmode="full", ragged=False, seed=19, Nv=30.0)
    with BaseSet(base) as bs:
        want = bs.estep_hyp(hyp, 6, want_pairs=False)["buf"]
        bs.set_stream(torch.cuda.current_stream().cuda_stream)
        K, r = hyp["eta"].shape
        c = np.ascontiguousarray
        Sr = c(hyp["Sr"], dtype=np.int32)
        H = [c(hyp["alpha"]), c(hyp["eta"]), c(np.transpose(hyp["epsilon"], (0, 2, 1))), c(hyp["m"]),
             c(np.transpose(hyp["W"], (0, 1, 3, 2))), c(hyp["v"]), c(hyp["lam"])]
        for _ in range(5):
            stats = torch.full((bs.stats_size(K, r),), float("nan"), dtype=torch.float64, device="cuda")
            check(lib().vbhem_estep_hyp_dev(bs._h, K, Sr.ctypes.data_as(ip), r, None, 1, 6,
                                            *[h.ctypes.data_as(dp) for h in H], C.c_void_p(stats.data_ptr()), None))
            got = (stats * 1.0).cpu().numpy()          # torch op on the same stream, no explicit synchronise
            assert np.array_equal(got, want)


def test_model_selection_grid_batched_equals_one_model_at_a_time():
    """BASELINE configs[3]: every (K,S) candidate's centres stacked in ONE fused E-step (one soft-max group per
    model, ragged state counts) must give each model what a call of its own gives."""
    from vbhem_b200 import synth
    base, _ = synth.make_base_set(300, S=4, d=2, C=3, covmode="full", seed=23, Nv=20.0)
    hyp, grp, models = synth.model_selection_grid(base, Ks=range(1, 4), Ss=range(1, 5), seed=23,
                                                  opt=dict(synth.DEFAULT_OPT, Nv=20.0))
    with BaseSet(base) as bs:
        allg = bs.estep_hyp(hyp, 5, group=grp)
        for g, (K, S) in enumerate(models):
            sel = np.where(grp == g)[0]
            one = bs.estep_hyp({k: v[sel] for k, v in hyp.items()}, 5)
            np.testing.assert_array_equal(allg["L_elbo"][:, sel], one["L_elbo"])
            np.testing.assert_array_equal(allg["zhat"][:, sel], one["zhat"])
            np.testing.assert_allclose(allg["Nj"][sel], one["Nj"], rtol=1e-13)
            np.testing.assert_allclose(allg["lb"][2 * g:2 * g + 2], one["lb"], rtol=1e-12)
            for k in ("nu1", "xi", "pr", "mu", "Mu"):
                np.testing.assert_allclose(allg[k][sel], one[k], rtol=1e-12, atol=1e-12 * np.abs(one[k]).max(), err_msg=k)


@pytest.mark.parametrize("sr_list", [[2, 8, 3, 1, 4, 6, 4], [3, 3], [1, 2, 3], [4, 5, 8]])
def test_centre_pairing_is_bit_identical_to_one_centre_per_tile(sr_list):
    """First pass: centres of <= 4 states are packed two per 8-state tile (block-diagonal transition weights,
    one row shift per half).  The packing adds exact zeros only, so L_elbo -- and with it everything
    downstream -- must be bit-identical to the one-centre-per-tile kernel, and both must match the oracle.
    Odd numbers of small centres exercise the partner-less tile."""
    K = len(sr_list)
    base, hyp, cen = make_case(N=200, K=K, Sb=8, Sr=8, d=2, covmode="full", ragged=True, seed=57, Nv=25.0,
                               sr_list=np.array(sr_list))
    with BaseSet(base) as bs:
        bs.set_kernel("halves")
        got = bs.estep(cen, 6)
        cp = bs.counters()
        bs.set_kernel("unpaired")
        ref = bs.estep(cen, 6)
    for k in ("L_elbo", "zhat", "buf"):
        assert np.array_equal(got[k], ref[k]), k
    assert cp["general_pairs"] == 0 and cp["split_tiles"] == 0
    check_fused(got, oracle_fused(base, cen, 6), 200)


@pytest.mark.parametrize("sr_list,tiles,split", [
    ([2, 8, 3, 1, 4, 6, 4], 4, 1),          # 8 | 6+2 | 4+4 | 3+1
    ([7, 1, 6, 2, 5, 3], 3, 3),             # every split class once
    ([5, 1, 5, 2, 7, 3, 6, 1, 1], 5, 4),    # 7+1 | 6+2 | 5+3 | 5+1 | 1+none: slots left empty beside a partner
    ([8, 8, 5, 7], 4, 0),                   # nothing to pair with
    ([6, 6, 2, 1], 2, 2)])
@pytest.mark.parametrize("covmode,d", [("full", 2), ("diag", 3)])
def test_split_tiles_match_one_centre_per_tile(sr_list, tiles, split, covmode, d):
    """First pass: a centre of 5..7 states shares its tile with one of <= 3..1 states (columns [0,k) and [k,8),
    two shifts per row).  The columns move, so sums are re-associated: equal to the one-centre-per-tile kernel
    to a few ulp (1e-13 here), equal to the oracle at the parity tolerance, and run-to-run identical."""
    K = len(sr_list)
    base, hyp, cen = make_case(N=150, K=K, Sb=8, Sr=8, d=d, covmode=covmode, ragged=True, seed=61, Nv=25.0,
                               sr_list=np.array(sr_list))
    with BaseSet(base) as bs:
        got = bs.estep(cen, 7)
        cp = bs.counters()
        again = bs.estep(cen, 7)
        bs.set_kernel("unpaired")
        ref = bs.estep(cen, 7)
    assert (cp["tiles"], cp["split_tiles"], cp["general_pairs"]) == (tiles, split, 0)
    for k in ("L_elbo", "zhat", "buf"):
        assert np.array_equal(got[k], again[k]), k
    np.testing.assert_allclose(got["L_elbo"], ref["L_elbo"], rtol=1e-13, atol=0)
    check_fused(got, oracle_fused(base, cen, 7), 150)


def test_estep_dev_returns_without_synchronising():
    """include/vbhem_b200.h: vbhem_estep_hyp_dev / vbhem_em_estep enqueue the whole E-step -- including the second
    pass, whose size (the pairs with Z > 1e-8) is known on the device only -- and return.  A long spin kernel is put
    on the stream first: if any call in the E-step waited for the stream, the spin would be over when it returns."""
    import ctypes as C
    import torch
    from vbhem_b200._lib import check, dp, ip, lib
    base, hyp, cen = make_case(N=3000, K=6, Sb=4, Sr=4, d=2, covmode="full", ragged=False, seed=23, Nv=30.0)
    with BaseSet(base) as bs:
        want = bs.estep_hyp(hyp, 6, want_pairs=False)["buf"]
        st = torch.cuda.current_stream()
        bs.set_stream(st.cuda_stream)
        K, r = hyp["eta"].shape
        c = np.ascontiguousarray
        Sr = c(hyp["Sr"], dtype=np.int32)
        H = [c(hyp["alpha"]), c(hyp["eta"]), c(np.transpose(hyp["epsilon"], (0, 2, 1))), c(hyp["m"]),
             c(np.transpose(hyp["W"], (0, 1, 3, 2))), c(hyp["v"]), c(hyp["lam"])]
        stats = torch.zeros(bs.stats_size(K, r), dtype=torch.float64, device="cuda")
        for _ in range(2):                                # warm: allocations of the first call may synchronise
            check(lib().vbhem_estep_hyp_dev(bs._h, K, Sr.ctypes.data_as(ip), r, None, 1, 6,
                                            *[h.ctypes.data_as(dp) for h in H], C.c_void_p(stats.data_ptr()), None))
        torch.cuda.synchronize()
        torch.cuda._sleep(int(2e9))                       # ~1 s of spinning on the stream
        check(lib().vbhem_estep_hyp_dev(bs._h, K, Sr.ctypes.data_as(ip), r, None, 1, 6,
                                        *[h.ctypes.data_as(dp) for h in H], C.c_void_p(stats.data_ptr()), None))
        assert not st.query(), "the E-step waited for its stream"
        torch.cuda.synchronize()
        assert np.array_equal(stats.cpu().numpy(), want)
        assert bs.counters()["active_pairs"] > 0          # read back on demand


def test_em_state_survives_other_esteps_on_the_handle():
    """the device-resident EM run keeps its own hyper-parameters / state counts; an ordinary E-step with a
    different centre set in between must neither corrupt it nor be silently mixed into its lower bound."""
    from vbhem_b200 import synth
    from vbhem_b200._lib import VbhemError
    base, hyp, cen = make_case(N=300, K=6, Sb=5, Sr=4, d=2, covmode="full", ragged=True, seed=31, Nv=20.0)
    _, hyp2, _ = make_case(N=300, K=3, Sb=5, Sr=3, d=2, covmode="full", ragged=False, seed=32, Nv=20.0)
    opt = dict(synth.DEFAULT_OPT, Nv=20.0, tau=5)
    with BaseSet(base) as bs:
        bs.em_begin(hyp, opt)
        want = bs.em_run(5, 3, 0.0)
        bs.em_begin(hyp, opt)
        bs.em_estep(5)
        bs.estep_hyp(hyp2, 5)                             # another centre set (K = 3) on the same handle
        with pytest.raises(VbhemError):
            bs.em_lower_bound()                           # its pre-pass is gone: refused, not wrong
        got = bs.em_run(5, 3, 0.0)                        # the run itself is intact
    np.testing.assert_array_equal(got["LogLs"], want["LogLs"])
    for k in ("alpha", "eta", "m", "W"):
        np.testing.assert_array_equal(got["hyp"][k], want["hyp"][k])


def _oracle_chunk(args):
    """Worker of test_full_size_parity_report: the oracle's per-pair outputs for a chunk of bases (packed)."""
    pbc, pc, tau = args
    from oracle import cref
    return cref.port_estep_pairs(pbc, pc, tau)


def test_full_size_parity_report():
    """at the BASELINE shape (N = 100 000 x K = 64, S = 8, tau = 10) compare 2 048 bases x 64
    centres ELEMENT-WISE with the oracle -- L_elbo and z_hat from the full run, the reduced statistics and the
    N-sized lower-bound terms from a fused E-step over the same sample -- and print the observed errors.
    Nv = 100 multiplies every L_elbo error by 100 before the soft-max, which is what makes z_hat the sensitive one."""
    import os
    from concurrent.futures import ProcessPoolExecutor
    from oracle import cref
    from vbhem_b200 import pack, synth
    from util import prepass_natural
    N, K, tau = 100_000, 64, 10
    base, hyp, _ = synth.north_star(N=N, K=K)
    cen = prepass_natural(hyp, "full")
    pb, pc = pack.pack_base(base), pack.pack_centres(cen)
    idx = np.arange(0, N, N // 2048)[:2048]
    with BaseSet(pb, packed=True) as bs:
        full = bs.estep(pc, tau, packed=True, raw=True)               # packed [K, N] rows
    sub = {k: (np.ascontiguousarray(v[idx]) if isinstance(v, np.ndarray) and v.shape[:1] == (N,) else v)
           for k, v in pb.items()}
    sub["N"] = len(idx)
    nw = min(os.cpu_count() or 1, 16)
    cuts = np.linspace(0, len(idx), nw + 1).astype(int)
    chunks = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        c = {k: (np.ascontiguousarray(v[a:b]) if isinstance(v, np.ndarray) and v.shape[:1] == (len(idx),) else v)
             for k, v in sub.items()}
        c["N"] = int(b - a)
        chunks.append((c, pc, tau))
    with ProcessPoolExecutor(nw) as ex:
        outs = list(ex.map(_oracle_chunk, chunks))
    o = {k: np.concatenate([x[k] for x in outs], axis=1) for k in outs[0]}
    zz = cref.port_zhat(o["L_elbo"], pc["logOmega"], sub["Ntilde"])
    red = cref.port_reduce(zz["Z"], o, pc["maxSr"], 2, 1)
    want = pack.stats_to_natural(dict(red, Nj=zz["Nj"], lb=zz["lb"]), 2, 1)
    with BaseSet(sub, packed=True) as bs:
        got = bs.estep(pc, tau, packed=True)
    eL = np.max(np.abs(full["L_elbo"][:, idx] - o["L_elbo"]) / np.abs(o["L_elbo"]))
    ez_abs = np.max(np.abs(full["zhat"][:, idx] - zz["zhat"]))
    ez_rel = np.max(np.abs(full["zhat"][:, idx] - zz["zhat"]) / np.maximum(zz["zhat"], 1e-12))
    es = {}
    for k in ("nu1", "xi", "pr", "mu", "Mu"):
        g, w = got[k], want[k]
        scale = np.maximum(np.max(np.abs(w), axis=tuple(range(1, w.ndim)), keepdims=True), 1e-300)
        es[k] = float(np.max(np.abs(g - w) / scale))
    eNj = float(np.max(np.abs(got["Nj"] + 1e-50 - want["Nj"])) / np.max(want["Nj"]))
    eLt1 = float(abs(got["lb"][0] - want["lb"][0]) / abs(want["lb"][0]))
    print(f"\nfull-size parity (2048 of 100000 bases x 64 centres): L_elbo rel {eL:.2e}, zhat abs {ez_abs:.2e} "
          f"rel(>=1e-12) {ez_rel:.2e}, Nj rel {eNj:.2e}, Lt1 rel {eLt1:.2e}, statistics {es}")
    assert eL <= RTOL
    np.testing.assert_allclose(full["zhat"][:, idx], zz["zhat"], rtol=RTOL, atol=1e-12)
    assert eNj <= RTOL and eLt1 <= RTOL and max(es.values()) <= RTOL
