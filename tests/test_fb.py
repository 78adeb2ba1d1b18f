"""VB-HMM forward-backward: the reference's src/hmm/vbhmm_fb_mex.c compiled in place, the
plain-C restatement (oracle/fb_oracle.c) and the numpy restatement of vbhmm_fb.m on CPU; the CUDA path
(vbhmm_fb through the C ABI) against them on the GPU.  Tolerance 1e-9, the authors' own cross-check bound
(vbhmm_fb.m:186-190)."""
import os

import numpy as np
import pytest

from oracle import cref, numpy_ref as nr

GOLD = os.path.join(os.path.dirname(__file__), "golden", "vbhmm_fb_ragged_d2_k4.npz")


def make_fb_case(N=12, K=4, dim=2, maxT=15, seed=3, ragged=True, with_empty=True):
    rng = np.random.default_rng(seed)
    centre = np.resize(np.array([256.0, 192.0, 128.0]), dim)
    lens = rng.integers(1, maxT + 1, size=N) if ragged else np.full(N, maxT)
    lens[0] = maxT
    if ragged and N > 2:
        lens[1] = 1
        if with_empty:
            lens[2] = 0
    means = rng.normal(size=(K, dim)) * 60 + centre
    data = []
    for T in lens:
        z = rng.integers(0, K, size=T)
        data.append(means[z] + rng.normal(size=(T, dim)) * 15)
    W = np.stack([np.linalg.inv(np.cov(rng.normal(size=(dim, 30)) * 20).reshape(dim, dim)) / 8 for _ in range(K)])
    vp = dict(v=5 + rng.uniform(size=K) * 10, W=W, epsilon=1 + rng.uniform(size=(K, K)) * 5,
              alpha=1 + rng.uniform(size=K) * 3, m=means + rng.normal(size=(K, dim)) * 5, beta=1 + rng.uniform(size=K) * 5)
    return data, vp


def assert_fb_close(got, want, rtol=1e-9):
    np.testing.assert_allclose(got["logrho_Saved"], want["logrho_Saved"], rtol=rtol, atol=1e-12)
    np.testing.assert_allclose(got["gamma_all"], want["gamma_all"], rtol=rtol, atol=1e-13)
    np.testing.assert_allclose(got["xi_sum"], want["xi_sum"], rtol=rtol, atol=1e-13)
    np.testing.assert_allclose(got["phi_norm"], want["phi_norm"], rtol=rtol, atol=1e-12)


@pytest.mark.parametrize("K,dim", [(4, 2), (1, 1), (3, 3), (8, 2)])
def test_port_and_numpy_match_reference_build(K, dim):
    if not cref.have_fb_ref():
        pytest.skip("reference build not present")
    data, vp = make_fb_case(N=9, K=K, dim=dim, seed=10 + K)
    par = nr.vbhmm_fb_prepass(vp)
    ps = cref.pack_sequences(data)
    ref, port, npy = cref.ref_vbhmm_fb(ps, par), cref.port_vbhmm_fb(ps, par), nr.vbhmm_fb(data, par)
    assert np.array_equal(port["logrho_Saved"], ref["logrho_Saved"])
    assert np.array_equal(port["gamma_all"], ref["gamma_all"])
    assert_fb_close(port, ref, rtol=1e-13)
    assert_fb_close(npy, ref, rtol=1e-11)
    # forward-backward invariants: responsibilities sum to one, pair marginals to T - 1
    T = np.array([len(x) for x in data])
    for n in range(len(data)):
        np.testing.assert_allclose(ref["gamma_all"][n, :T[n]].sum(1), 1.0, rtol=1e-12)
        np.testing.assert_allclose(ref["xi_sum"][n].sum(), max(T[n] - 1, 0), rtol=1e-12, atol=1e-13)
        assert not ref["gamma_all"][n, T[n]:].any() and not ref["logrho_Saved"][n, T[n]:].any()


def test_golden_fixture_pins_the_port():
    """tests/golden/vbhmm_fb_*.npz was generated from the reference build (tests/golden/make_golden.py)."""
    g = np.load(GOLD)
    data = [g[f"seq{n}"] for n in range(int(g["N"]))]
    vp = {k: g[k] for k in ("v", "W", "epsilon", "alpha", "m", "beta")}
    par = nr.vbhmm_fb_prepass(vp)
    port = cref.port_vbhmm_fb(cref.pack_sequences(data), par)
    assert_fb_close(port, {k: g[k] for k in ("logrho_Saved", "gamma_all", "xi_sum", "phi_norm")}, rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("K,dim,N,maxT", [(4, 2, 40, 15), (1, 1, 7, 5), (5, 2, 300, 50), (8, 3, 33, 9), (3, 2, 5, 1),
                                           (11, 2, 21, 8), (4, 5, 19, 6)])
def test_cuda_fb_matches_oracle(K, dim, N, maxT):
    from vbhem_b200 import fb
    data, vp = make_fb_case(N=N, K=K, dim=dim, maxT=maxT, seed=100 + K + dim)
    par = nr.vbhmm_fb_prepass(vp)
    want = cref.port_vbhmm_fb(cref.pack_sequences(data), par)
    got = fb.vbhmm_fb(data, vp)
    assert_fb_close(got, want)
    np.testing.assert_allclose(got["logLambdaTilde"], par["logLambdaTilde"], rtol=1e-13)
    np.testing.assert_allclose(got["logATilde"], par["logATilde"], rtol=1e-13)


@pytest.mark.gpu
def test_cuda_fb_golden_and_resident_handle():
    from vbhem_b200 import fb
    g = np.load(GOLD)
    data = [g[f"seq{n}"] for n in range(int(g["N"]))]
    vp = {k: g[k] for k in ("v", "W", "epsilon", "alpha", "m", "beta")}
    with fb.SequenceSet(data) as seqs:
        a = fb.vbhmm_fb(seqs, vp)
        vp2 = dict(vp, m=vp["m"] + 3.0)
        fb.vbhmm_fb(seqs, vp2)                      # another model on the resident data
        b = fb.vbhmm_fb(seqs, vp)
    assert_fb_close(a, {k: g[k] for k in ("logrho_Saved", "gamma_all", "xi_sum", "phi_norm")})
    for k in ("logrho_Saved", "gamma_all", "xi_sum", "phi_norm"):
        assert np.array_equal(a[k], b[k]), k       # bit-identical run to run


@pytest.mark.gpu
def test_cuda_fb_long_sequences_stay_normalised():
    """Size-independent properties at a length where an unscaled recursion would underflow."""
    from vbhem_b200 import fb
    data, vp = make_fb_case(N=64, K=5, dim=2, maxT=4000, seed=7, ragged=False)
    out = fb.vbhmm_fb(data, vp)
    assert np.isfinite(out["phi_norm"]).all()
    np.testing.assert_allclose(out["gamma_all"].sum(2), 1.0, rtol=1e-10)
    np.testing.assert_allclose(out["xi_sum"].sum((1, 2)), 3999.0, rtol=1e-10)
    # xi marginals agree with gamma: sum_to xi(from, to) = sum_{t < T} gamma(t, from)
    np.testing.assert_allclose(out["xi_sum"].sum(2), out["gamma_all"][:, :-1].sum(1), rtol=1e-9)


@pytest.mark.gpu
def test_cuda_fb_error_codes():
    import ctypes as C
    from vbhem_b200 import fb
    from vbhem_b200._lib import VbhemError, lib
    data, vp = make_fb_case(N=4, K=2, dim=2, maxT=4, seed=1)
    with fb.SequenceSet(data) as seqs:
        par = fb.prepass(vp)
        par_bad = dict(par, v=np.zeros(40), beta=np.ones(40), logLambdaTilde=np.zeros(40), pz1=np.ones(40),
                       m=np.zeros((40, 2)), W=np.zeros((40, 2, 2)), A=np.zeros((40, 40)))
        with pytest.raises(VbhemError) as ei:
            fb.fb_call(seqs, par_bad)
        assert ei.value.code == -1 and "vbhmm_fb_mex:invalidinput" in ei.value.msg
    h = C.c_void_p()
    assert lib().vbhmm_pack_data(0, 0, 2, None, None, None, C.byref(h)) == -1


needs_fb_gateway = pytest.mark.skipif(not os.path.exists(cref.FB_GATEWAY_SO), reason="fb gateway not built")


@needs_fb_gateway
def test_fb_gateway_argument_checks_need_no_gpu():
    """Checks fire before any CUDA call: a sequence longer than the caller's maxT is refused with the
    reference's error id family (the reference itself would write out of bounds there)."""
    data, vp = make_fb_case(N=4, K=3, dim=2, maxT=5, seed=2)
    par = nr.vbhmm_fb_prepass(vp)
    ps = cref.pack_sequences(data)
    with pytest.raises(RuntimeError):
        cref._fb("fb_gateway", dict(ps, maxT=ps["maxT"] - 1), par)


@needs_fb_gateway
@pytest.mark.gpu
def test_fb_gateway_matches_reference_golden_and_caches_data():
    g = np.load(GOLD)
    data = [g[f"seq{n}"] for n in range(int(g["N"]))]
    vp = {k: g[k] for k in ("v", "W", "epsilon", "alpha", "m", "beta")}
    par = nr.vbhmm_fb_prepass(vp)
    ps = cref.pack_sequences(data)
    want = {k: g[k] for k in ("logrho_Saved", "gamma_all", "xi_sum", "phi_norm")}
    assert_fb_close(cref.gateway_vbhmm_fb(ps, par), want)
    # a second model on the same data, then other data of another shape
    par2 = nr.vbhmm_fb_prepass(dict(vp, m=vp["m"] - 2.0))
    assert_fb_close(cref.gateway_vbhmm_fb(ps, par2), cref.port_vbhmm_fb(ps, par2))
    data3, vp3 = make_fb_case(N=20, K=2, dim=3, maxT=7, seed=44)
    ps3, par3 = cref.pack_sequences(data3), nr.vbhmm_fb_prepass(vp3)
    assert_fb_close(cref.gateway_vbhmm_fb(ps3, par3), cref.port_vbhmm_fb(ps3, par3))


@pytest.mark.gpu
@pytest.mark.parametrize("K,dim,N,maxT", [(5, 2, 400, 50), (1, 1, 9, 4), (8, 3, 70, 12), (3, 2, 6, 1)])
def test_cuda_fb_stats_match_vbhmm_em_sums(K, dim, N, maxT):
    """vbhmm_fb_stats (forward-backward fused with the sums of vbhmm_em.m:158-246) against those sums taken
    from the oracle's per-step outputs."""
    from vbhem_b200 import fb
    data, vp = make_fb_case(N=N, K=K, dim=dim, maxT=maxT, seed=200 + K + dim)
    want = nr.vbhmm_em_stats(data, cref.port_vbhmm_fb(cref.pack_sequences(data), nr.vbhmm_fb_prepass(vp)))
    with fb.SequenceSet(data) as seqs:
        got = fb.fb_stats(seqs, vp)
        again = fb.fb_stats(seqs, vp)
        lit = fb.vbhmm_fb(seqs, vp)                       # the literal pass still works on the same handle afterwards
    for k in ("Nk1", "Nk", "M", "xbar", "Lt63", "Lt64"):
        np.testing.assert_allclose(got[k], want[k], rtol=1e-9, atol=1e-12, err_msg=k)
        assert np.array_equal(got[k], again[k]), k         # fixed-order sums
    np.testing.assert_allclose(got["S"], want["S"], rtol=1e-9, atol=1e-9 * np.abs(want["S"]).max())
    np.testing.assert_allclose(lit["gamma_all"].sum((0, 1)) + 1e-50, got["Nk"], rtol=1e-12)


@pytest.mark.gpu
def test_cuda_fb_stats_rejects_generic_shapes():
    from vbhem_b200 import fb
    from vbhem_b200._lib import VbhemError
    data, vp = make_fb_case(N=5, K=11, dim=2, maxT=4, seed=5)
    with fb.SequenceSet(data) as seqs, pytest.raises(VbhemError) as ei:
        fb.fb_stats(seqs, vp)
    assert ei.value.code == -1


@pytest.mark.gpu
@pytest.mark.parametrize("ragged", [True, False])
def test_cuda_fb_many_chunks_elementwise(ragged):
    """More sequences than the persistent grid has threads (every warp takes several chunks from the device counter),
    lengths beyond the shared-memory ring (the parked alpha_hat goes through the L2 scratch and cp.async), ragged warps
    and uniform ones: a sample of sequences element-wise against the plain-C oracle, which is per sequence."""
    from vbhem_b200 import fb
    N, K, dim, maxT = 90_000, 5, 2, 41
    rng = np.random.default_rng(77)
    lens = rng.integers(0, maxT + 1, size=N) if ragged else np.full(N, maxT)
    lens[0] = maxT
    means = rng.normal(size=(K, dim)) * 60 + np.array([256.0, 192.0])
    z = rng.integers(0, K, size=(N, maxT))
    Xall = means[z] + rng.normal(size=(N, maxT, dim)) * 15
    data = [Xall[n, :lens[n]] for n in range(N)]
    _, vp = make_fb_case(N=4, K=K, dim=dim, maxT=5, seed=78)
    vp["m"] = means + rng.normal(size=(K, dim)) * 5
    got = fb.vbhmm_fb(data, vp)
    pick = np.unique(np.concatenate([[0, 1, 31, 32, N - 1, N - 2], rng.integers(0, N, size=600),
                                     np.arange(37_888 - 40, 37_888 + 40)]))      # around the grid's first full round
    want = cref.port_vbhmm_fb(cref.pack_sequences([data[n] for n in pick]), nr.vbhmm_fb_prepass(vp))
    sub = dict(logrho_Saved=got["logrho_Saved"][pick], gamma_all=got["gamma_all"][pick], xi_sum=got["xi_sum"][pick],
               phi_norm=got["phi_norm"][pick])
    # the sample's own maxT may be shorter than the full set's: compare the common steps, the rest must be zeros
    Ts = want["gamma_all"].shape[1]
    for k in ("logrho_Saved", "gamma_all"):
        assert not sub[k][:, Ts:].any(), k
        sub[k] = sub[k][:, :Ts]
    assert_fb_close(sub, want)
    g = got["gamma_all"]
    live = np.arange(maxT)[None, :] < lens[:, None]
    np.testing.assert_allclose(g.sum(2)[live], 1.0, rtol=1e-11)
    assert not g[~live].any() and not got["logrho_Saved"][~live].any()
    np.testing.assert_allclose(got["xi_sum"].sum((1, 2)), np.maximum(lens - 1, 0), rtol=1e-11, atol=1e-13)


@pytest.mark.gpu
def test_cuda_fb_full_size_properties():
    """BASELINE configs[4] scale (200 k of the 1 M sequences, T = 50, S = 5, d = 2): size-independent properties
    of the literal outputs, and the fused statistics against sums of those outputs."""
    from vbhem_b200 import fb
    import bench_fb
    N, T, K, dim = 200_000, 50, 5, 2
    Tn, offs, flat, vp = bench_fb.make_workload(N, T, K, dim, seed=5)
    with fb.SequenceSet.from_packed(Tn, offs, flat, dim) as seqs:
        par = fb.prepass(vp)
        out = fb.fb_call(seqs, par, want=("gamma_all", "xi_sum", "phi_norm"))
        st = fb.fb_stats(seqs, vp)
    g = out["gamma_all"]
    assert np.isfinite(out["phi_norm"]).all()
    np.testing.assert_allclose(g.sum(2), 1.0, rtol=1e-11)
    np.testing.assert_allclose(out["xi_sum"].sum((1, 2)), T - 1.0, rtol=1e-11)
    np.testing.assert_allclose(out["xi_sum"].sum(2), g[:, :-1].sum(1), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st["Nk"], g.sum((0, 1)) + 1e-50, rtol=1e-11)
    np.testing.assert_allclose(st["Nk1"], g[:, 0].sum(0) + 1e-50, rtol=1e-11)
    np.testing.assert_allclose(st["M"], out["xi_sum"].sum(0), rtol=1e-11)
    np.testing.assert_allclose(st["Lt64"], out["phi_norm"].sum(), rtol=1e-12)
    np.testing.assert_allclose(st["Nk"].sum(), N * T, rtol=1e-12)
    X = np.transpose(flat.reshape(N, dim, T), (0, 2, 1))
    xbar = np.einsum("ntk,ntd->kd", g, X) / st["Nk"][:, None]
    np.testing.assert_allclose(st["xbar"], xbar, rtol=1e-10)
