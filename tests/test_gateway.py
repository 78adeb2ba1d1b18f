"""The MEX gateway of this repo (mex/vbhem_hmm_bwd_fwd_mex.cpp) behind the same mx shim and flat driver as
the reference's gateway: identical call, identical cell/struct arguments, outputs compared with the reference
golden vectors.  The error-id checks need no GPU (they fire before any CUDA call)."""
import glob
import os

import numpy as np
import pytest

from oracle import cref
from vbhem_b200 import pack
from test_oracle import load_golden
from util import assert_pairs_close

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
              if not os.path.basename(p).startswith(("vbhmm_", "vbhem_em_")))   # E-step fixtures (others: test_fb.py, test_golden_em.py)
needs_gateway = pytest.mark.skipif(not os.path.exists(cref.GATEWAY_SO), reason="gateway not built")


@needs_gateway
def test_gateway_error_ids_match_reference():
    """/root/reference/src/vbhem/vbhem_hmm_bwd_fwd_mex.c:335-366."""
    assert cref.ref_call_malformed(0, "gateway") == (1, "MyToolbox:arrayProduct:nrhs")
    assert cref.ref_call_malformed(1, "gateway") == (1, "MyToolbox:arrayProduct:nlhs")
    assert cref.ref_call_malformed(2, "gateway") == (1, "vbhem_mex:invalidinput")
    assert cref.ref_call_malformed(3, "gateway") == (1, "vbhem_mex:invalidinput")


@needs_gateway
@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_gateway_matches_reference_golden(path):
    base, hyp, cen, want, tau, covmode = load_golden(path)
    pb, pc = pack.pack_base(base), pack.pack_centres(cen)
    got = pack.pairs_to_natural(cref.ref_estep_pairs(pb, pc, tau, which="gateway"), pb["d"], pb["covmode"])
    assert_pairs_close(got, want, rtol=1e-9, what="MEX gateway (CUDA) vs reference golden")


@needs_gateway
@pytest.mark.gpu
def test_gateway_handle_reuse_and_reshape():
    """Consecutive calls: same shape re-uploads into the cached handle, a new shape re-packs."""
    from util import make_case, oracle_pairs
    for seed, N, Sb in ((1, 6, 3), (2, 6, 3), (3, 9, 4)):
        base, hyp, cen = make_case(N=N, K=3, Sb=Sb, Sr=3, d=2, covmode="full", ragged=True, seed=seed)
        pb, pc = pack.pack_base(base), pack.pack_centres(cen)
        got = pack.pairs_to_natural(cref.ref_estep_pairs(pb, pc, 4, which="gateway"), 2, 1)
        assert_pairs_close(got, oracle_pairs(base, cen, 4), rtol=1e-9)


needs_fused = pytest.mark.skipif(not os.path.exists(cref.FUSED_SO), reason="fused gateway not built")


@needs_fused
def test_fused_gateway_argument_count_errors():
    assert cref.ref_call_malformed(0, "fused") == (1, "MyToolbox:arrayProduct:nrhs")
    assert cref.ref_call_malformed(1, "fused")[0] == 1


@needs_fused
@pytest.mark.gpu
@pytest.mark.parametrize("covmode,d", [("full", 2), ("diag", 3)])
def test_fused_gateway_matches_oracle(covmode, d):
    """vbhem_estep_fused_mex through the shim vs the oracle's literal path + numpy normalisation."""
    from oracle import numpy_ref as nr
    from util import make_case, oracle_fused
    base, hyp, cen = make_case(N=150, K=4, Sb=4, Sr=3, d=d, covmode=covmode, ragged=True, seed=29, Nv=12.0,
                               sr_list=[3, 1, 2, 3])
    pb, pc = pack.pack_base(base), pack.pack_centres(cen)
    got = cref.gateway_estep_fused(pb, pc, 5)
    want = oracle_fused(base, cen, 5)
    np.testing.assert_allclose(got["L_elbo"].T, want["L_elbo"], rtol=1e-9)
    np.testing.assert_allclose(got["zhat"].T, want["zhat"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got["Nj"], want["Nj"], rtol=1e-9)
    np.testing.assert_allclose(got["lb"][0], want["lb"][0], rtol=1e-9)
    for j in range(4):
        r = int(cen["Sr"][j])
        f = nr.finalize_stats(want["nu1"][j, :r], want["xi"][j, :r, :r], want["pr"][j, :r], want["mu"][j, :r],
                              want["Mu"][j, :r], covmode)
        np.testing.assert_allclose(got["Nj_rho1"][j, :r], f[0], rtol=1e-9)
        np.testing.assert_allclose(got["Nj_rho2rho"][j].T[:r, :r], f[1], rtol=1e-9, atol=1e-9 * np.abs(f[1]).max())
        np.testing.assert_allclose(got["Nj_rho"][j, :r], f[2], rtol=1e-9)
        np.testing.assert_allclose(got["y_bar"][j].T[:r], f[3], rtol=1e-9)
        SC = got["S_plus_C"][j].reshape(d, d, 3).transpose(2, 1, 0)[:r] if covmode == "full" else got["S_plus_C"][j].T[:r]
        np.testing.assert_allclose(SC, f[4], rtol=1e-7, atol=1e-9 * np.abs(f[4]).max())


needs_em = pytest.mark.skipif(not os.path.exists(cref.EM_SO), reason="EM-run gateway not built")


@needs_em
@pytest.mark.gpu
@pytest.mark.parametrize("covmode,d", [("full", 2), ("diag", 3)])
def test_em_run_gateway_equals_direct_em_run(covmode, d):
    """vbhem_em_run_mex through the shim (cell tree of base HMMs + flat hyper-parameter arrays in, hyper-parameters
    and LogLs out) must be the C ABI's vbhem_em_begin + vbhem_em_run + vbhem_em_fetch bit for bit, and its trace that
    of the oracle's EM loop."""
    from util import make_case
    from vbhem_b200 import synth
    from vbhem_b200.api import BaseSet
    base, hyp, cen = make_case(N=80, K=3, Sb=3, Sr=3, d=d, covmode=covmode, ragged=False, seed=33, Nv=15.0)
    opt = dict(synth.DEFAULT_OPT, Nv=15.0, tau=5, max_iter=12, minDiff=1e-6, m0=(256.0, 192.0, 128.0)[:d])
    got = cref.gateway_em_run(pack.pack_base(base), hyp, opt)
    with BaseSet(base) as bs:
        bs.em_begin(hyp, opt)
        want = bs.em_run(opt["tau"], opt["max_iter"], opt["minDiff"])
    assert got["iters"] == want["iters"] and got["unstable"] == want["unstable"]
    assert np.array_equal(got["LogLs"], want["LogLs"])
    for k in ("alpha", "eta", "epsilon", "m", "W", "v", "lam"):
        assert np.array_equal(got["hyp"][k], want["hyp"][k]), k
    if covmode == "full" and d == 2:
        from test_em_loop import oracle_em
        _, LL = oracle_em(base, hyp, opt)
        np.testing.assert_allclose(got["LogLs"], LL, rtol=1e-9)


@needs_em
def test_em_run_gateway_argument_errors():
    lib = cref._lib("em")
    assert cref.ref_call_malformed(0, "em") == (1, "MyToolbox:arrayProduct:nrhs")


needs_em_batch = pytest.mark.skipif(not os.path.exists(cref.EM_BATCH_SO), reason="batched-learning gateway not built")


@needs_em_batch
@pytest.mark.gpu
def test_em_batch_gateway_equals_direct_batch():
    """vbhmm_em_batch_mex through the shim (nested cell of subjects' sequences, stacked initial parameters) must be
    hmm.vbhmm_em_batch (vbhmm_batch_begin / run / fetch) bit for bit."""
    from oracle import numpy_ref as nr
    from test_hmm_em import INI, make_sequences
    from vbhem_b200 import fb, hmm
    ini = dict(INI, maxIter=25)
    subjects, inits, run_subject = [], [], []
    for s in range(3):
        data, gmm = make_sequences(n=5 + 2 * s, K=3, seed=80 + s, tmin=1)
        subjects.append(data)
        for trial in range(2):
            g = dict(gmm, mean=gmm["mean"] + 4.0 * trial)
            inits.append(nr.vbhmm_init_from_gmm(g, sum(len(x) for x in data), ini))
            run_subject.append(s)
    got = cref.gateway_em_batch(subjects, run_subject, 3, ini, inits)
    start = np.concatenate([[0], np.cumsum([len(x) for x in subjects])])
    with fb.SequenceSet([x for sub in subjects for x in sub]) as seqs:
        want, _ = hmm.vbhmm_em_batch(seqs, [(int(start[s]), int(start[s + 1])) for s in run_subject], 3, ini, inits)
    for g, w in zip(got, want):
        assert g["iter"] == w["iter"] and g["unstable"] == w["unstable"]
        assert np.array_equal(g["LLs"], w["LLs"])
        for k in ("alpha", "epsilon", "beta", "v", "m", "W"):
            assert np.array_equal(g["varpar"][k], w["varpar"][k]), k
        for k in ("N1", "N", "M"):
            assert np.array_equal(g[k], w[k]), k


@needs_em_batch
def test_em_batch_gateway_argument_errors():
    assert cref.ref_call_malformed(0, "em_batch") == (1, "MyToolbox:arrayProduct:nrhs")
