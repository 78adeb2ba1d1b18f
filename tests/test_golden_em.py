"""Whole EM runs against committed golden fixtures (tests/golden/make_golden_em.py): the VB-HMM learning loop with
the reference's own C forward-backward inside, and the VBHEM loop on the C E-step that is pinned bit for bit on the
reference.  CPU: the oracle loops reproduce their fixtures (numpy forward-backward against the reference-C one).
GPU: the CUDA loops -- single run, batched runs, host mirror and device-resident -- reproduce them with no oracle
code at run time."""
import glob
import os

import numpy as np
import pytest

import vbhem_b200  # noqa: F401

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HMM_FIX = sorted(glob.glob(os.path.join(GOLD, "vbhmm_em_*.npz")))
H3M_FIX = sorted(glob.glob(os.path.join(GOLD, "vbhem_em_*.npz")))
ids = lambda ps: [os.path.basename(p)[:-4] for p in ps]


def load_hmm(path):
    z = np.load(path)
    T = z["T"]
    offs = np.concatenate([[0], np.cumsum(T)])
    data = [z["flat"][offs[n]:offs[n + 1]] for n in range(len(T))]
    ini = {k[4:]: (z[k] if z[k].size > 1 else float(z[k][0])) for k in z.files if k.startswith("ini_")}
    ini["maxIter"] = int(ini["maxIter"])
    gmm = dict(prior=z["gmm_prior"], mean=z["gmm_mean"], cov=z["gmm_cov"])
    want = dict(LLs=z["LLs"], iter=int(z["iters"]), varpar={k[3:]: z[k] for k in z.files if k.startswith("vp_")},
                prior=z["prior"], trans=z["trans"], mean=z["mean"], cov=z["cov"])
    return data, int(z["K"]), ini, gmm, want


def load_h3m(path):
    z = np.load(path)
    base = {k[5:]: z[k] for k in z.files if k.startswith("base_")}
    hyp = {k[4:]: z[k] for k in z.files if k.startswith("hyp_")}
    fin = {k[4:]: z[k] for k in z.files if k.startswith("fin_")}
    opt = {k[4:]: (z[k] if z[k].size > 1 else float(z[k][0])) for k in z.files if k.startswith("opt_")}
    opt["tau"], opt["max_iter"] = int(opt["tau"]), int(opt["max_iter"])
    return base, hyp, fin, opt, z["LLs"]


def check_hmm(got, want):
    assert got["iter"] == want["iter"]
    np.testing.assert_allclose(got["LLs"], want["LLs"], rtol=1e-9)
    for k, v in want["varpar"].items():
        np.testing.assert_allclose(got["varpar"][k], v, rtol=1e-7, atol=1e-12, err_msg=k)
    for k in ("prior", "trans", "mean", "cov"):
        np.testing.assert_allclose(got[k], want[k], rtol=1e-7, atol=1e-12, err_msg=k)


@pytest.mark.parametrize("path", HMM_FIX, ids=ids(HMM_FIX))
def test_oracle_vbhmm_em_reproduces_golden(path):
    from oracle import numpy_ref as nr
    data, K, ini, gmm, want = load_hmm(path)
    vp0 = nr.vbhmm_init_from_gmm(gmm, sum(len(x) for x in data), ini)
    check_hmm(nr.vbhmm_em(data, vp0, ini), want)                  # numpy forward-backward vs the reference-C golden


@pytest.mark.parametrize("path", H3M_FIX, ids=ids(H3M_FIX))
def test_oracle_vbhem_em_reproduces_golden(path):
    from test_em_loop import oracle_em
    base, hyp, fin, opt, LLs = load_h3m(path)
    got, LL = oracle_em(base, hyp, opt)
    np.testing.assert_allclose(LL, LLs, rtol=1e-12)
    for k in ("alpha", "eta", "epsilon", "m", "W", "v", "lam"):
        np.testing.assert_allclose(got[k], fin[k], rtol=1e-10, err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("path", HMM_FIX, ids=ids(HMM_FIX))
def test_cuda_vbhmm_em_reproduces_golden(path):
    from vbhem_b200 import fb, hmm
    data, K, ini, gmm, want = load_hmm(path)
    check_hmm(hmm.vbhmm_em(data, K, dict(ini, initmode="initgmm", initgmm=gmm)), want)
    vp0 = hmm.vbhmm_init(data, K, dict(ini, initmode="initgmm", initgmm=gmm))
    with fb.SequenceSet(data) as seqs:                                # the same run twice in one device batch
        both, _ = hmm.vbhmm_em_batch(seqs, [(0, len(data))] * 2, K, ini, [vp0, vp0])
    for h in both:
        check_hmm(h, want)


@pytest.mark.gpu
@pytest.mark.parametrize("path", H3M_FIX, ids=ids(H3M_FIX))
def test_cuda_vbhem_em_reproduces_golden(path):
    from vbhem_b200 import h3m
    from vbhem_b200.api import BaseSet
    base, hyp, fin, opt, LLs = load_h3m(path)
    with BaseSet(base) as bs:
        for device_loop in (False, True):
            run = h3m.vbhem_h3m_c_step_fc(hyp, bs, opt, device_loop=device_loop)
            np.testing.assert_allclose(run["LogLs"], LLs, rtol=1e-9)
            for k in ("alpha", "eta", "epsilon", "m", "W", "v", "lam"):
                np.testing.assert_allclose(run["hyp"][k], fin[k], rtol=1e-7, atol=1e-12, err_msg=f"{k} {device_loop}")
