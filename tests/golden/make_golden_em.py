"""Generates tests/golden/vbhmm_em_*.npz and vbhem_em_*.npz: whole EM runs.

Run in the dev container (needs /root/reference for the native parts):
    python tests/golden/make_golden_em.py
* vbhmm_em_*: the VB-HMM learning loop (oracle/numpy_ref.vbhmm_em = restatement of src/hmm/vbhmm_em.m +
  vbhmm_em_lb.m) with the forward-backward pass of every iteration taken from the REFERENCE's own C
  (src/hmm/vbhmm_fb_mex.c compiled unmodified, oracle/_ref/libvbhmm_fb_ref.so).
* vbhem_em_*: the VBHEM EM loop (tests/test_em_loop.oracle_em) with the per-pair E-step of every iteration from
  oracle/estep_oracle.c, which tests/test_oracle.py pins bit for bit on the reference's src/vbhem/vbhem_hmm_bwd_fwd_mex.c.
Inputs and outputs are stored together, so the GPU box checks the CUDA loops without any oracle code at run time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import vbhem_b200  # noqa: E402,F401
from oracle import cref, numpy_ref as nr  # noqa: E402
from test_em_loop import OPT, oracle_em  # noqa: E402
from test_hmm_em import INI, make_sequences  # noqa: E402
from util import make_case  # noqa: E402


def ref_fb(data, par):
    assert cref.have_fb_ref(), "build oracle/_ref first (make -C oracle ref)"
    return cref.ref_vbhmm_fb(cref.pack_sequences(data), par)


def main():
    for name, (K, dim, seed) in {"vbhmm_em_k3_d2": (3, 2, 4), "vbhmm_em_k2_d3": (2, 3, 8)}.items():
        ini = dict(INI, mu0=[256.0, 192.0, 100.0][:dim])
        data, gmm = make_sequences(n=16, K=K, dim=dim, seed=seed, tmin=1)
        vp0 = nr.vbhmm_init_from_gmm(gmm, sum(len(x) for x in data), ini)
        out = nr.vbhmm_em(data, vp0, ini, fb=ref_fb)
        T = np.array([len(x) for x in data])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), flat=np.concatenate(data), T=T, K=K, dim=dim,
                            gmm_prior=gmm["prior"], gmm_mean=gmm["mean"], gmm_cov=gmm["cov"],
                            **{"ini_" + k: np.atleast_1d(np.asarray(v, float)) for k, v in ini.items()},
                            LLs=out["LLs"], iters=out["iter"], **{"vp_" + k: v for k, v in out["varpar"].items()},
                            prior=out["prior"], trans=out["trans"], mean=out["mean"], cov=out["cov"])
        print(name, out["iter"], out["LLs"][-1])
    for name, kw in {"vbhem_em_full_d2": dict(N=60, K=4, Sb=3, Sr=3, d=2, covmode="full", ragged=False, seed=23, Nv=20.0)}.items():
        base, hyp, cen = make_case(**kw)
        fin, LL = oracle_em(base, hyp, OPT)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), LLs=LL,
                            **{"base_" + k: v for k, v in base.items()}, **{"hyp_" + k: v for k, v in hyp.items()},
                            **{"fin_" + k: v for k, v in fin.items()},
                            **{"opt_" + k: np.atleast_1d(np.asarray(v, float)) for k, v in OPT.items()})
        print(name, len(LL), LL[-1])


if __name__ == "__main__":
    main()
