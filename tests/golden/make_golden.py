"""Generates tests/golden/*.npz from the REFERENCE's own native code.

Run in the dev container (needs /root/reference):
    python tests/golden/make_golden.py
It builds oracle/_ref/libvbhem_ref.so -- the reference's
src/vbhem/vbhem_hmm_bwd_fwd_mex.c compiled unmodified behind oracle/mxshim -- and
stores its six outputs for a few small seeded cases, together with the exact inputs.
The fixtures are what pins oracle/estep_oracle.c, oracle/numpy_ref.py and the CUDA
path on machines where /root/reference does not exist (the GPU box).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import vbhem_b200  # noqa: E402,F401
from oracle import cref  # noqa: E402
from util import make_case, oracle_pairs  # noqa: E402

CASES = {
    # name: (make_case kwargs, tau)
    "full_ragged_d2": (dict(N=7, K=3, Sb=4, Sr=3, d=2, covmode="full", ragged=True, seed=5), 5),
    "diag_ragged_d3": (dict(N=6, K=4, Sb=3, Sr=4, d=3, covmode="diag", ragged=True, seed=11), 4),
    "full_s8_d2_tau10": (dict(N=5, K=4, Sb=8, Sr=8, d=2, covmode="full", ragged=False, seed=1001), 10),
    "full_tau1": (dict(N=4, K=2, Sb=2, Sr=2, d=2, covmode="full", ragged=False, seed=3), 1),
    "exprmt1_shape": (dict(N=40, K=2, Sb=2, Sr=2, d=2, covmode="full", ragged=False, seed=7, Nv=100.0), 50),
    # BASELINE configs[1] (demo/vbdemo_face.m:46-63): tens of subjects' HMMs with <= 3 states, K up to 5, Nv = 10, tau = 5
    "face_shape": (dict(N=30, K=5, Sb=3, Sr=3, d=2, covmode="full", ragged=True, seed=17, Nv=10.0), 5),
}
ONLY = os.environ.get("GOLDEN_ONLY")   # regenerate one fixture without touching the others

if __name__ == "__main__":
    assert os.path.exists("/root/reference/src/vbhem/vbhem_hmm_bwd_fwd_mex.c"), "needs the reference checkout"
    cref.build()
    here = os.path.dirname(os.path.abspath(__file__))
    for name, (kw, tau) in CASES.items():
        if ONLY and name != ONLY:
            continue
        base, hyp, cen = make_case(**kw)
        out = oracle_pairs(base, cen, tau, which="ref")
        blob = {f"base_{k}": v for k, v in base.items()}
        blob.update({f"cen_{k}": v for k, v in cen.items()})
        blob.update({f"hyp_{k}": v for k, v in hyp.items()})
        blob.update({f"out_{k}": v for k, v in out.items()})
        blob["tau"] = np.int64(tau)
        blob["covmode"] = np.array(kw["covmode"])
        np.savez_compressed(os.path.join(here, name + ".npz"), **blob)
        print(name, "L_elbo[0,:]=", out["L_elbo"][0])

    if ONLY:
        sys.exit(0)
    # VB-HMM forward-backward (src/hmm/vbhmm_fb_mex.c compiled in place): ragged sequences incl. T = 1 and T = 0
    assert os.path.exists("/root/reference/src/hmm/vbhmm_fb_mex.c")
    from oracle import numpy_ref as nr
    from test_fb import make_fb_case
    data, vp = make_fb_case(N=12, K=4, dim=2, maxT=15, seed=3)
    out = cref.ref_vbhmm_fb(cref.pack_sequences(data), nr.vbhmm_fb_prepass(vp))
    blob = {f"seq{n}": x for n, x in enumerate(data)}
    blob.update(vp)
    blob.update(out)
    blob["N"] = np.int64(len(data))
    np.savez_compressed(os.path.join(here, "vbhmm_fb_ragged_d2_k4.npz"), **blob)
    print("vbhmm_fb_ragged_d2_k4 phi_norm[:3]=", out["phi_norm"][:3])
