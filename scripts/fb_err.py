"""Observed error of the CUDA forward-backward pass against the plain-C oracle (ragged sequences, K = 5, dim = 2)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vbhem_b200  # noqa
from vbhem_b200 import fb
from oracle import cref, numpy_ref as nr
from test_fb import make_fb_case
data, vp = make_fb_case(N=2000, K=5, dim=2, maxT=50, seed=5)
got = fb.vbhmm_fb(data, vp)
want = cref.port_vbhmm_fb(cref.pack_sequences(data), nr.vbhmm_fb_prepass(vp))
for k in ("logrho_Saved", "gamma_all", "xi_sum", "phi_norm"):
    g, w = got[k], want[k]
    print(k, "max abs %.3e" % np.max(np.abs(g - w)), "max rel (|w| > 1e-6) %.3e" % np.max(np.abs(g - w)[np.abs(w) > 1e-6] / np.abs(w)[np.abs(w) > 1e-6]))
