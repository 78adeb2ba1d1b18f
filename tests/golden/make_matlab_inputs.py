"""Inputs for tests/golden/make_golden_matlab.m (run by anyone with Octave or MATLAB and the reference toolbox):
a small seeded base set, an initial centre set and the options, written to matlab_stage_inputs.mat in the natural
layouts of vbhem_b200.pack (leading axis = HMM index).

    python tests/golden/make_matlab_inputs.py

The MATLAB-only stages of the path -- pre-pass, responsibilities, statistics reduction, lower bound, M-step
(vbhem_h3m_c_step_fc.m:118-165,271-283, vbhem_compute_Statistics.m, vbhemh3m_lb.m, vbhem_mstep_component.m) -- are
pinned in this repository by restatements only (neither MATLAB nor Octave exists in the build image); the .m script
turns the reference's own files into a fixture for them, and tests/test_matlab_golden.py compares when it exists."""
import os
import sys

import numpy as np
from scipy.io import savemat

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import vbhem_b200  # noqa: E402,F401
from vbhem_b200 import synth  # noqa: E402

OPT = dict(alpha0=1.0, eta0=1.0, epsilon0=1.0, lambda0=1.0, v0=5.0, W0=0.005, m0=(256.0, 192.0), Nv=20.0, tau=5,
           max_iter=25, minDiff=1e-6)


def main():
    base, _ = synth.make_base_set(30, S=3, d=2, C=3, covmode="full", seed=77, Nv=OPT["Nv"], ragged=True)
    hyp = synth.init_centres_baseem(base, 3, 3, seed=77, opt=dict(synth.DEFAULT_OPT, Nv=OPT["Nv"]))
    out = {"b_" + k: np.asarray(v, float) for k, v in base.items()}
    out.update({"r_" + k: np.asarray(v, float) for k, v in hyp.items()})
    out.update({"opt_" + k: np.asarray(v, float) for k, v in OPT.items()})
    savemat(os.path.join(HERE, "matlab_stage_inputs.mat"), out, do_compression=True)
    print("wrote", os.path.join(HERE, "matlab_stage_inputs.mat"))


if __name__ == "__main__":
    main()
