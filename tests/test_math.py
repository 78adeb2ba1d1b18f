"""CPU tier: the lattice kernels' own fp64 exp / log (csrc/vbhem_math.cuh) compiled for the host and compared
with libm.  They replace libdevice's exp / log (19 / 47 DFMA-equivalents per call on B200) with 7 / 6 fp64
instructions and carry the log-sum-exp shift in the integer exponent; this pins their accuracy -- Taylor degree 3 /
4 since round 2: 1.4e-13 relative for exp, 8e-15 absolute for log, four orders inside the 1e-9 parity budget
(csrc/gen_mathtab.py has the measured effect on L_elbo and z_hat at the full shape)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "clustering-hidden-markov-models-with-variational-bayesian-hierarchical-em_b200", "csrc")


def test_kernel_math_against_libm(tmp_path):
    exe = str(tmp_path / "math_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", CSRC, "-o", exe,
                    os.path.join(ROOT, "tests", "math_check.cpp")], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")
    exp_rel, log_abs_small, log_rel, lse_rel = map(float, out[0].split())
    assert exp_rel < 1.5e-13              # degree-3 remainder (u/2)^4 / 24 of mant * 2^e over [-700, 700]
    assert log_abs_small < 1.5e-14        # log s with |log s| < 32: degree-4 remainder 5.7e-15 + one ulp of 32 / u units
    assert log_rel < 6e-15                # c ln2 + log s, relative to max(|c ln2|, |log s|, 1)
    assert lse_rel < 1.5e-13              # a whole row (split, integer shift, scale, sum, log) down to G = -9e6
    assert out[1].split() == ["1"] * 8, out[1]
    assert out[2].split() == ["1"] * 4, out[2]   # row shifts are trusted up to |G| = 1e9 and flagged beyond


def test_math_tables_are_current(tmp_path):
    """vbhem_mathtab.h is generated; regenerate and compare so the committed header cannot drift."""
    import importlib.util
    import shutil
    src = os.path.join(CSRC, "gen_mathtab.py")
    work = tmp_path / "gen"
    work.mkdir()
    shutil.copy(src, work / "gen_mathtab.py")
    spec = importlib.util.spec_from_file_location("gen_mathtab", str(work / "gen_mathtab.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.HERE = str(work)
    mod.main()
    assert open(work / "vbhem_mathtab.h").read() == open(os.path.join(CSRC, "vbhem_mathtab.h")).read()
