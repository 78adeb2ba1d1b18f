"""Times vbhmm_fb on synthetic sequences (BASELINE.json configs[4] shape by default)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vbhem_b200  # noqa: F401,E402
from vbhem_b200 import fb  # noqa: E402
from vbhem_b200._lib import check, dp, lib  # noqa: E402

N, T, K, dim = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1_000_000, 50, 5, 2
RAGGED = "--ragged" in sys.argv          # lengths uniform in 1..T instead of all T (real recordings are ragged)
rng = np.random.default_rng(1001)
means = rng.uniform([0, 0], [512, 384], size=(K, dim))
z = rng.integers(0, K, size=(N, T))
X = means[z] + rng.normal(size=(N, T, dim)) * 15
flat = np.ascontiguousarray(np.transpose(X, (0, 2, 1))).ravel()      # each sequence [T x dim] column-major
Tn = np.full(N, T, np.int32)
offs = np.arange(N + 1, dtype=np.int64) * T * dim
if RAGGED:      # every sequence keeps its slot of T steps in `flat` but only its first Tn steps count: [Tn x dim] column-major
    Tn = rng.integers(1, T + 1, size=N).astype(np.int32)
    Xr = np.transpose(X, (0, 2, 1)).copy()                            # [N, dim, T]
    flat = np.zeros(N * T * dim)
    for d in range(dim):                                              # sequence n: dim runs of Tn[n] steps from offs[n]
        idx = offs[:-1, None] + d * Tn[:, None] + np.arange(T)[None, :]
        keep = np.arange(T)[None, :] < Tn[:, None]
        flat[idx[keep]] = Xr[:, d, :][keep]
vp = dict(v=5 + rng.uniform(size=K) * 10, W=np.stack([np.eye(dim) / (15.0 ** 2 * 8)] * K),
          epsilon=1 + rng.uniform(size=(K, K)) * 5, alpha=1 + rng.uniform(size=K) * 3, m=means, beta=np.full(K, 50.0))
par = fb.prepass(vp)
t0 = time.perf_counter()
seqs = fb.SequenceSet.from_packed(Tn, offs, flat, dim)
t_pack = time.perf_counter() - t0
c = np.ascontiguousarray
m, W, A = c(par["m"]), c(np.transpose(par["W"], (0, 2, 1))), c(par["A"].T)
P = lambda a: a.ctypes.data_as(dp)
ms = C.c_double()
times = []
for _ in range(6):
    check(lib().vbhmm_fb_dev(seqs._h, K, P(m), P(W), P(c(par["v"])), P(c(par["beta"])), P(c(par["logLambdaTilde"])),
                             C.c_double(par["const_denominator"]), P(c(par["pz1"])), P(A)))
    check(lib().vbhmm_fb_last_ms(seqs._h, C.byref(ms)))
    times.append(ms.value)
phi = np.zeros(N)
check(lib().vbhmm_fb_fetch(seqs._h, None, None, None, P(phi)))
steps = int(Tn.sum())
alg = steps * dim * 8 + 2 * N * T * K * 8 + N * K * K * 8 + N * 8
print(json.dumps(dict(N=N, T=T, K=K, dim=dim, ragged=RAGGED, sequence_steps=steps,
                      ns_per_sequence_step=min(times) * 1e6 / steps, pack_s=t_pack, ms=times, best_ms=min(times),
                      seq_per_s=N / (min(times) * 1e-3), algorithmic_GB=alg / 1e9,
                      achieved_GBs=alg / (min(times) * 1e-3) / 1e9, phi_mean=float(phi.mean()))))
