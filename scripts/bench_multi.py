#!/usr/bin/env python
"""Strong scaling of the north-star E-step through the single-process multi-GPU C ABI (vbhem_multi_*).

    python scripts/bench_multi.py [--bases 100000] [--devices 1,2,4,8] [--steps 20]

For every device count n: MultiBaseSet over devices 0..n-1 (bases split contiguously), then
  * estep_ms  -- vbhem_multi_estep_hyp without per-pair outputs: host call -> statistics back (wall) and the
                 device time of the slowest shard (vbhem_multi_last_ms: E-step + all-reduce, CUDA events);
  * iter_ms   -- vbhem_multi_em_step (E-step + all-reduce + lower bound + M-step, hyper-parameters resident).
Prints one JSON object.
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bases", type=int, default=100_000)
    ap.add_argument("--centres", type=int, default=64)
    ap.add_argument("--states", type=int, default=8)
    ap.add_argument("--tau", type=int, default=10)
    ap.add_argument("--devices", default="")
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    import ctypes as C
    import numpy as np
    import vbhem_b200  # noqa: F401
    from vbhem_b200 import pack, synth
    from vbhem_b200._lib import lib
    from vbhem_b200.api import MultiBaseSet
    n = C.c_int()
    lib().vbhem_device_count(C.byref(n))
    counts = [int(x) for x in a.devices.split(",")] if a.devices else [c for c in (1, 2, 4, 8) if c <= n.value]
    base, _ = synth.make_base_set(a.bases, S=a.states, d=2, C=a.centres, seed=1001)
    hyp = synth.init_centres_baseem(base, a.centres, a.states, seed=1001)
    pb = pack.pack_base(base)
    opt = dict(synth.DEFAULT_OPT, tau=a.tau)
    out = dict(bases=a.bases, centres=a.centres, states=a.states, tau=a.tau, runs=[])
    ref = None
    for nd in counts:
        with MultiBaseSet(pb, devices=list(range(nd)), packed=True) as ms:
            for _ in range(3):
                r = ms.estep_hyp(hyp, a.tau, want_pairs=False)
            wall, dev = [], []
            for _ in range(a.steps):
                t0 = time.perf_counter()
                r = ms.estep_hyp(hyp, a.tau, want_pairs=False)
                wall.append(1e3 * (time.perf_counter() - t0))
                dev.append(ms.last_ms())
            if ref is None:
                ref = r["buf"]
            err = float(np.max(np.abs(r["buf"] - ref)) / np.max(np.abs(ref)))
            ms.em_begin(hyp, opt)
            ms.em_step(a.tau)
            ms.em_begin(hyp, opt)
            it, lbs = [], []
            for _ in range(8):
                t0 = time.perf_counter()
                lb = ms.em_step(a.tau)
                it.append(1e3 * (time.perf_counter() - t0))
                lbs.append(float(lb[0, 10]))
            out["runs"].append(dict(devices=nd, estep_wall_ms=statistics.median(wall), estep_dev_ms=statistics.median(dev),
                                    estep_dev_min_ms=min(dev), iter_wall_ms=statistics.median(it), LB=lbs,
                                    stats_rel_err_vs_first=err, pairs_per_s=a.bases * a.centres / (statistics.median(dev) * 1e-3),
                                    active_pairs=int(ms.counters()["active_pairs"])))
    t1 = out["runs"][0]["estep_dev_ms"] * out["runs"][0]["devices"]
    for r in out["runs"]:
        r["strong_efficiency"] = t1 / (r["estep_dev_ms"] * r["devices"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
