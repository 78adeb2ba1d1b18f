import sys, time, cProfile, pstats
sys.path.insert(0, '/root/repo')
import numpy as np, vbhem_b200
from vbhem_b200 import h3m, hmm
import bench_exprmt1 as b
subjects, label = b.sample_data(1001)
hm = hmm.vbhmm_learn_batch(subjects, [2], dict(b.VBOPT, numTrials=5, seed=1001, gmm_init="device"))
keep = [dict(prior=h["prior"], trans=h["trans"], mean=h["mean"], cov=h["cov"], varpar={k: h["varpar"][k] for k in ("alpha", "epsilon", "beta")}) for h in hm]
h3m.vbhem_h3m_cluster(keep, [1, 2], [1, 2], dict(b.VBHEMOPT, trials=2, seed=1))
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
out = h3m.vbhem_h3m_cluster(keep, list(range(1, 7)), list(range(1, 6)), dict(b.VBHEMOPT, trials=100, seed=1001))
print('total', time.perf_counter() - t0, out['seconds'])
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
