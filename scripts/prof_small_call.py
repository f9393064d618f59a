"""Host-side cost of one small NLL+grad call (n_h = 160, one problem): cProfile over 3000 calls."""
import os, sys, time, cProfile, pstats, io
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from surmise_b200 import ops
w = bench.workload()
rows = np.random.default_rng(0).choice(2000, 160, replace=False)
data = ops.GPData(w['theta'][rows], w['pc'][rows, 0], np.zeros(160))
mean_d, std_d = ops._f64(torch, w['mean']), ops._f64(torch, w['std'])
h = w['hyps'][:1]
for _ in range(50):
    ops.gp_nll_grad(data, h, mean_d, std_d, 0.01, True)
t0 = time.perf_counter()
for _ in range(3000):
    ops.gp_nll_grad(data, h, mean_d, std_d, 0.01, True)
dt = (time.perf_counter() - t0) / 3000
print('per call %.1f us' % (dt * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(3000):
    ops.gp_nll_grad(data, h, mean_d, std_d, 0.01, True)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(14); print(s.getvalue()[:2600])
