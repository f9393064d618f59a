"""cProfile of one PCGPwM_B200 fit at config C3 (10 % missing outputs), after a warm-up fit."""
import os, sys, time, cProfile, pstats, io
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from synth import make_simulator
from surmise_b200.emulationmethods import PCGPwM_B200 as M
name = sys.argv[1] if len(sys.argv) > 1 else 'C3'
mode = sys.argv[2] if len(sys.argv) > 2 else 'reference'
c = {'C3': dict(n=4000, d=8, m=1000, k=32, missing=0.10), 'C4': dict(n=8000, d=10, m=2000, k=64, missing=0.0)}[name]
x, theta, f, truth = make_simulator(c['n'], c['d'], c['m'], c['k'], seed=0, missing=c['missing'])
np.random.seed(0); fi = {}
t0 = time.perf_counter(); M.fit(fi, x, theta, f, fit_mode=mode); torch.cuda.synchronize(); print('first', time.perf_counter() - t0)
np.random.seed(0); fi = {}
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
M.fit(fi, x, theta, f, fit_mode=mode)
torch.cuda.synchronize(); pr.disable()
print(name, mode, 'wall', time.perf_counter() - t0, 'evals', fi['_b200']['nevals'])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(30); print(s.getvalue()[:5500])
