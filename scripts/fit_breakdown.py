import os, sys, time, cProfile, pstats, io
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from surmise_b200.emulationmethods import PCGPwM_B200 as M
w = bench.workload()
for mode in ('reference', 'parallel'):
    np.random.seed(0); fi = {}
    M.fit(fi, w['x'], w['theta'], w['f'], fit_mode=mode)   # warm
    np.random.seed(0); fi = {}
    pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
    M.fit(fi, w['x'], w['theta'], w['f'], fit_mode=mode)
    torch.cuda.synchronize(); pr.disable()
    print(mode, 'wall', time.perf_counter() - t0)
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(22); print(s.getvalue()[:3800])
