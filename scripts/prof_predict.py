"""cProfile of PCGPwM_B200.predict (plugin level, numpy in / numpy out) on the C2 emulator."""
import os, sys, time, cProfile, pstats, io
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from surmise_b200.emulationmethods import PCGPwM_B200 as M
w = bench.workload()
np.random.seed(0)
fi = {'method': 'PCGPwM_B200'}
M.fit(fi, w['x'], w['theta'], w['f'], fit_mode='parallel')
for B, grad in ((256, True), (256, False), (2048, False)):
    tn = np.random.default_rng(1).uniform(0.05, 0.95, size=(B, 8))
    for _ in range(2):
        p = {}; M.predict(p, fi, w['x'], tn, return_grad=grad)
    pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
    for _ in range(5):
        p = {}; M.predict(p, fi, w['x'], tn, return_grad=grad)
    pr.disable(); dt = (time.perf_counter() - t0) / 5
    print('B', B, 'grad', grad, 'predict s', dt, 'bytes out', sum(v.nbytes for v in p.values() if hasattr(v, 'nbytes')) / 1e6, 'MB')
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(8); print(s.getvalue()[200:1800])
