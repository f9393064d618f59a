"""Phase clock of a PCGPwM_B200 fit: python scripts/fit_phases.py C3|C4|C2 [reference|parallel]"""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from synth import make_simulator
from surmise_b200.emulationmethods import PCGPwM_B200 as M
name = sys.argv[1] if len(sys.argv) > 1 else 'C3'
mode = sys.argv[2] if len(sys.argv) > 2 else 'reference'
c = {'C2': dict(n=2000, d=8, m=500, k=20, missing=0.0), 'C3': dict(n=4000, d=8, m=1000, k=32, missing=0.10),
     'C4': dict(n=8000, d=10, m=2000, k=64, missing=0.0)}[name]
x, theta, f, truth = make_simulator(c['n'], c['d'], c['m'], c['k'], seed=0, missing=c['missing'])
for rep in range(2):
    np.random.seed(0); fi = {}
    t0 = time.perf_counter(); M.fit(fi, x, theta, f, fit_mode=mode); torch.cuda.synchronize()
    print(json.dumps({'config': name, 'mode': mode, 'dag': os.environ.get('SB200_POTRF_DAG', '1'), 'rep': rep,
                      'wall': round(time.perf_counter() - t0, 4),
                      'phases': {k: round(v, 4) for k, v in fi['_b200']['timing'].items()},
                      'eig_path': fi['standardpcinfo'].get('_eig_path')}))
