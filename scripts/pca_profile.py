"""Time spent per primitive inside _pca_device.standardize (synchronised wrappers): python scripts/pca_profile.py [C2|C3|C4]"""
import os, sys, time, json, collections
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from synth import make_simulator
from surmise_b200 import _pca_device as P, ops

name = sys.argv[1] if len(sys.argv) > 1 else 'C4'
c = {'C2': dict(n=2000, d=8, m=500, k=20, missing=0.0), 'C3': dict(n=4000, d=8, m=1000, k=32, missing=0.10),
     'C4': dict(n=8000, d=10, m=2000, k=64, missing=0.0)}[name]
x, theta, f, truth = make_simulator(c['n'], c['d'], c['m'], c['k'], seed=0, missing=c['missing'])
fT = f.T                                      # (n, m) view of the (m, n) array, as PCGPwM_B200.fit holds it
ops._torch()
P.standardize(fT, 1e-3, 1e-5)
acc = collections.defaultdict(lambda: [0, 0.0])


def wrap(mod, fn):
    orig = getattr(mod, fn)

    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = orig(*a, **k)
        torch.cuda.synchronize()
        key = fn
        if fn == 'syevj':
            key = 'syevj_p%d' % a[0].shape[-1]
        e = acc[key]; e[0] += 1; e[1] += 1e3 * (time.perf_counter() - t0)
        return r
    setattr(mod, fn, w)


for fn in ('mm', 'syevj', 'col_stats', 'standardize', 'sumsq', 'spd_solve_small', 'dgemm', 'to_host'):
    wrap(ops, fn)
for fn in ('_upload', '_leading_block', '_full_eig', '_impute_schur', '_ridge_rows', '_orthonormalise'):
    wrap(P, fn)
torch.cuda.synchronize(); t0 = time.perf_counter()
spc, mof, rows = P.standardize(fT, 1e-3, 1e-5)
torch.cuda.synchronize(); t1 = time.perf_counter()
P.principal_components(spc, mof, rows, 1e-3, 1e-5)
torch.cuda.synchronize(); t2 = time.perf_counter()
print(json.dumps({'config': name, 'standardize_ms': 1e3 * (t1 - t0), 'pcs_ms': 1e3 * (t2 - t1), 'eig_path': spc['_eig_path'],
                  'U_shape': list(spc['U'].shape),
                  'parts': {k: [v[0], round(v[1], 2)] for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1])}}))
