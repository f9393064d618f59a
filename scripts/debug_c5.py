import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from synth import make_observation
from surmise_b200.emulationmethods import PCGPwM_B200 as M
from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
w = bench.workload()
np.random.seed(0)
fi = {}
M.fit(fi, w['x'], w['theta'], w['f'], fit_mode='parallel')
theta_true, y, yvar = make_observation(w['truth'], 8, seed=3)
class Emu: _info = fi
cal = {'yvar': yvar}
rng = np.random.default_rng(0)
tt = np.vstack([theta_true, np.clip(theta_true + 0.05 * rng.normal(size=(5, 8)), 0, 1), rng.uniform(size=(5, 8)),
                np.round(rng.uniform(size=(5, 8)))])
ll, dll = C.loglik_grad(cal, Emu, tt, y, w['x'])
print('loglik at true / near / random / corners:\n', np.round(ll[:, 0], 1))
p = {}
M.predict(p, fi, w['x'], tt)
tr = w['truth'](tt)
print('emulator rmse per theta:', np.round(np.sqrt(np.mean((p['mean'] - tr) ** 2, 0)), 3))
print('pred sd (mean over x):', np.round(np.sqrt(p['var']).mean(0), 3))
print('resid rms vs y:', np.round(np.sqrt(np.mean((p['mean'] - y[:, None]) ** 2, 0)), 3))
# finite-difference check of the gradient at the true theta
h = 1e-6
g = []
for i in range(8):
    t2 = tt[:1].copy(); t2[0, i] += h
    g.append((C.loglik(cal, Emu, t2, y, w['x'])[0, 0] - ll[0, 0]) / h)
print('grad analytic', np.round(dll[0], 2)); print('grad fd      ', np.round(g, 2))
