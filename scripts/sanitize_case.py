"""Smallest run that touches every kernel (for compute-sanitizer memcheck): small fit (sweep path), blocked path
(n = 300 > 207), predict with gradients, back-projection, Woodbury log-likelihood, standalone covmat."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from synth import make_simulator, make_observation
import surmise_b200
from surmise_b200 import ops
from surmise_b200.emulationmethods import PCGPwM_B200 as M
from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
x, theta, f, truth = make_simulator(90, 3, 14, 3, seed=1, missing=0.08)
np.random.seed(1)
fi = {}
M.fit(fi, x, theta, f)
p = {}
M.predict(p, fi, x, np.random.default_rng(0).uniform(size=(37, 3)), return_grad=True)
_, y, yvar = make_observation(truth, 3, seed=2)
class Emu: _info = fi
C.loglik_grad({'yvar': yvar}, Emu, np.random.default_rng(1).uniform(size=(19, 3)), y, x)
rng = np.random.default_rng(3)
th = rng.uniform(size=(300, 5)); G = rng.normal(size=(2, 300))
mean, lo, hi, std = M.hyper_prior(th, -10, -20, None)
ops.gp_nll_grad(ops.GPData(th, G, np.abs(G) * 0.01), np.tile(mean, (2, 1)), mean, std, 0.01, True)
ops.gp_nll_grad(ops.GPData(th, G, None), np.tile(mean, (2, 1)), mean, std, 0.01, False)
surmise_b200.covmat(th[:45], th[:33], np.zeros(6), return_gradhyp=True)
surmise_b200.covmat(th[:45], th[:33], np.zeros(6), return_gradx1=True)
torch.cuda.synchronize()
print('sanitize case done')
