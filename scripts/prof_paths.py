"""Short runs of the three hot paths for ncu launch lists: argv[1] in {nll, loglik, fit}."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from surmise_b200 import ops
from surmise_b200.emulationmethods import PCGPwM_B200 as M
from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
from synth import make_observation
which = sys.argv[1]
w = bench.workload()
dev = torch.device('cuda')
if which == 'nll':
    data = ops.GPData(w['theta'], np.ascontiguousarray(w['pc'].T), None)
    hyp = torch.from_numpy(w['hyps']).cuda()
    for _ in range(3):
        ops.gp_nll_grad(data, hyp, w['mean'], w['std'], 0.01, True)
elif which == 'nllsmall':
    rows = np.random.default_rng(0).choice(2000, 160, replace=False)
    data = ops.GPData(w['theta'][rows], np.ascontiguousarray(w['pc'][rows].T), None)
    hyp = torch.from_numpy(w['hyps']).cuda()
    for _ in range(3):
        ops.gp_nll_grad(data, hyp, w['mean'], w['std'], 0.01, True)
    d1 = ops.GPData(w['theta'][rows], w['pc'][rows, 0], None)
    for _ in range(3):
        ops.gp_nll_grad(d1, hyp[:1], w['mean'], w['std'], 0.01, True)
else:
    np.random.seed(0)
    fi = {'method': 'PCGPwM_B200'}
    M.fit(fi, w['x'], w['theta'], w['f'])
    if which == 'loglik':
        _, y, yvar = make_observation(w['truth'], 8, seed=3)
        class Emu: _info = fi
        th = torch.from_numpy(np.random.default_rng(1).uniform(size=(2048, 8))).cuda()
        for _ in range(3):
            C.loglik_device({'yvar': yvar}, Emu, th, y, w['x'], return_grad=True)
torch.cuda.synchronize()
print('ok')
