"""theta log-likelihood throughput (C2 emulator) at a few population sizes, with and without gradient."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from surmise_b200.emulationmethods import PCGPwM_B200 as M
from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
from synth import make_observation
w = bench.workload()
np.random.seed(0)
fi = {'method': 'PCGPwM_B200'}
M.fit(fi, w['x'], w['theta'], w['f'])
_, y, yvar = make_observation(w['truth'], 8, seed=3)


class Emu:
    _info = fi


cal = {'yvar': yvar}
for B in (264, 2048, 8192):
    th = torch.from_numpy(np.random.default_rng(1).uniform(size=(B, 8))).cuda()
    for g in (False, True):
        for _ in range(3):
            C.loglik_device(cal, Emu, th, y, w['x'], return_grad=g)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            C.loglik_device(cal, Emu, th, y, w['x'], return_grad=g)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print('B=%d grad=%d  %.3f ms  %.0f theta/s' % (B, g, ms, B / ms * 1e3))
