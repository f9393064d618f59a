"""A short device-resident PT-LMC run at the C5 population (264 chains on the C2 emulator) for an ncu launch list."""
import os, sys, time
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
w['theta_true'], y, yvar = make_observation(w['truth'], 8, seed=3)
res = bench.c5_run(C, fi, w, y, yvar, torch, int(sys.argv[1]) if len(sys.argv) > 1 else 60)
print(res)
