"""One MCMC step of the device-resident PT-LMC at the C5 population (B = 264 on the C2 emulator): event-timed
evaluator launch; run under `ncu --metrics gpu__time_duration.sum` for the per-kernel list."""
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
_, y, yvar = make_observation(w['truth'], 8, seed=3)
class Emu: _info = fi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 264
dl = C.DeviceLogpost({'yvar': yvar}, Emu, w['x'], y, None)
ev = dl.evaluator(B, 8)
ev.thetap.copy_(torch.rand((B, 8), dtype=torch.float64, device='cuda'))
for _ in range(3):
    ev.launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ev.launch()
e1.record(); torch.cuda.synchronize()
print('B=%d evaluator launch: %.3f ms (%d distinct factors, k=%d)' % (B, e0.elapsed_time(e1) / 20, fi['_b200']['Linv'].shape[0], len(fi['emulist'])))
