"""Short run of the theta log-likelihood path only (emulator state built from given hyper-parameters, no fit)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from surmise_b200.emulationmethods import PCGPwM_B200 as M
from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
from synth import make_observation
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
grad = (sys.argv[2] != '0') if len(sys.argv) > 2 else True
w = bench.workload()
k = w['k']
hyps = w['hyps'].copy()
hypinds = np.arange(k)
hyps[14:] = hyps[:k - 14]                    # 14 distinct factors, as the fitted C2 emulator has
hypinds[14:] = np.arange(k - 14)
spc = {kk: w['spc'][kk] for kk in ('offset', 'scale', 'extravar')}
emulist, state = M.import_state(w['theta'], w['pc'], np.zeros_like(w['pc']), hyps, hypinds, w['pcs']['pcti'], spc)
fi = {'_b200': state, 'emulist': emulist, 'x': w['x'], 'theta': w['theta'], 'pcti': w['pcs']['pcti'], 'standardpcinfo': spc}
_, y, yvar = make_observation(w['truth'], 8, seed=3)


class Emu:
    _info = fi


th = torch.from_numpy(np.random.default_rng(1).uniform(size=(B, 8))).cuda()
for _ in range(3):
    C.loglik_device({'yvar': yvar}, Emu, th, y, w['x'], return_grad=grad)
torch.cuda.synchronize()
print('ok')
