"""cProfile of the C5 sampling loop (PTLMC_B200, B = 264) for a short chain: where the host time per step goes."""
import os, sys, time, cProfile, pstats, io
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from synth import make_observation
from surmise_b200.emulationmethods import PCGPwM_B200 as M
from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
spc = int(sys.argv[1]) if len(sys.argv) > 1 else 600
w = bench.workload()
np.random.seed(0)
fi = {'method': 'PCGPwM_B200'}
M.fit(fi, w['x'], w['theta'], w['f'], fit_mode='parallel', merge_leaders=bool(os.environ.get('MERGE')))
theta_true, y, yvar = make_observation(w['truth'], 8, seed=3)


class prior:
    @staticmethod
    def lpdf(th):
        th = np.atleast_2d(th)
        return np.where(np.all((th >= 0) & (th <= 1), axis=1), 0.0, -np.inf).reshape(-1, 1)

    @staticmethod
    def lpdf_grad(th):
        return np.zeros_like(np.atleast_2d(th))

    @staticmethod
    def rnd(n):
        return np.random.uniform(size=(n, 8))


class Emu:
    _info = fi
    _emulator__theta = w['theta']


cal = {'thetaprior': prior, 'yvar': yvar}
np.random.seed(1)
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
C.fit(cal, Emu, w['x'], y, sampler='PTLMC_B200', numsamp=2000, numtemps=8, numchain=256, sampperchain=spc, maxtemp=30)
torch.cuda.synchronize(); pr.disable()
dt = time.perf_counter() - t0
steps = int(np.ceil(spc * 0.5)) + spc
print('wall', dt, 'steps', steps, 'ms/step (incl. pre-optimiser)', dt / steps * 1e3)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28); print(s.getvalue()[:6000])
