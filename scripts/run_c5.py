"""Config C5: directbayeswoodbury calibration with a PTLMC sampler, 256 chains + 8 temperatures (B = 264 per step),
on the C2 emulator (n=2000, d=8, m=500, k=20).  usage: run_c5.py [sampperchain]  -> gpurun_out/c5.json"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from synth import make_observation
from surmise_b200.emulationmethods import PCGPwM_B200 as M
from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
spc = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
w = bench.workload()
np.random.seed(0)
fi = {'method': 'PCGPwM_B200'}
t0 = time.perf_counter(); M.fit(fi, w['x'], w['theta'], w['f'], fit_mode='parallel'); t_fit = time.perf_counter() - t0
theta_true, y, yvar = make_observation(w['truth'], 8, seed=3)
stats = {'calls': 0, 'rows': 0, 't_prior': 0.0}


class prior:                      # uniform on [0,1]^d; analytic gradient supplied
    @staticmethod
    def lpdf(th):
        t0 = time.perf_counter()
        th = np.atleast_2d(th)
        r = np.where(np.all((th >= 0) & (th <= 1), axis=1), 0.0, -np.inf).reshape(-1, 1)
        stats['t_prior'] += time.perf_counter() - t0
        return r

    @staticmethod
    def lpdf_grad(th):
        return np.zeros_like(np.atleast_2d(th))

    @staticmethod
    def rnd(n):
        return np.random.uniform(size=(n, 8))


class Emu:
    _info = fi
    _emulator__theta = w['theta']

    @staticmethod
    def predict(x=None, theta=None, args={}):
        raise NotImplementedError


cal = {'thetaprior': prior, 'yvar': yvar}
orig = C._evaluate


def counted(fitinfo, emu, theta, y_, x_, return_grad):
    stats['calls'] += 1; stats['rows'] += np.atleast_2d(theta).shape[0]
    return orig(fitinfo, emu, theta, y_, x_, return_grad)


C._evaluate = counted
np.random.seed(1)
t0 = time.perf_counter()
C.fit(cal, Emu, w['x'], y, sampler='PTLMC_B200', numsamp=2000, numtemps=8, numchain=256, sampperchain=spc, maxtemp=30)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
steps = int(np.ceil(spc * 0.5)) + spc
post = cal['thetarnd']
res = {'config': 'C5: PTLMC_B200, numchain=256, numtemps=8 (B=264 per step), sampperchain=%d (%d steps) on the C2 emulator' % (spc, steps),
       'emulator_fit_s': t_fit, 'calibration_wall_s': dt, 'loglik_calls': stats['calls'], 'theta_evaluated': stats['rows'],
       'theta_per_s_end_to_end': stats['rows'] / dt, 'ms_per_mcmc_step_end_to_end': dt / steps * 1e3,
       'host_prior_s': stats['t_prior'], 'posterior_mean': post.mean(0).tolist(), 'theta_true': theta_true[0].tolist(),
       'posterior_sd': post.std(0).tolist(),
       'fraction_of_draws_within_0.1_of_truth': float(np.mean(np.all(np.abs(post - theta_true[0]) < 0.1, axis=1))),
       'best_logpost_in_sample': float(np.max(C.loglik(cal, Emu, post[:512], y, w['x']))),
       'logpost_at_truth': float(C.loglik(cal, Emu, theta_true, y, w['x'])[0, 0]),
       'abs_err_over_sd': (np.abs(post.mean(0) - theta_true[0]) / post.std(0)).tolist()}
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'c5.json'), 'w'), indent=1)
