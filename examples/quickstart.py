"""Quick start: fit a PCGPwM_B200 emulator, predict, and calibrate with directbayeswoodbury_B200 on a B200.

With surmise installed (or `baseline/_ref` on sys.path) everything goes through surmise's own `emulator` /
`calibrator` classes -- only the method names differ from a CPU run.  Without surmise the same plugin modules are
driven directly (that is all the facades do).

    python examples/quickstart.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from synth import make_simulator, make_observation          # noqa: E402  (seeded synthetic simulator)

import surmise_b200                                          # noqa: E402

x, theta, f, truth = make_simulator(n=1000, d=4, m=120, k=10, seed=0)      # f: (m, n) simulator outputs
theta_true, y, yvar = make_observation(truth, 4, seed=1)                    # one noisy observation of truth(theta_true)


class prior:                                                 # uniform on the unit box
    @staticmethod
    def lpdf(th):
        th = np.atleast_2d(th)
        return np.where(np.all((th >= 0) & (th <= 1), axis=1), 0.0, -np.inf).reshape(-1, 1)

    @staticmethod
    def lpdf_grad(th):
        return np.zeros_like(np.atleast_2d(th))

    @staticmethod
    def rnd(n):
        return np.random.uniform(size=(n, 4))


try:
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if os.path.isdir(ref):
        sys.path.insert(0, ref)
    from surmise.emulation import emulator
    from surmise.calibration import calibrator
    surmise_b200.register()
    t0 = time.perf_counter()
    emu = emulator(x, theta, f, method='PCGPwM_B200', args={'fit_mode': 'parallel'})
    print('fit            %.3f s   (%d components)' % (time.perf_counter() - t0, len(emu._info['emulist'])))
    tn = np.random.default_rng(2).uniform(0.1, 0.9, size=(256, 4))
    t0 = time.perf_counter()
    pred = emu.predict(x, tn, args={'return_grad': True})
    print('predict        %.4f s  rmse/sd %.4f' % (time.perf_counter() - t0,
                                                   np.sqrt(np.mean((pred.mean() - truth(tn)) ** 2)) / np.std(truth(tn))))
    np.random.seed(3)
    t0 = time.perf_counter()
    cal = calibrator(emu, y, x, prior, method='directbayeswoodbury_B200', yvar=yvar,
                     args={'sampler': 'PTLMC_B200', 'numsamp': 2000, 'numtemps': 16, 'numchain': 112, 'sampperchain': 1500})
    draws = cal.theta.rnd(2000)
    print('calibrate      %.2f s   posterior mean %s  sd %s  (truth %s)' % (
        time.perf_counter() - t0, np.round(draws.mean(0), 3), np.round(draws.std(0), 3), np.round(theta_true[0], 3)))
except ImportError:
    from surmise_b200.emulationmethods import PCGPwM_B200 as E
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    fitinfo = {'method': 'PCGPwM_B200'}
    t0 = time.perf_counter()
    E.fit(fitinfo, x, theta, f, fit_mode='parallel')
    print('fit            %.3f s   (%d components)' % (time.perf_counter() - t0, len(fitinfo['emulist'])))
    tn = np.random.default_rng(2).uniform(0.1, 0.9, size=(256, 4))
    p = {}
    t0 = time.perf_counter()
    E.predict(p, fitinfo, x, tn, return_grad=True)
    print('predict        %.4f s  rmse/sd %.4f' % (time.perf_counter() - t0,
                                                   np.sqrt(np.mean((p['mean'] - truth(tn)) ** 2)) / np.std(truth(tn))))

    class emu:                                               # what the calibration plugin reads from an emulator object
        _info = fitinfo
        _emulator__theta = theta
    cal = {'thetaprior': prior, 'yvar': yvar}
    np.random.seed(3)
    t0 = time.perf_counter()
    C.fit(cal, emu, x, y, sampler='PTLMC_B200', numsamp=2000, numtemps=16, numchain=112, sampperchain=1500)
    print('calibrate      %.2f s   posterior mean %s  sd %s  (truth %s)' % (
        time.perf_counter() - t0, np.round(cal['thetarnd'].mean(0), 3), np.round(cal['thetarnd'].std(0), 3),
        np.round(theta_true[0], 3)))
