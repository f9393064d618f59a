"""Golden fixture for the calibrator's generic path (an emulator outside the PCGPwM family): outputs of the REAL
reference's directbayeswoodbury.loglik (calibrationmethods/directbayeswoodbury.py:261-306) for
  (a) an emulator fitted by the reference with method='PCGP' on a seeded synthetic simulator, and
  (b) a synthetic emulator object whose variance exceeds sum covxhalf^2, so that the batch-wide test fires,
together with the emulator outputs the likelihood consumed (mean / var / covxhalf at the proposals and at the training
thetas).  Build-container only:

    python tests/golden/make_golden_any_emulator.py            # imports the reference from baseline/_ref
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))

from surmise.emulation import emulator                                    # noqa: E402
import surmise.calibrationmethods.directbayeswoodbury as ref_dbw          # noqa: E402
from synth import make_simulator, make_observation                        # noqa: E402
from oracle import woodbury_oracle as wo                                  # noqa: E402


class _Pred:
    def __init__(self, mean, var, cxh):
        self._m, self._v, self._c = mean, var, cxh

    def mean(self):
        return self._m

    def var(self):
        return self._v

    def covxhalf(self):
        return self._c


class SyntheticEmulator:
    def __init__(self, m, K, d, seed, var_scale):
        rng = np.random.default_rng(seed)
        self.Wm, self.Wc = rng.standard_normal((m, d)), 0.3 * rng.standard_normal((m, K, d))
        self.theta0 = rng.uniform(size=(11, d))
        self.var_scale = var_scale

    def predict(self, x=None, theta=None):
        th = self.theta0 if theta is None else np.atleast_2d(theta)
        mean = self.Wm @ th.T
        cxh = np.einsum('mkd,bd->mbk', self.Wc, np.cos(th)) + 0.1
        var = self.var_scale * np.sum(cxh ** 2, 2) + 0.01 * (self.var_scale > 1)
        return _Pred(mean, var, cxh)


def record(tag, emu, x, theta, y, yvar, out):
    ll = ref_dbw.loglik({'yvar': yvar}, emu, theta, y, x)
    p, old = emu.predict(x, theta), emu.predict(x)
    d = {'mean': p.mean(), 'var': p.var(), 'covxhalf': p.covxhalf(), 'old_var': old.var(), 'old_covxhalf': old.covxhalf(),
         'y': np.squeeze(y), 'yvar': np.squeeze(yvar), 'theta': theta, 'loglik': ll}
    for k, v in d.items():
        out[tag + '_' + k] = np.asarray(v, dtype=np.float64)
    llo, fired = wo.loglik_from_prediction(d['mean'], d['var'], d['covxhalf'], yvar, y, d['old_var'], d['old_covxhalf'])
    out[tag + '_fired'] = np.array(fired)
    print('%-6s B=%d m=%d k=%d fired=%s  oracle vs reference %.2e' % (
        tag, theta.shape[0], d['mean'].shape[0], d['covxhalf'].shape[2], fired, np.max(np.abs(llo - ll) / np.abs(ll))))


def main():
    out = {}
    # (a) a real emulator of another method
    x, theta, f, truth = make_simulator(60, 3, 18, 3, seed=4)
    emu = emulator(x=x, theta=theta, f=f, method='PCGP')
    _, y, yvar = make_observation(truth, 3, seed=6)
    tn = np.random.default_rng(2).uniform(0.1, 0.9, size=(12, 3))
    record('pcgp', emu, x, tn, y, yvar, out)
    # (b) synthetic outputs for which the batch-wide variance test fires
    rng = np.random.default_rng(33)
    m = 29
    syn = SyntheticEmulator(m, 7, 3, seed=36, var_scale=3.0)
    record('fired', syn, None, rng.uniform(size=(33, 3)), rng.standard_normal(m), 0.05 + 0.1 * rng.random(m), out)
    np.savez_compressed(os.path.join(HERE, 'any_emulator.npz'), **out)


if __name__ == '__main__':
    main()
