"""Golden fit/predict of the reference's indGP (build-container only, see make_golden.py).

    PYTHONPATH=/tmp/ref/src python tests/golden/make_golden_indgp.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from surmise.emulation import emulator                      # noqa: E402
from synth import make_simulator                            # noqa: E402

out = {}
x, theta, f, truth = make_simulator(60, 2, 7, 3, seed=61)
rng = np.random.default_rng(62)
fm = f.copy()
fm[1, rng.choice(60, 6, replace=False)] = np.nan            # rows 1 and 2 share one missing pattern
fm[2, np.isnan(fm[1])] = np.nan
fm[4, rng.choice(60, 9, replace=False)] = np.nan
tn = rng.uniform(0.1, 0.9, size=(9, 2))
for tag, ff in (('full', f), ('miss', fm)):
    emu = emulator(x, theta, ff, method='indGP')
    fi = emu._info
    p = {}
    emu.method.predict(p, fi, x, tn, return_grad=True)
    out[tag + '_f'] = ff
    out[tag + '_hyp'] = np.array([e['hyp'] for e in fi['emulist']])
    out[tag + '_sig2'] = np.array([e['sig2'] for e in fi['emulist']])
    for k in ('mean', 'var', 'predvecs', 'predvars', 'covxhalf'):      # the reference drops its gradient arrays
        out[tag + '_' + k] = p[k]
    print(tag, out[tag + '_hyp'][0], p['covxhalf'].shape, float(np.abs(p['mean'] - truth(tn)).max()))
p1 = {}
emu.method.predict(p1, fi, x, tn[:1], return_grad=True)
out['miss_one_mean'], out['miss_one_covxhalf'] = p1['mean'], p1['covxhalf']
out['x'], out['theta'], out['thetanew'] = x, theta, tn
np.savez_compressed(os.path.join(HERE, 'indgp.npz'), **out)
