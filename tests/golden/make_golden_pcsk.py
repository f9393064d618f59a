"""Golden fit/predict of the reference's PCSK (build-container only, see make_golden.py).

    PYTHONPATH=/tmp/ref/src python tests/golden/make_golden_pcsk.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from surmise.emulation import emulator                      # noqa: E402
from synth import make_simulator                            # noqa: E402

x, theta, f, truth = make_simulator(120, 2, 16, 3, seed=71)
rng = np.random.default_rng(72)
simsd = 0.02 * np.abs(f) + 0.01                              # (m, n) simulation standard deviations
f = f + simsd * rng.normal(size=f.shape)
np.random.seed(71)
emu = emulator(x, theta, f, method='PCSK', args={'simsd': simsd})
fi = emu._info
tn = rng.uniform(0.1, 0.9, size=(11, 2))
p = {}
emu.method.predict(p, fi, x, tn, return_grad=True)
out = {'x': x, 'theta': theta, 'f': f, 'simsd': simsd, 'thetanew': tn,
       'hyp': np.array([e['hyp'] for e in fi['emulist']]), 'hypind': np.array([e['hypind'] for e in fi['emulist']]),
       'sig2': np.array([e['sig2'] for e in fi['emulist']]), 'pc': fi['pc'], 'pcstdvar': fi['pcstdvar'],
       'desc': np.array(fi['param_desc'])}
for k in ('mean', 'var', 'predvecs', 'predvars', 'covxhalf', 'mean_gradtheta', 'predvecs_gradtheta',
          'predvars_gradtheta', 'covxhalf_gradtheta'):
    out['p_' + k] = p[k]
print(out['hyp'].shape, out['hypind'], float(np.abs(p['mean'] - truth(tn)).max()))
np.savez_compressed(os.path.join(HERE, 'pcsk.npz'), **out)
