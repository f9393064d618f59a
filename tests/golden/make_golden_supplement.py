"""Golden outputs of the reference's PCGPwM.supplementtheta (build-container only, see make_golden.py).

    PYTHONPATH=/tmp/ref/src python tests/golden/make_golden_supplement.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from surmise.emulation import emulator                      # noqa: E402
import surmise.emulationmethods.PCGPwM as ref               # noqa: E402
from synth import make_simulator                            # noqa: E402

x, theta, f, truth = make_simulator(40, 2, 12, 3, seed=31, missing=0.1)
np.random.seed(31)
emu = emulator(x, theta, f, method='PCGPwM')
rng = np.random.default_rng(32)
choices = rng.uniform(size=(9, 2))
new = rng.uniform(size=(5, 2))
pending = rng.uniform(size=f.shape) < 0.05
out = {'theta': theta, 'choices': choices, 'new': new, 'pending': pending}
cases = {'plain': dict(costs=np.ones(9), kw={}),
         'costs': dict(costs=np.array([1, 2, -1, 0.5, 3, 1, 1, 2, 4.0]), kw={}),
         'pending': dict(costs=np.linspace(1, 2, 9),
                         kw={'pending': pending, 'includepending': True, 'costpending': 0.3})}
for name, c in cases.items():
    th, info = ref.supplementtheta(emu._info, 4, new, choices.copy(), c['costs'].copy(), None, **c['kw'])
    out[name + '_costs'] = c['costs']
    out[name + '_theta'] = th
    out[name + '_crit'] = info['crit']
    if 'obviatesugg' in info:
        out[name + '_obviate'] = info['obviatesugg']
    print(name, th[:, 0], info)
np.savez_compressed(os.path.join(HERE, 'supplement.npz'), **out)
