"""Golden fit/predict of the reference's PCGPwImpute (build-container only, see make_golden.py).

    PYTHONPATH=/tmp/ref/src python tests/golden/make_golden_impute.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from surmise.emulation import emulator                      # noqa: E402
from synth import make_simulator                            # noqa: E402

out = {}
for tag, method in (('knn', 'KNN'), ('br', 'BayesianRidge')):
    x, theta, f, truth = make_simulator(80, 2, 14, 3, seed=41, missing=0.08)
    np.random.seed(41)
    emu = emulator(x, theta, f, method='PCGPwImpute', args={'completionmethod': method})
    fi = emu._info
    tn = np.random.default_rng(42).uniform(0.1, 0.9, size=(12, 2))
    p = emu.predict(x, tn)
    out[tag + '_f'] = f
    out[tag + '_fcompleted'] = fi['f']
    out[tag + '_hyp'] = np.array([e['hyp'] for e in fi['emulist']])
    out[tag + '_hypind'] = np.array([e['hypind'] for e in fi['emulist']])
    out[tag + '_mean'] = p.mean()
    out[tag + '_var'] = p.var()
    out[tag + '_desc'] = np.array(fi['param_desc'])
    print(tag, out[tag + '_hyp'].shape, out[tag + '_hypind'], float(np.abs(p.mean() - truth(tn)).max()))
out['x'], out['theta'], out['thetanew'] = x, theta, tn
np.savez_compressed(os.path.join(HERE, 'impute.npz'), **out)
