"""Golden chains for the sampler plugins, produced by the REAL reference samplers on analytic targets.

Build-container only (see make_golden.py for how /tmp/ref is built):

    PYTHONPATH=/tmp/ref/src python tests/golden/make_golden_samplers.py

Writes samplers.npz: the draws of surmise.utilitiesmethods.metropolis_hastings.sampler and LMC.sampler with numpy's
global generator seeded as recorded, on tests/synth.py's boxed_logpost / gaussian_logpost.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE)))

import surmise.utilitiesmethods.metropolis_hastings as ref_mh      # noqa: E402
import surmise.utilitiesmethods.LMC as ref_lmc                     # noqa: E402
from synth import gaussian_logpost, boxed_logpost, box_draws       # noqa: E402

out = {}
for step in ('normal', 'uniform'):
    np.random.seed(11)
    r = ref_mh.sampler(boxed_logpost, box_draws, numsamp=400, stepType=step, burnSamples=150)
    out['mh_%s_theta' % step] = r['theta']
    out['mh_%s_acc' % step] = r['acc_rate']
    out['mh_%s_lpostlist' % step] = r['lpostlist']
    print('MH', step, 'acc', r['acc_rate'], 'mean', r['theta'].mean(0))

np.random.seed(12)
r = ref_lmc.sampler(gaussian_logpost, box_draws, numsamp=500)
out['lmc_grad_theta'] = r['theta']
out['lmc_grad_logpost_shape'] = np.array(r['logpost'].shape)
print('LMC grad mean', r['theta'].mean(0), 'sd', r['theta'].std(0), r['logpost'].shape)

np.random.seed(13)
r = ref_lmc.sampler(lambda th: gaussian_logpost(th, return_grad=False), box_draws, numsamp=500)
out['lmc_nograd_theta'] = r['theta']
print('LMC nograd mean', r['theta'].mean(0), 'sd', r['theta'].std(0), r['logpost'].shape)
np.savez_compressed(os.path.join(HERE, 'samplers.npz'), **out)
