"""torchrun --nproc-per-node N scripts/dist_check.py : NCCL checks of the sharded paths on N GPUs.
  1. fit_mode='parallel' with PCs sharded over ranks == the same fit on one rank (hyper-parameter table);
  2. theta population sharded over ranks == unsharded log-likelihood (+ gradient), incl. the batch-wide flag."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from synth import make_simulator, make_observation  # noqa: E402
from surmise_b200.emulationmethods import PCGPwM_B200 as M  # noqa: E402
from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
x, theta, f, truth = make_simulator(400, 3, 40, 8, seed=3)
np.random.seed(5)
fi = {'method': 'PCGPwM_B200'}
M.fit(fi, x, theta, f, fit_mode='parallel')
hyp = np.array([e['hyp'] for e in fi['emulist']])
inds = np.array([e['hypind'] for e in fi['emulist']])
# reference: the same Jacobi fit computed by rank 0 alone (no process group visible -> world = 1)
gathered = [None] * world
dist.all_gather_object(gathered, (hyp, inds))
same_everywhere = all(np.array_equal(g[0], hyp) and np.array_equal(g[1], inds) for g in gathered)

_, y, yvar = make_observation(truth, 3, seed=8)


class Emu:
    _info = fi


tn = np.random.default_rng(1).uniform(size=(37, 3))
pred_sharded = {}
M.predict(pred_sharded, fi, x, tn, return_grad=True)       # factors came from their owner ranks
full = C._packed({'yvar': yvar}, Emu, tn, y, x, True)
shard = C._evaluate({'yvar': yvar, 'shard_theta': True}, Emu, tn, y, x, True)
res = {'rank': rank, 'world': world, 'fit_identical_on_all_ranks': bool(same_everywhere),
       'sharded_loglik_max_abs_diff': float(np.max(np.abs(full - shard))), 'k': int(hyp.shape[0])}
print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    # single-rank Jacobi fit for comparison (process group gone -> Comm() sees world = 1)
    np.random.seed(5)
    f1 = {'method': 'PCGPwM_B200'}
    M.fit(f1, x, theta, f, fit_mode='parallel')
    h1 = np.array([e['hyp'] for e in f1['emulist']])
    p1 = {}
    M.predict(p1, f1, x, tn, return_grad=True)
    print(json.dumps({'sharded_vs_single_rank_hyp_max_abs_diff': float(np.max(np.abs(h1 - hyp))),
                      'sharded_factor_predict_max_abs_diff': float(max(np.max(np.abs(p1[k] - pred_sharded[k]))
                                                                    for k in ('mean', 'var', 'mean_gradtheta'))),
                      'distinct_factors': int(fi['_b200']['Linv'].shape[0]),
                      'same_hypind': bool(np.array_equal(inds, [e['hypind'] for e in f1['emulist']]))}), flush=True)
