"""C5 through the calibrator on N ranks (PC-sharded factors), once with the NCCL all-gather and once with the
peer-memory exchange: torchrun --nproc-per-node N scripts/c5_exchange_ab.py [kept draws per chain]"""
import os, sys, json
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from surmise_b200.emulationmethods import PCGPwM_B200 as M
from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
from synth import make_observation

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
keep = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
w2 = bench.workload(seed=0)
d = w2['theta'].shape[1]
w2['theta_true'], y, yvar = make_observation(w2['truth'], d, seed=3)
for mode in ('0', 'auto', '0', 'auto'):
    os.environ['SB200_PEER_EXCHANGE'] = mode
    np.random.seed(0)
    f2 = {'method': 'PCGPwM_B200'}
    M.fit(f2, w2['x'], w2['theta'], w2['f'], distributed=False)
    res = bench.c5_run(C, f2, w2, y, yvar, torch, keep, shard_factors=True)
    if rank == 0:
        print(json.dumps({'ranks': world, 'mode': mode, 'wall_s': res['wall_s'], 'theta_per_s': res['theta_per_s_end_to_end'],
                          'chain': res['device_chain']}), flush=True)
    dist.barrier()
dist.destroy_process_group()
