"""Plugin fit (fit_mode='parallel') at config C4 or C3 on the ranks of the current torch.distributed job (or one
process): PCs and the final factors are sharded over the ranks.  usage: [torchrun ...] fit_scaling.py C4"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from synth import make_simulator
from surmise_b200.emulationmethods import PCGPwM_B200 as M
name = sys.argv[1] if len(sys.argv) > 1 else 'C4'
c = {'C2': dict(n=2000, d=8, m=500, k=20, missing=0.0), 'C3': dict(n=4000, d=8, m=1000, k=32, missing=0.10),
     'C4': dict(n=8000, d=10, m=2000, k=64, missing=0.0)}[name]
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
if world > 1:
    import torch.distributed as dist
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
x, theta, f, truth = make_simulator(c['n'], c['d'], c['m'], c['k'], seed=0, missing=c['missing'])
walls = []
for rep in range(2):
    np.random.seed(0)
    fi = {'method': 'PCGPwM_B200'}
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    M.fit(fi, x, theta, f, fit_mode='parallel')
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    walls.append(time.perf_counter() - t0)
if rank == 0:
    tn = np.random.default_rng(1).uniform(0.05, 0.95, size=(64, c['d']))
    p = {}
    M.predict(p, fi, x, tn)
    rmse = float(np.sqrt(np.mean((p['mean'] - truth(tn)) ** 2)) / np.std(truth(tn)))
    print(json.dumps({'config': name, 'world': world, 'fit_wall_first_s': walls[0], 'fit_wall_s': walls[1],
                      'k': len(fi['emulist']), 'factors': int(fi['_b200']['Linv'].shape[0]),
                      'holdout_rmse_over_sd': rmse}), flush=True)
if world > 1:
    dist.destroy_process_group()
