"""Per-phase CUDA-event times of one sharded evaluator launch (predict on the local factors, exchange, Woodbury) on N
ranks: torchrun --nproc-per-node N scripts/c5_phase_times.py"""
import os, sys, json
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from surmise_b200 import ops
from surmise_b200.emulationmethods import PCGPwM_B200 as M
from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
from synth import make_observation

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
w2 = bench.workload(seed=0)
d = w2['theta'].shape[1]
_, y, yvar = make_observation(w2['truth'], d, seed=3)
np.random.seed(0)
f2 = {'method': 'PCGPwM_B200'}
M.fit(f2, w2['x'], w2['theta'], w2['f'], distributed=False)
M.shard_state(f2)
class Emu: _info = f2
dl = C.DeviceLogpost({'yvar': yvar}, Emu, w2['x'], y, None)
B = 264
ev = dl.evaluator(B, d)
ev.thetap.copy_(torch.rand((B, d), dtype=torch.float64, device='cuda', generator=torch.Generator(device='cuda').manual_seed(1)))
for _ in range(5):
    ev.launch()
torch.cuda.synchronize(); dist.barrier()
names = ['predict_local', 'exchange', 'woodbury', 'whole_launch']
acc = {k: 0.0 for k in names}
reps = 40
E = lambda: torch.cuda.Event(enable_timing=True)
for _ in range(reps):
    blk = ev._blocks[ev.peer['par'] if ev.peer else 0]
    e = [E() for _ in range(4)]
    e[0].record()
    if ev.kl:
        ops.gp_predict(ev.st['local'], ev.thetap, True, out=blk['pred_local'], population=B, out_ld=ev.kmax)
    e[1].record()
    if ev.peer is None:
        dist.all_gather_into_tensor(ev.gathered, blk['local'])
    else:
        from surmise_b200._lib import lib
        pr = ev.peer; par = pr['par']; pr['par'] ^= 1
        st = ops._stream(torch)
        lib.sb200_peer_allgather_post(blk['local'].data_ptr(), ev.chunk, pr['ptrs'].data_ptr(), (par * world + rank) * ev.chunk,
                                      pr['flag_off'], rank, world, pr['epoch'].data_ptr(), pr['arrive'].data_ptr(), st)
        lib.sb200_peer_allgather_wait(pr['buf'].data_ptr() + 8 * pr['flag_off'], rank, world, pr['epoch'].data_ptr(),
                                      pr['error'].data_ptr(), st)
    e[2].record()
    ops.woodbury_loglik(ev.consts, *blk['pred'], out=(ev.ll, ev.dll, ev.fired), row_valid=ev.valid, segments=ev.segments, seg_pos=ev.seg_pos)
    e[3].record()
    torch.cuda.synchronize()
    acc['predict_local'] += e[0].elapsed_time(e[1]); acc['exchange'] += e[1].elapsed_time(e[2])
    acc['woodbury'] += e[2].elapsed_time(e[3]); acc['whole_launch'] += e[0].elapsed_time(e[3])
    dist.barrier()
out = {k: round(v / reps * 1e3, 1) for k, v in acc.items()}
out.update(rank=rank, world=world, local_factors=int(ev.st['local']['Linv'].shape[0]), local_pcs=ev.kl, kmax=ev.kmax,
           exchange='peer' if ev.peer else 'nccl', unit='us')
gath = [None] * world
dist.all_gather_object(gath, out)
if rank == 0:
    for g in gath:
        print(json.dumps(g))
dist.destroy_process_group()
