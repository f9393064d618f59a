"""torchrun --nproc-per-node N scripts/dist_check.py : NCCL checks of the sharded paths on N GPUs (prints JSON lines).
  1. fit_mode='parallel', factors='owner': PCs sharded over ranks, full-data factors kept by their owners, PC-sharded
     predict == the same fit on one rank (hyper-parameter table bit-identical, predictions to rounding);
  2. factors='replicated' + theta population sharded over ranks == unsharded log-likelihood (+ gradient);
  3. shard_state(): a replicated emulator re-laid out by owner gives the same log-likelihoods, also through the
     fixed-buffer evaluator of the device-resident sampler (segmented Woodbury input, in-place all-gather);
  4. the device-resident PT-LMC chain with sharded factors == the same chain with replicated factors (same seed)."""
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
tn = np.random.default_rng(1).uniform(size=(37, 3))
_, y, yvar = make_observation(truth, 3, seed=8)


def emu_of(fi):
    class Emu:
        _info = fi
        _emulator__theta = theta
    return Emu


def say(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


# single-rank reference results (no collective: distributed=False)
np.random.seed(5)
f1 = {'method': 'PCGPwM_B200'}
M.fit(f1, x, theta, f, fit_mode='parallel', distributed=False)
p1 = {}
M.predict(p1, f1, x, tn, return_grad=True)
full = C._packed({'yvar': yvar}, emu_of(f1), tn, y, x, True)

# 1. owner layout
np.random.seed(5)
fo = {'method': 'PCGPwM_B200'}
M.fit(fo, x, theta, f, fit_mode='parallel')
po = {}
M.predict(po, fo, x, tn, return_grad=True)
hyp1 = np.array([e['hyp'] for e in f1['emulist']])
hypo = np.array([e['hyp'] for e in fo['emulist']])
ll_owner = C._packed({'yvar': yvar}, emu_of(fo), tn, y, x, True)
say(check='owner_fit', world=world, hyp_max_abs_diff=float(np.max(np.abs(hyp1 - hypo))),
    same_hypind=bool(np.array_equal([e['hypind'] for e in f1['emulist']], [e['hypind'] for e in fo['emulist']])),
    pw_max_rel=float(max(np.max(np.abs(a['pw'] - b['pw'])) / np.max(np.abs(a['pw'])) for a, b in zip(f1['emulist'], fo['emulist']))),
    predict_max_rel={k: float(np.max(np.abs(p1[k] - po[k])) / np.max(np.abs(p1[k]))) for k in ('mean', 'var', 'mean_gradtheta', 'covxhalf')},
    loglik_max_abs_diff=float(np.max(np.abs(full - ll_owner))), factors_total=int(fo['_b200']['nf_total']),
    factors_here=int(fo['_b200']['Linv'].shape[0]), timing=fo['_b200']['timing'])

# 2. replicated factors + theta sharding
np.random.seed(5)
fr = {'method': 'PCGPwM_B200'}
M.fit(fr, x, theta, f, fit_mode='parallel', factors='replicated')
shard = C._evaluate({'yvar': yvar, 'shard_theta': True}, emu_of(fr), tn, y, x, True)
say(check='theta_sharded', loglik_max_abs_diff=float(np.max(np.abs(full - shard))))

# 3. shard_state on the single-rank fit + the sampler's evaluator
f3 = dict(f1)
M.shard_state(f3)
ll3 = C._packed({'yvar': yvar}, emu_of(f3), tn, y, x, True)
dl = C.DeviceLogpost({'yvar': yvar}, emu_of(f3), x, y, None)
ev = dl.evaluator(tn.shape[0], 3)
ev.thetap.copy_(torch.from_numpy(tn).cuda())
ev.launch()
torch.cuda.synchronize()
got = torch.cat([ev.ll[:, None], ev.dll], 1).cpu().numpy()
say(check='shard_state', loglik_max_abs_diff=float(np.max(np.abs(full - ll3))),
    evaluator_max_rel=float(np.max(np.abs(got - full) / (np.abs(full) + 1e-300))), padded_components=int(ev.world * ev.kmax), exchange='peer memory' if ev.peer else 'nccl')


# 4. device-resident chain: sharded factors vs replicated factors, same seed
class prior:
    box = (np.zeros(3), np.ones(3))

    @staticmethod
    def lpdf(th):
        th = np.atleast_2d(th)
        return np.where(np.all((th >= 0) & (th <= 1), axis=1), 0.0, -np.inf).reshape(-1, 1)

    @staticmethod
    def lpdf_grad(th):
        return np.zeros_like(np.atleast_2d(th))

    @staticmethod
    def rnd(n):
        return np.random.uniform(size=(n, 3))


post = {}
for name, fi, extra in (('replicated', f1, {}), ('sharded', dict(f1), {'shard_factors': True}),
                        ('sharded_peer', dict(f1), {'shard_factors': True})):
    # the sharded chain once with the NCCL all-gather and once with the exchange over NVLink peer memory (csrc/peer.cu)
    os.environ['SB200_PEER_EXCHANGE'] = '1' if name == 'sharded_peer' else '0'
    cal = {'thetaprior': prior, 'yvar': yvar}
    np.random.seed(11)
    C.fit(cal, emu_of(fi), x, y, sampler='PTLMC_B200', numsamp=2000, numtemps=6, numchain=26, sampperchain=300, **extra)
    post[name] = (cal['thetarnd'], cal.get('_b200sampler'))
same = [None] * world
dist.all_gather_object(same, post['sharded'][0].tobytes())
say(check='device_chain', replicated_stats=post['replicated'][1], sharded_stats=post['sharded'][1],
    posterior_mean_replicated=post['replicated'][0].mean(0).tolist(), posterior_mean_sharded=post['sharded'][0].mean(0).tolist(),
    max_abs_diff_draws=float(np.max(np.abs(post['replicated'][0] - post['sharded'][0]))),
    peer_exchange_stats=post['sharded_peer'][1],
    max_abs_diff_draws_peer_exchange=float(np.max(np.abs(post['replicated'][0] - post['sharded_peer'][0]))),
    identical_on_all_ranks=bool(all(s == same[0] for s in same)))
dist.barrier()
dist.destroy_process_group()
