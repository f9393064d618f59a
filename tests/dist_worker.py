"""Worker functions for tests/test_dist_gloo.py (spawned processes, gloo backend, CPU only).

The sharding plumbing of surmise_b200/_dist.py is exercised with CPU stand-ins for the device
evaluators: the numpy oracle plays the role of the CUDA likelihood (test infrastructure only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def _problem():
    from synth import make_simulator
    from oracle import pcgpwm_oracle as po
    x, theta, f, truth = make_simulator(70, 2, 16, 4, seed=4)
    fi = {'f': f.T.copy(), 'mof': None, 'mofrows': None, 'epsilonPC': 0.001, 'epsilonImpute': 10e-6}
    po.standardize_f(fi)
    po.principal_components(fi)
    return theta, fi['pc']


def cpu_fit_one_factory(theta, pc, nh):
    """Same contract as _GPFitter.fit_one, likelihood evaluated by the oracle on the host."""
    import scipy.optimize as spo
    from oracle import pcgpwm_oracle as po
    mean, lo, hi, std = po.hyper_regulariser(theta, -10, -20)
    p = mean.shape[0]
    bar = po.share_threshold(p)

    def fit_one(j, cand, cand_ids, rows):
        sub = {'theta': theta[rows], 'g': pc[rows, j], 'gvar': np.zeros(len(rows)), 'sig2ofconst': 0.01,
               'hypregmean': mean, 'hypregstd': std}
        hyp, lead = mean, -1
        L0 = po.negloglik_and_grad_chol(mean, sub, False)[0]
        if cand is not None:
            for c in range(cand.shape[0]):
                L1 = po.negloglik_and_grad_chol(cand[c], sub, False)[0]
                if L1 < L0:
                    hyp, L0, lead = cand[c], L1, cand_ids[c]

        def scaled(v):
            val, grad, _ = po.negloglik_and_grad_chol(mean + v * std, sub, True)
            return val, grad * std
        res = spo.minimize(scaled, (hyp - mean) / std, method='L-BFGS-B', jac=True, options={'gtol': 0.1},
                           bounds=spo.Bounds((lo - mean) / std, (hi - mean) / std))
        if lead > -0.5 and 2 * (L0 - res.fun) < bar:
            return hyp, lead
        return mean + res.x * std, -1
    return fit_one, p


def run_sweeps():
    from surmise_b200 import _dist
    theta, pc = _problem()
    nh = 40
    fit_one, p = cpu_fit_one_factory(theta, pc, nh)
    np.random.seed(123)
    comm = _dist.Comm()
    hyps, inds, rows = _dist.jacobi_sweeps(fit_one, pc.shape[1], p, comm,
                                           lambda j: np.random.choice(theta.shape[0], nh, replace=False))
    return hyps, inds, np.array(rows)


def run_sharded(B):
    from surmise_b200 import _dist
    rng = np.random.default_rng(0)
    theta = rng.uniform(size=(B, 3))

    def fn(th):                                   # stands in for the fused device log-likelihood (+ gradient)
        return np.hstack([np.sin(th.sum(1, keepdims=True)), np.cos(th) * th])
    return _dist.sharded_rows(fn, theta, _dist.Comm()), fn(theta)


def worker(rank, world, port, outdir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        hyps, inds, rows = run_sweeps()
        shard = {B: run_sharded(B) for B in (1, 2, 7, 64)}
        from surmise_b200 import _dist
        comm = _dist.Comm()
        ragged = comm.allgather_rows(np.full((rank + 1, 3), float(rank)), [r + 1 for r in range(world)])
        # gather_blocks: each rank owns a contiguous block of a batch of "factors" (here 5 blocks of 3 x 4 numbers)
        import torch
        sizes, starts = _dist.block_partition(5, world)
        lo, hi = int(starts[rank]), int(starts[rank + 1])
        full = torch.arange(5 * 12, dtype=torch.float64).reshape(5, 3, 4)
        blocks = _dist.gather_blocks(comm, full[lo:hi].clone() if hi > lo else None, sizes,
                                     torch.full((5, 3, 4), -1.0, dtype=torch.float64)).numpy()
        # owner layout: components follow the rank that owns their factor; every rank contributes rows for its own
        # components (here the component id repeated) and the gathered table, re-ordered, must be in component order
        fidx = np.array([0, 1, 0, 2, 3, 1, 4, 4, 2, 0, 3])
        plan = _dist.owner_plan(fidx, starts, rank)
        mine = torch.from_numpy(plan['pcs_local'].astype(np.float64))[:, None].repeat(1, 3)
        owner_table = comm.allgather_tensor(mine, plan['counts'])[torch.from_numpy(plan['inverse'])].numpy()
        np.savez(os.path.join(outdir, 'rank%d.npz' % rank), hyps=hyps, inds=inds, rows=rows, ragged=ragged,
                 blocks=blocks, owner_table=owner_table,
                 **{'got%d' % B: v[0] for B, v in shard.items()}, **{'want%d' % B: v[1] for B, v in shard.items()})
    finally:
        dist.destroy_process_group()
