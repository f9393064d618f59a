"""bench.py -- PCGPwM / directbayeswoodbury hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N = 1  (BASELINE.json configs[1], "C2"): PCGPwM n=2000 design points, d=8, m=500 outputs, k=20 principal components.
    One STEP = the penalised negative log-likelihood AND its gradient for all k components at full n
    (PCGPwM.__negloglik + __negloglikgrad, PCGPwM.py:850-907), i.e. k "evals".
N > 1  (BASELINE.json configs[3], "C4", STRONG scaling): ONE fixed problem -- n=8000, d=10, m=2000, k=64 -- whose 64
    components are block-partitioned over the ranks (independent principal components); one STEP = NLL+grad
    of all 64 components at full n followed by the NCCL all-gather of the (k, p+2) result table, which is the only
    exchange the sharded fit needs.  Rank 0 also times the SAME 64-component step alone on its GPU
    (`strong_scaling.single_gpu_same_work`), so efficiency = value / (N * that) is on identical work.
    value : evals/s with theta / PC scores resident in HBM (device timed, CUDA events, max over ranks)
    e2e   : the same call with HOST buffers: pinned host -> device copies of theta, scores, hyper-parameters and the
            device -> host read of (nll, grad) inside the timed region
BASELINE.json's other metrics ride in "secondary": plugin fit wall-seconds at C2 / C3 / C4 (the C4 fit sharded over the
ranks when N > 1, with its phase times), calibration theta-log-likelihood evals/s, the C5 PT-LMC run through the
calibrator plugin, potrf / trtri throughput at n=8000.

--impl reference: the unmodified reference (baseline/_ref) on the host cores for the same metric and config; at N = 1
it also times the reference's own `emulator(..., method='PCGPwM')` fit at C2 and `loglik` / `loglik_grad` at B = 264, and
leaves them in .bench_reference.json so that the b200 arm run on the same box can quote same-box ratios.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

C2 = dict(n=2000, d=8, m=500, k=20)
C3 = dict(n=4000, d=8, m=1000, k=32, missing=0.10)
C4 = dict(n=8000, d=10, m=2000, k=64)
METRIC = 'PCGPwM NLL+grad evals/s'
UNIT = 'evals/s'
REF_CACHE = os.path.join(ROOT, '.bench_reference.json')
REF_EMU = os.path.join(ROOT, '.bench_reference_emu.npz')


def workload(seed=0, **shape):
    """Seeded synthetic simulator -> theta, PC scores, hyper-parameters, regulariser (host PCA)."""
    from synth import make_simulator
    from surmise_b200 import _pca
    from surmise_b200.emulationmethods.PCGPwM_B200 import hyper_prior
    cfg = dict(C2)
    cfg.update(shape)
    x, theta, f, truth = make_simulator(cfg['n'], cfg['d'], cfg['m'], cfg['k'], seed=seed)
    spc, mof, mofrows = _pca.standardize(f.T, 0.001, 10e-6)
    pcs = _pca.principal_components(spc['fs'], spc['U'], spc['S'], mof, mofrows, 0.001, 10e-6)
    mean, lo, hi, std = hyper_prior(theta, -10, -20, None)
    k = pcs['pc'].shape[1]
    hyps = trial_hyps(mean, std, k, seed)
    return dict(x=x, theta=theta, f=f, truth=truth, pc=pcs['pc'], hyps=hyps, mean=mean, std=std, lo=lo, hi=hi, k=k,
                spc=spc, pcs=pcs)


def trial_hyps(mean, std, k, seed):
    rng = np.random.default_rng(seed + 1)
    hyps = mean + 0.25 * std * rng.normal(size=(k, mean.shape[0]))
    hyps[:, -1] = -10.0 + 0.5 * rng.normal(size=k)
    return hyps


def workload_c4(seed=0):
    """The C4 problem (n=8000, d=10, m=2000, k=64): simulator on the host, standardisation / PCA on this rank's GPU."""
    from synth import make_simulator
    from surmise_b200 import _pca_device
    from surmise_b200.emulationmethods.PCGPwM_B200 import hyper_prior
    x, theta, f, truth = make_simulator(C4['n'], C4['d'], C4['m'], C4['k'], seed=seed)
    spc, mof, mofrows = _pca_device.standardize(f.T, 0.001, 10e-6)
    pcs = _pca_device.principal_components(spc, mof, mofrows, 0.001, 10e-6)
    mean, lo, hi, std = hyper_prior(theta, -10, -20, None)
    k = pcs['pc'].shape[1]
    return dict(x=x, theta=theta, f=f, truth=truth, pc=pcs['pc'], hyps=trial_hyps(mean, std, k, seed), mean=mean,
                std=std, k=k)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) >= 7:
                self.rows.append(parts)

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.rows)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference itself (baseline/_ref) or, when it is not installed, the oracle's restatement
# ------------------------------------------------------------------------------------------------------------------
def reference_package():
    """Puts baseline/_ref (the pip-installed, unmodified reference; git-ignored, travels with the snapshot) on the
    path.  Returns True when `import surmise` then resolves to it."""
    ref_dir = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_dir, 'surmise')):
        return False
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import surmise  # noqa: F401
        return True
    except Exception:
        return False


def cpu_functions():
    """(negloglik, negloglikgrad, kind) of the CPU arm.

    kind 'reference': the unmodified reference's own PCGPwM.__negloglik / __negloglikgrad from baseline/_ref.
    kind 'port': otherwise the oracle's restatement of the same eigh-based algorithm (oracle/pcgpwm_oracle.py)."""
    if reference_package():
        try:
            import importlib
            mod = importlib.import_module('surmise.emulationmethods.PCGPwM')
            return getattr(mod, '__negloglik'), getattr(mod, '__negloglikgrad'), 'reference'
        except Exception:
            pass
    from oracle import pcgpwm_oracle as po
    return po.negloglik_eigh, po.negloglikgrad_eigh, 'port'


def cpu_nll_grad_sample(w, npcs=1, rows=None):
    """Reference-algorithm NLL + gradient (eigh form, PCGPwM.py:850-907) for `npcs` components on the host
    (optionally on a row sub-sample of the design)."""
    nll_fn, grad_fn, _kind = cpu_functions()
    theta = w['theta'] if rows is None else w['theta'][rows]
    n = theta.shape[0]
    t0 = time.perf_counter()
    out = []
    for j in range(npcs):
        g = w['pc'][:, j] if rows is None else w['pc'][rows, j]
        info = {'theta': theta, 'g': g, 'gvar': np.zeros(n), 'sig2ofconst': 0.01,
                'hypregmean': w['mean'], 'hypregstd': w['std']}
        out.append((nll_fn(w['hyps'][j], info), grad_fn(w['hyps'][j], info)))
    return npcs / (time.perf_counter() - t0), out


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core it can, so the
    BLAS/LAPACK pools behind numpy are raised to the cores this process may run on."""
    try:
        from threadpoolctl import threadpool_limits
        n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
        threadpool_limits(limits=n)
    except Exception:
        pass


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        ths = [p.get('num_threads', 1) for p in threadpool_info() if p.get('user_api') == 'blas']
        return max(ths) if ths else os.cpu_count()
    except Exception:
        return os.cpu_count()


def reference_plugin_metrics(w):
    """BASELINE.json's other two metrics through the reference's own public API on the host (C2):
    `emulator(x, theta, f, method='PCGPwM')` wall-seconds (PCGPwM.py:12-137) and directbayeswoodbury `loglik` /
    `loglik_grad` (directbayeswoodbury.py:261-383) at the C5 population B = 264."""
    from synth import make_observation
    out = {}
    if not reference_package():
        return {'unavailable': 'baseline/_ref is not installed on this box'}
    import importlib
    from surmise.emulation import emulator
    np.random.seed(0)
    t0 = time.perf_counter()
    emu = emulator(w['x'], w['theta'], w['f'], method='PCGPwM')
    out['pcgpwm_fit_wall_s'] = {'value': time.perf_counter() - t0, 'unit': 's', 'higher_is_better': False,
                                'k': len(emu._info['emulist']),
                                'distinct_hyperparameter_sets': int(len({e['hypind'] for e in emu._info['emulist']})),
                                'call': "emulator(x, theta, f, method='PCGPwM'), reference defaults"}
    dbw = importlib.import_module('surmise.calibrationmethods.directbayeswoodbury')
    _, y, yvar = make_observation(w['truth'], w['theta'].shape[1], seed=3)
    th = np.random.default_rng(9).uniform(size=(264, w['theta'].shape[1]))
    cal = {'yvar': yvar}
    dbw.loglik(cal, emu, th[:8], y, w['x'])                      # untimed: pages everything in
    t0 = time.perf_counter()
    dbw.loglik(cal, emu, th, y, w['x'])
    t1 = time.perf_counter()
    dbw.loglik_grad(cal, emu, th, y, w['x'])
    t2 = time.perf_counter()
    out['theta_loglik_evals_per_s_B264_nograd'] = {'value': 264 / (t1 - t0), 'unit': 'theta/s'}
    out['theta_loglik_evals_per_s_B264_grad'] = {'value': 264 / (t2 - t1), 'unit': 'theta/s'}
    # leave the reference-FITTED C2 emulator (what PCGPwM_B200.adopt needs) and the reference's own outputs at the 264
    # proposals on disk: the b200 arm adopts it and reports config-size parity (secondary.c2_adopted_parity)
    try:
        ll = dbw.loglik(cal, emu, th, y, w['x'])
        ll2, dll = dbw.loglik_grad(cal, emu, th, y, w['x'])
        pr = emu.predict(w['x'], th)._info
        fi, spc = emu._info, emu._info['standardpcinfo']
        np.savez(REF_EMU, theta=fi['theta'], pc=fi['pc'], pcti=fi['pcti'], unscaled_pcstdvar=fi['unscaled_pcstdvar'],
                 eta=fi['eta'], dampalpha=fi['dampalpha'], offset=spc['offset'], scale=spc['scale'],
                 extravar=spc['extravar'], hyp=np.array([e['hyp'] for e in fi['emulist']]),
                 hypind=np.array([e['hypind'] for e in fi['emulist']]).astype(int),
                 condR=np.array([np.linalg.cond(fi['emulist'][j]['R']) for j in sorted({e['hypind'] for e in fi['emulist']})]),
                 thetanew=th, y=y, yvar=yvar, loglik=ll, loglik_g=ll2, dloglik=dll, mean=pr['mean'], var=pr['var'],
                 predvecs=pr['predvecs'], predvars=pr['predvars'])
    except Exception as exc:
        out['emulator_dump_error'] = repr(exc)
    return out


def run_reference(args):
    """CPU arm: the reference's NLL + gradient (baseline/_ref when installed, else the oracle's restatement of the
    same algorithm), numpy + LAPACK eigh on all host threads, on the config the b200 arm uses at this --gpus."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    kind = cpu_functions()[2]
    use_all_host_threads()
    what = ("surmise's own PCGPwM.__negloglik + __negloglikgrad from baseline/_ref" if kind == 'reference' else
            'eigh-based numpy port of the reference algorithm (oracle/pcgpwm_oracle.py), checked bit-exact against surmise')
    secondary = {}
    if args.gpus <= 1:
        w = workload()
        rows, scale, extrap = None, 1.0, False
        cfg = {'workload': 'C2: PCGPwM n=2000 d=8 m=500 k=20; NLL+grad at full n (PCGPwM.py:850-907)',
               'note': 'each step = 1 of the 20 components (bounded sample); ' + what}
        sample = '1 component x 1 NLL+grad evaluation per step'
    else:
        # the b200 arm's N > 1 config is C4 (n = 8000): one reference evaluation there is two 8000 x 8000 eigh plus
        # 13 n^3 products, minutes per step; the bounded sample is the same evaluation on a seeded 2000-row sub-sample
        # of the C4 design, scaled by (8000/2000)^3 -- the cost law of eigh and of the gradient's products
        from synth import make_simulator
        from surmise_b200.emulationmethods.PCGPwM_B200 import hyper_prior
        x, theta, f, truth = make_simulator(C4['n'], C4['d'], 64, 8, seed=0)       # theta as the b200 arm draws it
        rng = np.random.default_rng(1)
        pc = np.sin(theta @ rng.normal(size=(C4['d'], 1))) + 0.01 * rng.normal(size=(C4['n'], 1))
        mean, lo, hi, std = hyper_prior(theta, -10, -20, None)
        w = dict(theta=theta, pc=pc, mean=mean, std=std, hyps=trial_hyps(mean, std, 1, 0))
        rows = np.sort(np.random.default_rng(2).choice(C4['n'], 2000, replace=False))
        scale, extrap = (C4['n'] / 2000.0) ** 3, True
        cfg = {'workload': 'C4: PCGPwM n=8000 d=10 m=2000 k=64; NLL+grad at full n (PCGPwM.py:850-907)',
               'note': 'EXTRAPOLATED: each step = 1 component on a 2000-row sub-sample of the C4 design, its time '
                       'multiplied by (8000/2000)^3 = 64 (cost law of eigh and of the n^3 gradient products); ' + what}
        sample = '1 component, 2000 of 8000 rows, time x 64 (extrapolated)'
    for _ in range(1 if args.warmup >= 1 else 0):                # one untimed pass is enough to page LAPACK in
        cpu_nll_grad_sample(w, 1, rows)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_nll_grad_sample(w, 1, rows)
    dt = (time.perf_counter() - t0) / args.steps * scale
    val = 1.0 / dt
    if args.gpus <= 1 and not args.skip_secondary:
        try:
            secondary = reference_plugin_metrics(w)
        except Exception as exc:                                  # the headline must survive a failure here
            secondary = {'error': repr(exc)}
        try:
            json.dump({'when': time.time(), 'cores': host_threads(), 'nll_grad_evals_per_s': val,
                       'secondary': secondary}, open(REF_CACHE, 'w'))
        except OSError:
            pass
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
            'scaling': 'weak' if args.gpus <= 1 else 'strong', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': cfg, 'extrapolated': extrap,
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': host_threads(), 'kind': kind, 'sample': sample},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'secondary': secondary}
    print(json.dumps(line), flush=True)


def reference_cache():
    """Same-box numbers left by `--impl reference` (run by the driver right before this arm), if fresh."""
    try:
        c = json.load(open(REF_CACHE))
        if time.time() - c.get('when', 0) < 6 * 3600:
            return c
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------------------------
# b200 arm
# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--skip-secondary', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    from surmise_b200 import ops, _lib, _dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (B200); there is no CPU fallback for the product arm')

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    W = max(args.warmup, 3)
    K = args.steps
    strong = world > 1

    # N = 1: C2, all k components on this GPU.  N > 1: ONE C4 problem (same seed on every rank), components
    # block-partitioned over the ranks.
    w = workload_c4(seed=0) if strong else workload(seed=0)
    n, d, k = w['theta'].shape[0], w['theta'].shape[1], w['k']
    p = d + 3
    sizes, starts = _dist.block_partition(k, world)
    lo, hi = int(starts[rank]), int(starts[rank + 1])
    kmax = max(sizes)
    theta_d = torch.from_numpy(w['theta']).to(dev)
    pc_all = np.ascontiguousarray(w['pc'].T)                      # (k, n)
    pc_d = torch.from_numpy(pc_all[lo:hi]).to(dev)
    hyp_d = torch.from_numpy(w['hyps'][lo:hi]).to(dev)
    mean_d, std_d = torch.from_numpy(w['mean']).to(dev), torch.from_numpy(w['std']).to(dev)
    data = ops.GPData(theta_d, pc_d, None)
    padded = torch.zeros((kmax, p + 2), dtype=torch.float64, device=dev)
    table = torch.empty((world * kmax, p + 2), dtype=torch.float64, device=dev) if strong else None

    def exchange(res):
        """the (k, p+2) all-gather of the PC-sharded fit (NCCL): [nll | grad | info] of every component, everywhere"""
        padded[:res.shape[0]] = res
        dist.all_gather_into_tensor(table, padded)
        return table

    def step_device():
        res = ops.gp_nll_grad(data, hyp_d, mean_d, std_d, 0.01, True, return_device=True)
        return exchange(res) if strong else res

    # pinned host staging for the end-to-end leg
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    theta_h, pc_h, hyp_h = pin(w['theta']), pin(pc_all[lo:hi]), pin(w['hyps'][lo:hi])
    out_h = torch.empty((world * kmax if strong else k, p + 2), dtype=torch.float64).pin_memory()
    h2d = (theta_h.numel() + pc_h.numel() + hyp_h.numel()) * 8
    d2h = out_h.numel() * 8

    def step_e2e():
        th = theta_h.to(dev, non_blocking=True)
        pc = pc_h.to(dev, non_blocking=True)
        hy = hyp_h.to(dev, non_blocking=True)
        res = ops.gp_nll_grad(ops.GPData(th, pc, None), hy, mean_d, std_d, 0.01, True, return_device=True)
        if strong:
            res = exchange(res)
        out_h.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    for _ in range(W):
        step_device()
    l0 = _lib.lib.sb200_launch_count()
    clk = ClockSampler(local)
    clk.__enter__()                                   # sampled across the device-timed, e2e and roofline legs
    ms_dev = timed(step_device, K)
    launches = (_lib.lib.sb200_launch_count() - l0)
    res = step_device()
    torch.cuda.synchronize()
    info_bad = int((res[:, -1] != 0).sum().item())

    for _ in range(W):
        step_e2e()
    ms_e2e = timed(step_e2e, K)

    # sustained leg: the K timed steps of the headline can be short (C2: 9 ms each); the same step repeated for >= 2 s
    n_sus = max(K, int(np.ceil(2000.0 / ms_dev)))
    ms_sus = timed(step_device, n_sus)

    # roofline leg: per-launch CUDA-event timing of the dominant kernel (the DMMA GEMM) over K more steps
    import ctypes
    # two passes: the first fills the library's pool of timing events (creating two events per launch keeps the host
    # behind the device on the small launches, and the gaps land inside the brackets); the second is the measurement
    for _pass in range(2):
        _lib.lib.sb200_gemm_profile_enable(1)
        barrier()
        for _ in range(K):
            step_device()
        barrier()
        _lib.lib.sb200_gemm_profile_enable(0)
        g_ms, g_fl = ctypes.c_double(0), ctypes.c_double(0)
        g_n = _lib.lib.sb200_gemm_profile_collect(ctypes.byref(g_ms), ctypes.byref(g_fl))
        c_ms, c_fl = ctypes.c_double(0), ctypes.c_double(0)
        c_n = _lib.lib.sb200_kernel_profile_collect(1, ctypes.byref(c_ms), ctypes.byref(c_fl))   # task-graph Cholesky launches
    clk.__exit__(None, None, None)

    # strong scaling reference point: the SAME k-component step on one GPU (rank 0 alone, the others wait)
    single = None
    if strong:
        if rank == 0:
            single = single_gpu_same_work(w, theta_d, mean_d, std_d, dev, torch, ops)
        ops.release_workspaces()
        torch.cuda.empty_cache()
        barrier()

    # FP64 peak measured in this run (MEASURED_PEAKS.json carries no FP64 entry): cuBLAS DGEMM 8192^3
    peak = None
    if rank == 0:
        A = torch.randn((8192, 8192), dtype=torch.float64, device=dev)
        Bm = torch.randn((8192, 8192), dtype=torch.float64, device=dev)
        best = 1e9
        for i in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(A, Bm)
            e1.record()
            torch.cuda.synchronize()
            if i > 0:
                best = min(best, e0.elapsed_time(e1))
        peak = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
        del A, Bm
        torch.cuda.empty_cache()
    barrier()

    secondary = {}
    if not args.skip_secondary:
        try:
            if strong:
                secondary = secondary_multi(w, dev, torch, dist, ops, world, rank, peak)
            elif rank == 0:
                secondary = secondary_single(w, dev, torch, ops, peak)
        except Exception as exc:                       # the headline line must survive a failure in the extras
            import traceback
            secondary = {'error': repr(exc), 'traceback': traceback.format_exc()[-1500:]}
    barrier()

    if rank == 0:
        evals = k
        traffic, traffic_note = None, None
        try:                                            # DRAM bytes of the DMMA GEMM launches of one step, from the committed ncu run
            tj = json.load(open(os.path.join(ROOT, 'profiles', 'dram_traffic_r2.json')))['kernels']
            tj = next(v for k, v in tj.items() if k.startswith('dgemm'))
            if not strong:
                traffic = (tj['dram_read_bytes'] + tj['dram_write_bytes']) / tj['launches']
                traffic_note = ('ncu dram__bytes_read+write summed over the %d DMMA GEMM launches of one C2 step / launches '
                                '(profiles/dram_traffic_r2.json): %.2f GB per step'
                                % (tj['launches'], (tj['dram_read_bytes'] + tj['dram_write_bytes']) / 1e9))
        except Exception:
            pass
        ach = (g_fl.value / (g_ms.value * 1e-3) / 1e12) if g_ms.value > 0 else None
        cpu = None
        if world == 1:
            use_all_host_threads()
            cpu_nll_grad_sample(w, 1)                   # untimed: pages the reference and LAPACK in
            v, ref_out = cpu_nll_grad_sample(w, 1)
            got = res[0].cpu().numpy()
            nerr = abs(got[0] - ref_out[0][0]) / (abs(ref_out[0][0]) + n)
            gerr = float(np.max(np.abs(got[1:1 + p] - ref_out[0][1])) / np.max(np.abs(ref_out[0][1])))
            cpu = {'value': v, 'unit': UNIT, 'cores': host_threads(), 'kind': cpu_functions()[2],
                   'sample': '1 of %d components, one NLL+grad evaluation after one untimed call (eigh form, numpy/LAPACK)' % k,
                   'parity_vs_gpu': {'nll_rel': nerr, 'grad_rel': gerr}}
        name = 'C4' if strong else 'C2'
        line = {
            'metric': METRIC, 'value': evals / (ms_dev * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': K,
            'warmup': W, 'ms_per_step': ms_dev, 'higher_is_better': True, 'scaling': 'strong' if strong else 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': '%s: PCGPwM n=%d d=%d m=%d k=%d; one step = NLL+grad of all k components at full n '
                                   '(PCGPwM.py:850-907)' % (name, n, d, w['f'].shape[0], k),
                       'parallelism': ('ONE fixed problem, its %d components block-partitioned over %d ranks (%s per rank); '
                                       'NCCL all-gather of the (k,p+2) result table every step' % (k, world, sizes))
                       if strong else 'single GPU, k components batched per launch',
                       'l2': 'working set per step and rank (augmented matrices + inverses = %.2f GB) exceeds the 126 MB L2'
                             % (2 * (hi - lo) * n * n * 8 / 1e9),
                       'breakdown_info_nonzero': info_bad},
            'e2e': {'value': evals / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d * world,
                    'd2h_bytes_per_step': d2h * world, 'ms_per_step': ms_e2e},
            'sustained': {'value': evals / (ms_sus * 1e-3), 'unit': UNIT, 'steps': n_sus, 'seconds': n_sus * ms_sus * 1e-3,
                          'note': 'the same device-timed step repeated for >= 2 s'},
            'gpu_launches': int(launches) * world,
            'gpu_launches_per_rank': int(launches),
            'clocks': clk.summary(),
            'roofline': {'bound': 'tensor', 'kernel': 'dgemm_tma_kernel / dgemm_dmma_kernel (FP64 DMMA.8x8x4; TMA-staged where the operands allow)', 'achieved': ach,
                         'peak': peak, 'unit': 'TFLOP/s', 'frac': (ach / peak) if (ach and peak) else None,
                         'traffic': traffic, 'traffic_note': traffic_note, 'launches_timed': g_n, 'gemm_ms_per_step': g_ms.value / K,
                         'gemm_share_of_step': (g_ms.value / K) / ms_dev,
                         'algorithmic_flop_per_step': g_fl.value / K,
                         'whole_step_tflops': (hi - lo) * float(n) ** 3 / (ms_dev * 1e-3) / 1e12,
                         'whole_step_frac': ((hi - lo) * float(n) ** 3 / (ms_dev * 1e-3) / 1e12 / peak) if peak else None,
                         'measured_on': 'rank 0' if strong else 'the GPU',
                         'peak_source': 'cuBLAS DGEMM 8192^3 via torch.matmul, best of 5, measured in this run '
                                        '(MEASURED_PEAKS.json has no FP64 entry)'},
            'roofline_kernels': roofline_kernels(g_n, g_ms.value, g_fl.value, c_n, c_ms.value, c_fl.value, K, ms_dev, peak),
            'cpu_baseline': cpu,
            'secondary': secondary,
        }
        if strong:
            line['strong_scaling'] = {
                'single_gpu_same_work': single,
                'note': 'single_gpu_same_work.value is the same %d-component step run by rank 0 alone in this very run; '
                        'scaling efficiency on identical work = value / (n_gpus * single_gpu_same_work.value)' % k}
        ref = reference_cache()
        if ref is not None and world == 1:
            line['same_box_reference'] = ref
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def roofline_kernels(g_n, g_ms, g_fl, c_n, c_ms, c_fl, K, ms_step, peak):
    """Per-kernel roofline entries of the timed step.  The two FP64 tensor-core kernels are timed live with CUDA events
    around every launch (algorithmic flop / event time); the FP64-ALU / HBM kernels carry the figures of the committed
    ncu --set full captures (profiles/ncu_kernels_r2.json, same command line), marked as such."""
    out = []
    for name, n, ms, fl, what in (
            ('dgemm_tma_kernel + dgemm_dmma_kernel', g_n, g_ms, g_fl, 'triangular inverse, L^-T L^-1, predict products (and the recursion\'s '
                                                   'trailing updates where the task-graph kernel is not used)'),
            ('potrf_dag_kernel', c_n, c_ms, c_fl, 'batched Cholesky as one persistent task-graph launch (n^3/3 + rhs flop)')):
        if n and ms > 0:
            ach = fl / (ms * 1e-3) / 1e12
            out.append({'kernel': name, 'bound': 'tensor', 'what': what, 'launches_per_step': n / K, 'ms_per_step': ms / K,
                        'share_of_step': ms / K / ms_step, 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s',
                        'frac': ach / peak if peak else None, 'timed': 'CUDA events around every launch, this run'})
    try:
        extra = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_kernels_r2.json')))['kernels']
        for e in extra:
            e = dict(e)
            e['timed'] = 'ncu --set full capture committed under profiles/ (not this run)'
            out.append(e)
    except Exception:
        pass
    return out


def single_gpu_same_work(w, theta_d, mean_d, std_d, dev, torch, ops, chunk=32):
    """All k components of the strong-scaling problem on ONE GPU (chunks of 32 components keep the workspace at
    ~40 GB): 1 untimed + 2 timed steps."""
    k = w['k']
    pc = torch.from_numpy(np.ascontiguousarray(w['pc'].T)).to(dev)
    hy = torch.from_numpy(w['hyps']).to(dev)

    def step():
        return [ops.gp_nll_grad(ops.GPData(theta_d, pc[s:s + chunk], None), hy[s:s + chunk], mean_d, std_d, 0.01, True,
                                return_device=True) for s in range(0, k, chunk)]
    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {'value': k / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'steps': reps,
            'whole_step_tflops': k * float(theta_d.shape[0]) ** 3 / (ms * 1e-3) / 1e12}


def _fit_twice(M, torch, x, theta, f, barrier=None, **kw):
    """(first-call seconds, warm seconds = the better of two further calls, fitinfo of the last call).  All calls fit the
    SAME seeded problem; a warm call can still pay a cudaMalloc of a few hundred MB for the factor buffer (0.1 s at C2)
    when torch's caching allocator has split the block the previous call returned -- hence the better of two."""
    first, best, best_fi = None, None, None
    for rep in range(3):                              # first call, then the better of two warm calls
        np.random.seed(1)
        fi = {'method': 'PCGPwM_B200'}
        if barrier is not None:
            barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        M.fit(fi, x, theta, f, **kw)
        torch.cuda.synchronize()
        if barrier is not None:
            barrier()
        wall = time.perf_counter() - t0
        if rep == 0:
            first = wall
        elif best is None or wall < best:
            best, best_fi = wall, fi
        del fi
    return first, best, best_fi


def _holdout(M, fi, x, truth, d, B=128):
    tn = np.random.default_rng(5).uniform(0.05, 0.95, size=(B, d))
    pred = {}
    M.predict(pred, fi, x, tn)
    tr = truth(tn)
    return float(np.sqrt(np.mean((pred['mean'] - tr) ** 2)) / np.std(tr))


def _loglik_rates(C, cal, Emu, y, x, d, dev, torch, sizes=(264, 2048, 16384)):
    out = {}
    rng = np.random.default_rng(9)
    for B in sizes:
        th_h = torch.from_numpy(rng.uniform(size=(B, d))).pin_memory()
        th_d = th_h.to(dev)
        for grad in (False, True):
            def run_dev():
                return C.loglik_device(cal, Emu, th_d, y, x, return_grad=grad)

            def run_e2e():
                t = th_h.to(dev, non_blocking=True)
                ll, dll, _ = C.loglik_device(cal, Emu, t, y, x, return_grad=grad)
                r = torch.cat([ll[:, None], dll], 1) if grad else ll
                return r.cpu()
            res = {}
            for name, fn in (('value', run_dev), ('e2e', run_e2e)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                res[name] = B / (e0.elapsed_time(e1) * 1e-3 / reps)
            out['theta_loglik_evals_per_s_B%d_%s' % (B, 'grad' if grad else 'nograd')] = {
                'value': res['value'], 'e2e': res['e2e'], 'unit': 'theta/s'}
    return out


def c5_run(C, fi, w, y, yvar, torch, sampperchain, sampler='PTLMC_B200', **extra):
    """Config C5 through the calibration plugin: PT-LMC, 256 chains + 8 temperatures (B = 264 per step) on the C2
    emulator, uniform prior on [0,1]^d; `sampperchain` kept draws per chain (+50 % tuning steps)."""
    d = w['theta'].shape[1]
    stats = {'rows': 0}

    class prior:
        @staticmethod
        def lpdf(th):
            th = np.atleast_2d(th)
            return np.where(np.all((th >= 0) & (th <= 1), axis=1), 0.0, -np.inf).reshape(-1, 1)

        @staticmethod
        def lpdf_grad(th):
            return np.zeros_like(np.atleast_2d(th))

        @staticmethod
        def rnd(nn):
            return np.random.uniform(size=(nn, d))

        box = (np.zeros(d), np.ones(d))              # lets a device-resident sampler evaluate the prior itself

    class Emu:
        _info = fi
        _emulator__theta = w['theta']

    cal = {'thetaprior': prior, 'yvar': yvar}
    orig = C._evaluate

    def counted(fitinfo, emu, theta, y_, x_, return_grad):
        stats['rows'] += np.atleast_2d(theta).shape[0]
        return orig(fitinfo, emu, theta, y_, x_, return_grad)
    C._evaluate = counted
    try:
        np.random.seed(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        C.fit(cal, Emu, w['x'], y, sampler=sampler, numsamp=2000, numtemps=8, numchain=256,
              sampperchain=sampperchain, maxtemp=30, **extra)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    finally:
        C._evaluate = orig
    steps = int(np.ceil(sampperchain * 0.5)) + sampperchain
    dev_rows = int(cal.get('_b200sampler', {}).get('theta_evaluated_on_device', 0))
    rows = stats['rows'] + dev_rows
    post = cal['thetarnd']
    return {'sampler': sampler, 'mcmc_steps': steps, 'population': 264, 'wall_s': dt, 'theta_evaluated': rows,
            'theta_per_s_end_to_end': rows / dt, 'ms_per_mcmc_step_end_to_end': dt / steps * 1e3,
            'posterior_mean_abs_err': float(np.mean(np.abs(post.mean(0) - w['theta_true'][0]))),
            'distinct_factors': int(fi['_b200'].get('nf_total', fi['_b200']['Linv'].shape[0])),
            'device_chain': {k: v for k, v in cal.get('_b200sampler', {}).items()
                             if k in ('cuda_graph', 'steps_per_graph', 'accept_rate', 'swap_rate', 'preoptimise_s', 'exchange',
                                      'setup_and_capture_s', 'loop_s', 'ms_per_step_in_loop')}}


def adopted_parity(M, C, w):
    """Config-size parity on this box: the C2 emulator FITTED BY THE REFERENCE (left on disk by `--impl reference`, which
    the driver runs right before this arm) is adopted onto the device; prediction and log-likelihoods at the C5
    population (B = 264) are compared with the reference's own outputs.  None when the dump is absent or stale."""
    if not os.path.exists(REF_EMU) or time.time() - os.path.getmtime(REF_EMU) > 6 * 3600:
        return None
    g = dict(np.load(REF_EMU))
    k = g['hyp'].shape[0]
    fi = {'method': 'PCGPwM', 'theta': g['theta'], 'x': w['x'], 'pc': g['pc'], 'pcti': g['pcti'],
          'unscaled_pcstdvar': g['unscaled_pcstdvar'], 'eta': float(g['eta']), 'dampalpha': float(g['dampalpha']),
          'standardpcinfo': {'offset': g['offset'], 'scale': g['scale'], 'extravar': g['extravar']},
          'emulist': [{'hyp': g['hyp'][j], 'hypind': int(g['hypind'][j])} for j in range(k)]}
    M.adopt(fi)

    class Emu:
        _info = fi
    cal = {'yvar': g['yvar']}
    pred = {}
    M.predict(pred, fi, w['x'], g['thetanew'])
    ll = C.loglik(cal, Emu, g['thetanew'], g['y'], w['x'])
    ll2, dll = C.loglik_grad(cal, Emu, g['thetanew'], g['y'], w['x'])
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    return {'what': 'reference-fitted C2 emulator (n=2000, k=%d, %d distinct hyper-parameter sets) adopted with '
                    'PCGPwM_B200.adopt; B=264 proposals; relative to the reference outputs computed on this box' % (k, len(set(g['hypind'].tolist()))),
            'cond_R_max': float(g['condR'].max()),
            'mean_rel': rel(pred['mean'], g['mean']), 'var_rel': rel(pred['var'], g['var']),
            'predvecs_rel': rel(pred['predvecs'], g['predvecs']),
            'predvars_rel': float(np.max(np.abs(pred['predvars'] - g['predvars']) / g['predvars'])),
            'loglik_rel_per_theta_max': float(np.max(np.abs(ll - g['loglik']) / np.abs(g['loglik']))),
            'loglik_gradcall_rel_max': float(np.max(np.abs(ll2 - g['loglik_g']) / np.abs(g['loglik_g']))),
            'dloglik_rel': rel(dll, g['dloglik']),
            'bars': 'north_star: mean/var 1e-6, log-likelihood per theta 1e-9'}


def secondary_single(w, dev, torch, ops, peak):
    """BASELINE.json's other metrics on one GPU: plugin fit wall-seconds at C2 / C3 / C4, calibration theta
    log-likelihood evals/s, the C5 sampler run, potrf / trtri at n = 8000."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    from synth import make_observation, make_simulator
    out = {}
    ref = (reference_cache() or {}).get('secondary', {})
    d = w['theta'].shape[1]
    # (i) PCGPwM fit wall-seconds through the plugin entry point, host numpy in (reference defaults)
    first, warm, fi = _fit_twice(M, torch, w['x'], w['theta'], w['f'])
    out['pcgpwm_fit_wall_s'] = {'value': warm, 'first_call_s': first, 'unit': 's', 'higher_is_better': False,
                                'nll_evals': fi['_b200']['nevals'], 'distinct_factors': int(fi['_b200']['Linv'].shape[0]),
                                'holdout_rmse_over_sd': _holdout(M, fi, w['x'], w['truth'], d),
                                'phases_s': fi['_b200']['timing'], 'lbfgsb_driver': fi['_b200']['lbfgsb_driver'],
                                'note': 'C2, reference defaults (n_h = 160 subset for the optimiser, 3 Gauss-Seidel '
                                        'sweeps); wall clock incl. PCA and L-BFGS-B'}
    rfit = ref.get('pcgpwm_fit_wall_s', {}).get('value')
    if rfit:
        out['pcgpwm_fit_wall_s']['same_box_reference_s'] = rfit
        out['pcgpwm_fit_wall_s']['vs_cpu'] = rfit / warm
    firstp, warmp, fp = _fit_twice(M, torch, w['x'], w['theta'], w['f'], fit_mode='parallel', distributed=False)
    out['pcgpwm_fit_wall_s_parallel_mode'] = {'value': warmp, 'first_call_s': firstp, 'unit': 's', 'higher_is_better': False,
                                              'nll_evals': fp['_b200']['nevals'],
                                              'distinct_factors': int(fp['_b200']['Linv'].shape[0]),
                                              'holdout_rmse_over_sd': _holdout(M, fp, w['x'], w['truth'], d),
                                              'phases_s': fp['_b200']['timing'],
                                              'note': "C2, fit_mode='parallel': Jacobi sweeps, the k optimisers in lock step"}
    if rfit:
        out['pcgpwm_fit_wall_s_parallel_mode']['vs_cpu'] = rfit / warmp
    # (ii) calibration theta log-likelihood evals/s on the reference-mode emulator (+ gradient, as PTLMC/LMC request)
    theta_true, y, yvar = make_observation(w['truth'], d, seed=3)
    w['theta_true'] = theta_true

    class Emu:
        _info = fi
    cal = {'yvar': yvar}
    out.update(_loglik_rates(C, cal, Emu, y, w['x'], d, dev, torch))
    for key in ('theta_loglik_evals_per_s_B264_nograd', 'theta_loglik_evals_per_s_B264_grad'):
        rv = ref.get(key, {}).get('value')
        if rv:
            out[key]['same_box_reference'] = rv
            out[key]['vs_cpu'] = out[key]['e2e'] / rv
    out['calibration_note'] = ('C2 emulator fitted above with the reference defaults (k=%d PCs, %d distinct factors); '
                               'B=264 is the PTLMC population of config C5 (256 chains + 8 temperatures)'
                               % (len(fi['emulist']), int(fi['_b200']['Linv'].shape[0])))
    try:
        par = adopted_parity(M, C, w)
        if par is not None:
            out['c2_adopted_parity'] = par
    except Exception as exc:
        out['c2_adopted_parity'] = {'error': repr(exc)}
    # (iii) C5 through the calibrator plugin (bounded: 2000 kept draws per chain = 3000 MCMC steps of 264 chains)
    try:
        out['c5_ptlmc'] = c5_run(C, fi, w, y, yvar, torch, 2000)
    except Exception as exc:
        out['c5_ptlmc'] = {'error': repr(exc)}
    del fi, fp
    ops.release_workspaces()
    torch.cuda.empty_cache()
    # (iv) C3: 10 % missing outputs
    x3, th3, f3, tr3 = make_simulator(C3['n'], C3['d'], C3['m'], C3['k'], seed=0, missing=C3['missing'])
    first, warm, f3i = _fit_twice(M, torch, x3, th3, f3)
    out['c3_fit_wall_s'] = {'value': warm, 'first_call_s': first, 'unit': 's', 'k': len(f3i['emulist']),
                            'holdout_rmse_over_sd': _holdout(M, f3i, x3, tr3, C3['d']), 'phases_s': f3i['_b200']['timing'],
                            'config': 'C3: n=4000 d=8 m=1000 k=32, 10 % NaN, reference defaults'}
    del f3i
    ops.release_workspaces()
    torch.cuda.empty_cache()
    # (v) C4 on one GPU: plugin fit (parallel mode) and the strong-scaling step of `--gpus N`
    x4, th4, f4, tr4 = make_simulator(C4['n'], C4['d'], C4['m'], C4['k'], seed=0)
    first, warm, f4i = _fit_twice(M, torch, x4, th4, f4, fit_mode='parallel', distributed=False)
    out['c4_fit_wall_s'] = {'value': warm, 'first_call_s': first, 'unit': 's', 'k': len(f4i['emulist']),
                            'distinct_factors': int(f4i['_b200']['Linv'].shape[0]),
                            'holdout_rmse_over_sd': _holdout(M, f4i, x4, tr4, C4['d']), 'phases_s': f4i['_b200']['timing'],
                            'config': "C4: n=8000 d=10 m=2000 k=64, fit_mode='parallel', one GPU"}
    del f4i
    ops.release_workspaces()
    torch.cuda.empty_cache()
    w4 = workload_c4(seed=0)
    out['c4_nll_grad_single_gpu'] = single_gpu_same_work(
        w4, torch.from_numpy(w4['theta']).to(dev), torch.from_numpy(w4['mean']).to(dev),
        torch.from_numpy(w4['std']).to(dev), dev, torch, ops)
    out['c4_nll_grad_single_gpu']['note'] = 'the step `bench.py --gpus N` (N > 1) shards over the ranks, on one GPU'
    del w4
    ops.release_workspaces()
    torch.cuda.empty_cache()
    out.update(cholesky_rates(torch, ops, dev, peak))
    return out


def cholesky_rates(torch, ops, dev, peak):
    """Batched blocked Cholesky and triangular inverse at the C4 order (n = 8000, batch 8) and the C2 shape."""
    out = {}
    g = torch.Generator(device=dev).manual_seed(0)
    for n, nb in ((2000, 20), (8000, 8)):
        X = torch.randn((n, n + 16), dtype=torch.float64, device=dev, generator=g)
        A0 = torch.tril(X @ X.T / n + 1e-3 * torch.eye(n, dtype=torch.float64, device=dev))
        del X
        lda = (n + 7) // 8 * 8
        buf = torch.zeros((nb, n, lda), dtype=torch.float64, device=dev)

        def best_ms(fn, reps=3):
            ts = []
            for _ in range(reps + 1):
                buf[:, :, :n] = A0
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return float(np.median(ts[1:])), r
        tp, (Linv, _info) = best_ms(lambda: ops.potrf_batched(buf, n, n))
        buf[:, :, :n] = A0
        Linv, _info = ops.potrf_batched(buf, n, n)
        torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.trtri_batched(buf, Linv, n)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        tt = float(np.median(ts[1:]))
        fl = nb * float(n) ** 3 / 3
        for name, ms in (('potrf', tp), ('trtri', tt)):
            tf = fl / (ms * 1e-3) / 1e12
            out['%s_n%d_b%d' % (name, n, nb)] = {'ms': ms, 'tflops': tf, 'frac_of_fp64_peak': (tf / peak) if peak else None}
        del buf, Linv, A0
        torch.cuda.empty_cache()
    return out


def secondary_multi(w, dev, torch, dist, ops, world, rank, peak):
    """N > 1: the C4 PLUGIN fit sharded over the ranks (searches by component, full-data factors with their owners,
    PC-sharded predict), and calibration on the box: theta populations and the C5 sampler."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    from synth import make_observation
    out = {}

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()
    first, warm, fi = _fit_twice(M, torch, w['x'], w['theta'], w['f'], barrier=barrier, fit_mode='parallel')
    phases = fi['_b200']['timing']
    allp = [None] * world
    dist.all_gather_object(allp, phases)
    rmse = _holdout(M, fi, w['x'], w['truth'], w['theta'].shape[1])          # collective: PC-sharded predict
    out['c4_fit_wall_s'] = {'value': warm, 'first_call_s': first, 'unit': 's', 'higher_is_better': False,
                            'k': len(fi['emulist']), 'distinct_factors': int(fi['_b200']['nf_total']),
                            'factors_on_this_rank': int(fi['_b200']['Linv'].shape[0]),
                            'holdout_rmse_over_sd': rmse,
                            'phases_s_max_over_ranks': {key: max(ph.get(key, 0.0) for ph in allp) for key in phases},
                            'phases_s_rank0': phases,
                            'config': "C4 plugin fit, fit_mode='parallel', PCs and full-data factors sharded over %d ranks "
                                      "(factors='owner'); wall clock between barriers" % world}
    del fi
    ops.release_workspaces()
    torch.cuda.empty_cache()
    barrier()
    # calibration on the C2 emulator: (a) replicated factors, every rank its own theta population (weak);
    # (b) the C5 sampler with the factors sharded over the ranks (PC-sharded predict, one all-gather per step)
    w2 = workload(seed=0)
    d = w2['theta'].shape[1]
    theta_true, y, yvar = make_observation(w2['truth'], d, seed=3)
    w2['theta_true'] = theta_true
    np.random.seed(0)
    f2 = {'method': 'PCGPwM_B200'}
    M.fit(f2, w2['x'], w2['theta'], w2['f'], distributed=False)              # reference defaults, replicated

    class Emu:
        _info = f2
    cal = {'yvar': yvar}
    B = 2048
    th = torch.from_numpy(np.random.default_rng(100 + rank).uniform(size=(B, d))).to(dev)
    rates = {}
    for grad in (False, True):
        for _ in range(3):
            C.loglik_device(cal, Emu, th, y, w2['x'], return_grad=grad)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            C.loglik_device(cal, Emu, th, y, w2['x'], return_grad=grad)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        rates['grad' if grad else 'nograd'] = {'value': world * B / (float(ms.item()) * 1e-3), 'unit': 'theta/s',
                                               'B_per_rank': B, 'distinct_factors': int(f2['_b200']['Linv'].shape[0])}
    out['theta_loglik_evals_per_s_all_ranks'] = rates
    try:
        out['c5_ptlmc'] = c5_run(C, f2, w2, y, yvar, torch, 2000, shard_factors=True)
        out['c5_ptlmc']['ranks'] = world
    except Exception as exc:
        import traceback
        out['c5_ptlmc'] = {'error': repr(exc), 'traceback': traceback.format_exc()[-1200:]}
    return out


if __name__ == '__main__':
    main()
