"""bench.py -- PCGPwM / directbayeswoodbury hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Headline workload (BASELINE.json configs[1], "C2"): PCGPwM n=2000 design points, d=8, m=500 outputs,
k=20 principal components.  One STEP = the penalised negative log-likelihood AND its gradient for all
k components at full n (PCGPwM.__negloglik + __negloglikgrad, PCGPwM.py:850-907), i.e. k "evals".
    value : evals/s with theta / PC scores resident in HBM (device timed, CUDA events)
    e2e   : same call with HOST buffers: pinned host -> device copies of theta, scores, hyper-parameters and
            the device -> host read of (nll, grad) inside the timed region
Secondary metrics of BASELINE.json (fit wall-s, calibration theta-loglik evals/s) ride in "secondary".
With N > 1 (torchrun) every rank evaluates its own k components (independent PCs, weak scaling) and the
(k, p+1) results are all-gathered over NCCL, which is the only exchange the sharded fit needs.
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
METRIC = 'PCGPwM NLL+grad evals/s'
UNIT = 'evals/s'


def workload(seed=0, **shape):
    """Seeded synthetic simulator -> theta, PC scores, hyper-parameters, regulariser."""
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
    rng = np.random.default_rng(seed + 1)
    hyps = mean + 0.25 * std * rng.normal(size=(k, mean.shape[0]))
    hyps[:, -1] = -10.0 + 0.5 * rng.normal(size=k)
    return dict(x=x, theta=theta, f=f, truth=truth, pc=pcs['pc'], hyps=hyps, mean=mean, std=std, lo=lo, hi=hi, k=k,
                spc=spc, pcs=pcs)


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


def cpu_functions():
    """(negloglik, negloglikgrad, kind) of the CPU arm.

    kind 'reference': the unmodified reference's own PCGPwM.__negloglik / __negloglikgrad, when it is installed in the
    git-ignored baseline/_ref (pip install --target baseline/_ref; it travels to the GPU box with the snapshot).
    kind 'port': otherwise the oracle's restatement of the same eigh-based algorithm (oracle/pcgpwm_oracle.py)."""
    ref_dir = os.path.join(ROOT, 'baseline', '_ref')
    if os.path.isdir(os.path.join(ref_dir, 'surmise')):
        try:
            if ref_dir not in sys.path:
                sys.path.insert(0, ref_dir)
            import importlib
            mod = importlib.import_module('surmise.emulationmethods.PCGPwM')
            return getattr(mod, '__negloglik'), getattr(mod, '__negloglikgrad'), 'reference'
        except Exception:
            pass
    from oracle import pcgpwm_oracle as po
    return po.negloglik_eigh, po.negloglikgrad_eigh, 'port'


def cpu_nll_grad_sample(w, npcs=1):
    """Reference-algorithm NLL + gradient (eigh form, PCGPwM.py:850-907) for `npcs` components on the host."""
    nll_fn, grad_fn, _kind = cpu_functions()
    n = w['theta'].shape[0]
    t0 = time.perf_counter()
    out = []
    for j in range(npcs):
        info = {'theta': w['theta'], 'g': w['pc'][:, j], 'gvar': np.zeros(n), 'sig2ofconst': 0.01,
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


def run_reference(args):
    """CPU arm: the reference's NLL + gradient (baseline/_ref when installed, else the oracle's restatement of the
    same algorithm), numpy + LAPACK eigh on all host threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    w = workload()
    kind = cpu_functions()[2]
    use_all_host_threads()
    for _ in range(args.warmup if args.warmup < 2 else 1):      # one untimed pass is enough to page LAPACK in
        cpu_nll_grad_sample(w, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_nll_grad_sample(w, 1)
    dt = (time.perf_counter() - t0) / args.steps
    val = 1.0 / dt
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'C2: PCGPwM n=2000 d=8 m=500 k=20; NLL+grad at full n (PCGPwM.py:850-907)',
                       'note': 'each step = 1 of the 20 components (bounded sample); ' +
                               ("surmise's own PCGPwM.__negloglik + __negloglikgrad from baseline/_ref"
                                if kind == 'reference' else
                                'eigh-based numpy port of the reference algorithm (oracle/pcgpwm_oracle.py), '
                                'checked bit-exact against surmise')},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': host_threads(), 'kind': kind,
                             'sample': '1 component x 1 NLL+grad evaluation per step'},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


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
    from surmise_b200 import ops, _lib
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

    w = workload(seed=rank)                       # every rank owns k different components (weak scaling)
    n, d, k = w['theta'].shape[0], w['theta'].shape[1], w['k']
    p = d + 3
    theta_d = torch.from_numpy(w['theta']).to(dev)
    pc_d = torch.from_numpy(np.ascontiguousarray(w['pc'].T)).to(dev)
    hyp_d = torch.from_numpy(w['hyps']).to(dev)
    mean_d, std_d = torch.from_numpy(w['mean']).to(dev), torch.from_numpy(w['std']).to(dev)
    data = ops.GPData(theta_d, pc_d, None)
    gathered = [torch.empty((k, p + 2), dtype=torch.float64, device=dev) for _ in range(world)] if world > 1 else None

    def step_device():
        res = ops.gp_nll_grad(data, hyp_d, mean_d, std_d, 0.01, True, return_device=True)
        if world > 1:
            dist.all_gather(gathered, res)        # the (k, p+1) exchange of the PC-sharded fit
        return res

    # pinned host staging for the end-to-end leg
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    theta_h, pc_h, hyp_h = pin(w['theta']), pin(w['pc'].T), pin(w['hyps'])
    out_h = torch.empty((k, p + 2), dtype=torch.float64).pin_memory()
    h2d = theta_h.numel() * 8 + pc_h.numel() * 8 + hyp_h.numel() * 8
    d2h = out_h.numel() * 8

    def step_e2e():
        th = theta_h.to(dev, non_blocking=True)
        pc = pc_h.to(dev, non_blocking=True)
        hy = hyp_h.to(dev, non_blocking=True)
        res = ops.gp_nll_grad(ops.GPData(th, pc, None), hy, mean_d, std_d, 0.01, True, return_device=True)
        if world > 1:
            dist.all_gather(gathered, res)
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

    # roofline leg: per-launch CUDA-event timing of the dominant kernel (the DMMA GEMM) over K more steps
    import ctypes
    _lib.lib.sb200_gemm_profile_enable(1)
    barrier()
    for _ in range(K):
        step_device()
    barrier()
    _lib.lib.sb200_gemm_profile_enable(0)
    g_ms, g_fl = ctypes.c_double(0), ctypes.c_double(0)
    g_n = _lib.lib.sb200_gemm_profile_collect(ctypes.byref(g_ms), ctypes.byref(g_fl))
    clk.__exit__(None, None, None)

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
    if rank == 0 and not args.skip_secondary:
        secondary = secondary_metrics(w, dev, torch, ops, world)
    barrier()
    if world > 1 and not args.skip_secondary:
        # calibration throughput over the box: every rank evaluates its own slice of a theta population
        # (independent proposals); whole-job value = sum over ranks / max time over ranks
        tot = calibration_scaling(w, dev, torch, dist, world)
        if rank == 0:
            secondary['theta_loglik_evals_per_s_all_ranks'] = tot
    barrier()

    if rank == 0:
        evals = k * world
        traffic, traffic_note = None, None
        try:                                            # DRAM bytes of the DMMA GEMM launches of one step, from the committed ncu run
            tj = json.load(open(os.path.join(ROOT, 'profiles', 'dram_traffic_r1.json')))['kernels']['dgemm_dmma_kernel']
            traffic = (tj['dram_read_bytes'] + tj['dram_write_bytes']) / tj['launches']
            traffic_note = ('ncu dram__bytes_read+write summed over the %d DMMA GEMM launches of one step / launches '
                            '(profiles/dram_traffic_r1.json): %.2f GB per step' % (tj['launches'], (tj['dram_read_bytes'] + tj['dram_write_bytes']) / 1e9))
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
        line = {
            'metric': METRIC, 'value': evals / (ms_dev * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': K,
            'warmup': W, 'ms_per_step': ms_dev, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'C2: PCGPwM n=%d d=%d m=500 k=%d; one step = NLL+grad of all k components at full n '
                                   '(PCGPwM.py:850-907)' % (n, d, k),
                       'parallelism': 'independent principal components per rank; NCCL allgather of (k,p+1) results'
                       if world > 1 else 'single GPU, k components batched per launch',
                       'l2': 'working set per step (k augmented matrices + k inverses = %.2f GB) exceeds the 126 MB L2'
                             % (2 * k * n * n * 8 / 1e9),
                       'breakdown_info_nonzero': info_bad},
            'e2e': {'value': evals / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e},
            'gpu_launches': int(launches),
            'clocks': clk.summary(),
            'roofline': {'bound': 'tensor', 'kernel': 'dgemm_dmma_kernel (FP64 DMMA.8x8x4)', 'achieved': ach,
                         'peak': peak, 'unit': 'TFLOP/s', 'frac': (ach / peak) if (ach and peak) else None,
                         'traffic': traffic, 'traffic_note': traffic_note, 'launches_timed': g_n, 'gemm_ms_per_step': g_ms.value / K,
                         'gemm_share_of_step': (g_ms.value / K) / ms_dev,
                         'algorithmic_flop_per_step': g_fl.value / K,
                         'whole_step_tflops': k * float(n) ** 3 / (ms_dev * 1e-3) / 1e12,
                         'peak_source': 'cuBLAS DGEMM 8192^3 via torch.matmul, best of 5, measured in this run '
                                        '(MEASURED_PEAKS.json has no FP64 entry)'},
            'cpu_baseline': cpu,
            'secondary': secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def calibration_scaling(w, dev, torch, dist, world):
    """Log-likelihood (+gradient) evaluations/s summed over ranks: each rank fits the same C2 emulator (replicated
    factors) and evaluates its own B/world slice of a population of B = 2048 * world proposals."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    from synth import make_observation
    np.random.seed(0)
    fi = {'method': 'PCGPwM_B200'}
    M.fit(fi, w['x'], w['theta'], w['f'], fit_mode='parallel', distributed=False)
    _, y, yvar = make_observation(w['truth'], w['theta'].shape[1], seed=3)

    class Emu:
        _info = fi
    cal = {'yvar': yvar}
    B = 2048
    th = torch.from_numpy(np.random.default_rng(100 + dist.get_rank()).uniform(size=(B, w['theta'].shape[1]))).to(dev)
    out = {}
    for grad in (False, True):
        for _ in range(3):
            C.loglik_device(cal, Emu, th, y, w['x'], return_grad=grad)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            C.loglik_device(cal, Emu, th, y, w['x'], return_grad=grad)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out['grad' if grad else 'nograd'] = {'value': world * B / (float(ms.item()) * 1e-3), 'unit': 'theta/s',
                                             'B_per_rank': B, 'distinct_factors': int(fi['_b200']['Linv'].shape[0])}
    return out


def secondary_metrics(w, dev, torch, ops, world):
    """BASELINE.json's other two metrics on the same C2 problem, rank 0, one GPU."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    from synth import make_observation
    out = {}
    # (i) PCGPwM fit wall-seconds through the plugin entry point, host numpy in (reference defaults)
    fits = []
    for rep in range(3):                                   # value = the last (warm) call; all calls are reported
        np.random.seed(rep)
        fi = {'method': 'PCGPwM_B200'}
        t0 = time.perf_counter()
        M.fit(fi, w['x'], w['theta'], w['f'])
        torch.cuda.synchronize()
        fits.append(time.perf_counter() - t0)
    pfits = []
    for rep in range(2):                                   # Jacobi sweeps, all PCs' optimisers batched per launch
        np.random.seed(rep)
        fp = {'method': 'PCGPwM_B200'}
        t0 = time.perf_counter()
        M.fit(fp, w['x'], w['theta'], w['f'], fit_mode='parallel', distributed=False)
        torch.cuda.synchronize()
        pfits.append(time.perf_counter() - t0)
    tn = np.random.default_rng(5).uniform(0.05, 0.95, size=(256, w['theta'].shape[1]))
    pred = {}
    M.predict(pred, fp, w['x'], tn)
    rmse_p = float(np.sqrt(np.mean((pred['mean'] - w['truth'](tn)) ** 2)) / np.std(w['truth'](tn)))
    out['pcgpwm_fit_wall_s_parallel_mode'] = {'value': pfits[-1], 'first_call_s': pfits[0], 'unit': 's',
                                              'higher_is_better': False, 'nll_evals': fp['_b200']['nevals'],
                                              'holdout_rmse_over_sd': rmse_p,
                                              'note': "fit_mode='parallel': Jacobi sweeps, the k optimisers run in lock step and "
                                                      'their NLL+grad calls are one batched launch'}
    pred = {}
    M.predict(pred, fi, w['x'], tn)
    rmse = float(np.sqrt(np.mean((pred['mean'] - w['truth'](tn)) ** 2)) / np.std(w['truth'](tn)))
    out['pcgpwm_fit_wall_s'] = {'value': fits[-1], 'first_call_s': fits[0], 'all_calls_s': fits, 'unit': 's', 'higher_is_better': False,
                                'nll_evals': fi['_b200']['nevals'], 'distinct_factors': int(fi['_b200']['Linv'].shape[0]),
                                'holdout_rmse_over_sd': rmse,
                                'note': 'reference defaults (n_h = 160 subset for the optimiser, 3 sweeps); wall clock incl. '
                                        'host PCA and scipy L-BFGS-B; surmise CPU reference measured 104-109 s (BASELINE.md)'}
    # (ii) calibration theta log-likelihood evals/s on that emulator (+ gradient, as PTLMC/LMC request)
    _, y, yvar = make_observation(w['truth'], w['theta'].shape[1], seed=3)

    class Emu:
        _info = fi
    cal = {'yvar': yvar}
    rng = np.random.default_rng(9)
    for B in (264, 2048, 16384):
        th_h = torch.from_numpy(rng.uniform(size=(B, w['theta'].shape[1]))).pin_memory()
        th_d = th_h.to(dev)
        for grad in (False, True):
            def run_dev():
                return C.loglik_device(cal, Emu, th_d, y, w['x'], return_grad=grad)

            def run_e2e():
                t = th_h.to(dev, non_blocking=True)
                ll, dll, _ = C.loglik_device(cal, Emu, t, y, w['x'], return_grad=grad)
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
    out['calibration_note'] = ('emulator fitted above (k=%d PCs, %d distinct factors); B=264 is the PTLMC population of '
                               'config C5 (256 chains + 8 temperatures); reference CPU: 272 / 118 theta/s without / '
                               'with gradient (BASELINE.md)' % (len(fi['emulist']), int(fi['_b200']['Linv'].shape[0])))
    return out


if __name__ == '__main__':
    main()
