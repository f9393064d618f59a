"""End-to-end plugin fits at the BASELINE.json configurations (synthetic simulator of SURVEY 8d).
usage: run_configs.py C1 C2 C3 C4   -> gpurun_out/configs.json"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from synth import make_simulator, make_observation
from surmise_b200.emulationmethods import PCGPwM_B200 as M
from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
CFG = {'C1': dict(n=200, d=3, m=50, k=8, missing=0.0), 'C2': dict(n=2000, d=8, m=500, k=20, missing=0.0),
       'C3': dict(n=4000, d=8, m=1000, k=32, missing=0.10), 'C4': dict(n=8000, d=10, m=2000, k=64, missing=0.0)}
out = {}
path = os.path.join(ROOT, 'gpurun_out', 'configs.json')
if os.path.exists(path):
    out = json.load(open(path))
for name in sys.argv[1:]:
    c = CFG[name]
    x, theta, f, truth = make_simulator(c['n'], c['d'], c['m'], c['k'], seed=0, missing=c['missing'])
    res = {}
    for mode in ('reference', 'parallel'):
        walls = []
        for rep in range(2):                           # first call (library / allocator warm-up included), then warm
            np.random.seed(0)
            fi = {'method': 'PCGPwM_B200'}
            torch.cuda.reset_peak_memory_stats()
            t0 = time.perf_counter()
            M.fit(fi, x, theta, f, fit_mode=mode)
            torch.cuda.synchronize()
            walls.append(time.perf_counter() - t0)
        dt = walls[1]
        tn = np.random.default_rng(1).uniform(0.05, 0.95, size=(256, c['d']))
        tps = []
        for rep in range(2):
            t1 = time.perf_counter()
            p = {}
            M.predict(p, fi, x, tn, return_grad=True)
            tps.append(time.perf_counter() - t1)
        tp = tps[1]
        tr = truth(tn)
        rmse = float(np.sqrt(np.mean((p['mean'] - tr) ** 2)) / np.std(tr))
        res[mode] = {'fit_wall_s': dt, 'fit_wall_first_call_s': walls[0], 'predict256_grad_s': tp,
                     'predict256_grad_first_call_s': tps[0], 'k': len(fi['emulist']), 'factors': int(fi['_b200']['Linv'].shape[0]),
                     'nll_evals': fi['_b200']['nevals'], 'holdout_rmse_over_sd': rmse,
                     'peak_gpu_GB': torch.cuda.max_memory_allocated() / 1e9,
                     'clamped': int(sum(e['gvar_clamped'] for e in fi['emulist']))}
        print(name, mode, res[mode], flush=True)
        if name in ('C3', 'C4') and mode == 'reference':
            break          # one fit is enough at the big sizes
    _, y, yvar = make_observation(truth, c['d'], seed=3)
    class Emu: _info = fi
    cal = {'yvar': yvar}
    for B in (264, 2048):
        th = torch.from_numpy(np.random.default_rng(2).uniform(size=(B, c['d']))).cuda()
        for _ in range(2):
            C.loglik_device(cal, Emu, th, y, x, return_grad=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            C.loglik_device(cal, Emu, th, y, x, return_grad=True)
        torch.cuda.synchronize()
        res['loglik_grad_theta_per_s_B%d' % B] = B / ((time.perf_counter() - t0) / 3)
    out[name] = res
    print(name, {k: v for k, v in res.items() if k.startswith('loglik')}, flush=True)
    json.dump(out, open(path, 'w'), indent=1)
    del fi
    torch.cuda.empty_cache()
