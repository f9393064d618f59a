"""Where the 'pca' phase of a PCGPwM_B200 fit goes (config C4 by default): every sub-step of _pca_device timed with a
device synchronisation on both sides.  Usage: python scripts/pca_breakdown.py [C2|C3|C4]"""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from synth import make_simulator
from surmise_b200 import _pca_device as P, ops

name = sys.argv[1] if len(sys.argv) > 1 else 'C4'
c = {'C2': dict(n=2000, d=8, m=500, k=20, missing=0.0), 'C3': dict(n=4000, d=8, m=1000, k=32, missing=0.10),
     'C4': dict(n=8000, d=10, m=2000, k=64, missing=0.0)}[name]
x, theta, f, truth = make_simulator(c['n'], c['d'], c['m'], c['k'], seed=0, missing=c['missing'])
f = np.ascontiguousarray(f.T)                     # (n, m) as the method module receives it
ops._torch()
dev = torch.device('cuda', 0)


def timed(label, fn, out):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize(); out[label] = round(1e3 * (time.perf_counter() - t0), 3)
    return r


for rep in range(2):
    t = {}
    timed('whole_standardize', lambda: P.standardize(f, 1e-3, 1e-3), t)
    spc, mof, mofrows = P.standardize(f, 1e-3, 1e-3)
    timed('whole_principal_components', lambda: P.principal_components(spc, mof, mofrows, 1e-3, 1e-3), t)
    if c['missing'] == 0:
        off = timed('host_mean', lambda: np.mean(f, axis=0), t)
        sc = timed('host_std', lambda: np.std(f, axis=0), t)
        fs_np = timed('host_scale', lambda: (f - off) / sc, t)
        fs = timed('h2d_pageable', lambda: torch.from_numpy(np.ascontiguousarray(fs_np)).to(dev), t)
        fd = timed('h2d_raw_f', lambda: torch.from_numpy(f).to(dev), t)
        timed('device_mean_std_scale', lambda: (fd - fd.mean(0)) / fd.std(0, unbiased=False), t)
        G = timed('gram_cublas', lambda: fs.T @ fs, t)
        lamU = timed('eigh_cusolver', lambda: torch.linalg.eigh(G), t)
        lam, U = lamU
        Up = U[:, -c['k']:]
        timed('resid_extravar', lambda: ((fs - (fs @ Up) @ Up.T) ** 2).mean(0).cpu(), t)
        timed('pc_scores', lambda: fs @ Up, t)
        timed('to_host_fs_U', lambda: ops.to_host([fs, U]), t)
        timed('to_host_fs_U_again', lambda: ops.to_host([fs, U]), t)
        Gm = torch.empty_like(G)
        timed('gram_sb200_dgemm', lambda: ops.dgemm(fs, fs, Gm, False, False, c['m'], c['m'], c['n'], lds=(c['m'], c['m'], c['m'])), t)
        t['gram_max_abs_diff'] = float((Gm - G).abs().max())
    print(json.dumps({'config': name, 'rep': rep, 'ms': t}))
