"""Kernel-level timings on one B200 (CUDA events): DMMA GEMM vs cuBLAS DGEMM, batched Cholesky,
NLL+grad at the reference's n_h and at full n, predict + Woodbury at the C5 batch shapes.
Writes gpurun_out/microbench.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from surmise_b200 import ops  # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return float(np.median(ts)), float(np.min(ts))


out = {}
dev = torch.device('cuda')
g = torch.Generator(device='cuda').manual_seed(0)

# FP64 peak proxy: cuBLAS DGEMM
for n in (4096, 8192):
    A = torch.randn((n, n), dtype=torch.float64, device=dev, generator=g)
    B = torch.randn((n, n), dtype=torch.float64, device=dev, generator=g)
    t, tb = timeit(lambda: torch.matmul(A, B), reps=5)
    out[f'cublas_dgemm_{n}'] = {'s': t, 'tflops': 2 * n ** 3 / t / 1e12, 'tflops_best': 2 * n ** 3 / tb / 1e12}
    C = torch.empty((n, n), dtype=torch.float64, device=dev)
    for (akc, bkc) in ((True, True), (True, False), (False, False)):
        t, tb = timeit(lambda: ops.dgemm(A, B, C, akc, bkc, n, n, n, lds=(n, n, n)), reps=5)
        out[f'dmma_dgemm_{n}_{int(akc)}{int(bkc)}'] = {'s': t, 'tflops': 2 * n ** 3 / t / 1e12,
                                                        'tflops_best': 2 * n ** 3 / tb / 1e12}
    del A, B, C
print(json.dumps(out, indent=1), flush=True)

# batched Cholesky (+ inverse) at the BASELINE orders
for n, nb in ((2000, 20), (4000, 8), (8000, 8)):
    X = torch.randn((n, n + 16), dtype=torch.float64, device=dev, generator=g)
    A0 = torch.tril(X @ X.T / n + 1e-3 * torch.eye(n, dtype=torch.float64, device=dev))
    del X
    lda = (n + 7) // 8 * 8
    buf = torch.zeros((nb, n, lda), dtype=torch.float64, device=dev)

    def run_potrf():
        buf[:, :, :n] = A0
        ops.potrf_batched(buf, n, n)

    def fill_only():
        buf[:, :, :n] = A0
    tf, _ = timeit(fill_only, reps=3, warm=1)
    tp, _ = timeit(run_potrf, reps=3, warm=1)
    tp -= tf
    out[f'potrf_n{n}_b{nb}'] = {'s': tp, 'tflops': nb * n ** 3 / 3 / tp / 1e12}
    Linv, _ = ops.potrf_batched(buf, n, n)
    tt, _ = timeit(lambda: ops.trtri_batched(buf, Linv, n), reps=3, warm=1)
    out[f'trtri_n{n}_b{nb}'] = {'s': tt, 'tflops': nb * n ** 3 / 3 / tt / 1e12}
    del buf, Linv, A0
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1), flush=True)

# NLL(+grad) throughput
from synth import make_simulator  # noqa: E402
from surmise_b200.emulationmethods import PCGPwM_B200 as M  # noqa: E402
for n, d, nb in ((160, 8, 1), (160, 8, 20), (160, 8, 148), (200, 10, 64), (2000, 8, 20), (4000, 8, 8)):
    rng = np.random.default_rng(n)
    theta = rng.uniform(size=(n, d))
    G = rng.normal(size=(nb, n))
    mean, lo, hi, std = M.hyper_prior(theta, -10, -20, None)
    hyps = np.tile(mean, (nb, 1)) + 0.01 * rng.normal(size=(nb, mean.shape[0]))
    data = ops.GPData(theta, G, None)
    hyp_t = torch.from_numpy(hyps).cuda()
    for wg in (0, 1):
        t, _ = timeit(lambda: ops.gp_nll_grad(data, hyp_t, mean, std, 0.01, bool(wg)), reps=5, warm=2)
        flop = nb * (n ** 3 / 3 + (2 * n ** 3 / 3 if wg else 0))
        out[f'nll_n{n}_d{d}_b{nb}_grad{wg}'] = {'s': t, 'evals_per_s': nb / t, 'tflops': flop / t / 1e12}
print(json.dumps(out, indent=1), flush=True)

# predict + Woodbury at the C5 shapes (n=2000, d=8, k=20, m=500), worst case: every PC its own factor
n, d, k, m = 2000, 8, 20, 500
x, theta, f, truth = make_simulator(n, d, m, k, seed=0)
rng = np.random.default_rng(0)
mean, lo, hi, std = M.hyper_prior(theta, -10, -20, None)
hyps = np.tile(mean, (k, 1)) + 0.05 * rng.normal(size=(k, mean.shape[0]))
hyps[:, -1] = -10
pc = rng.normal(size=(n, k))
spc = {'offset': np.zeros(m), 'scale': np.ones(m), 'extravar': np.full(m, 1e-6)}
pcti = rng.normal(size=(m, k))
t0, _ = timeit(lambda: M.import_state(theta, pc, np.zeros((n, k)), hyps, np.arange(k), pcti, spc), reps=2, warm=1)
out['factorize_C2_k20'] = {'s': t0}
emulist, state = M.import_state(theta, pc, np.zeros((n, k)), hyps, np.arange(k), pcti, spc)
fitinfo = {'_b200': state, 'x': x}
from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C  # noqa: E402


class Emu:
    _info = fitinfo


cal = {'yvar': np.full(m, 0.0025)}
y = rng.normal(size=m)
for B in (1, 264, 2048, 16384):
    tn = torch.from_numpy(rng.uniform(size=(B, d))).cuda()
    for wg in (False, True):
        t, _ = timeit(lambda: C.loglik_device(cal, Emu, tn, y, x, return_grad=wg), reps=3, warm=1)
        flop = k * n * n * B * (2 if wg else 1)
        out[f'loglik_B{B}_grad{int(wg)}'] = {'s': t, 'theta_per_s': B / t, 'tflops': flop / t / 1e12}
# standalone Matern covariance (HBM-write bound): R only, R + hyper-parameter gradient planes, R + d/dx1 planes
for (n1, n2, dd) in ((8000, 8000, 10), (2000, 2000, 8), (2048, 2000, 8)):
    x1 = torch.from_numpy(rng.uniform(size=(n1, dd))).cuda()
    x2 = torch.from_numpy(rng.uniform(size=(n2, dd))).cuda()
    gam = torch.from_numpy(np.append(rng.normal(size=dd) * 0.3, -0.5)).cuda()
    for mode, planes in ((0, 0), (1, dd + 1), (2, dd)):
        if n1 * n2 * (planes + 1) * 8 > 8e9:
            continue
        t, tb = timeit(lambda: ops.covmat_device(x1, x2, gam, mode), reps=5)
        nbytes = 8.0 * n1 * n2 * (1 + planes) + 8.0 * dd * (n1 + n2)
        out[f'covmat_{n1}x{n2}_d{dd}_mode{mode}'] = {'s': t, 'GBps': nbytes / t / 1e9, 'GBps_best': nbytes / tb / 1e9,
                                                     'pairs_per_s': n1 * n2 / t}
print(json.dumps(out, indent=1), flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'microbench.json'), 'w'), indent=1)
