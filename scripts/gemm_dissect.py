"""Where the batched trailing update loses against the large square product: vary one factor at a time."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from surmise_b200 import ops
dev = torch.device('cuda')


def timeit(fn, reps=7, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e-3)
    return float(np.median(ts))


def tiles_lower(M, N):
    return sum(max(0, -(-(M - 64 * j) // 128)) for j in range(-(-N // 64)))


out = {}
nb = 20
for M, K in ((1000, 1000), (1024, 1024), (1000, 4000), (1000, 250), (2000, 1000)):
    P = torch.randn((nb, M, K), dtype=torch.float64, device=dev)
    Cb = torch.zeros((nb, M, M), dtype=torch.float64, device=dev)
    kt = -(-K // 16) * 16
    for name, beta, flags in (('lower_beta1', 1.0, 16), ('lower_beta0', 0.0, 16), ('full_beta1', 1.0, 0), ('full_beta0', 0.0, 0)):
        t = timeit(lambda: ops.dgemm(P, P, Cb, True, True, M, M, K, alpha=-1.0, beta=beta, flags=flags, batch=nb,
                                     strides=(M * K, M * K, M * M), lds=(K, K, M)))
        ntile = tiles_lower(M, M) if flags else (-(-M // 128)) * (-(-M // 64))
        execf = nb * ntile * 128 * 64 * kt * 2
        out['M%d_K%d_%s' % (M, K, name)] = {'us': round(t * 1e6, 1), 'exec_TF': round(execf / t / 1e12, 2),
                                           'waves': round(nb * ntile / 296, 2)}
    del P, Cb
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'gemm_dissect.json'), 'w'), indent=1)
