"""Per-launch DMMA GEMM throughput inside one C2 NLL+grad step (CUDA events around every launch)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from surmise_b200 import ops, _lib
w = bench.workload()
data = ops.GPData(w['theta'], np.ascontiguousarray(w['pc'].T), None)
hyp = torch.from_numpy(w['hyps']).cuda()
for _ in range(3):
    ops.gp_nll_grad(data, hyp, w['mean'], w['std'], 0.01, True)
torch.cuda.synchronize()
_lib.lib.sb200_gemm_profile_enable(1)
ops.gp_nll_grad(data, hyp, w['mean'], w['std'], 0.01, True)
torch.cuda.synchronize()
_lib.lib.sb200_gemm_profile_enable(0)
cap = 4096
ms = (ctypes.c_double * cap)(); fl = (ctypes.c_double * cap)()
n = _lib.lib.sb200_gemm_profile_dump(ms, fl, cap)
tot_ms = sum(ms[i] for i in range(n)); tot_fl = sum(fl[i] for i in range(n))
print('launches', n, 'gemm ms', round(tot_ms, 3), 'TF', round(tot_fl / tot_ms / 1e9, 2))
cum = 0
for i in range(n):
    cum += ms[i]
    print('%3d  %8.1f us  %8.3f GF  %6.2f TF   cum %.2f ms' % (i, ms[i] * 1e3, fl[i] / 1e9, fl[i] / ms[i] / 1e9 if ms[i] > 0 else 0, cum))
