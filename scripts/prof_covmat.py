"""Standalone sb200_matern_covmat at the C4 shape (8000 x 8000, d = 10), R only and R + gradient planes: for ncu."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from surmise_b200 import ops
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.uniform(size=(8000, 10))).cuda()
g = torch.from_numpy(np.r_[rng.uniform(-1, 1, size=10), -3.0]).cuda()
for mode in (0, 1):
    ops.covmat_device(x, x, g, grad_mode=mode)
torch.cuda.synchronize()
print('ok')
