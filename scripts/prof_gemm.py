"""A few representative DMMA GEMM launches for ncu --set full (see profiles/)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from surmise_b200 import ops
dev = torch.device('cuda')
n = 4096
A = torch.randn((n, n), dtype=torch.float64, device=dev)
B = torch.randn((n, n), dtype=torch.float64, device=dev)
C = torch.zeros((n, n), dtype=torch.float64, device=dev)
for rep in range(2):
    ops.dgemm(A, B, C, True, True, n, n, n, lds=(n, n, n))        # both k-contiguous (SYRK / panel form)
    ops.dgemm(A, B, C, False, False, n, n, n, lds=(n, n, n))      # both row-contiguous (lauum form)
# mid-size batched SYRK like the C2 top-level trailing update: M=N=1000, K=1024, batch 20, lower only
nb, M, K = 20, 1000, 1024
P = torch.randn((nb, M, K), dtype=torch.float64, device=dev)
Cb = torch.zeros((nb, M, M), dtype=torch.float64, device=dev)
for rep in range(2):
    ops.dgemm(P, P, Cb, True, True, M, M, K, alpha=-1.0, beta=1.0, flags=16, batch=nb, strides=(M * K, M * K, M * M),
              lds=(K, K, M))
torch.cuda.synchronize()
print('ok')
