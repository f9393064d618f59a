"""Cycle accounting of the task-graph Cholesky (sb200_debug_dag_stats): share of CTA time spent waiting on dependencies,
in the 64x64 factorisations and in the off-diagonal solves.  python scripts/dag_stats.py [n batch]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from surmise_b200 import ops
from surmise_b200._lib import lib

shapes = [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) > 2 else [(2000, 20), (8000, 8)]
for n, nb in shapes:
    g = torch.Generator(device='cuda').manual_seed(0)
    X = torch.randn((n, n + 16), dtype=torch.float64, device='cuda', generator=g)
    A0 = torch.tril(X @ X.T / n + 1e-3 * torch.eye(n, dtype=torch.float64, device='cuda'))
    del X
    buf = torch.empty((nb, n, n), dtype=torch.float64, device='cuda')
    stats = torch.zeros((296, 8), dtype=torch.int64, device='cuda')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3):
        buf[:] = A0
        torch.cuda.synchronize()
        lib.sb200_debug_dag_stats(stats.data_ptr() if it == 2 else None)
        e0.record()
        ops.potrf_batched(buf, n, n)
        e1.record()
        torch.cuda.synchronize()
    lib.sb200_debug_dag_stats(None)
    s = stats.cpu().double()
    s = s[s[:, 0] > 0]
    if s.shape[0] == 0:
        print('n=%d batch=%d  %.3f ms  (recursive path, no task-graph statistics)' % (n, nb, e0.elapsed_time(e1)))
        continue
    tot = s[:, 0].sum()
    nd = max(float(s[:, 7].sum()), 1.0)
    print('   per diagonal task: finish %.0f cycles, of which factor+inverse %.0f, rows below %.0f'
          % (s[:, 2].sum() / nd, s[:, 5].sum() / nd, s[:, 6].sum() / nd))
    print('n=%d batch=%d  %.3f ms  %.2f TFLOP/s  ctas=%d  cycles/cta mean %.0f max %.0f | wait %.1f%%  diag %.1f%%  solve %.1f%%  '
          'tasks/cta %.1f' % (n, nb, e0.elapsed_time(e1), nb * n ** 3 / 3 / e0.elapsed_time(e1) / 1e9, s.shape[0],
                              s[:, 0].mean(), s[:, 0].max(), 100 * s[:, 1].sum() / tot, 100 * s[:, 2].sum() / tot,
                              100 * s[:, 3].sum() / tot, s[:, 4].mean()))
    del buf, A0
