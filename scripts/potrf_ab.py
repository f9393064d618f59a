"""potrf_batched timing at a few (n, batch) shapes; run with SB200_PANEL_FUSE=0/1 to compare the fused leaf pair."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from surmise_b200 import ops

for n, nb in [(2000, 1), (2000, 4), (2000, 20), (4000, 8), (8000, 1), (8000, 8)]:
    g = torch.Generator(device='cuda').manual_seed(0)
    X = torch.randn((n, n + 16), dtype=torch.float64, device='cuda', generator=g)
    A0 = torch.tril(X @ X.T / n + 1e-3 * torch.eye(n, dtype=torch.float64, device='cuda'))
    del X
    buf = torch.empty((nb, n, n), dtype=torch.float64, device='cuda')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for it in range(5):
        buf[:] = A0
        torch.cuda.synchronize()
        e0.record()
        ops.potrf_batched(buf, n, n)
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1))
    print('dag=%s fuse=%s n=%d batch=%d  %.3f ms  %.2f TFLOP/s' % (os.environ.get('SB200_POTRF_DAG', '1'), os.environ.get('SB200_PANEL_FUSE', '1'), n, nb, best, nb * n ** 3 / 3 / best / 1e9))
    del buf, A0
