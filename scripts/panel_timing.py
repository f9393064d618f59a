import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from surmise_b200 import ops, _lib
n, nb = 2000, 20
X = torch.randn((n, n + 16), dtype=torch.float64, device='cuda')
A0 = torch.tril(X @ X.T / n + 1e-3 * torch.eye(n, dtype=torch.float64, device='cuda'))
buf = torch.zeros((nb, n, n), dtype=torch.float64, device='cuda')
st = torch.zeros(8, dtype=torch.int64, device='cuda')
_lib.lib.sb200_debug_panel_stamps(st.data_ptr())
for _ in range(2):
    buf[:] = A0
    ops.potrf_batched(buf, n, n)
torch.cuda.synchronize()
_lib.lib.sb200_debug_panel_stamps(None)
s = st.cpu().numpy()
names = ['stage + block load', 'phase1 factor', 'publish', 'phase2 inverse', 'writeback']   # slot 6 (phase 3) is not written by the last, row-less panel
print({nm: int(s[i + 1] - s[i]) for i, nm in enumerate(names)}, 'total to writeback', int(s[5] - s[0]), 'cycles')
