import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from surmise_b200 import ops, _lib
from surmise_b200.emulationmethods.PCGPwM_B200 import hyper_prior
for n, d in ((160, 8), (60, 3), (200, 10)):
    rng = np.random.default_rng(0)
    theta = rng.uniform(size=(n, d)); g = rng.normal(size=n)
    mean, lo, hi, std = hyper_prior(theta, -10, -20, None)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    th, gd, hy, mu, sd = t(theta), t(g), t(mean[None, :]), t(mean), t(std)
    nll = torch.zeros(1, dtype=torch.float64, device='cuda'); grad = torch.zeros((1, d + 3), dtype=torch.float64, device='cuda')
    info = torch.zeros(1, dtype=torch.int32, device='cuda'); st = torch.zeros(8, dtype=torch.int64, device='cuda')
    ws = torch.empty(8 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(3):
        rc = _lib.lib.sb200_gp_nll_small_profile(1, n, d, th.data_ptr(), gd.data_ptr(), hy.data_ptr(), mu.data_ptr(), sd.data_ptr(),
                                                 nll.data_ptr(), grad.data_ptr(), info.data_ptr(), st.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    s = st.cpu().numpy()
    names = ['load regs', 'sweep', 'store+nll']
    print(n, d, rc, {nm: int(s[i + 1] - s[i]) for i, nm in enumerate(names)}, 'sweep-kernel cycles', int(s[3] - s[0]))
