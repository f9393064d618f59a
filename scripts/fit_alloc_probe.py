"""Why does the warm C2 plugin fit sometimes spend 0.1 s in its 'factor' phase?  Repeats the fit and prints, per call,
the phase clock, the allocator's cudaMalloc count and the garbage collector's state; then times a bare cudaMalloc of
the factor buffer's size."""
import gc
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import bench
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    w = bench.workload(seed=0)
    keep = os.environ.get('KEEP', '0') == '1'
    held = []
    for rep in range(8):
        np.random.seed(1)
        fi = {'method': 'PCGPwM_B200'}
        s0 = torch.cuda.memory_stats()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        M.fit(fi, w['x'], w['theta'], w['f'])
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        s1 = torch.cuda.memory_stats()
        ph = fi['_b200']['timing']
        print('rep %d wall %.4f factor %.4f search %.4f  cudaMallocs %d  reserved %.0f MB  gc %s' % (
            rep, wall, ph['factor'], ph['search'], s1.get('num_device_alloc', 0) - s0.get('num_device_alloc', 0),
            s1.get('reserved_bytes.all.current', 0) / 2**20, gc.get_count()), flush=True)
        if keep and rep % 2 == 0:
            held.append(fi)
        del fi
        if rep == 4:
            t0 = time.perf_counter()
            n = gc.collect()
            print('gc.collect: %d objects, %.4f s' % (n, time.perf_counter() - t0))
    torch.cuda.empty_cache()
    for mb in (64, 448, 448, 2048):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b = torch.empty(mb * 2**20, dtype=torch.uint8, device='cuda')
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        del b
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print('cudaMalloc %d MB: %.4f s, cudaFree %.4f s' % (mb, t1 - t0, t2 - t1))


if __name__ == '__main__':
    main()
