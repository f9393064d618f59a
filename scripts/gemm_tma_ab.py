"""cp.async (tile 7) vs TMA-staged (tile 8: 64x64, tile 9: 128x64) DMMA GEMM on the shapes of the C2 step.
python scripts/gemm_tma_ab.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from surmise_b200 import ops
from surmise_b200._lib import lib

dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(0)


def run(name, a_kc, b_kc, M, N, K, batch, flags=0):
    A = torch.randn((batch, M, K) if a_kc else (batch, K, M), dtype=torch.float64, device=dev, generator=g)
    B = torch.randn((batch, N, K) if b_kc else (batch, K, N), dtype=torch.float64, device=dev, generator=g)
    if flags & 1: A = torch.tril(A) if a_kc else torch.triu(A)
    if flags & 2: A = torch.triu(A) if a_kc else torch.tril(A)
    C = torch.zeros((batch, M, N), dtype=torch.float64, device=dev)
    fl = 2.0 * M * N * K * batch * (0.5 if flags & (1 | 2 | 4 | 8) else 1.0) * (0.5 if flags & 16 else 1.0)
    out = []
    ref = None
    for tile in (7, 8, 9):
        best = 1e9
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.dgemm(A, B, C, a_kc, b_kc, M, N, K, batch=batch, flags=flags, tile=tile,
                      strides=(A.stride(0), B.stride(0), M * N), lds=(A.shape[2], B.shape[2], N))
            e1.record(); torch.cuda.synchronize()
            if it: best = min(best, e0.elapsed_time(e1))
        if ref is None: ref = C.clone()
        err = float((torch.tril(C) - torch.tril(ref)).abs().max()) if flags & 16 else float((C - ref).abs().max())
        out.append('tile %d: %.3f ms %.1f TF (diff %.1e)' % (tile, best, fl / best / 1e9, err))
    print('%-34s %s' % (name, ' | '.join(out)), 'tma launches', lib.sb200_tma_gemm_count())


run('square 4096 kc,kc', True, True, 4096, 4096, 4096, 1)
run('square 4096 rc,rc', False, False, 4096, 4096, 4096, 1)
run('lauum 2000 x20 rc,rc A_UPPER|C_LOWER', False, False, 2000, 2000, 2000, 20, flags=2 | 16)
run('trtri top 976x1024x1024 x20 kc,rc', True, False, 976, 1024, 1024, 20)
run('trtri A_LOWER 976x1024x976 x20', True, False, 976, 1024, 976, 20, flags=1)
run('trtri level 256 x80 kc,rc', True, False, 256, 256, 256, 80)
run('trtri level 64 x320 kc,rc', True, False, 64, 64, 64, 320)
run('predict 2000x264x2000 x14 kc,kc', True, True, 2000, 264, 2000, 14)
run('update 1000x1000x1024 x20 C_LOWER', True, True, 1000, 1000, 1024, 20, flags=16)
