"""GPU: the hot path at BASELINE.json's full sizes, checked through size-independent properties (the oracle does
not finish in seconds there): derivative consistency, batch / order invariance, interpolation, factor residuals."""
import numpy as np
import pytest

from gpu_util import rel, report
from synth import make_simulator, make_observation

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from surmise_b200 import ops as _ops
    return _ops


@pytest.fixture(scope='module')
def c2():
    """Config C2: n=2000, d=8, m=500, k=20 -- PCA'd scores, prior, hyper-parameters near the prior mean."""
    from surmise_b200 import _pca
    from surmise_b200.emulationmethods.PCGPwM_B200 import hyper_prior
    x, theta, f, truth = make_simulator(2000, 8, 500, 20, seed=0)
    spc, mof, mofrows = _pca.standardize(f.T, 0.001, 10e-6)
    pcs = _pca.principal_components(spc['fs'], spc['U'], spc['S'], mof, mofrows, 0.001, 10e-6)
    mean, lo, hi, std = hyper_prior(theta, -10, -20, None)
    rng = np.random.default_rng(1)
    k = pcs['pc'].shape[1]
    hyps = mean + 0.25 * std * rng.normal(size=(k, mean.shape[0]))
    hyps[:, -1] = -10.0 + 0.5 * rng.normal(size=k)
    return dict(x=x, theta=theta, f=f, truth=truth, spc=spc, pcs=pcs, mean=mean, std=std, hyps=hyps, k=k)


def test_c2_nll_gradient_is_the_derivative_of_the_nll(ops, c2):
    """Directional central differences of the batched NLL at full n against the analytic gradient."""
    data = ops.GPData(c2['theta'], np.ascontiguousarray(c2['pcs']['pc'].T), None)
    h0 = c2['hyps']
    nll, grad, info = ops.gp_nll_grad(data, h0, c2['mean'], c2['std'], 0.01, True)
    assert not info.any()
    rng = np.random.default_rng(2)
    v = rng.normal(size=h0.shape) * c2['std']
    v[:, -2] = 0.0                                   # h_v has no effect without imputation variance
    eps = 1e-5
    up, _, _ = ops.gp_nll_grad(data, h0 + eps * v, c2['mean'], c2['std'], 0.01, False)
    dn, _, _ = ops.gp_nll_grad(data, h0 - eps * v, c2['mean'], c2['std'], 0.01, False)
    fd = (up - dn) / (2 * eps)
    an = np.sum(grad * v, axis=1)
    err = float(np.max(np.abs(fd - an) / (np.abs(an) + 1.0)))
    report('fullsize_nll_fd', n=2000, k=c2['k'], err=err)
    assert err < 1e-5


def test_c2_nll_is_independent_of_the_batch(ops, c2):
    """A component evaluated alone, in the full batch, and in a permuted batch gives the same bits."""
    pcT = np.ascontiguousarray(c2['pcs']['pc'].T)
    full = ops.gp_nll_grad(ops.GPData(c2['theta'], pcT, None), c2['hyps'], c2['mean'], c2['std'], 0.01, True)
    perm = np.random.default_rng(3).permutation(c2['k'])
    shuf = ops.gp_nll_grad(ops.GPData(c2['theta'], pcT[perm], None), c2['hyps'][perm], c2['mean'], c2['std'], 0.01, True)
    assert np.array_equal(full[0][perm], shuf[0]) and np.array_equal(full[1][perm], shuf[1])
    one = ops.gp_nll_grad(ops.GPData(c2['theta'], pcT[7], None), c2['hyps'][7:8], c2['mean'], c2['std'], 0.01, True)
    assert one[0][0] == full[0][7] and np.array_equal(one[1][0], full[1][7])


@pytest.fixture(scope='module')
def c2_state(c2):
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    k = c2['k']
    hyps, hypinds = c2['hyps'].copy(), np.arange(k)
    hyps[14:] = hyps[:k - 14]                        # 14 distinct factors, as a fitted C2 emulator has
    hypinds[14:] = np.arange(k - 14)
    spc = {kk: c2['spc'][kk] for kk in ('offset', 'scale', 'extravar')}
    emulist, state = M.import_state(c2['theta'], c2['pcs']['pc'], np.zeros_like(c2['pcs']['pc']), hyps, hypinds,
                                    c2['pcs']['pcti'], spc)
    return M, {'_b200': state, 'emulist': emulist, 'x': c2['x'], 'theta': c2['theta'], 'pcti': c2['pcs']['pcti'],
               'standardpcinfo': spc}


def test_c2_factors_invert_the_covariance(ops, c2, c2_state):
    """Linv' Linv R = I for the stored factors (R rebuilt by the standalone covariance kernel)."""
    import torch
    _, fi = c2_state
    st = fi['_b200']
    n = c2['theta'].shape[0]
    worst = 0.0
    for f in (0, 5, 13):
        j = int(np.where(st['fidx'].cpu().numpy() == f)[0][0])
        hyp = fi['emulist'][j]['hyp']
        R0, _ = ops.covmat_device(st['theta'], st['theta'], torch.from_numpy(hyp[:-2]).cuda(), 0)
        nug = fi['emulist'][j]['nug']
        R = (1 - nug) * R0 + nug * torch.eye(n, dtype=torch.float64, device='cuda')
        L = st['Linv'][f][:, :n]
        resid = (L.T @ (L @ R) - torch.eye(n, dtype=torch.float64, device='cuda')).abs().max()
        worst = max(worst, float(resid.cpu()))
    report('fullsize_factor_residual', n=n, resid=worst)
    assert worst < 1e-7                                  # cond(R) ~ 1e6 at these hyper-parameters


def test_c2_predict_interpolates_and_is_order_invariant(ops, c2, c2_state):
    """At design points the emulator reproduces the simulator output (up to the nugget and the discarded
    components) with near-zero PC variance; permuting the theta batch permutes the results bit for bit."""
    M, fi = c2_state
    rows = np.random.default_rng(4).choice(2000, 96, replace=False)
    p = {}
    M.predict(p, fi, c2['x'], c2['theta'][rows], return_grad=True)
    # exact algebra at a design point: (1 - nug) R0[row] pw = g[row] - nug pw[row]
    nug = np.array([e['nug'] for e in fi['emulist']])
    pw = np.array([e['pw'] for e in fi['emulist']]).T                     # (n, k)
    want = c2['pcs']['pc'][rows] - nug * pw[rows]
    err = float(np.max(np.abs(p['predvecs'] - want)) / np.max(np.abs(want)))
    sig2 = np.array([e['sig2'] for e in fi['emulist']])
    pvmax = float(np.max(p['predvars'] / sig2))
    report('fullsize_interpolation', identity_err=err, predvar_over_sig2=pvmax)
    assert err < 1e-8 and pvmax < 1e-3
    perm = np.random.default_rng(5).permutation(96)
    q = {}
    M.predict(q, fi, c2['x'], c2['theta'][rows][perm], return_grad=True)
    for key in ('mean', 'var', 'mean_gradtheta'):
        assert np.array_equal(p[key][:, perm], q[key])
    assert np.array_equal(p['covxhalf_gradtheta'][:, perm], q['covxhalf_gradtheta'])


def test_c2_loglik_gradient_is_the_derivative_of_the_loglik(ops, c2, c2_state):
    """Central differences of the fused log-likelihood over a B = 2048 population against its analytic gradient,
    and chunk invariance (the same theta evaluated in populations of different size)."""
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    M, fi = c2_state
    _, y, yvar = make_observation(c2['truth'], 8, seed=3)

    class Emu:
        _info = fi
    cal = {'yvar': yvar}
    rng = np.random.default_rng(6)
    th = rng.uniform(0.1, 0.9, size=(2048, 8))
    ll, dll = C.loglik_grad(cal, Emu, th, y, c2['x'])
    v = rng.normal(size=th.shape)
    eps = 1e-6
    up = C.loglik(cal, Emu, th + eps * v, y, c2['x'])[:, 0]
    dn = C.loglik(cal, Emu, th - eps * v, y, c2['x'])[:, 0]
    fd = (up - dn) / (2 * eps)
    an = np.sum(dll * v, axis=1)
    # the reference's analytic gradient (reproduced here) carries (1 - nug) twice on the mean part
    # (PCGPwM.py:236, 268), so it differs from the true derivative by O(nug) ~ 5e-5 relative
    scale = np.abs(an) + np.mean(np.abs(an))
    err = float(np.median(np.abs(fd - an) / scale))
    worst = float(np.max(np.abs(fd - an) / scale))
    report('fullsize_loglik_fd', B=2048, median=err, worst=worst)
    assert err < 5e-4 and worst < 2e-2
    sub = C.loglik_grad(cal, Emu, th[100:364], y, c2['x'])
    assert rel(sub[0], ll[100:364]) < 1e-12 and rel(sub[1], dll[100:364]) < 1e-10


def test_c4_order_factorisation_residual(ops):
    """n = 8000 (config C4): L L' = A and Linv L = I on two problems."""
    import torch
    n, nb = 8000, 2
    g = torch.Generator(device='cuda').manual_seed(1)
    bufs = []
    A0 = []
    for _ in range(nb):
        X = torch.randn((n, n + 32), dtype=torch.float64, device='cuda', generator=g)
        A = X @ X.T / n + 1e-3 * torch.eye(n, dtype=torch.float64, device='cuda')
        A0.append(A)
        bufs.append(torch.tril(A))
        del X
    buf = torch.stack(bufs)
    del bufs
    Linv, info = ops.potrf_batched(buf, n, n)
    ops.trtri_batched(buf, Linv, n)
    torch.cuda.synchronize()
    assert not info.cpu().numpy().any()
    e1 = e2 = 0.0
    eye = torch.eye(n, dtype=torch.float64, device='cuda')
    for b in range(nb):
        L = torch.tril(buf[b])
        e1 = max(e1, float(((L @ L.T - A0[b]).abs().max() / A0[b].abs().max()).cpu()))
        e2 = max(e2, float(((Linv[b][:, :n] @ L - eye).abs().max()).cpu()))
    report('chol_c4', n=n, LLt=e1, LinvL=e2)
    assert e1 < 1e-13 and e2 < 1e-8
