"""GPU parity tests proper: the CUDA hot path (through the C ABI) against the reference's golden
vectors (tests/golden/, generated from bandframework/surmise itself) and against the numpy oracle
on seeded inputs.  Tolerances: BASELINE.json north_star -- NLL/gradient 1e-9, predictive mean /
variance 1e-6, log-likelihood per theta 1e-9 (relative), widened only where cond(R) * eps exceeds
the bar (the reference's own eigh-vs-Cholesky agreement is bounded by that product)."""
import numpy as np
import pytest

from gpu_util import rel, report, dev
from oracle import pcgpwm_oracle as po
from oracle import woodbury_oracle as wo

pytestmark = pytest.mark.gpu
EPS = 2.2e-16


@pytest.fixture(scope='module')
def ops():
    from surmise_b200 import ops as o
    o._torch()
    return o


# ---------------------------------------------------------------------------------- covmat (K1-K3)
@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_covmat_golden(ops, golden, tag):
    import surmise_b200
    g = golden('covmat')
    x1, x2, gam = g[f'{tag}_x1'], g[f'{tag}_x2'], g[f'{tag}_gam']
    R = surmise_b200.covmat(x1, x2, gam)
    _, dRh = surmise_b200.covmat(x1, x2, gam, return_gradhyp=True)
    _, dRx = surmise_b200.covmat(x1, x2, gam, return_gradx1=True)
    e = (rel(R, g[f'{tag}_R']), rel(dRh, g[f'{tag}_dRhyp']), rel(dRx, g[f'{tag}_dRx1']))
    report('covmat_golden_' + tag, R=e[0], dRhyp=e[1], dRx1=e[2])
    assert max(e) < 1e-13
    assert dRh.shape == g[f'{tag}_dRhyp'].shape and dRx.shape == g[f'{tag}_dRx1'].shape


def test_covmat_1d_and_ragged(ops, golden):
    import surmise_b200
    g = golden('covmat')
    assert rel(surmise_b200.covmat(g['e_x1'], g['e_x2'], g['e_gam']), g['e_R']) < 1e-13
    rng = np.random.default_rng(4)
    for n1, n2, d in [(1, 1, 1), (33, 65, 16), (257, 31, 5), (1000, 999, 10)]:
        x1, x2, gam = rng.uniform(size=(n1, d)), rng.uniform(size=(n2, d)), rng.normal(size=d + 1)
        want, dwant = po.matern_covmat(x1, x2, gam, 'hyp')
        got, dgot = surmise_b200.covmat(x1, x2, gam, return_gradhyp=True)
        assert rel(got, want) < 1e-13 and rel(dgot, dwant) < 1e-13
    with pytest.raises(Exception):
        surmise_b200.covmat(rng.uniform(size=(4, 17)), rng.uniform(size=(4, 17)), np.zeros(18))   # d > 16


# ---------------------------------------------------------------------------------- NLL + gradient (K4, K5)
@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_nll_grad_golden(ops, golden, tag):
    g = golden('nll')
    data = ops.GPData(g[f'{tag}_theta'], g[f'{tag}_g'], g[f'{tag}_gvar'])
    hyps = g[f'{tag}_hyps']
    n = g[f'{tag}_g'].shape[0]
    nll, grad, info = ops.gp_nll_grad(data, hyps, g[f'{tag}_mean'], g[f'{tag}_std'], 0.01, True)
    nll0, _, info0 = ops.gp_nll_grad(data, hyps, g[f'{tag}_mean'], g[f'{tag}_std'], 0.01, False)
    assert np.all(info == 0) and np.all(info0 == 0)
    assert np.array_equal(nll, nll0)                       # value-only path is the same arithmetic
    # The distance to the reference's stored values must be explained by the reference's OWN distance from the exact
    # value (40-digit evaluation, tests/golden/arbiter.npz; its eigh-based inverse loses cond * eps), times 5, above a
    # floor of 1e-11 / 1e-10 -- never by a function of cond(R).  The bar itself (1e-9 against the exact value) is
    # asserted in test_gpu_arbiter.py.
    arb = golden('arbiter')
    for i in range(hyps.shape[0]):
        e_v = abs(nll[i] - g[f'{tag}_nll'][i]) / (abs(g[f'{tag}_nll'][i]) + n)
        e_g = rel(grad[i], g[f'{tag}_grad'][i])
        report('nll_golden_' + tag, i=i, nll=e_v, grad=e_g, cond=g[f'{tag}_cond'][i])
        assert e_v < max(1e-11, 5 * arb[f'nll_{tag}_ref_err'][i]), (tag, i, e_v)
        assert e_g < max(1e-10, 5 * arb[f'nllgrad_{tag}_ref_err'][i]), (tag, i, e_g)


@pytest.mark.parametrize('n,d,nb', [(160, 8, 7), (200, 10, 3), (61, 3, 4), (700, 8, 2)])
def test_nll_grad_vs_oracle_batched(ops, n, d, nb):
    rng = np.random.default_rng(n + d)
    theta = rng.uniform(size=(n, d))
    g = np.sin(theta @ rng.normal(size=d)) + 0.05 * rng.normal(size=n)
    gvar = rng.uniform(size=n) * (rng.uniform(size=n) < 0.2) * 0.3
    mean, LB, UB, std = po.hyper_regulariser(theta, -10, -20)
    hyps = np.array([mean] + [LB + rng.uniform(0.3, 0.7, size=mean.shape) * (UB - LB) for _ in range(nb - 1)])
    hyps[:, -1] = np.maximum(hyps[:, -1], -12)             # keep cond(R) where 1e-9 is attainable
    info = {'theta': theta, 'g': g, 'gvar': gvar, 'sig2ofconst': 0.01, 'hypregmean': mean, 'hypregstd': std}
    nll, grad, code = ops.gp_nll_grad(ops.GPData(theta, g, gvar), hyps, mean, std, 0.01, True)
    assert np.all(code == 0)
    for i in range(nb):
        v = po.negloglik_eigh(hyps[i], info)
        dv = po.negloglikgrad_eigh(hyps[i], info)
        cond = np.linalg.cond(po.assemble_gp_matrix(theta, hyps[i], gvar)[0])
        e_v, e_g = abs(nll[i] - v) / (abs(v) + n), rel(grad[i], dv)
        report('nll_oracle', n=n, d=d, i=i, nll=e_v, grad=e_g, cond=cond)
        assert e_v < 1e-9 and e_g < 1e-9                    # the north-star bar (measured: <= 3e-12 / 4e-11)


@pytest.mark.parametrize('n,d,nb', [(61, 3, 2), (130, 4, 3), (208, 5, 2), (700, 8, 2), (1000, 6, 3), (2000, 8, 2)])
def test_nll_grad_does_not_read_uninitialised_workspace(ops, n, d, nb):
    """The blocked path no longer clears its inverse-factor buffer (only the tiles the factorisation and the trtri
    levels write are ever read): with the workspace poisoned by NaN before the call the results are finite, identical
    to a call on a zeroed workspace, and within the bar of the oracle (PCGPwM.py:850-907)."""
    rng = np.random.default_rng(n)
    theta = rng.uniform(size=(n, d))
    G = np.sin(theta @ rng.normal(size=(d, nb))).T + 0.05 * rng.normal(size=(nb, n))
    GV = rng.uniform(size=(nb, n)) * (rng.uniform(size=(nb, n)) < 0.2) * 0.3
    mean, LB, UB, std = po.hyper_regulariser(theta, -10, -20)
    hyps = np.array([mean] + [LB + rng.uniform(0.3, 0.7, size=mean.shape) * (UB - LB) for _ in range(nb - 1)])
    hyps[:, -1] = np.maximum(hyps[:, -1], -12)
    data = ops.GPData(theta, G, GV)
    ops.force_blocked_path(True)
    try:
        ops.gp_nll_grad(data, hyps, mean, std, 0.01, True)                     # sizes the workspace
        out = []
        for byte in (0, 255):                                                   # 0xFF.. is a NaN in every double
            ops.workspace(1).fill_(byte)
            out.append(ops.gp_nll_grad(data, hyps, mean, std, 0.01, True))
    finally:
        ops.force_blocked_path(False)
    (v0, g0, c0), (v1, g1, c1) = out
    assert np.all(c0 == 0) and np.all(c1 == 0)
    assert np.all(np.isfinite(v1)) and np.all(np.isfinite(g1))
    assert np.array_equal(v0, v1) and np.array_equal(g0, g1)
    if n <= 1000:
        i = nb - 1
        info = {'theta': theta, 'g': G[i], 'gvar': GV[i], 'sig2ofconst': 0.01, 'hypregmean': mean, 'hypregstd': std}
        v, dv = po.negloglik_eigh(hyps[i], info), po.negloglikgrad_eigh(hyps[i], info)
        assert abs(v1[i] - v) / (abs(v) + n) < 1e-9 and rel(g1[i], dv) < 1e-9


@pytest.mark.parametrize('n,d', [(30, 2), (63, 3), (64, 5), (160, 8), (200, 10), (207, 16)])
def test_nll_small_fused_kernel_equals_blocked_path(ops, n, d):
    """The single-launch sweep kernel (n <= 207) and the blocked DMMA path are two routes to the same numbers."""
    assert ops.small_max_n() == 207
    rng = np.random.default_rng(n)
    theta = rng.uniform(size=(n, d))
    G = np.sin(theta @ rng.normal(size=(d, 3))).T + 0.05 * rng.normal(size=(3, n))
    GV = rng.uniform(size=(3, n)) * (rng.uniform(size=(3, n)) < 0.2) * 0.3
    mean, LB, UB, std = po.hyper_regulariser(theta, -10, -20)
    hyps = np.array([mean, LB + 0.4 * (UB - LB), LB + 0.6 * (UB - LB)])
    hyps[:, -1] = np.maximum(hyps[:, -1], -12)
    data = ops.GPData(theta, G, GV)
    a = ops.gp_nll_grad(data, hyps, mean, std, 0.01, True)
    a0 = ops.gp_nll_grad(data, hyps, mean, std, 0.01, False)
    ops.force_blocked_path(True)
    try:
        b = ops.gp_nll_grad(data, hyps, mean, std, 0.01, True)
    finally:
        ops.force_blocked_path(False)
    assert np.all(a[2] == 0) and np.all(b[2] == 0)
    assert np.array_equal(a[0], a0[0])
    e_v = float(np.max(np.abs(a[0] - b[0]) / (np.abs(b[0]) + n)))
    e_g = rel(a[1], b[1])
    report('nll_small_vs_blocked', n=n, d=d, nll=e_v, grad=e_g)
    assert e_v < 1e-10 and e_g < 1e-7


def test_nll_per_problem_data_and_breakdown(ops):
    """Different (g, gvar) per problem; an indefinite R must come back as info > 0, not a crash."""
    rng = np.random.default_rng(9)
    n, d, nb = 90, 4, 3
    theta = rng.uniform(size=(n, d))
    G = rng.normal(size=(nb, n))
    GV = np.abs(rng.normal(size=(nb, n))) * 0.1
    GV[2, 5] = -50.0                                       # forces a negative pivot
    mean, LB, UB, std = po.hyper_regulariser(theta, -10, -20)
    hyps = np.tile(mean, (nb, 1))
    nll, grad, code = ops.gp_nll_grad(ops.GPData(theta, G, GV), hyps, mean, std, 0.01, True)
    assert code[0] == 0 and code[1] == 0 and code[2] > 0
    for i in range(2):
        info = {'theta': theta, 'g': G[i], 'gvar': GV[i], 'sig2ofconst': 0.01, 'hypregmean': mean, 'hypregstd': std}
        assert abs(nll[i] - po.negloglik_eigh(hyps[i], info)) < 1e-9 * n
        assert rel(grad[i], po.negloglikgrad_eigh(hyps[i], info)) < 1e-8


# ---------------------------------------------------------------------------------- factorise + predict (K6-K9)
def _load_state(golden_npz):
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    g = golden_npz
    gv = po.damp_gvar(g['st_unscaled_pcstdvar'], 10, 0.3)
    spc = {'offset': g['st_offset'], 'scale': g['st_scale'], 'extravar': g['st_extravar']}
    emulist, state = M.import_state(g['st_theta'], g['st_pc'], gv, g['st_hyp'], g['st_hypind'], g['st_pcti'], spc)
    fitinfo = {'_b200': state, 'emulist': emulist, 'x': g['in_x'], 'theta': g['st_theta'], 'pcti': g['st_pcti'],
               'standardpcinfo': spc}
    return M, fitinfo


# Discrepancy against the REFERENCE's stored outputs, measured on a B200 (profiles/parity_report_r1.txt), per fixture
# and quantity.  The asserted bound is 5 x these (a regression of one digit fails), not a function of cond(R).  They are
# dominated by the reference's own eigh-based inverse: tests/golden/arbiter.npz (40-digit evaluation of the same
# formulas) puts the reference itself at exactly these distances from the exact values -- see test_gpu_arbiter.py.
# 'd8' (cond 9e4, the conditioning of config C2) carries BASELINE.json's bars themselves: 1e-9 on the log-likelihood.
MEASURED_VS_REFERENCE = {
    'c1': dict(pw=1.93e-8, sig2=8.2e-12, predvecs=6.4e-10, predvars=2.7e-9, dpv=9.4e-11, dpvar=3.1e-10, mean=4.1e-10,
               var=1.5e-10, dmean=7.0e-11, cxh=2.5e-10, dcxh=6.7e-10, ll=3.2e-9, dll=7.8e-10),
    'misspd': dict(pw=9.4e-10, sig2=2.3e-12, predvecs=5.3e-10, predvars=5.0e-10, dpv=3.2e-11, dpvar=6.6e-10,
                   mean=2.6e-10, var=2.4e-10, dmean=3.3e-11, cxh=7.3e-11, dcxh=7.7e-10, ll=3.1e-9, dll=6.7e-10),
    'trig': dict(pw=6.7e-8, sig2=9.9e-11, predvecs=3.8e-9, predvars=5.1e-8, dpv=1.3e-9, dpvar=2.7e-9, mean=3.0e-9,
                 var=8.5e-11, dmean=1.1e-9, cxh=1.8e-9, dcxh=6.8e-9, ll=3.4e-8, dll=3.1e-9),
    'd1': dict(pw=6.3e-6, sig2=2.5e-8, predvecs=3.4e-7, predvars=2.5e-6, dpv=4.2e-9, dpvar=1.7e-7, mean=1.7e-7,
               var=2.7e-7, dmean=3.8e-9, cxh=2.3e-7, dcxh=9.6e-7, ll=3.0e-6, dll=2.7e-7),
}
BARS_D8 = dict(pw=2e-10, sig2=2e-10, predvecs=2e-10, predvars=2e-10, dpv=2e-10, dpvar=2e-10, mean=2e-10, var=2e-10,
               dmean=2e-10, cxh=2e-10, dcxh=2e-10, ll=2e-10, dll=2e-10)      # asserted at 5 x = 1e-9, the north-star bar


def _bound(tag, key):
    return 5 * (BARS_D8 if tag == 'd8' else MEASURED_VS_REFERENCE[tag])[key]


@pytest.mark.parametrize('tag', ['c1', 'misspd', 'trig', 'd1', 'd8'])
def test_factorise_and_predict_golden(ops, golden, tag):
    g = golden('emu_' + tag)
    M, fitinfo = _load_state(g)
    cond = float(g['st_condR'].max())
    pw = np.array([e['pw'] for e in fitinfo['emulist']]).T
    sig2 = np.array([e['sig2'] for e in fitinfo['emulist']])
    e_pw, e_s2 = rel(pw, g['st_pw']), rel(sig2, g['st_sig2'])
    pred = {}
    M.predict(pred, fitinfo, g['in_x'], g['thetanew'], return_grad=True)
    errs = dict(pw=e_pw, sig2=e_s2, predvecs=rel(pred['predvecs'], g['predvecs']),
                predvars=float(np.max(np.abs(pred['predvars'] - g['predvars']) / g['predvars'])),
                dpv=rel(pred['predvecs_gradtheta'], g['predvecs_grad']),
                dpvar=rel(pred['predvars_gradtheta'], g['predvars_grad']),
                mean=rel(pred['mean'], g['mean']), var=rel(pred['var'], g['var']),
                dmean=rel(pred['mean_gradtheta'], g['mean_grad']),
                cxh=rel(pred['covxhalf'][:, :2, :], g['covxhalf_sample']),
                dcxh=rel(pred['covxhalf_gradtheta'][:, :2, :, :], g['covxhalf_grad_sample']))
    report('predict_golden_' + tag, cond=cond, **errs)
    for key, err in errs.items():
        assert err < _bound(tag, key), (tag, key, err, _bound(tag, key))
    # BASELINE bar: fitted predictive mean / variance within 1e-6 relative -- on every fixture
    assert errs['mean'] < 1e-6 and errs['var'] < 1e-6
    assert pred['mean'].shape == g['mean'].shape and pred['covxhalf'].shape[2] == g['st_hyp'].shape[0]
    # no-gradient call gives the same mean/variance
    pred2 = {}
    M.predict(pred2, fitinfo, g['in_x'], g['thetanew'])
    assert rel(pred2['mean'], pred['mean']) < 1e-12 and rel(pred2['var'], pred['var']) < 1e-12
    # at the design points the variance is a cancellation: looser, reported
    pd = {}
    M.predict(pd, fitinfo, g['in_x'], g['st_theta'][:8])
    e_dp = rel(pd['predvecs'], g['design_predvecs'])
    report('predict_design_' + tag, predvecs=e_dp,
           predvars_abs=float(np.max(np.abs(pd['predvars'] - g['design_predvars']))))
    assert e_dp < 5 * {'c1': 1.3e-9, 'misspd': 5.9e-10, 'trig': 2.9e-9, 'd1': 2.9e-7, 'd8': 2e-10}[tag]


def test_predict_row_subset_and_edge_shapes(ops, golden):
    g = golden('emu_c1')
    M, fitinfo = _load_state(g)
    x = g['in_x']
    rows = np.array([7, 3, 20])
    xs = np.vstack([x[rows], [[12345.0]]])                  # last row matches nothing -> NaN (PCGPwM.py:273-280)
    pred = {}
    M.predict(pred, fitinfo, xs, g['thetanew'][:5])
    assert pred['mean'].shape == (4, 5)
    assert np.all(np.isnan(pred['mean'][3])) and np.all(np.isfinite(pred['mean'][:3]))
    assert rel(pred['mean'][:3], g['mean'][rows][:, :5]) < 1e-6
    one = {}
    M.predict(one, fitinfo, x, g['thetanew'][0], return_grad=True)      # 1-D theta (B = 1)
    assert one['mean'].shape == (x.shape[0], 1) and one['covxhalf_gradtheta'].shape[1] == 1
    assert rel(one['mean'][:, 0], g['mean'][:, 0]) < 1e-6


# ---------------------------------------------------------------------------------- Woodbury log-lik (K10, K11)
class _Emu:
    """Just enough of surmise's emulator object for the calibration plugin (._info, .predict)."""

    def __init__(self, method, fitinfo):
        self.method, self._info = method, fitinfo


@pytest.mark.parametrize('tag', ['c1', 'misspd', 'trig', 'd1', 'd8'])
def test_woodbury_loglik_golden(ops, golden, tag):
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    g = golden('emu_' + tag)
    M, fitinfo = _load_state(g)
    emu = _Emu(M, fitinfo)
    cal = {'yvar': g['yvar']}
    ll = C.loglik(cal, emu, g['thetanew'], g['y'], g['in_x'])
    ll2, dll = C.loglik_grad(cal, emu, g['thetanew'], g['y'], g['in_x'])
    _, _, fired = C.loglik_device(cal, emu, g['thetanew'], g['y'], g['in_x'])
    cond = float(g['st_condR'].max())
    e_l = float(np.max(np.abs(ll - g['loglik']) / np.abs(g['loglik'])))
    e_d = rel(dll, g['dloglik'])
    report('loglik_golden_' + tag, ll=e_l, dll=e_d, fired=int(fired.cpu()[0]), cond=cond)
    assert bool(int(fired.cpu()[0])) == bool(g['fired'])
    assert ll.shape == g['loglik'].shape and dll.shape == g['dloglik'].shape
    assert np.allclose(ll, ll2, rtol=1e-12, atol=0)
    assert e_l < _bound(tag, 'll') and e_d < _bound(tag, 'dll'), (tag, e_l, e_d)


@pytest.mark.parametrize('tag', ['c1', 'misspd'])
def test_calibration_adopts_reference_fitted_emulator(ops, golden, tag):
    """A fitinfo laid out as the reference's PCGPwM.fit leaves it (no device state): directbayeswoodbury_B200
    factors it on first use and reproduces the reference's log-likelihoods."""
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    g = golden('emu_' + tag)
    k = g['st_hyp'].shape[0]
    fitinfo = {'method': 'PCGPwM', 'theta': g['st_theta'], 'x': g['in_x'], 'pc': g['st_pc'], 'pcti': g['st_pcti'],
               'unscaled_pcstdvar': g['st_unscaled_pcstdvar'], 'eta': 10, 'dampalpha': 0.3,
               'standardpcinfo': {'offset': g['st_offset'], 'scale': g['st_scale'], 'extravar': g['st_extravar']},
               'emulist': [{'hyp': g['st_hyp'][j], 'hypind': int(g['st_hypind'][j])} for j in range(k)]}
    emu = _Emu(M, fitinfo)
    ll, dll = C.loglik_grad({'yvar': g['yvar']}, emu, g['thetanew'], g['y'], g['in_x'])
    assert '_b200' in fitinfo
    e_l = float(np.max(np.abs(ll - g['loglik']) / np.abs(g['loglik'])))
    report('loglik_adopted_' + tag, ll=e_l, dll=rel(dll, g['dloglik']))
    assert e_l < _bound(tag, 'll') and rel(dll, g['dloglik']) < _bound(tag, 'dll')
    pred = {}
    M.predict(pred, fitinfo, g['in_x'], g['thetanew'])
    assert rel(pred['mean'], g['mean']) < 1e-6 and rel(pred['var'], g['var']) < 1e-6
    with pytest.raises(ValueError):
        M.adopt({'theta': g['st_theta']})


def test_woodbury_kernel_isolated(ops, golden):
    """Feed the reference's own PC-space predictions to the kernel: 1e-9 per theta, no GP error."""
    for tag in ['c1', 'trig', 'd1']:
        g = golden('emu_' + tag)
        phi = g['st_pcti'] * g['st_scale'][:, None]
        yv = g['yvar']
        ev = g['st_extravar']
        sinv = np.stack([1 / yv, 1 / (yv + np.abs(ev))])
        G = np.stack([phi.T @ (phi * s[:, None]) for s in sinv])
        consts = {'phiT': dev(phi.T.copy()), 'resid0': dev(g['y'] - g['st_offset']), 'extravar': dev(ev),
                  'sinv': dev(sinv), 'G': dev(G)}
        ll, dll, fired = ops.woodbury_loglik(consts, dev(g['predvecs']), dev(g['predvars']),
                                             dev(g['predvecs_grad']), dev(g['predvars_grad']))
        e_l = float(np.max(np.abs(ll.cpu().numpy()[:, None] - g['loglik']) / np.abs(g['loglik'])))
        e_d = rel(dll.cpu().numpy(), g['dloglik'])
        report('woodbury_isolated_' + tag, ll=e_l, dll=e_d)
        assert bool(int(fired.cpu()[0])) == bool(g['fired'])
        assert e_l < 1e-9 and e_d < 1e-8


@pytest.mark.parametrize('world', [2, 3])
def test_woodbury_segmented_layout_equals_dense(ops, golden, world):
    """The layout an all-gather of PC-sharded predictions produces (one (B, kmax[, d]) block per rank, padded
    components with predvar 1 and zero loadings) gives the dense result bit for bit in ll and to rounding in dll."""
    import torch
    g = golden('emu_c1')
    phi = g['st_pcti'] * g['st_scale'][:, None]
    yv, ev = g['yvar'], g['st_extravar']
    sinv = np.stack([1 / yv, 1 / (yv + np.abs(ev))])
    G = np.stack([phi.T @ (phi * s[:, None]) for s in sinv])
    consts = {'phiT': dev(phi.T.copy()), 'resid0': dev(g['y'] - g['st_offset']), 'extravar': dev(ev),
              'sinv': dev(sinv), 'G': dev(G)}
    pv, pvar, dpv, dpvar = (np.asarray(g[k]) for k in ('predvecs', 'predvars', 'predvecs_grad', 'predvars_grad'))
    B, K = pv.shape
    d = dpv.shape[2]
    ll0, dll0, _ = ops.woodbury_loglik(consts, dev(pv), dev(pvar), dev(dpv), dev(dpvar), variant=0)
    counts = [len(c) for c in np.array_split(np.arange(K), world)]
    kmax = max(counts)
    chunk = B * kmax * (2 + 2 * d)
    gathered = np.zeros((world, chunk))
    pos, o, j = [], B * kmax, 0
    for r, cnt in enumerate(counts):
        blk = gathered[r]
        a, b_, c, e = (blk[:o].reshape(B, kmax), blk[o:2 * o].reshape(B, kmax), blk[2 * o:2 * o + o * d].reshape(B, kmax, d),
                       blk[2 * o + o * d:].reshape(B, kmax, d))
        b_[...] = 1.0
        a[:, :cnt], b_[:, :cnt], c[:, :cnt], e[:, :cnt] = pv[:, j:j + cnt], pvar[:, j:j + cnt], dpv[:, j:j + cnt], dpvar[:, j:j + cnt]
        pos += [r * kmax + i for i in range(cnt)]
        j += cnt
    Kp = world * kmax
    phiT = np.zeros((Kp, phi.shape[0]))
    phiT[pos] = phi.T
    Gp = np.zeros((2, Kp, Kp))
    Gp[:, np.array(pos)[:, None], np.array(pos)[None, :]] = G
    gt = dev(gathered.reshape(-1))
    blk0 = gt[:chunk]
    views = (blk0[:o].view(B, kmax), blk0[o:2 * o].view(B, kmax), blk0[2 * o:2 * o + o * d].view(B, kmax, d),
             blk0[2 * o + o * d:].view(B, kmax, d))
    ll1, dll1, _ = ops.woodbury_loglik(dict(consts, phiT=dev(phiT), G=dev(Gp)), *views, variant=0, segments=(kmax, chunk))
    torch.cuda.synchronize()
    e_l = float(np.max(np.abs((ll1 - ll0).cpu().numpy()) / np.abs(ll0.cpu().numpy())))
    e_d = rel(dll1.cpu().numpy(), dll0.cpu().numpy())
    report('woodbury_segmented_w%d' % world, ll=e_l, dll=e_d)
    assert e_l < 1e-12 and e_d < 1e-10
    # the same layout read through a position table: only the real components are visited, the constants stay unpadded
    ll2, dll2, _ = ops.woodbury_loglik(consts, *views, variant=0, segments=(kmax, chunk),
                                       seg_pos=dev(np.array(pos, dtype=np.int32)))
    torch.cuda.synchronize()
    assert float(np.max(np.abs((ll2 - ll0).cpu().numpy()) / np.abs(ll0.cpu().numpy()))) < 1e-13
    assert rel(dll2.cpu().numpy(), dll0.cpu().numpy()) < 1e-12


def test_woodbury_global_scratch_path_equals_shared_memory_path(ops, golden):
    """k x k systems beyond shared memory live in global scratch (include/surmise_b200.h: sb200_woodbury_loglik_ws);
    the arithmetic is the same code, so on shapes both paths take the results agree bit for bit -- also when the
    workspace only holds a few proposals at a time."""
    import torch
    from surmise_b200._lib import lib, check
    g = golden('emu_c1')
    phi = g['st_pcti'] * g['st_scale'][:, None]
    yv, ev = g['yvar'], g['st_extravar']
    sinv = np.stack([1 / yv, 1 / (yv + np.abs(ev))])
    G = np.stack([phi.T @ (phi * s[:, None]) for s in sinv])
    consts = {'phiT': dev(phi.T.copy()), 'resid0': dev(g['y'] - g['st_offset']), 'extravar': dev(ev),
              'sinv': dev(sinv), 'G': dev(G)}
    args = [dev(np.asarray(g[k])) for k in ('predvecs', 'predvars', 'predvecs_grad', 'predvars_grad')]
    B, K = args[0].shape
    m, d = phi.shape[0], args[2].shape[2]
    ll0, dll0, f0 = ops.woodbury_loglik(consts, *args)
    assert lib.sb200_woodbury_workspace(B, K, m) == 0
    ops.woodbury_force_global(True)
    try:
        need = lib.sb200_woodbury_workspace(B, K, m)
        assert need == B * 2 * K * K * 8
        ll1, dll1, f1 = ops.woodbury_loglik(consts, *args)
        # a workspace for three proposals: ceil(B / 3) launches
        ws = torch.empty(3 * 2 * K * K * 8 + 8, dtype=torch.uint8, device='cuda')
        ll2, dll2 = torch.empty_like(ll0), torch.empty_like(dll0)
        f2 = torch.empty(1, dtype=torch.int32, device='cuda')
        p = lambda t: t.data_ptr()
        check(lib.sb200_woodbury_loglik_ws(B, K, m, d, p(args[0]), p(args[1]), p(args[2]), p(args[3]), p(consts['phiT']),
                                           p(consts['resid0']), p(consts['extravar']), p(consts['sinv']), p(consts['G']),
                                           1, -1, p(ll2), p(dll2), p(f2), None, 0, 0, None, p(ws), ws.numel(),
                                           torch.cuda.current_stream().cuda_stream), 'sb200_woodbury_loglik_ws')
        # without a workspace the shape is an argument error, not a silent wrong answer
        rc = lib.sb200_woodbury_loglik(B, K, m, d, p(args[0]), p(args[1]), p(args[2]), p(args[3]), p(consts['phiT']),
                                       p(consts['resid0']), p(consts['extravar']), p(consts['sinv']), p(consts['G']),
                                       1, -1, p(ll2), p(dll2), p(f2), None, 0, 0, None,
                                       torch.cuda.current_stream().cuda_stream)
        assert rc < 0
    finally:
        ops.woodbury_force_global(False)
    torch.cuda.synchronize()
    for a, b in ((ll0, ll1), (dll0, dll1), (ll0, ll2), (dll0, dll2)):
        assert torch.equal(a, b)
    assert int(f0.cpu()[0]) == int(f1.cpu()[0]) == int(f2.cpu()[0])


def test_woodbury_more_components_than_shared_memory_holds(ops):
    """k = 150 principal components at m = 400 (2 k^2 doubles = 360 KB > 227 KB): global-scratch path against the
    oracle's k-space restatement of directbayeswoodbury.py:261-383, and against the m-space Gaussian density."""
    rng = np.random.default_rng(11)
    B, K, m, d = 9, 150, 400, 3
    phi = rng.standard_normal((m, K)) / np.sqrt(K)
    offset, ev = rng.standard_normal(m), 1e-3 * rng.random(m)
    yvar, y = 0.05 + 0.1 * rng.random(m), rng.standard_normal(m)
    pv, pvar = rng.standard_normal((B, K)), 0.05 + rng.random((B, K))
    dpv, dpvar = rng.standard_normal((B, K, d)), 0.1 * rng.standard_normal((B, K, d))
    ll_o, dll_o, fired_o = wo.loglik_pcspace(pv, pvar, phi, offset, ev, yvar, y, dpv, dpvar)
    cs = wo.woodbury_constants(phi, offset, ev, yvar, y)
    consts = {'phiT': dev(phi.T.copy()), 'resid0': dev(y - offset), 'extravar': dev(ev),
              'sinv': dev(np.stack([c['sinv'] for c in cs])), 'G': dev(np.stack([c['G'] for c in cs]))}
    from surmise_b200._lib import lib
    assert lib.sb200_woodbury_workspace(B, K, m) == B * 2 * K * K * 8
    ll, dll, fired = ops.woodbury_loglik(consts, dev(pv), dev(pvar), dev(dpv), dev(dpvar))
    ll, dll = ll.cpu().numpy(), dll.cpu().numpy()
    assert bool(int(fired.cpu()[0])) == fired_o
    e_l, e_d = float(np.max(np.abs(ll - ll_o[:, 0]) / np.abs(ll_o[:, 0]))), rel(dll, dll_o)
    report('woodbury_k150', ll=e_l, dll=e_d)
    assert e_l < 1e-11 and e_d < 1e-9
    # the density itself: r ~ N(0, Sigma + Phi diag(v) Phi'), up to the theta-independent -1/2 log|Sigma| - m/2 log 2 pi
    ov = yvar + (np.abs(ev) if fired_o else 0.0)
    for b in range(B):
        r = y - offset - phi @ pv[b]
        S = np.diag(ov) + (phi * pvar[b]) @ phi.T
        ref = -0.5 * (np.linalg.slogdet(S)[1] - np.sum(np.log(ov))) - 0.5 * r @ np.linalg.solve(S, r)
        assert abs(ll[b] - ref) < 1e-9 * abs(ref)


class _AnyPrediction:
    def __init__(self, mean, var, cxh):
        self._m, self._v, self._c = mean, var, cxh

    def mean(self):
        return self._m

    def var(self):
        return self._v

    def covxhalf(self):
        return self._c


class _AnyEmulator:
    """An emulator outside the PCGPwM family, as the calibrator sees it: predict(x, theta) with the accessor API of
    surmise.emulation.prediction; predict(x) alone = at its own training thetas (emulation.py)."""

    def __init__(self, m, K, d, seed, var_scale):
        rng = np.random.default_rng(seed)
        self.Wm, self.Wc = rng.standard_normal((m, d)), 0.3 * rng.standard_normal((m, K, d))
        self.theta0 = rng.uniform(size=(11, d))
        self.var_scale = var_scale
        self._info = {'method': 'some_other_method'}

    def predict(self, x=None, theta=None):
        th = self.theta0 if theta is None else np.atleast_2d(theta)
        mean = self.Wm @ th.T                                                     # (m, B)
        cxh = np.einsum('mkd,bd->mbk', self.Wc, np.cos(th)) + 0.1                 # (m, B, K)
        var = self.var_scale * np.sum(cxh ** 2, 2) + 0.01 * (self.var_scale > 1)
        return _AnyPrediction(mean, var, cxh)


@pytest.mark.parametrize('m,K,B,var_scale', [(29, 7, 33, 1.0), (29, 7, 33, 3.0), (5, 1, 1, 1.0), (200, 40, 9, 1.0)])
def test_loglik_for_an_emulator_outside_the_pcgpwm_family(ops, m, K, B, var_scale):
    """directbayeswoodbury.loglik as written (dbw.py:261-306) on the device for any emulator's mean / var / covxhalf:
    against the oracle's restatement (with and without the batch-wide variance test firing) and against the Gaussian
    density itself; the gradient form is refused, not approximated."""
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    d = 3
    emu = _AnyEmulator(m, K, d, seed=m + K, var_scale=var_scale)
    rng = np.random.default_rng(B)
    theta = rng.uniform(size=(B, d))
    y = rng.standard_normal(m)
    yvar = 0.05 + 0.1 * rng.random(m)
    p, old = emu.predict(None, theta), emu.predict(None)
    ll_o, fired_o = wo.loglik_from_prediction(p.mean(), p.var(), p.covxhalf(), yvar, y, old.var(), old.covxhalf())
    assert fired_o == (var_scale > 1)
    ll = C.loglik({'yvar': yvar}, emu, theta, y, None)
    assert ll.shape == (B, 1)
    e_l = float(np.max(np.abs(ll - ll_o) / np.abs(ll_o)))
    report('loglik_any_emulator', m=m, K=K, B=B, fired=float(fired_o), ll=e_l)
    assert e_l < 1e-10
    obsvar = yvar + (np.mean(np.abs(old.var() - np.sum(old.covxhalf() ** 2, 2)), 1) if fired_o else 0.0)
    for b in range(min(B, 4)):
        r = y - p.mean()[:, b]
        Cb = p.covxhalf()[:, b, :]
        S = np.diag(obsvar) + Cb @ Cb.T
        ref = -0.5 * (np.linalg.slogdet(S)[1] - np.sum(np.log(obsvar))) - 0.5 * r @ np.linalg.solve(S, r)
        assert abs(ll[b, 0] - ref) < 1e-9 * abs(ref)
    with pytest.raises(ValueError):
        C.loglik_grad({'yvar': yvar}, emu, theta, y, None)
    with pytest.raises(ValueError):
        C.loglik({}, emu, theta, y, None)


class _ReplayEmulator:
    """Hands the calibrator the emulator outputs stored in a golden fixture (at the proposals; at the training thetas
    when predict is called without theta, as surmise.emulation.emulator.predict does)."""
    _info = {'method': 'replayed outputs of another method'}

    def __init__(self, g, tag):
        self.g, self.tag = g, tag

    def predict(self, x=None, theta=None):
        k = self.tag + ('_old_' if theta is None else '_')
        return _AnyPrediction(None if theta is None else self.g[self.tag + '_mean'], self.g[k + 'var'], self.g[k + 'covxhalf'])


@pytest.mark.parametrize('tag', ['pcgp', 'fired'])
def test_loglik_for_another_emulator_against_the_reference(ops, golden, tag):
    """The generic path against the reference's own directbayeswoodbury.loglik outputs (dbw.py:261-306) for a
    reference-fitted PCGP emulator and for outputs that make the variance test fire: 1e-9 per theta."""
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    g = golden('any_emulator')
    ll = C.loglik({'yvar': g[tag + '_yvar']}, _ReplayEmulator(g, tag), g[tag + '_theta'], g[tag + '_y'], None)
    ref = g[tag + '_loglik']
    assert ll.shape == ref.shape
    e_l = float(np.max(np.abs(ll - ref) / np.abs(ref)))
    report('loglik_any_emulator_golden_' + tag, ll=e_l)
    assert e_l < 1e-9


def test_logpost_closure_contract(ops, golden):
    """Shapes / -inf handling of the closure handed to samplers (dbw.py:100-129)."""
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    g = golden('emu_c1')
    M, fitinfo = _load_state(g)
    emu = _Emu(M, fitinfo)

    class prior:
        @staticmethod
        def lpdf(theta):
            inside = np.all((theta >= 0) & (theta <= 1), axis=1)
            return np.where(inside, 0.0, -np.inf).reshape(-1, 1)

        @staticmethod
        def rnd(n):
            return np.random.default_rng(0).uniform(size=(n, 3))
    cal = {'yvar': g['yvar'], 'thetaprior': prior}
    logpost, draw = C.make_logpost(cal, emu, g['in_x'], g['y'])
    th = g['thetanew'].copy()
    th[1, 0] = 1.5                                          # outside the prior support
    lp, dlp = logpost(th)
    assert lp.shape == (th.shape[0], 1) and dlp.shape == th.shape
    assert np.isneginf(lp[1, 0]) and np.all(np.isfinite(np.delete(lp[:, 0], 1)))
    lp2 = logpost(th, return_grad=False)
    assert np.allclose(np.delete(lp2[:, 0], 1), np.delete(lp[:, 0], 1), rtol=1e-12)
    assert np.allclose(np.delete(lp[:, 0], 1), np.delete(g['loglik'][:, 0], 1), rtol=1e-7)
    assert draw(10).shape == (10, 3)


def test_edge_shapes_single_pc_max_dim_tiny_n(ops):
    """k = 1 (dbw.py:293-301, 371-374 special-case it), d = 16 (the compiled maximum), very small n, B = 1."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    rng = np.random.default_rng(2)
    n, d, m = 40, 16, 7
    theta = rng.uniform(size=(n, d))
    load = rng.normal(size=m)
    f = (np.sin(theta @ rng.normal(size=d))[:, None] * load[None, :] + 1e-6 * rng.normal(size=(n, m))).T   # rank 1
    x = np.arange(m, dtype=float)[:, None]
    np.random.seed(2)
    fi = {}
    M.fit(fi, x, theta, f)
    assert len(fi['emulist']) == 1
    # oracle at the same hyper-parameters
    e = fi['emulist'][0]
    o = po.final_factorisation(theta, fi['pc'][:, 0], np.zeros(n), e['hyp'])
    o['hypind'] = 0
    fo = {'theta': theta, 'emulist': [o], 'pcti': fi['pcti'], 'standardpcinfo': fi['standardpcinfo']}
    for B in (1, 5):
        tn = rng.uniform(size=(B, d))
        pg, pr = {}, {}
        M.predict(pg, fi, x, tn, return_grad=True)
        po.predict(pr, fo, x, tn, return_grad=True)
        assert pg['covxhalf'].shape == (m, B, 1) and pg['covxhalf_gradtheta'].shape == (m, B, 1, d)
        assert rel(pg['mean'], pr['mean']) < 1e-7 and rel(pg['var'], pr['var']) < 1e-6
        assert rel(pg['mean_gradtheta'], pr['mean_gradtheta']) < 1e-6
        y = pr['mean'][:, 0] + 0.01
        yvar = np.full(m, 1e-3)

        class Emu:
            _info = fi
        ll, dll = C.loglik_grad({'yvar': yvar}, Emu, tn, y, x)
        oll, odll = wo.loglik_mspace(fo, yvar, tn, y, return_grad=True)
        assert np.max(np.abs(ll - oll) / np.abs(oll)) < 1e-6 and rel(dll, odll) < 1e-5
    # n = 2 design points, B larger than n
    th2 = rng.uniform(size=(2, 3))
    mean, lo, hi, std = M.hyper_prior(rng.uniform(size=(30, 3)), -10, -20, None)
    g2 = np.array([0.3, -0.2])
    info = {'theta': th2, 'g': g2, 'gvar': np.zeros(2), 'sig2ofconst': 0.01, 'hypregmean': mean, 'hypregstd': std}
    v, gr, code = ops.gp_nll_grad(ops.GPData(th2, g2, np.zeros(2)), mean[None, :], mean, std, 0.01, True)
    assert code[0] == 0 and abs(v[0] - po.negloglik_eigh(mean, info)) < 1e-12
    assert rel(gr[0], po.negloglikgrad_eigh(mean, info)) < 1e-10
