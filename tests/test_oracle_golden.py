"""CPU: the numpy oracle reproduces the committed outputs of the real reference (tests/golden/)."""
import numpy as np
import pytest

from oracle import pcgpwm_oracle as po
from oracle import woodbury_oracle as wo


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_covmat_matches_reference(golden, tag):
    g = golden('covmat')
    x1, x2, gam = g[f'{tag}_x1'], g[f'{tag}_x2'], g[f'{tag}_gam']
    assert rel(po.matern_covmat(x1, x2, gam), g[f'{tag}_R']) < 1e-14
    assert rel(po.matern_covmat(x1, x2, gam, 'hyp')[1], g[f'{tag}_dRhyp']) < 1e-14
    assert rel(po.matern_covmat(x1, x2, gam, 'x1')[1], g[f'{tag}_dRx1']) < 1e-14


def test_covmat_1d_input(golden):
    g = golden('covmat')
    assert rel(po.matern_covmat(g['e_x1'], g['e_x2'], g['e_gam']), g['e_R']) < 1e-14


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_negloglik_and_grad(golden, tag):
    g = golden('nll')
    info = {'theta': g[f'{tag}_theta'], 'g': g[f'{tag}_g'], 'gvar': g[f'{tag}_gvar'], 'sig2ofconst': 0.01,
            'hypregmean': g[f'{tag}_mean'], 'hypregstd': g[f'{tag}_std']}
    for h, nll, grad, cond in zip(g[f'{tag}_hyps'], g[f'{tag}_nll'], g[f'{tag}_grad'], g[f'{tag}_cond']):
        assert abs(po.negloglik_eigh(h, info) - nll) <= 1e-12 * abs(nll)
        assert rel(po.negloglikgrad_eigh(h, info), grad) < 1e-10
        # Cholesky restatement (the GPU formulation): tolerance scales with cond(R)
        tol = max(1e-11, 1e-15 * cond)
        v, dv, code = po.negloglik_and_grad_chol(h, info)
        assert code == 0
        # NLL is a difference of O(n) terms (log-det vs n log sig2): scale the bar by n
        assert abs(v - nll) <= tol * (abs(nll) + info['g'].shape[0])
        assert rel(dv, grad) < 10 * tol


def test_standardize_and_pcs_with_missing(golden):
    g = golden('standardize')
    f = g['f']
    fi = {'f': f.T.copy(), 'epsilonPC': 0.001, 'epsilonImpute': 10e-6}
    fi['mof'] = np.logical_not(np.isfinite(fi['f']))
    fi['mofrows'] = np.where(np.any(fi['mof'] > 0.5, 1))[0]
    po.standardize_f(fi)
    po.principal_components(fi)
    spc = fi['standardpcinfo']
    assert rel(spc['offset'], g['offset']) < 1e-14 and rel(spc['scale'], g['scale']) < 1e-14
    assert rel(spc['fs'], g['fs']) < 1e-9
    assert rel(spc['extravar'], g['extravar']) < 1e-9
    assert rel(fi['pc'], g['pc']) < 1e-9 and rel(fi['pcti'], g['pcti']) < 1e-9
    assert np.max(np.abs(fi['unscaled_pcstdvar'] - g['unscaled_pcstdvar'])) < 1e-9


def _ref_state_as_fitinfo(g):
    """Rebuild an oracle fitinfo from the reference's fitted hyper-parameters."""
    theta = g['st_theta']
    k = g['st_hyp'].shape[0]
    emulist = []
    for j in range(k):
        gv = po.damp_gvar(g['st_unscaled_pcstdvar'][:, j], 10, 0.3)
        e = po.final_factorisation(theta, g['st_pc'][:, j], gv, g['st_hyp'][j])
        e['hypind'] = int(g['st_hypind'][j])
        emulist.append(e)
    return {'theta': theta, 'emulist': emulist, 'pcti': g['st_pcti'],
            'standardpcinfo': {'offset': g['st_offset'], 'scale': g['st_scale'], 'extravar': g['st_extravar']}}


@pytest.mark.parametrize('tag', ['c1', 'misspd', 'trig', 'd1'])
def test_predict_and_loglik_from_reference_state(golden, tag):
    g = golden('emu_' + tag)
    fi = _ref_state_as_fitinfo(g)
    # re-factorising at the reference's hyper-parameters: agreement is bounded by cond(R) * eps
    tol = max(1e-8, 100 * 2.2e-16 * float(g['st_condR'].max()))
    assert rel(np.array([e['pw'] for e in fi['emulist']]).T, g['st_pw']) < 10 * tol
    assert rel(np.array([e['sig2'] for e in fi['emulist']]), g['st_sig2']) < tol
    pred = {}
    po.predict(pred, fi, None, g['thetanew'], return_grad=True)
    assert rel(pred['mean'], g['mean']) < tol
    assert rel(pred['var'], g['var']) < 10 * tol
    assert rel(pred['mean_gradtheta'], g['mean_grad']) < 10 * tol
    assert rel(pred['covxhalf'][:, :2, :], g['covxhalf_sample']) < 10 * tol
    ll, dll = wo.loglik_mspace(fi, g['yvar'], g['thetanew'], g['y'], return_grad=True)
    assert np.max(np.abs(ll - g['loglik_g']) / np.abs(g['loglik_g'])) < 10 * tol
    assert rel(dll, g['dloglik']) < 100 * tol
    # k-space restatement fed with the reference's own PC-space predictions: tight
    spc = fi['standardpcinfo']
    phi = fi['pcti'] * spc['scale'][:, None]
    kll, kdll, fired = wo.loglik_pcspace(g['predvecs'], g['predvars'], phi, spc['offset'], spc['extravar'],
                                         g['yvar'], g['y'], g['predvecs_grad'], g['predvars_grad'])
    assert fired == bool(g['fired'])
    assert np.max(np.abs(kll - g['loglik']) / np.abs(g['loglik'])) < 1e-11
    assert rel(kdll, g['dloglik']) < 1e-9


def test_trigger_case_actually_fires(golden):
    assert bool(golden('emu_trig')['fired']) and not bool(golden('emu_c1')['fired'])


def test_seeded_fit_reproduces_reference_hyperparameters(golden):
    """Whole fit (RNG stream included) on the small d=1 case: PCGPwM.py:12-137."""
    g = golden('emu_d1')
    np.random.seed(13)
    fi = {}
    po.fit(fi, g['in_x'], g['in_theta'], g['in_f'])
    assert rel(np.array([e['hyp'] for e in fi['emulist']]), g['st_hyp']) < 1e-8
    assert [int(e['hypind']) for e in fi['emulist']] == [int(v) for v in g['st_hypind']]


@pytest.mark.parametrize('tag', ['c1', 'misspd', 'trig', 'd1', 'd8'])
def test_arbiter_fixture_is_consistent_with_the_goldens(golden, tag):
    """tests/golden/arbiter.npz (40-digit evaluation of the reference's formulas, make_golden_arbiter.py): the stored
    reference-vs-exact errors are what the goldens give, both Cholesky-form and reference stay within cond * eps of
    the exact values, and the batch-wide variance flag agrees."""
    g, a = golden('emu_' + tag), golden('arbiter')
    cond = float(g['st_condR'].max())
    e_ll = float(np.max(np.abs(g['loglik'] - a[tag + '_loglik']) / np.abs(a[tag + '_loglik'])))
    assert abs(e_ll - float(a[tag + '_ref_err_loglik'])) <= 1e-3 * e_ll + 1e-18
    assert bool(a[tag + '_fired']) == bool(g['fired'])
    assert e_ll < 20 * 2.2e-16 * cond
    states = [po.factor_state_chol(g['st_theta'], g['st_pc'][:, j], po.damp_gvar(g['st_unscaled_pcstdvar'][:, j], 10, 0.3),
                                   g['st_hyp'][j]) for j in range(g['st_hyp'].shape[0])]
    pv, pvar = po.predict_pcspace_chol(g['st_theta'], states, g['st_hypind'], g['thetanew'])[:2]
    assert np.max(np.abs(pv - a[tag + '_predvecs'])) / np.max(np.abs(a[tag + '_predvecs'])) < 20 * 2.2e-16 * cond
    assert np.max(np.abs(pvar - a[tag + '_predvars']) / a[tag + '_predvars']) < 50 * 2.2e-16 * cond


@pytest.mark.parametrize('tag', ['pcgp', 'fired'])
def test_loglik_from_any_emulators_prediction(golden, tag):
    """The oracle's restatement of directbayeswoodbury.loglik for an arbitrary emulator (dbw.py:261-306) against the
    reference's own outputs: a reference-fitted PCGP emulator, and synthetic outputs for which the variance test fires
    (tests/golden/make_golden_any_emulator.py)."""
    g = golden('any_emulator')
    f = lambda k: g[tag + '_' + k]
    ll, fired = wo.loglik_from_prediction(f('mean'), f('var'), f('covxhalf'), f('yvar'), f('y'), f('old_var'), f('old_covxhalf'))
    assert fired == bool(f('fired')) == (tag == 'fired')
    assert ll.shape == f('loglik').shape
    assert np.max(np.abs(ll - f('loglik')) / np.abs(f('loglik'))) < 1e-13

