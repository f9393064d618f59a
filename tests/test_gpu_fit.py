"""GPU: whole PCGPwM_B200 fits through the plugin entry points (the calls surmise's facade makes)."""
import numpy as np
import pytest

from gpu_util import rel, report
from oracle import pcgpwm_oracle as po
from synth import make_simulator

pytestmark = pytest.mark.gpu


def _facade_fit(method, x, theta, f, **args):
    """What surmise.emulation.emulator.fit does: fitinfo dict in, method.fit(fitinfo, x, theta, f, **args)."""
    fitinfo = {'method': 'PCGPwM_B200'}
    method.fit(fitinfo, x, theta, f, **args)
    return fitinfo


def test_fit_c1_matches_oracle_fit_and_truth():
    """Config C1 (n=200, d=3, m=50): same seed, same RNG consumption as the reference (PCGPwM.py:744-748)."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    x, theta, f, truth = make_simulator(200, 3, 50, 8, seed=3)
    np.random.seed(3)
    fi = _facade_fit(M, x, theta, f)
    np.random.seed(3)
    fo = {}
    po.fit(fo, x, theta, f)
    k = len(fi['emulist'])
    assert k == len(fo['emulist']) == 8
    hyp_g = np.array([e['hyp'] for e in fi['emulist']])
    hyp_o = np.array([e['hyp'] for e in fo['emulist']])
    hi_g = [e['hypind'] for e in fi['emulist']]
    hi_o = [int(e['hypind']) for e in fo['emulist']]
    rng = np.random.default_rng(0)
    tn = rng.uniform(0.05, 0.95, size=(40, 3))
    pg, pr = {}, {}
    M.predict(pg, fi, x, tn)
    po.predict(pr, fo, x, tn)
    e_mean = rel(pg['mean'], pr['mean'])
    e_truth = float(np.sqrt(np.mean((pg['mean'] - truth(tn)) ** 2)) / np.std(truth(tn)))
    report('fit_c1', hyp=float(np.max(np.abs(hyp_g - hyp_o))), same_hypind=float(hi_g == hi_o), mean_vs_oracle=e_mean,
           rmse_vs_truth=e_truth, nevals=fi['_b200']['nevals'])
    assert e_truth < 0.05                                   # the emulator actually emulates
    assert e_mean < 1e-3                                    # and lands where the reference's optimiser lands
    for key in ('theta', 'x', 'param_desc', 'pct', 'pcti', 'pcstdvar', 'standardpcinfo', 'emulist', 'pc'):
        assert key in fi
    assert isinstance(fi['param_desc'], str) and 'number of GP components' in fi['param_desc']


def test_fit_with_missing_outputs_and_refit():
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    x, theta, f, truth = make_simulator(150, 2, 30, 4, seed=7, missing=0.10)
    np.random.seed(7)
    fi = _facade_fit(M, x, theta, f)
    assert fi['mof'] is not None and fi['mof'].sum() > 0
    tn = np.random.default_rng(1).uniform(0.1, 0.9, size=(25, 2))
    p = {}
    M.predict(p, fi, x, tn, return_grad=True)
    assert np.all(np.isfinite(p['mean'])) and np.all(p['var'] > 0)
    e_truth = float(np.sqrt(np.mean((p['mean'] - truth(tn)) ** 2)) / np.std(truth(tn)))
    report('fit_missing', rmse_vs_truth=e_truth, k=len(fi['emulist']))
    assert e_truth < 0.1
    # finite-difference check of the predictive-mean gradient
    h = 1e-6
    tp = tn.copy()
    tp[:, 0] += h
    p2 = {}
    M.predict(p2, fi, x, tp)
    fd = (p2['mean'] - p['mean']) / h
    assert rel(fd, p['mean_gradtheta'][:, :, 0]) < 1e-3
    # refit on the same fitinfo warm-starts from the old emulist (PCGPwM.py:680-685)
    M.fit(fi, x, theta, f)
    assert len(fi['emulist']) >= 4


def test_fit_full_data_likelihood_option():
    """nhyptrain='all' trains on all n points: exercises the large-matrix regime inside the optimiser."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    x, theta, f, truth = make_simulator(300, 3, 20, 3, seed=11)
    np.random.seed(11)
    fi = _facade_fit(M, x, theta, f, nhyptrain='all')
    tn = np.random.default_rng(2).uniform(0.1, 0.9, size=(20, 3))
    p = {}
    M.predict(p, fi, x, tn)
    assert float(np.sqrt(np.mean((p['mean'] - truth(tn)) ** 2)) / np.std(truth(tn))) < 0.05


def test_user_supplied_basis_and_errors():
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    x, theta, f, _ = make_simulator(80, 2, 12, 3, seed=5)
    U = np.linalg.svd(f - f.mean(1, keepdims=True), full_matrices=False)[0][:, :3]
    fi = _facade_fit(M, x, theta, f, standardpcinfo={'U': U})
    assert len(fi['emulist']) == 3
    with pytest.raises(AttributeError):
        _facade_fit(M, x, theta, f, standardpcinfo={})
    fconst = f.copy()
    fconst[0, :] = 1.0
    with pytest.raises(ValueError):
        _facade_fit(M, x, theta, fconst)
