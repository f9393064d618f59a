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


def test_fit_parallel_mode_single_rank():
    """fit_mode='parallel' (Jacobi sweeps, the shardable form) on one GPU: accurate and self-consistent."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    x, theta, f, truth = make_simulator(200, 3, 50, 8, seed=3)
    np.random.seed(3)
    fi = _facade_fit(M, x, theta, f, fit_mode='parallel')
    tn = np.random.default_rng(0).uniform(0.05, 0.95, size=(40, 3))
    p = {}
    M.predict(p, fi, x, tn)
    e_truth = float(np.sqrt(np.mean((p['mean'] - truth(tn)) ** 2)) / np.std(truth(tn)))
    inds = [e['hypind'] for e in fi['emulist']]
    report('fit_parallel', rmse_vs_truth=e_truth, leaders=len(set(inds)))
    assert e_truth < 0.05
    assert all(inds[j] <= j for j in range(len(inds)))
    for j, e in enumerate(fi['emulist']):                          # sharing PCs carry their leader's hyper-parameters
        assert np.array_equal(e['hyp'], fi['emulist'][e['hypind']]['hyp'])
    with pytest.raises(ValueError):
        _facade_fit(M, x, theta, f, fit_mode='nonsense')


class _ShimPrediction:
    def __init__(self, info):
        self._info = info

    def mean(self):
        return self._info['mean']

    def covxhalf(self):
        return self._info['covxhalf']

    def __call__(self):
        return self.mean()


class _ShimEmulator:
    """The three things surmise's emulator object offers to a calibration method."""

    def __init__(self, method, fitinfo, theta):
        self.method, self._info, self._emulator__theta = method, fitinfo, theta

    def predict(self, x=None, theta=None, args={}):
        info = {}
        self.method.predict(info, self._info, x, self._info['theta'] if theta is None else theta, **args)
        return _ShimPrediction(info)


def test_calibration_fit_through_sampler_contract(monkeypatch):
    """directbayeswoodbury_B200.fit drives a sampler exactly as surmise.utilities.sampler would (dbw.py:149-162).
    The GPU box has no surmise, so a stand-in `surmise.utilities.sampler` with the same constructor contract
    (utilities.py:36-71) runs a plain random-walk Metropolis on the closure it is given."""
    import sys
    import types
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    from synth import make_observation

    class sampler:
        def __init__(self, logpost_func, draw_func, sampler='metropolis_hastings', numsamp=400, **opts):
            rng = np.random.default_rng(0)
            th = draw_func(64)
            lp = logpost_func(th, return_grad=False)
            cur, curlp = th[np.argmax(lp)].copy(), float(np.max(lp))
            draws = []
            for it in range(numsamp + 200):
                prop = cur + 0.03 * rng.normal(size=cur.shape)
                plp = float(logpost_func(prop[None, :], return_grad=False)[0, 0])
                if np.log(rng.uniform()) < plp - curlp:
                    cur, curlp = prop, plp
                if it >= 200:
                    draws.append(cur.copy())
            self.sampler_info = {'theta': np.array(draws)}
    if 'surmise' not in sys.modules:
        pkg = types.ModuleType('surmise')
        util = types.ModuleType('surmise.utilities')
        util.sampler = sampler
        pkg.utilities = util
        monkeypatch.setitem(sys.modules, 'surmise', pkg)
        monkeypatch.setitem(sys.modules, 'surmise.utilities', util)
    else:
        import surmise.utilities
        monkeypatch.setattr(surmise.utilities, 'sampler', sampler)

    x, theta, f, truth = make_simulator(150, 2, 25, 4, seed=21)
    np.random.seed(21)
    fi = _facade_fit(M, x, theta, f)
    emu = _ShimEmulator(M, fi, theta)
    theta_true, y, yvar = make_observation(truth, 2, seed=22)

    class prior:
        @staticmethod
        def lpdf(th):
            th = np.atleast_2d(th)
            return np.where(np.all((th >= 0) & (th <= 1), axis=1), 0.0, -np.inf).reshape(-1, 1)

        @staticmethod
        def rnd(n):
            return np.random.uniform(size=(n, 2))
    cal = {'thetaprior': prior, 'yvar': yvar}
    C.fit(cal, emu, x, y, sampler='metropolis_hastings', numsamp=400)
    assert cal['thetarnd'].shape == (400, 2) and np.isfinite(cal['lpdfapproxnorm'])
    err = np.abs(cal['thetarnd'].mean(0) - theta_true[0])
    report('calibration_fit', err0=float(err[0]), err1=float(err[1]))
    assert np.all(err < 0.1)                                        # posterior sits on the generating theta
    lp = C.thetalpdf(cal, cal['thetarnd'][:5])
    assert lp.shape == (5, 1) and np.all(np.isfinite(lp))
    pred = {}
    C.predict(pred, cal, emu, x)
    assert pred['rnd'].shape == (400, x.shape[0]) and pred['mean'].shape == (x.shape[0],)
    assert np.sqrt(np.mean((pred['mean'] - truth(theta_true)[:, 0]) ** 2)) < 0.2
    assert C.thetarnd(cal, 7).shape == (7, 2)
    with pytest.raises(ValueError):
        C.loglik({}, emu, cal['thetarnd'][:2], y, x)                # yvar is mandatory (dbw.py:266-269)


def test_lockstep_batching_equals_sequential_jacobi():
    """Batching the per-PC optimisers' likelihood calls into one launch must not change any number."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    x, theta, f, truth = make_simulator(260, 3, 30, 6, seed=17)
    outs = []
    default = M._GPFitter.lockstep
    assert default == 'coroutine' and M._LbfgsbCore.available()
    for lock in ('coroutine', 'threads', False):     # generators + one scheduler / threads + rendezvous / sequential
        M._GPFitter.lockstep = lock
        try:
            np.random.seed(17)
            fi = _facade_fit(M, x, theta, f, fit_mode='parallel')
        finally:
            M._GPFitter.lockstep = default
        outs.append((np.array([e['hyp'] for e in fi['emulist']]), [e['hypind'] for e in fi['emulist']],
                     fi['_b200']['nevals']))
    for other in outs[1:]:
        assert np.array_equal(outs[0][0], other[0]) and outs[0][1] == other[1] and outs[0][2] == other[2]


def test_device_pca_matches_host_pca(golden):
    """Imputation / PCA on the device (Gram eigh + batched solves) against the host implementation and, through it,
    the reference internals (PCGPwM.py:533-654).  The 40 ridge iterations amplify rounding (the reference's own
    GEMM-vs-SYRK choice moves fs by 2e-5), hence the tolerances; complete data is tight."""
    from surmise_b200 import _pca, _pca_device
    g = golden('standardize')
    spc_d, mof, rows = _pca_device.standardize(g['f'].T, 0.001, 10e-6)
    pcs_d = _pca_device.principal_components(spc_d, mof, rows, 0.001, 10e-6)
    assert pcs_d['pc'].shape == g['pc'].shape
    sign = np.sign(np.sum(pcs_d['pc'] * g['pc'], 0))
    e_fs = float(np.max(np.abs(spc_d['fs'] - g['fs'])))
    e_pc = float(np.max(np.abs(pcs_d['pc'] * sign - g['pc'])) / np.max(np.abs(g['pc'])))
    e_ev = float(np.max(np.abs(spc_d['extravar'] - g['extravar'])) / np.max(g['extravar']))
    e_sv = float(np.max(np.abs(pcs_d['unscaled_pcstdvar'] - g['unscaled_pcstdvar'])))
    report('device_pca_missing', fs=e_fs, pc=e_pc, extravar=e_ev, pcstdvar=e_sv)
    # 5 x the discrepancies measured on a B200 (profiles/parity_report_r1.txt: 9.8e-6, 8.2e-7, 3.2e-5, 1.8e-4)
    assert e_fs < 5e-5 and e_pc < 5e-6 and e_ev < 2e-4 and e_sv < 1e-3
    c = golden('emu_c1')
    spc_c, mof, rows = _pca_device.standardize(c['in_f'].T, 0.001, 10e-6)
    assert mof is None
    pcs_c = _pca_device.principal_components(spc_c, None, None, 0.001, 10e-6)
    sign = np.sign(np.sum(pcs_c['pc'] * c['st_pc'], 0))
    e_pc = float(np.max(np.abs(pcs_c['pc'] * sign - c['st_pc'])) / np.max(np.abs(c['st_pc'])))
    e_ti = float(np.max(np.abs(pcs_c['pcti'] * sign - c['st_pcti'])) / np.max(np.abs(c['st_pcti'])))
    e_ev = float(np.max(np.abs(spc_c['extravar'] - c['st_extravar'])) / np.max(c['st_extravar']))
    report('device_pca_complete', pc=e_pc, pcti=e_ti, extravar=e_ev)
    assert e_pc < 1e-13 and e_ti < 3e-12 and e_ev < 3e-11        # measured 1.3e-15, 4.2e-13, 5.7e-12


def test_fit_host_pca_option_agrees_with_device_pca():
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    x, theta, f, truth = make_simulator(150, 2, 30, 4, seed=7, missing=0.10)
    tn = np.random.default_rng(1).uniform(0.1, 0.9, size=(20, 2))
    means = []
    for mode in ('device', 'host'):
        np.random.seed(7)
        fi = _facade_fit(M, x, theta, f, pca=mode)
        p = {}
        M.predict(p, fi, x, tn)
        means.append(p['mean'])
    report('fit_host_vs_device_pca', mean=rel(means[0], means[1]))
    assert rel(means[0], means[1]) < 1e-2
    with pytest.raises(ValueError):
        _facade_fit(M, x, theta, f, pca='nowhere')


def test_fitted_state_survives_pickle_and_predictlpdf_runs():
    """emu.save_to / load_from pickle the whole _info dict (helper.py:9-23): the device state must round-trip."""
    import pickle
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    x, theta, f, truth = make_simulator(120, 2, 20, 3, seed=31)
    np.random.seed(31)
    fi = _facade_fit(M, x, theta, f)
    tn = np.random.default_rng(3).uniform(0.1, 0.9, size=(9, 2))
    p0 = {}
    M.predict(p0, fi, x, tn, return_grad=True)
    fi2 = pickle.loads(pickle.dumps(fi))
    p1 = {}
    M.predict(p1, fi2, x, tn, return_grad=True)
    assert np.array_equal(p0['mean'], p1['mean']) and np.array_equal(p0['covxhalf_gradtheta'], p1['covxhalf_gradtheta'])
    lp, dlp = M.predictlpdf(p1, truth(tn), addvar=0.01)
    assert lp.shape == (9, 1) and dlp.shape == (9, 2) and np.all(np.isfinite(lp))
    h = 1e-6                                                    # finite-difference check of the lpdf gradient
    tp = tn.copy()
    tp[:, 1] += h
    p2 = {}
    M.predict(p2, fi, x, tp)
    p2['return_grad'] = False
    fd = (M.predictlpdf(p2, truth(tn), addvar=0.01) - lp)[:, 0] / h
    # the reference's analytic lpdf gradient (reproduced here) is only approximately the derivative of its lpdf
    assert rel(fd, dlp[:, 1]) < 5e-2


def test_calibration_with_ptlmc_b200_sampler():
    """End to end: PCGPwM_B200 emulator -> directbayeswoodbury_B200 -> PTLMC_B200 (config C5 in miniature)."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    from synth import make_observation
    x, theta, f, truth = make_simulator(150, 2, 25, 4, seed=21)
    np.random.seed(21)
    fi = _facade_fit(M, x, theta, f)
    emu = _ShimEmulator(M, fi, theta)
    theta_true, y, yvar = make_observation(truth, 2, seed=22)
    ncalls = {'n': 0, 'rows': 0}

    class prior:
        @staticmethod
        def lpdf(th):
            th = np.atleast_2d(th)
            ncalls['n'] += 1
            ncalls['rows'] += th.shape[0]
            return np.where(np.all((th >= 0) & (th <= 1), axis=1), 0.0, -np.inf).reshape(-1, 1)

        @staticmethod
        def lpdf_grad(th):
            return np.zeros_like(np.atleast_2d(th))

        @staticmethod
        def rnd(n):
            return np.random.uniform(size=(n, 2))
    cal = {'thetaprior': prior, 'yvar': yvar}
    np.random.seed(5)
    C.fit(cal, emu, x, y, sampler='PTLMC_B200', numsamp=1000, numtemps=6, numchain=18, sampperchain=120)
    assert cal['thetarnd'].shape == (1000, 2)
    err = np.abs(cal['thetarnd'].mean(0) - theta_true[0])
    report('calibration_ptlmc_b200', err0=float(err[0]), err1=float(err[1]), mean_batch=ncalls['rows'] / ncalls['n'])
    assert np.all(err < 0.1)
    assert ncalls['rows'] / ncalls['n'] > 10                # batched log-posterior calls, also in the pre-optimiser


@pytest.mark.gpu
@pytest.mark.parametrize('name,opts', [('LMC_B200', {'numsamp': 800}),
                                       ('metropolis_hastings_B200', {'numsamp': 4000, 'numchain': 128,
                                                                     'burnSamples': 300,
                                                                     'stepParam': np.array([0.03, 0.03])})])
def test_calibration_with_batched_sampler_plugins(name, opts):
    """directbayeswoodbury_B200 driven by the other two batched sampler plugins."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    from synth import make_observation
    x, theta, f, truth = make_simulator(150, 2, 25, 4, seed=21)
    np.random.seed(21)
    fi = _facade_fit(M, x, theta, f)
    emu = _ShimEmulator(M, fi, theta)
    theta_true, y, yvar = make_observation(truth, 2, seed=22)
    rows = []

    class prior:
        @staticmethod
        def lpdf(th):
            th = np.atleast_2d(th)
            rows.append(th.shape[0])
            return np.where(np.all((th >= 0) & (th <= 1), axis=1), 0.0, -np.inf).reshape(-1, 1)

        @staticmethod
        def lpdf_grad(th):
            return np.zeros_like(np.atleast_2d(th))

        @staticmethod
        def rnd(n):
            return np.random.uniform(size=(n, 2))
    cal = {'thetaprior': prior, 'yvar': yvar}
    np.random.seed(6)
    C.fit(cal, emu, x, y, sampler=name, **opts)
    assert cal['thetarnd'].shape == (opts['numsamp'], 2)
    err = np.abs(cal['thetarnd'].mean(0) - theta_true[0])
    report('calibration_' + name, err0=float(err[0]), err1=float(err[1]), mean_batch=float(np.mean(rows)))
    assert np.all(err < 0.1)
    assert np.mean(rows) > 10


@pytest.mark.gpu
@pytest.mark.parametrize('tag,method', [('knn', 'KNN'), ('br', 'BayesianRidge')])
def test_pcgpwimpute_b200_matches_reference_fit(golden, tag, method):
    """PCGPwImpute_B200 (sklearn completion + the device PCGPwM path) against a seeded fit and
    prediction of the reference's PCGPwImpute (tests/golden/impute.npz)."""
    from surmise_b200.emulationmethods import PCGPwImpute_B200 as M
    g = golden('impute')
    np.random.seed(41)
    fi = {'method': 'PCGPwImpute_B200'}
    M.fit(fi, g['x'], g['theta'], g[tag + '_f'], completionmethod=method)
    assert np.allclose(fi['f'], g[tag + '_fcompleted'], rtol=1e-10, atol=1e-12)
    hyp = np.array([e['hyp'] for e in fi['emulist']])
    assert hyp.shape == g[tag + '_hyp'].shape
    assert np.array_equal([e['hypind'] for e in fi['emulist']], g[tag + '_hypind'])
    assert fi['param_desc'].startswith('\tcompletion method: IterativeImputer:' + method)
    p = {}
    M.predict(p, fi, g['x'], g['thetanew'])
    e_h, e_m, e_v = rel(hyp, g[tag + '_hyp']), rel(p['mean'], g[tag + '_mean']), rel(p['var'], g[tag + '_var'])
    report('pcgpwimpute_' + tag, hyp=e_h, mean=e_m, var=e_v)
    assert e_h < 1e-5 and e_m < 1e-6 and e_v < 1e-5
    with pytest.raises(ValueError):
        M.fit({}, g['x'], g['theta'], g[tag + '_f'], completionmethod='nope')


@pytest.mark.gpu
def test_mlbayeswoodbury_b200_adds_classifier_term():
    """mlbayeswoodbury_B200 = directbayeswoodbury_B200's device log-likelihood + log acceptance probability of a
    classifier over quadratic features (mlbayeswoodbury.py:111-164); SURVEY 8f row 4."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    from surmise_b200.calibrationmethods import mlbayeswoodbury_B200 as ML
    from synth import make_observation
    x, theta, f, truth = make_simulator(120, 2, 20, 3, seed=51)
    np.random.seed(51)
    fi = _facade_fit(M, x, theta, f)
    emu = _ShimEmulator(M, fi, theta)
    theta_true, y, yvar = make_observation(truth, 2, seed=52)

    class prior:
        @staticmethod
        def lpdf(th):
            th = np.atleast_2d(th)
            return (-0.5 * np.sum((th - 0.5) ** 2, 1) / 0.3 ** 2
                    + np.where(np.all((th >= 0) & (th <= 1), axis=1), 0.0, -np.inf)).reshape(-1, 1)

        @staticmethod
        def rnd(n):
            return np.random.uniform(size=(n, 2))

    class clf:                                          # logistic acceptance model over [t1, t2, t1^2, t1 t2, t2^2]
        w = np.array([0.5, -1.0, 0.3, 0.8, -0.4])

        @classmethod
        def predict_proba(cls, feats):
            assert feats.shape[1] == 5
            p = 1 / (1 + np.exp(-(feats @ cls.w)))
            return np.stack([1 - p, p], axis=1)
    assert np.allclose(ML.quadratic_features(np.array([[2.0, 3.0]])), [[2, 3, 4, 6, 9]])
    th = np.vstack([np.random.default_rng(1).uniform(0.05, 0.95, size=(9, 2)), [[1.5, 0.5]]])
    cal = {'thetaprior': prior, 'yvar': yvar}
    base, _ = C.make_logpost(cal, emu, x, y, fd_step=1e-3)
    np.random.seed(3)
    ML.fit(cal, emu, x, y, clf_method=clf, sampler='metropolis_hastings_B200', numsamp=300, numchain=32,
           burnSamples=100, stepParam=np.array([0.03, 0.03]))
    assert cal['thetarnd'].shape == (300, 2) and cal['lpdfapproxnorm_un'].shape == (300, 1)
    # the stored un-normalised log-posterior is prior + device log-likelihood + classifier term
    t = cal['thetarnd'][:20]
    want = base(t, return_grad=False) + np.log(clf.predict_proba(ML.quadratic_features(t))[:, 1]).reshape(-1, 1)
    assert rel(cal['lpdfapproxnorm_un'][:20], want) < 1e-9
    # myusedir=False: the closure handed to the sampler must not return gradients
    seen = {}

    def fake_sampler(logpost_func, draw_func, **kw):
        out = logpost_func(th)
        seen['tuple'] = isinstance(out, tuple)
        seen['val'] = out[0] if isinstance(out, tuple) else out

        class R:
            sampler_info = {'theta': th[:9]}
        return R()
    import surmise_b200.utilities as U
    saved = U.sampler
    try:
        import sys
        U.sampler = fake_sampler
        if 'surmise.utilities' in sys.modules:
            pytest.skip('surmise installed: dispatcher not patchable here')
        ML.fit(dict(cal), emu, x, y, clf_method=clf, myusedir=False)
        assert seen['tuple'] is False and np.isneginf(seen['val'][-1, 0])
        ML.fit(dict(cal), emu, x, y, clf_method=clf, myusedir=True)
        assert seen['tuple'] is True
    finally:
        U.sampler = saved


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['full', 'miss'])
def test_indgp_b200_matches_reference_fit(golden, tag):
    """indGP_B200 (one GP per output row on the PCGPwM kernels) against a fit and prediction of
    the reference's indGP (tests/golden/indgp.npz); 'miss' has rows observed at different parameter subsets."""
    from surmise_b200.emulationmethods import indGP_B200 as M
    g = golden('indgp')
    fi = {'method': 'indGP_B200'}
    M.fit(fi, g['x'], g['theta'], g[tag + '_f'])
    hyp = np.array([e['hyp'] for e in fi['emulist']])
    sig2 = np.array([e['sig2'] for e in fi['emulist']])
    p = {}
    M.predict(p, fi, g['x'], g['thetanew'], return_grad=True)
    errs = {'hyp': rel(hyp, g[tag + '_hyp']), 'sig2': rel(sig2, g[tag + '_sig2'])}
    for k in ('mean', 'var', 'predvecs', 'predvars', 'covxhalf'):
        assert p[k].shape == g[tag + '_' + k].shape
        errs[k] = rel(p[k], g[tag + '_' + k])
    report('indgp_' + tag, **errs)
    assert errs['hyp'] < 1e-5 and errs['mean'] < 1e-6 and errs['var'] < 1e-5 and errs['covxhalf'] < 1e-5
    # the gradient arrays (extra keys here) against central differences of the prediction
    h = 1e-6
    tp, tm = g['thetanew'].copy(), g['thetanew'].copy()
    tp[:, 0] += h
    tm[:, 0] -= h
    pp, pm = {}, {}
    M.predict(pp, fi, g['x'], tp, return_covx=False)
    M.predict(pm, fi, g['x'], tm, return_covx=False)
    nug = np.array([e['nug'] for e in fi['emulist']])
    fd = ((pp['predvecs'] - pm['predvecs']) / (2 * h)).T                      # (B, nx)
    assert rel(p['predvecs_gradtheta'][:, :, 0], fd * (1 - nug)) < 1e-5       # (1 - nug) twice, as the reference
    fdv = ((pp['predvars'] - pm['predvars']) / (2 * h)).T
    assert rel(p['predvars_gradtheta'][:, :, 0], fdv) < 1e-4
    if tag == 'miss':
        p1 = {}
        M.predict(p1, fi, g['x'], g['thetanew'][:1])
        assert rel(p1['mean'], g['miss_one_mean']) < 1e-6 and p1['covxhalf'].shape == g['miss_one_covxhalf'].shape
        assert len(fi['_b200']['groups']) == 3


@pytest.mark.gpu
def test_pcsk_b200_matches_reference_fit(golden):
    """PCSK_B200 (PCGPwM's device search with PCSK's constants) against a seeded fit and prediction
    of the reference's PCSK (tests/golden/pcsk.npz)."""
    from surmise_b200.emulationmethods import PCSK_B200 as M
    g = golden('pcsk')
    np.random.seed(71)
    fi = {'method': 'PCSK_B200'}
    M.fit(fi, g['x'], g['theta'], g['f'], simsd=g['simsd'])
    hyp = np.array([e['hyp'] for e in fi['emulist']])
    assert hyp.shape == g['hyp'].shape
    assert np.array_equal([e['hypind'] for e in fi['emulist']], g['hypind'])
    sign = np.sign(np.sum(fi['pc'] * g['pc'], 0))
    assert rel(fi['pc'] * sign, g['pc']) < 1e-9 and rel(fi['pcstdvar'], g['pcstdvar']) < 1e-9
    p = {}
    M.predict(p, fi, g['x'], g['thetanew'], return_grad=True)
    errs = {'hyp': rel(hyp, g['hyp']), 'sig2': rel([e['sig2'] for e in fi['emulist']], g['sig2'])}
    for k in ('mean', 'var', 'mean_gradtheta'):
        assert p[k].shape == g['p_' + k].shape
        errs[k] = rel(p[k], g['p_' + k])
    errs['predvars'] = rel(p['predvars'], g['p_predvars'])
    errs['covx'] = rel(np.einsum('xbk,ybk->bxy', p['covxhalf'], p['covxhalf']),
                       np.einsum('xbk,ybk->bxy', g['p_covxhalf'], g['p_covxhalf']))      # sign-free
    report('pcsk', **errs)
    assert errs['hyp'] < 1e-5 and errs['mean'] < 1e-6 and errs['var'] < 1e-5 and errs['mean_gradtheta'] < 1e-5
    assert errs['covx'] < 1e-5 and str(g['desc'])[:60] == fi['param_desc'][:60]
    with pytest.raises(AssertionError):
        M.fit({}, g['x'], g['theta'], g['f'])


@pytest.mark.gpu
def test_merge_leaders_option_reduces_distinct_factors():
    """fit_mode='parallel' with merge_leaders=True: fewer full-data factors, same interface, comparable accuracy."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    x, theta, f, truth = make_simulator(400, 3, 40, 8, seed=3)
    tn = np.random.default_rng(1).uniform(0.1, 0.9, size=(50, 3))
    res = {}
    for merge in (False, True):
        np.random.seed(5)
        fi = _facade_fit(M, x, theta, f, fit_mode='parallel', merge_leaders=merge)
        p = {}
        M.predict(p, fi, x, tn)
        inds = np.array([e['hypind'] for e in fi['emulist']])
        assert np.all(inds <= np.arange(len(inds))) and np.all(inds[inds] == inds)      # leaders point at themselves
        for j, e in enumerate(fi['emulist']):
            assert np.array_equal(e['hyp'], fi['emulist'][inds[j]]['hyp'])
        res[merge] = (int(fi['_b200']['Linv'].shape[0]), float(np.sqrt(np.mean((p['mean'] - truth(tn)) ** 2)) / np.std(truth(tn))))
    report('merge_leaders', factors_off=res[False][0], factors_on=res[True][0], rmse_off=res[False][1], rmse_on=res[True][1])
    assert res[True][0] <= res[False][0] and res[True][1] < 2 * res[False][1] + 0.01


def test_clamped_branch_against_the_references_abs_eigenvalues(golden):
    """where R + e^{h_v} diag(gvar) is slightly indefinite (negative rounding noise in the imputation
    variance), the reference keeps going with |eigenvalues| (PCGPwM.py:824-827, 840-844) and PCGPwM_B200 clamps the
    negative variances at 0 and refactors.  This quantifies the difference in what comes out of predict, on the
    reference-fitted missing-data golden with noise-level negative entries injected until the factorisation fails."""
    import warnings
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    g = golden('emu_misspd')
    theta, pc = g['st_theta'], g['st_pc']
    gv = po.damp_gvar(g['st_unscaled_pcstdvar'], 10, 0.3)
    hyp = g['st_hyp']
    k = pc.shape[1]
    # smallest eigenvalue of component 0's covariance, then three negative entries that push it just below zero
    base = po.final_factorisation(theta, pc[:, 0], gv[:, 0], hyp[0])
    lam_min = float(np.linalg.eigvalsh(base['R'])[0])
    assert lam_min > 0
    gv_bad = gv.copy()
    rows = [5, 40, 77]
    gv_bad[rows, 0] = -3.0 * lam_min / np.exp(hyp[0][-2])
    ref_state = {'theta': theta, 'pcti': g['st_pcti'], 'standardpcinfo': {'offset': g['st_offset'], 'scale': g['st_scale'],
                                                                          'extravar': g['st_extravar']}, 'emulist': []}
    for j in range(k):
        o = po.final_factorisation(theta, pc[:, j], gv_bad[:, j], hyp[int(g['st_hypind'][j])])
        o['hypind'] = int(g['st_hypind'][j])
        ref_state['emulist'].append(o)
    assert np.linalg.eigvalsh(ref_state['emulist'][0]['R'])[0] < 0          # the reference's |W| branch is live
    want = {}
    po.predict(want, ref_state, g['in_x'], g['thetanew'])
    spc = ref_state['standardpcinfo']
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter('always')
        emulist, state = M.import_state(theta, pc, gv_bad, hyp, g['st_hypind'], g['st_pcti'], spc)
    assert any('clamped' in str(w.message) for w in caught) and emulist[0]['gvar_clamped']
    got = {}
    M.predict(got, {'_b200': state, 'emulist': emulist, 'x': g['in_x'], 'theta': theta, 'pcti': g['st_pcti'],
                    'standardpcinfo': spc}, g['in_x'], g['thetanew'])
    e_mean = rel(got['mean'], want['mean'])
    e_var = rel(got['var'], want['var'])
    report('clamped_vs_abs_eigenvalues', lam_min=lam_min, mean=e_mean, var=e_var)
    # Both treatments replace a matrix that is not a covariance by a nearby one that is.  Measured on a B200: the
    # predictive mean agrees to 8.7e-6 of its range; the variance differs by up to 0.11 of the largest predictive variance
    # -- the reference divides by |lambda| of the eigenvalue that crossed zero (here -1.2e-5), which inflates the
    # variance along that one direction, the clamp does not.  Neither is "right"; the bound below is what a user of
    # this branch can rely on (the branch warns, and fitinfo['emulist'][j]['gvar_clamped'] records it).
    assert e_mean < 5e-5 and e_var < 0.3
