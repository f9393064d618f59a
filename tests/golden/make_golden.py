"""Generate the golden fixtures in this directory by running the REAL reference.

Build-container only (the GPU box has no /root/reference):

    mkdir -p /tmp/ref && cp -r /root/reference/{src,setup.py,README.rst,pyproject.toml} /tmp/ref/
    (cd /tmp/ref && python setup.py build_ext --inplace)
    PYTHONPATH=/tmp/ref/src python tests/golden/make_golden.py

Every array written here is an output of bandframework/surmise's own code
(matern_covmat.covmat, PCGPwM.__negloglik/__negloglikgrad/fit/predict,
directbayeswoodbury.loglik/loglik_grad) on the seeded inputs stored next to it.
The script also prints the oracle-vs-reference discrepancies as a sanity report.
"""
import os
import sys

import numpy as np
import scipy.stats as sps

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from surmise.emulation import emulator                                    # noqa: E402
from surmise.calibration import calibrator                                # noqa: E402
from surmise.emulationsupport.matern_covmat import covmat as ref_covmat   # noqa: E402
import surmise.emulationmethods.PCGPwM as ref_pcgpwm                      # noqa: E402
import surmise.calibrationmethods.directbayeswoodbury as ref_dbw          # noqa: E402

from synth import make_simulator, make_observation                        # noqa: E402
from oracle import pcgpwm_oracle as po                                    # noqa: E402
from oracle import woodbury_oracle as wo                                  # noqa: E402

ref_nll = getattr(ref_pcgpwm, '__negloglik')
ref_nllgrad = getattr(ref_pcgpwm, '__negloglikgrad')


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def golden_covmat():
    rng = np.random.default_rng(11)
    out = {}
    for tag, (n1, n2, d) in {'a': (7, 5, 3), 'b': (33, 33, 8), 'c': (4, 6, 1)}.items():
        x1 = rng.uniform(size=(n1, d))
        x2 = x1.copy() if n1 == n2 else rng.uniform(size=(n2, d))
        gam = rng.normal(size=d + 1) * 0.7
        R = ref_covmat(x1, x2, gam)
        _, dRh = ref_covmat(x1, x2, gam, return_gradhyp=True)
        _, dRx = ref_covmat(x1, x2, gam, return_gradx1=True)
        out.update({f'{tag}_x1': x1, f'{tag}_x2': x2, f'{tag}_gam': gam, f'{tag}_R': R,
                    f'{tag}_dRhyp': np.ascontiguousarray(dRh), f'{tag}_dRx1': np.ascontiguousarray(dRx)})
        print('covmat', tag, rel(po.matern_covmat(x1, x2, gam), R),
              rel(po.matern_covmat(x1, x2, gam, 'hyp')[1], dRh),
              rel(po.matern_covmat(x1, x2, gam, 'x1')[1], dRx))
    # 1-D x1 edge case (matern_covmat.pyx:29-30)
    x1 = rng.uniform(size=3)
    x2 = rng.uniform(size=(5, 3))
    gam = rng.normal(size=4) * 0.5
    out.update({'e_x1': x1, 'e_x2': x2, 'e_gam': gam, 'e_R': ref_covmat(x1, x2, gam)})
    np.savez_compressed(os.path.join(HERE, 'covmat.npz'), **out)


def golden_nll():
    rng = np.random.default_rng(12)
    out = {}
    for tag, (n, d, with_gvar) in {'a': (40, 3, False), 'b': (96, 8, True), 'c': (25, 1, True)}.items():
        theta = rng.uniform(size=(n, d))
        g = np.sin(3 * theta.sum(1)) + 0.1 * rng.normal(size=n)
        gvar = (rng.uniform(size=n) * (rng.uniform(size=n) < 0.3) * 0.4) if with_gvar else np.zeros(n)
        mean, LB, UB, std = po.hyper_regulariser(theta, -10, -20)
        info = {'theta': theta, 'g': g, 'gvar': gvar, 'sig2ofconst': 0.01,
                'hypregmean': mean, 'hypregstd': std}
        hyps = [mean.copy()]
        for _ in range(5):
            hyps.append(LB + rng.uniform(0.25, 0.75, size=mean.shape) * (UB - LB))
        hyps = np.array(hyps)
        nll = np.array([ref_nll(h, info) for h in hyps])
        grad = np.array([ref_nllgrad(h, info) for h in hyps])
        cond = np.array([np.linalg.cond(po.assemble_gp_matrix(theta, h, gvar)[0]) for h in hyps])
        out.update({f'{tag}_theta': theta, f'{tag}_g': g, f'{tag}_gvar': gvar, f'{tag}_mean': mean,
                    f'{tag}_std': std, f'{tag}_hyps': hyps, f'{tag}_nll': nll, f'{tag}_grad': grad,
                    f'{tag}_cond': cond})
        e1 = max(abs(po.negloglik_eigh(h, info) - v) / abs(v) for h, v in zip(hyps, nll))
        e2 = max(rel(po.negloglikgrad_eigh(h, info), v) for h, v in zip(hyps, grad))
        e3 = max(abs(po.negloglik_and_grad_chol(h, info)[0] - v) / abs(v) for h, v in zip(hyps, nll))
        e4 = max(rel(po.negloglik_and_grad_chol(h, info)[1], v) for h, v in zip(hyps, grad))
        print('nll', tag, 'eigh', e1, e2, 'chol', e3, e4, 'cond max %.2e' % cond.max())
    np.savez_compressed(os.path.join(HERE, 'nll.npz'), **out)


def _emu_state(emu):
    info = emu._info
    el = info['emulist']
    spc = info['standardpcinfo']
    st = {'theta': info['theta'], 'pc': info['pc'], 'pcti': info['pcti'], 'pct': info['pct'],
          'pcw': info['pcw'], 'unscaled_pcstdvar': info['unscaled_pcstdvar'],
          'offset': spc['offset'], 'scale': spc['scale'], 'extravar': spc['extravar'],
          'hyp': np.array([e['hyp'] for e in el]), 'hypind': np.array([e['hypind'] for e in el]),
          'nug': np.array([e['nug'] for e in el]), 'sig2': np.array([e['sig2'] for e in el]),
          'pw': np.array([e['pw'] for e in el]).T,
          'condR': np.array([np.linalg.cond(e['R']) for e in el])}
    return st


def golden_emulator(tag, n, d, m, k, seed, missing, nB, fit_args=None, yvar_scale=1.0):
    x, theta, f, truth = make_simulator(n, d, m, k, seed=seed, missing=missing)
    np.random.seed(seed)
    emu = emulator(x, theta, f, method='PCGPwM', args=fit_args or {})
    st = _emu_state(emu)
    # oracle fit on the same RNG stream
    np.random.seed(seed)
    ofit = {}
    po.fit(ofit, x, theta, f, **(fit_args or {}))
    print(tag, 'k =', st['hyp'].shape[0], 'hypind', st['hypind'], 'cond max %.2e' % st['condR'].max())
    print(tag, 'oracle fit vs ref: hyp', rel(np.array([e['hyp'] for e in ofit['emulist']]), st['hyp']),
          'pw', rel(np.array([e['pw'] for e in ofit['emulist']]).T, st['pw']),
          'pc', rel(ofit['pc'], st['pc']), 'extravar', rel(ofit['standardpcinfo']['extravar'], st['extravar']))
    rng = np.random.default_rng(seed + 100)
    thetanew = rng.uniform(0.05, 0.95, size=(nB, d))
    p = emu.predict(x, thetanew, args={'return_grad': True})
    pi = p._info
    out = {f'in_{a}': b for a, b in (('x', x), ('theta', theta), ('f', f))}
    out.update({f'st_{a}': b for a, b in st.items()})
    out.update({'thetanew': thetanew, 'predvecs': pi['predvecs'], 'predvars': pi['predvars'],
                'predvecs_grad': pi['predvecs_gradtheta'], 'predvars_grad': pi['predvars_gradtheta'],
                'mean': pi['mean'], 'var': pi['var'], 'mean_grad': pi['mean_gradtheta'],
                'covxhalf_sample': np.ascontiguousarray(pi['covxhalf'][:, :2, :]),
                'covxhalf_grad_sample': np.ascontiguousarray(pi['covxhalf_gradtheta'][:, :2, :, :])})
    opred = {}
    po.predict(opred, emu._info, x, thetanew, return_grad=True)
    print(tag, 'oracle predict on ref state: mean', rel(opred['mean'], pi['mean']), 'var', rel(opred['var'], pi['var']),
          'dmean', rel(opred['mean_gradtheta'], pi['mean_gradtheta']),
          'dcovx', rel(opred['covxhalf_gradtheta'], pi['covxhalf_gradtheta']),
          'dpvar', rel(opred['predvars_gradtheta'], pi['predvars_gradtheta']))
    # Cholesky-form restatement at the reference's fitted hyper-parameters (what the GPU computes)
    try:
        states = [po.factor_state_chol(theta, st['pc'][:, j], po.damp_gvar(st['unscaled_pcstdvar'][:, j], 10, 0.3),
                                       st['hyp'][j]) for j in range(st['hyp'].shape[0])]
        c = po.predict_pcspace_chol(theta, states, st['hypind'], thetanew, return_grad=True)
        print(tag, 'chol-form vs ref: pw', rel(np.array([s_['pw'] for s_ in states]).T, st['pw']),
              'sig2', rel(np.array([s_['sig2'] for s_ in states]), st['sig2']),
              'predvecs', rel(c[0], pi['predvecs']), 'predvars', float(np.max(np.abs(c[1] - pi['predvars']) / pi['predvars'])),
              'dpv', rel(c[2], pi['predvecs_gradtheta']), 'dpvar', rel(c[3], pi['predvars_gradtheta']))
    except np.linalg.LinAlgError:
        print(tag, 'chol-form: R not positive definite at the reference hyper-parameters')
    # at-design predictions (cancellation regime), first 8 design points
    pdz = emu.predict(x, theta[:8])._info
    out.update({'design_predvecs': pdz['predvecs'], 'design_predvars': pdz['predvars']})
    # predictive log-density of new simulator output (PCGPwM.predictlpdf, PCGPwM.py:331-364)
    flp = truth(thetanew) + 0.01 * np.random.default_rng(seed + 7).normal(size=(x.shape[0], nB))
    lp, dlp = ref_pcgpwm.predictlpdf(pi, flp, addvar=0.003)
    out.update({'lpdf_f': flp, 'lpdf': lp, 'dlpdf': dlp})
    olp, odlp = po.predictlpdf(opred, flp, addvar=0.003)
    print(tag, 'oracle predictlpdf', rel(olp, lp), rel(odlp, dlp))
    # calibration log-likelihood
    theta_true, y, yvar = make_observation(truth, d, seed=seed + 1)
    yvar = yvar * yvar_scale
    cfit = {'yvar': yvar}
    ll = ref_dbw.loglik(cfit, emu, thetanew, y, x)
    ll2, dll = ref_dbw.loglik_grad(cfit, emu, thetanew, y, x)
    fired = wo.extravar_trigger(pi['var'], pi['covxhalf'])
    out.update({'y': y, 'yvar': yvar, 'loglik': ll, 'loglik_g': ll2, 'dloglik': dll, 'fired': np.array(fired)})
    oll, odll = wo.loglik_mspace(emu._info, yvar, thetanew, y, return_grad=True)
    spc = emu._info['standardpcinfo']
    phi = emu._info['pcti'] * spc['scale'][:, None]
    kll, kdll, kf = wo.loglik_pcspace(pi['predvecs'], pi['predvars'], phi, spc['offset'], spc['extravar'],
                                      yvar, y, pi['predvecs_gradtheta'], pi['predvars_gradtheta'])
    print(tag, 'trigger fired', fired, kf, 'loglik: mspace', rel(oll, ll), rel(odll, dll),
          'pcspace', float(np.max(np.abs(kll - ll) / np.abs(ll))), rel(kdll, dll))
    np.savez_compressed(os.path.join(HERE, f'emu_{tag}.npz'), **out)
    return emu


def golden_standardize():
    """Imputation / PCA internals on a small NaN case (PCGPwM.py:533-654)."""
    x, theta, f, _ = make_simulator(60, 2, 30, 4, seed=5, missing=0.15)
    fi = {'f': f.T.copy(), 'epsilonPC': 0.001, 'epsilonImpute': 10e-6}
    fi['mof'] = np.logical_not(np.isfinite(fi['f']))
    fi['mofrows'] = np.where(np.any(fi['mof'] > 0.5, 1))[0]
    getattr(ref_pcgpwm, '__standardizef')(fi)
    getattr(ref_pcgpwm, '__PCs')(fi)
    spc = fi['standardpcinfo']
    np.savez_compressed(os.path.join(HERE, 'standardize.npz'), f=f, offset=spc['offset'], scale=spc['scale'],
                        fs=spc['fs'], extravar=spc['extravar'], S=spc['S'], pcw=fi['pcw'], pc=fi['pc'],
                        pct=fi['pct'], pcti=fi['pcti'], unscaled_pcstdvar=fi['unscaled_pcstdvar'])
    fo = {'f': f.T.copy(), 'epsilonPC': 0.001, 'epsilonImpute': 10e-6}
    fo['mof'] = np.logical_not(np.isfinite(fo['f']))
    fo['mofrows'] = np.where(np.any(fo['mof'] > 0.5, 1))[0]
    po.standardize_f(fo)
    po.principal_components(fo)
    print('standardize: fs', rel(fo['standardpcinfo']['fs'], spc['fs']), 'pc', rel(fo['pc'], fi['pc']),
          'pcstdvar', rel(fo['unscaled_pcstdvar'], fi['unscaled_pcstdvar']))


def golden_d8():
    """A C2-like case (d = 8): cond(R) ~ 1e5-1e6, the conditioning SURVEY measured at C2 -- here BASELINE.json's 1e-9 bar on
    the per-theta log-likelihood must hold end to end (tests/test_gpu_path.py)."""
    golden_emulator('d8', n=300, d=8, m=40, k=10, seed=21, missing=0.0, nB=16)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'd8':      # PYTHONPATH=baseline/_ref python tests/golden/make_golden.py d8
        golden_d8()
        sys.exit(0)
    golden_covmat()
    golden_nll()
    golden_standardize()
    golden_emulator('c1', n=200, d=3, m=50, k=8, seed=3, missing=0.0, nB=24)
    golden_emulator('miss', n=120, d=2, m=30, k=4, seed=7, missing=0.10, nB=16)
    # same data, but the nugget floor / variance constant keep R positive definite despite the
    # rounding-noise (+-5e-8) entries of unscaled_pcstdvar -> usable for Cholesky parity
    golden_emulator('misspd', n=120, d=2, m=30, k=4, seed=7, missing=0.10, nB=16,
                    fit_args={'lognugLB': -12, 'varconstant': float(np.exp(-2))})
    # coarse epsilonPC -> extravar > 1e-4 -> the batch-wide obsvar trigger fires (dbw.py:277-284)
    golden_emulator('trig', n=100, d=2, m=24, k=6, seed=9, missing=0.0, nB=12,
                    fit_args={'epsilonPC': 0.5}, yvar_scale=1.0)
    golden_emulator('d1', n=60, d=1, m=21, k=3, seed=13, missing=0.0, nB=10)
    golden_d8()
