"""indGP_B200 -- surmise's indGP emulation method (one independent GP per output location) on the B200 path.

Same contract as surmise/emulationmethods/indGP.py (`fit(fitinfo, x, theta, f, lognugmean, lognugLB)`,
`predict(predinfo, fitinfo, x, theta, return_grad=, return_covx=)`; indGP.py:9-330): every row of f gets its own GP
on the parameters at which that row was observed, with d + 2 hyper-parameters [gamma_0..gamma_d, logit nugget]
(no variance constant) and sig2ofconst = 1 (indGP.py:123-172).  SURVEY.md section 8(f) row 4: the covariance, the
penalised likelihood + gradient, the factorisation and the prediction are the PCGPwM kernels --
  * rows with the same missing pattern share theta, so their likelihoods are ONE batched device call per optimiser
    step (the L-BFGS-B runs of a group are advanced together by one scheduler over scipy's own L-BFGS-B core); the kernel's log-variance-constant slot is pinned at its
    prior mean, where it contributes nothing;
  * one batched factorisation per group gives Linv, pw and sig2; prediction reuses sb200_gp_predict per group.
The reference's quirks are kept: `covxhalf` has shape (nx, B, B) with sqrt(var[k, b]) on the diagonal of its last two
axes (indGP.py:301-308).  The reference computes `predvecs_gradtheta` / `predvars_gradtheta` with return_grad and then
drops them (indGP.py:253-292, 296-299); here they are stored in predinfo as extra keys, (B, nx, d), with the
reference's (1 - nug)-twice convention.
Registered by surmise_b200.register() as surmise.emulationmethods.indGP_B200.
"""
import threading
import warnings
from pprint import pformat

import numpy as np
import scipy.optimize as spo

from ..utilitiesmethods.PTLMC_B200 import _Lockstep

SIG2_OF_CONST = 1.0                     # indGP.py:123
_MAX_LOCKSTEP = 256                     # optimiser threads per lock-step batch


def _ops():
    from .. import ops
    return ops


def hyper_prior(theta, lognugmean, lognugLB):
    """Regulariser mean / bounds / std of the d + 2 hyper-parameters (indGP.py:126-135)."""
    base = 0.5 * np.log(theta.shape[1]) + np.log(np.std(theta, 0))
    mean = np.append(base, (0, lognugmean))
    lo = np.append(-4 + base, (-12, lognugLB))
    hi = np.append(4 + base, (2, -8))
    std = (hi - lo) / 8
    std[-2] = 2
    std[-1] = 4
    return mean, lo, hi, std


def _widen(a, fill):
    """(.., d+2) indGP hyper-parameter layout -> the kernels' (.., d+3) layout with the variance-constant slot."""
    a = np.asarray(a, dtype=np.float64)
    return np.concatenate([a[..., :-1], np.full(a.shape[:-1] + (1,), fill), a[..., -1:]], axis=-1)


def _fit_group(theta_g, G, lognugmean, lognugLB):
    """All GPs that share the design theta_g (n_g, d); G: their standardised outputs (K, n_g)."""
    ops = _ops()
    torch = ops._torch()
    K, n = G.shape
    d = theta_g.shape[1]
    mean, lo, hi, std = hyper_prior(theta_g, lognugmean, lognugLB)
    mean3, std3 = _widen(mean, 0.0), _widen(std, 1.0)
    dev = torch.device('cuda', torch.cuda.current_device())
    theta_dev = torch.from_numpy(np.ascontiguousarray(theta_g)).to(dev)
    G_dev = torch.from_numpy(np.ascontiguousarray(G)).to(dev)
    mean_dev, std_dev = torch.from_numpy(mean3).to(dev), torch.from_numpy(std3).to(dev)
    bounds = spo.Bounds((lo - mean) / std, (hi - mean) / std)
    hyps = np.empty((K, d + 2))
    nevals = 0

    def batched(reqs):
        """reqs (nreq, 1 + p): [row of G | scaled hyper-parameters] of every optimiser that is waiting."""
        rows = reqs[:, 0].astype(np.int64)
        data = ops.GPData(theta_dev, G_dev[torch.from_numpy(rows).to(dev)], None)
        # variance-constant slot pinned at -1e-8: its prior term (1e-8 + h - 0) / 1 vanishes (PCGPwM.py:866)
        hyp3 = _widen(mean + reqs[:, 1:] * std, -1e-8)
        nll, grad, info = ops.gp_nll_grad(data, hyp3, mean_dev, std_dev, SIG2_OF_CONST, True)
        bad = info != 0                                           # not positive definite: infeasible point
        g = np.delete(grad, d + 1, axis=1) * std
        g[bad] = 0.0
        return np.where(bad, 1e30, nll), g

    from .PCGPwM_B200 import _LbfgsbCore, minimize_many
    if _LbfgsbCore.available():
        # all optimisers of the group advanced by one scheduler (scipy's own L-BFGS-B core, reverse communication)
        X, _fun, nfev = minimize_many(lambda idx, Xs: batched(np.column_stack([idx.astype(np.float64), Xs])),
                                      np.zeros((K, d + 2)), bounds.lb, bounds.ub, 0.1)
        hyps[:] = mean + X * std
        nevals += int(nfev.sum())
    else:
        for c0 in range(0, K, _MAX_LOCKSTEP):                       # fallback: one thread per optimiser
            members = list(range(c0, min(c0 + _MAX_LOCKSTEP, K)))
            ev = _Lockstep(batched, len(members))
            results = {}

            def work(j, ev=ev, results=results):
                try:
                    results[j] = spo.minimize(lambda xs: ev(np.concatenate(([float(j)], xs))), np.zeros(d + 2),
                                              method='L-BFGS-B', jac=True, bounds=bounds, options={'gtol': 0.1})
                finally:
                    ev.retire()
            threads = [threading.Thread(target=work, args=(j,)) for j in members]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            for j in members:
                hyps[j] = mean + results[j].x * std
                nevals += results[j].nfev
    # full factorisation of every GP of the group in one batch (indGP.py:158-171)
    data = ops.GPData(theta_dev, G_dev, None)
    Linv, pw, sig2, _ldh, info = ops.gp_factorize(data, _widen(hyps, 0.0), SIG2_OF_CONST)
    if np.any(info != 0):
        raise ValueError('indGP_B200: covariance not positive definite at the fitted hyper-parameters.')
    nug = np.exp(hyps[:, -1]) / (1 + np.exp(hyps[:, -1]))
    ar = torch.arange(K, dtype=torch.int32, device=dev)
    state = {'theta': theta_dev, 'Linv': Linv, 'pw': pw, 'sig2': sig2, 'nug': torch.from_numpy(nug).to(dev),
             'hypcov': torch.from_numpy(np.ascontiguousarray(hyps[:, :-1])).to(dev), 'fidx': ar, 'fu': ar.clone()}
    pw_h, sig2_h = ops.to_host([pw, sig2])
    subinfos = []
    for j in range(K):
        subinfos.append({'hyp': hyps[j], 'hypcov': hyps[j, :-1], 'nug': nug[j], 'sig2': sig2_h[j], 'pw': pw_h[j],
                         'theta': theta_g, 'g': G[j], 'sig2ofconst': SIG2_OF_CONST, 'hypregmean': mean,
                         'hypregLB': lo, 'hypregUB': hi, 'hypregstd': std})
    return subinfos, state, nevals


def fit(fitinfo, x, theta, f, lognugmean=-10, lognugLB=-20, **kwargs):
    """One GP per row of f (indGP.py:9-58)."""
    _ops()._torch()                                # fail loudly without CUDA
    f = np.asarray(f, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64)
    mof = np.logical_not(np.isfinite(f))
    fitinfo['mof'] = mof
    fitinfo['mofrows'] = np.where(np.any(mof > 0.5, 1))[0]
    fitinfo['mofall'] = np.where(np.all(mof, 1))[0]
    if len(fitinfo['mofall']) > 0:
        warnings.warn('Some rows are completely missing. An emulator will still be built, but will return NaN '
                      'predictions at those corresponding locations. If you are to proceed with calibration with '
                      'surmise, remove these rows.', stacklevel=2)
    nx = x.shape[0]
    fitinfo['theta'], fitinfo['f'], fitinfo['x'], fitinfo['nx'] = theta, f, x, nx
    buildinds = np.delete(np.arange(nx), fitinfo['mofall'])
    fitinfo['buildinds'] = buildinds
    # standardisation per row over its finite entries (indGP.py:61-92)
    offset, scale = np.full(nx, np.nan), np.full(nx, np.nan)
    fnan = np.where(mof, np.nan, f)
    for k in buildinds:
        offset[k] = np.nanmean(fnan[k])
        scale[k] = np.nanstd(fnan[k])
    if (scale == 0).any():
        raise ValueError('One of the rows in f is non-varying.')
    fs = ((f.T - offset) / scale).T
    fitinfo['offset'], fitinfo['scale'], fitinfo['fs'] = offset, scale, fs
    # rows observed at the same parameters form one batched group
    groups = {}
    for j in buildinds:
        groups.setdefault(mof[j].tobytes(), []).append(int(j))
    emulist = [dict() for _ in range(nx)]
    states, nevals = [], 0
    for rows in groups.values():
        keep = ~mof[rows[0]]
        subinfos, state, ne = _fit_group(theta[keep], fs[np.ix_(rows, np.where(keep)[0])], lognugmean, lognugLB)
        nevals += ne
        for j, info in zip(rows, subinfos):
            emulist[j] = info
        states.append({'rows': np.array(rows), 'state': state})
    fitinfo['emulist'] = emulist
    fitinfo['_b200'] = {'groups': states, 'nevals': nevals}
    built = [emulist[k] for k in buildinds]
    fitinfo['param_desc'] = ('\tnumber of GP components:\t{:d}\n'
                             '\tGP parameters, following Gramacy (ch.5, 2022) notations:\n'
                             '\t\tlengthscales (in log):\n\t\t\t{:s}\n'
                             '\t\tnuggets (in log):\t{:s}\n').format(
        len(emulist), pformat(np.array([e['hypcov'] for e in built])).replace('\n', '\n\t\t\t'),
        pformat(['{:.3f}'.format(np.log(e['nug'])) for e in built]))
    return


def predict(predinfo, fitinfo, x, theta, **kwargs):
    """Predictive mean / variance per output location (indGP.py:215-310), same keys and shapes."""
    ops = _ops()
    return_grad = kwargs.get('return_grad', False) is True
    return_covx = not (kwargs.get('return_covx', True) is False)
    theta = np.asarray(theta, dtype=np.float64)
    if theta.ndim < 2:
        theta = theta.reshape(1, -1)
    B, d = theta.shape
    nx = len(fitinfo['emulist'])
    predvecs = np.full((nx, B), np.nan)
    predvars = np.full((nx, B), np.nan)
    if return_grad:
        dvecs = np.zeros((B, nx, d))
        dvars = np.zeros((B, nx, d))
    for grp in fitinfo['_b200']['groups']:
        rows = grp['rows']
        pv, pvar, dpv, dpvar = ops.gp_predict(grp['state'], theta, return_grad)
        pv_h, pvar_h, dpv_h, dpvar_h = ops.to_host([pv, pvar, dpv, dpvar])
        predvecs[rows] = pv_h.T
        predvars[rows] = pvar_h.T
        if return_grad:
            dvecs[:, rows, :] = dpv_h
            dvars[:, rows, :] = dpvar_h
    predinfo['mean'] = (predvecs.T * fitinfo['scale'] + fitinfo['offset']).T
    predinfo['var'] = (predvars.T * (fitinfo['scale'] ** 2)).T
    predinfo['predvecs'] = predvecs
    predinfo['predvars'] = predvars
    if return_grad:
        predinfo['predvecs_gradtheta'] = dvecs
        predinfo['predvars_gradtheta'] = dvars
    if return_covx:
        CH = np.sqrt(predinfo['var'])
        cxh = np.zeros((nx, B, B))
        ii = np.arange(B)
        cxh[:, ii, ii] = CH                            # np.diag(CH[k]) per location (indGP.py:303-308)
        predinfo['covxhalf'] = cxh
    return
