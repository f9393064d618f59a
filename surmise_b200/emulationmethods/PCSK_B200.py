"""PCSK_B200 -- surmise's PCSK emulation method (principal components with stochastic kriging) on the B200 path.

Same contract as surmise/emulationmethods/PCSK.py (`fit(fitinfo, x, theta, f, epsilonPC, lognugmean, lognugLB,
varconstant, dampalpha, eta, simsd, numpcs, standardpcinfo, verbose)`, `predict`; PCSK.py:9-317).  `simsd` (m, n),
the standard deviation of the stochastic simulator at every output, is required; it decides which components are
kept and is carried as `pcstdvar`, but the GP likelihood itself has no per-point variance term (PCSK.py:565-616).
SURVEY.md section 8(f) row 4: the per-component hyper-parameter search is PCGPwM's with other constants -- prior and
bounds (PCSK.py:467-479), min(25 d, n) training points (481), two sweeps (440), gtol 0.05 (524), no skip test (503),
share bar 1.1 (p + 2 sqrt p) (531-532), sig2ofconst 1e-5 for a search without leaders (458) -- so it runs on
PCGPwM_B200's device likelihood, factorisation and predict kernels; `predict` is PCGPwM_B200.predict (the
reference's two predict functions are the same computation, PCSK.py:132-317 vs PCGPwM.py:140-328).
Registered by surmise_b200.register() as surmise.emulationmethods.PCSK_B200.
"""
import numpy as np

from . import PCGPwM_B200 as _pcgpwm


def hyper_prior(theta):
    """PCSK.py:467-479: [length-scales, scale, log var-const, logit nugget]."""
    base = 0.5 * np.log(theta.shape[1]) + np.log(np.std(theta, 0))
    mean = np.append(base, (0, 0, -17))
    lo = np.append(-4 + base, (-12, -3, -20))
    hi = np.append(4 + base, (6, 3, -12))
    return mean, lo, hi, (hi - lo) / 4


def standardize(f):
    """Column-wise standardisation of the (n, m) outputs, PCSK has no imputation (PCSK.py:353-386): the offset is the
    mean over the finite cells, the scale their standard deviation inflated by 1/sqrt(observed fraction)."""
    seen = np.isfinite(f) | np.isinf(f)                       # only NaN counts as missing, as in the reference
    count = seen.sum(axis=0)
    offset = np.nansum(f, axis=0) / count
    centred = np.where(seen, f - offset, 0.0)
    scale = np.sqrt((centred ** 2).sum(axis=0) / count) / np.sqrt(count / f.shape[0])
    if not np.all(scale != 0):
        raise ValueError("You have a row that is non-varying.")
    return dict(offset=offset, scale=scale, fs=(f - offset) / scale, extravar=np.zeros_like(scale))


def principal_components(fitinfo, simsd):
    """Loadings, scores and the per-component share of simulation noise (PCSK.py:389-424).  A user-supplied basis is
    taken as is; otherwise the left singular vectors of fs' are kept where S^2 > epsilonPC and the standardised
    simulation variance, projected on the component, stays below 8 times the component's own variance (or the first
    `numpcs` of them when that is positive)."""
    spc = fitinfo['standardpcinfo']
    fs = spc['fs']
    noise = np.square(simsd.T / spc['scale'])                 # (n, m) simulation variances in standardised units

    def noise_share(basis):
        return (noise @ np.square(basis)) / np.var(fs @ basis, axis=0)

    if 'U' in spc:
        basis, weights = np.array(spc['U'], copy=True), np.array(spc['S'], copy=True)
    else:
        U, S, _vt = np.linalg.svd(fs.T, full_matrices=False)
        if fitinfo['numpcs'] > 0:
            chosen = np.arange(fitinfo['numpcs'])
        else:
            chosen = np.flatnonzero((noise_share(U).mean(axis=0) < 8.0) & (S ** 2 > fitinfo['epsilonPC']))
        basis, weights = U[:, chosen], S[chosen]
    scores = fs @ basis
    fitinfo.update(pcw=weights, pcto=basis.copy(), pct=basis, pcti=basis, pc=scores,
                   unscaled_pcstdvar=(noise @ np.square(basis)) / scores.var(axis=0))


def fit(fitinfo, x, theta, f, epsilonPC=0.001, lognugmean=-10, lognugLB=-20, varconstant=None, dampalpha=0.3,
        eta=10, simsd=None, numpcs=-1, standardpcinfo=None, verbose=0, **kwargs):
    """PCSK.fit (PCSK.py:9-129) with the GP searches and factorisations on the device."""
    _pcgpwm._ops()._torch()                        # fail loudly without CUDA
    assert simsd is not None, 'Variable `simsd` must be provided for PCSK method.'
    f = np.asarray(f, dtype=np.float64).T
    theta = np.asarray(theta, dtype=np.float64)
    fitinfo['epsilonPC'] = epsilonPC
    fitinfo['dampalpha'] = dampalpha
    fitinfo['eta'] = eta
    fitinfo['theta'] = theta
    fitinfo['f'] = f
    fitinfo['x'] = x
    fitinfo['numpcs'] = numpcs
    if standardpcinfo is None:
        fitinfo['standardpcinfo'] = standardize(f)
    else:
        fitinfo['standardpcinfo'] = standardpcinfo
        fitinfo['standardpcinfo']['fs'] = (f - standardpcinfo['offset']) / standardpcinfo['scale']
    principal_components(fitinfo, np.asarray(simsd, dtype=np.float64))
    numpcs = fitinfo['pc'].shape[1]
    fitinfo['numpcs'] = numpcs
    if verbose > 0:
        print(fitinfo.get('method', 'PCSK_B200'), 'considering ', numpcs, 'PCs')
    variant = {'prior': hyper_prior(theta), 'nh_rule': lambda d, n: int(np.max(np.min((25 * d, n)))),
               'share_bar': lambda p: 1.1 * (p + 2 * np.sqrt(p)), 'sweeps': 2, 'gtol': 0.05, 'skip_test': False,
               'use_gvar': False, 's0_noleader': 0.00001}
    fitter = _pcgpwm._GPFitter(fitinfo, theta, lognugmean, lognugLB, None, None, variant=variant)
    emulist = fitter.run(previous=fitinfo.get('emulist'))
    fitinfo['varc_status'] = 'fixed' if varconstant is not None else 'optimized'
    fitinfo['pcstdvar'] = fitinfo['unscaled_pcstdvar']
    fitinfo['emulist'] = emulist
    fitinfo['mof'], fitinfo['mofrows'] = None, None
    fitinfo['_b200'] = fitter.device_state(emulist, fitinfo['standardpcinfo'], fitinfo['pcti'])
    fitinfo['param_desc'] = _pcgpwm._param_desc(fitinfo)
    return


def predict(predinfo, fitinfo, x, theta, **kwargs):
    return _pcgpwm.predict(predinfo, fitinfo, x, theta, **kwargs)
