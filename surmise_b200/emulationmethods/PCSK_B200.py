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
    """PCSK.py:353-386 for complete data: column offset / scale, no imputation, extravar = 0."""
    offset = np.zeros(f.shape[1])
    scale = np.zeros(f.shape[1])
    for k in range(f.shape[1]):
        offset[k] = np.nanmean(f[:, k])
        scale[k] = np.nanstd(f[:, k]) / np.sqrt(1 - np.isnan(f[:, k]).mean())
        if scale[k] == 0:
            raise ValueError("You have a row that is non-varying.")
    return {'offset': offset, 'scale': scale, 'fs': (f - offset) / scale, 'extravar': 0 * scale}


def principal_components(fitinfo, simsd):
    """PCSK.py:389-424: components kept where S^2 > epsilonPC and the simulation noise does not dominate."""
    spc = fitinfo['standardpcinfo']
    fs = spc['fs']
    stdvarsadj = (simsd.T / spc['scale']) ** 2
    if 'U' in spc:
        pct, pcw = 1 * spc['U'], 1 * spc['S']
    else:
        U, S, _ = np.linalg.svd(fs.T, full_matrices=False)
        pc = fs @ U
        ucpcsc = (stdvarsadj @ (U ** 2)) / (pc.var(0))
        if fitinfo['numpcs'] <= 0:
            keep = (ucpcsc.mean(0) < 8.) * (S ** 2 > fitinfo['epsilonPC'])
        else:
            keep = range(0, fitinfo['numpcs'])
        pct = U[:, keep]
        pcw = np.sqrt((S ** 2)[keep])
    fitinfo['pcw'] = pcw
    fitinfo['pcto'] = 1 * pct
    fitinfo['pct'] = pct
    fitinfo['pcti'] = pct
    fitinfo['pc'] = fs @ pct
    fitinfo['unscaled_pcstdvar'] = (stdvarsadj @ (pct ** 2)) / (fitinfo['pc'].var(0))


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
