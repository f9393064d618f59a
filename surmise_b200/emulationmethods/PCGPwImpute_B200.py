"""PCGPwImpute_B200 -- surmise's PCGPwImpute emulation method on the B200 path.

Same contract as surmise/emulationmethods/PCGPwImpute.py (`fit(fitinfo, x, theta, f, epsilonPC, lognugmean, lognugLB,
completionmethod, verbose)`, `predict`, `predictlpdf`; PCGPwImpute.py:23-134): missing simulation outputs are first
completed with scikit-learn's IterativeImputer (KNN, BayesianRidge or RandomForest regressor -- host code, the
reference's own dependency), then the completed data go through the PCGPwM machinery with the variance constant fixed
at 1, no damping (dampalpha = 0) and eta = 1000 (PCGPwImpute.py:86-101).  That machinery -- PCA, the per-component GP
fits, factorisation, prediction -- is PCGPwM_B200's device path; nothing here runs a
GP on the CPU.  Registered by surmise_b200.register() as surmise.emulationmethods.PCGPwImpute_B200.
"""
import numpy as np

from . import PCGPwM_B200 as _pcgpwm

methodoptionstr = ("\nTry one of the following: "
                   "\n'KNN' (k-nearest neighbor method), "
                   "\n'BayesianRidge' (Bayesian ridge regression), "
                   "\n'RandomForest' (random forest method).")


def _complete(f, completionmethod):
    """f (n, m) with NaN/inf -> completed copy (PCGPwImpute.py:106-126)."""
    from sklearn.experimental import enable_iterative_imputer  # noqa: F401
    from sklearn.impute import IterativeImputer
    if completionmethod == 'KNN':
        from sklearn.neighbors import KNeighborsRegressor
        estimator = KNeighborsRegressor()
    elif completionmethod == 'BayesianRidge':
        from sklearn.linear_model import BayesianRidge
        estimator = BayesianRidge()
    elif completionmethod == 'RandomForest':
        from sklearn.ensemble import RandomForestRegressor
        estimator = RandomForestRegressor()
    else:
        raise ValueError('Specify completion method.' + methodoptionstr)
    fhat = IterativeImputer(estimator=estimator).fit_transform(f)
    if not np.isfinite(fhat).all():
        raise ValueError('Completion method (sklearn.impute.IterativeImputer) failed.')
    return fhat


def fit(fitinfo, x, theta, f, epsilonPC=0.001, lognugmean=-10, lognugLB=-20, completionmethod='KNN', verbose=0,
        **kwargs):
    """Complete f, then fit as PCGPwM_B200 with varconstant = 1, dampalpha = 0, eta = 1000.

    Extra keyword arguments (nhyptrain, fit_mode, pca, distributed) are PCGPwM_B200.fit's."""
    _pcgpwm._ops()._torch()                        # fail loudly without CUDA before the (slow) completion
    f = np.asarray(f, dtype=np.float64)
    fitinfo['completionmethod'] = ''
    if not np.all(np.isfinite(f)):
        fin = f.T.copy()
        fin[~np.isfinite(fin)] = np.nan            # IterativeImputer treats NaN as missing
        f = _complete(fin, completionmethod).T
        fitinfo['completionmethod'] = 'IterativeImputer:{:s}'.format(completionmethod)
        if verbose > 0:
            print('Completing f with method IterativeImputer:{:s}.'.format(completionmethod))
    elif verbose > 0:
        print('No missing values identified... Proceeding with PCGP method.')
    for k in ('epsilonImpute', 'varconstant', 'dampalpha', 'eta'):
        kwargs.pop(k, None)
    _pcgpwm.fit(fitinfo, x, theta, f, epsilonPC=epsilonPC, lognugmean=lognugmean, lognugLB=lognugLB, varconstant=1,
                dampalpha=0, eta=1000, verbose=verbose, **kwargs)
    fitinfo['param_desc'] = '\tcompletion method: {:s}\n{:s}'.format(fitinfo['completionmethod'],
                                                                    fitinfo['param_desc'])
    return


def predict(predinfo, fitinfo, x, theta, **kwargs):
    return _pcgpwm.predict(predinfo, fitinfo, x, theta, **kwargs)


def predictlpdf(predinfo, f, addvar=0, **kwargs):
    return _pcgpwm.predictlpdf(predinfo, f, addvar, **kwargs)
