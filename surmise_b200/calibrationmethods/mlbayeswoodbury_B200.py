"""mlbayeswoodbury_B200 -- surmise's mlbayeswoodbury calibration method on the B200 path.

Same contract as surmise/calibrationmethods/mlbayeswoodbury.py (`fit(fitinfo, emu, x, y, clf_method=None,
myusedir=True, **sampler_args)`, `predict`, `thetarnd`, `thetalpdf`, `loglik`, `loglik_grad`;
mlbayeswoodbury.py:7-417): directbayeswoodbury's Woodbury log-likelihood, plus an optional classifier
`clf_method` (anything with `predict_proba` over [theta, theta_i theta_j for i <= j]) whose log acceptance
probability is added to the log-posterior (mlbayeswoodbury.py:111-131), a 1e-3 finite-difference step for priors
without `lpdf_grad` (mlbayeswoodbury.py:92-108), and the un-normalised log-posterior of the draws kept as
fitinfo['lpdfapproxnorm_un'].  The log-likelihood and its gradient are directbayeswoodbury_B200's fused device
evaluation; the classifier term is host code, as in the reference.
Registered by surmise_b200.register() as surmise.calibrationmethods.mlbayeswoodbury_B200.
"""
import numpy as np

from . import directbayeswoodbury_B200 as _dbw
from .directbayeswoodbury_B200 import predict, thetarnd, thetalpdf, loglik_grad   # noqa: F401  (same functions)


def loglik(fitinfo, emu, theta, y, x, args=None):
    """(B, 1) log-likelihoods (mlbayeswoodbury.py:297-342; `args` is accepted and unused, as there)."""
    return _dbw.loglik(fitinfo, emu, theta, y, x)


def quadratic_features(theta):
    """[theta, theta_i * theta_j for i <= j] in the reference's column order (mlbayeswoodbury.py:112-117)."""
    theta = np.asarray(theta, dtype=np.float64)
    d = theta.shape[1]
    cols = [theta] + [(theta[:, i] * theta[:, j])[:, None] for i in range(d) for j in range(i, d)]
    return np.concatenate(cols, axis=1)


def fit(fitinfo, emu, x, y, clf_method=None, myusedir=True, **bayeswoodbury_args):
    """Posterior sampling (mlbayeswoodbury.py:7-200)."""
    try:
        from surmise.utilities import sampler        # the reference's dispatcher, used unmodified
    except ImportError:
        from ..utilities import sampler
    if 'shard_theta' in bayeswoodbury_args:
        fitinfo['shard_theta'] = bool(bayeswoodbury_args.pop('shard_theta'))
    _dbw.woodbury_constants(fitinfo, emu, x, y)
    extra = None
    if clf_method is not None:
        def logprobacce(theta):
            prob = clf_method.predict_proba(quadratic_features(theta))[:, 1]
            return np.reshape(np.log(prob), (len(theta), 1))

        def logprobacce_grad(theta):                 # forward differences, step 1e-6 (mlbayeswoodbury.py:123-131)
            base = logprobacce(theta)
            out = np.zeros((base.shape[0], theta.shape[1]))
            for i in range(theta.shape[1]):
                prop = 1 * theta
                prop[:, i] += 10 ** (-6)
                out[:, i] = np.squeeze((logprobacce(prop) - base) * (10 ** 6))
            return out
        extra = (logprobacce, logprobacce_grad)
    # the PCGPwM_B200 emulator always provides gradients, so `myusedir` alone decides (mlbayeswoodbury.py:81-91)
    logpost, draw_func = _dbw.make_logpost(fitinfo, emu, x, y, fd_step=10 ** (-3), extra_logp=extra,
                                           with_grad=bool(myusedir))
    sampler_obj = sampler(logpost_func=logpost, draw_func=draw_func, **bayeswoodbury_args)
    theta = sampler_obj.sampler_info['theta']
    ladj = logpost(theta, return_grad=False)
    mladj = np.max(ladj)
    fitinfo['lpdfapproxnorm'] = np.log(np.mean(np.exp(ladj - mladj))) + mladj
    fitinfo['lpdfapproxnorm_un'] = ladj
    fitinfo['thetarnd'] = theta
    fitinfo['y'] = y
    fitinfo['x'] = x
    fitinfo['emu'] = emu
    return
