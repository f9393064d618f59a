"""metropolis_hastings_B200 -- random-walk Metropolis sampler plugin with a population of chains.

Same algorithm, options and return value as surmise/utilitiesmethods/metropolis_hastings.py
(`sampler(logpost_func, draw_func, numsamp, theta0, stepType, stepParam, burnSamples, verbose)` ->
{'theta', 'acc_rate', 'lpostlist'}, metropolis_hastings.py:7-104).  The reference advances ONE chain and calls the
log-posterior with a single theta per step, which on the B200 path is a latency-bound launch per sample.  Here
`numchain` independent chains advance in lock step, so each step is one batched evaluation of `numchain` proposals
(SURVEY.md section 8(f) row 1):
  * numchain = 1 (default) consumes numpy's global random stream in the reference's order and reproduces the
    reference chain sample for sample;
  * numchain = C > 1 runs C chains of burnSamples + ceil(numsamp / C) steps from C starting points and returns the
    first numsamp retained draws, chain after chain.
The reference's step distributions are kept, including its 'uniform' step, which draws from U(-0.5, 0) * stepParam
(scipy's uniform(loc=-0.5, scale=0.5), metropolis_hastings.py:71-72).
Registered by surmise_b200.register() as surmise.utilitiesmethods.metropolis_hastings_B200.
"""
import numpy as np


def _values(logpost_func, theta):
    """(B,) log-posterior values without gradients (the closure returns (B, 1))."""
    try:
        out = logpost_func(theta, return_grad=False)
    except TypeError:
        out = logpost_func(theta)
    if isinstance(out, tuple):
        out = out[0]
    return np.asarray(out, dtype=np.float64).reshape(theta.shape[0])


def sampler(logpost_func, draw_func, numsamp=2000, theta0=None, stepType='normal', stepParam=None,
            burnSamples=1000, verbose=False, numchain=1, **mh_options):
    """Random-walk Metropolis; see the module docstring.  theta0: (numchain, d) starting points (or None)."""
    if stepType not in ('normal', 'uniform'):
        raise ValueError("stepType must be 'normal' or 'uniform'")
    numchain = int(numchain)
    if numchain < 1:
        raise ValueError('numchain must be positive')
    if stepParam is None:
        stepParam = np.std(draw_func(burnSamples), axis=0)
    stepParam = np.asarray(stepParam, dtype=np.float64)
    if theta0 is None:
        theta0 = draw_func(numchain)
    theta0 = np.atleast_2d(np.asarray(theta0, dtype=np.float64))
    if theta0.shape[0] != numchain:
        if theta0.shape[0] < numchain:
            raise ValueError('theta0 needs one row per chain')
        theta0 = theta0[:numchain]
    d = theta0.shape[1]
    keep = -(-numsamp // numchain)                      # retained steps per chain
    total = burnSamples + keep
    chain = np.empty((total, numchain, d))
    lpost = np.empty((total, numchain))
    proposed = np.empty((total - 1, numchain))
    chain[0] = theta0
    lpost[0] = _values(logpost_func, theta0)
    accepted = np.zeros(numchain)
    for i in range(1, total):
        if verbose and i % 30000 == 0:
            print('At sample {}, acceptance rate is {}.'.format(i, accepted.sum() / (i * numchain)))
        if stepType == 'normal':
            step = np.random.standard_normal(size=(numchain, d))
        else:
            step = np.random.uniform(0.0, 1.0, size=(numchain, d)) * 0.5 - 0.5
        cand = chain[i - 1] + stepParam * step
        lp = _values(logpost_func, cand)
        finite = np.isfinite(lp)
        if numchain == 1:                               # reference order: the uniform is only drawn for a finite value
            take = finite & (np.random.uniform(size=1) < np.exp(np.minimum(0.0, lp - lpost[i - 1]))
                             if finite[0] else finite)
        else:
            u = np.random.uniform(size=numchain)
            with np.errstate(invalid='ignore'):
                take = finite & (u < np.exp(np.minimum(0.0, lp - lpost[i - 1])))
        chain[i] = np.where(take[:, None], cand, chain[i - 1])
        lpost[i] = np.where(take, lp, lpost[i - 1])
        proposed[i - 1] = lp
        if i >= burnSamples:
            accepted += take
    kept = chain[burnSamples:].transpose(1, 0, 2).reshape(-1, d)[:numsamp]
    info = {'theta': kept, 'acc_rate': float(accepted.sum()) / (keep * numchain),
            'lpostlist': proposed[:, 0] if numchain == 1 else proposed}
    if verbose:
        print('Final Acceptance Rate: ', info['acc_rate'])
    return info
