"""LMC_B200 -- Metropolis-adjusted Langevin sampler plugin, batched for the B200 path.

Same algorithm, options and return value as surmise/utilitiesmethods/LMC.py (`sampler(logpost_func, draw_func,
numsamp, theta0)` -> {'theta', 'logpost'}, LMC.py:43-315): importance-resampled starting cloud (LMC.py:115-138),
ten bounded L-BFGS-B polishing runs (LMC.py:140-206), then rounds of `numchain` Langevin chains whose step size
and length adapt to the acceptance rate and the effective sample size (LMC.py:208-309).  The population steps were
already one batched log-posterior call per step in the reference; what it ran one theta at a time was the
polishing stage -- value and gradient in separate B = 1 calls.  Here the ten optimisers advance in lock step, one
batched value+gradient call per optimiser step.  numpy's global random stream is
consumed in the reference's order, so with the same seed and a log-posterior whose rows do not depend on the batch
the draws are the reference's.
Extra options: numchain (100), maxiters (10), numsamppc (200) -- constants in the reference (LMC.py:212-214).
Registered by surmise_b200.register() as surmise.utilitiesmethods.LMC_B200.
"""
import threading

import numpy as np
import scipy.optimize as spo

from .PTLMC_B200 import _Lockstep


def _rho(tau):
    return 2 * (1 + (np.exp(2 * tau) - 1) / (np.exp(2 * tau) + 1))


def _lockstep_minimise(starts, fun, bounds, use_jac, options=None):
    """scipy L-BFGS-B from every row of `starts` at once; fun(X (B, d)) -> (values (B,), grads (B, d) or None)."""
    ev = _Lockstep(fun, len(starts))
    results = [None] * len(starts)

    def work(i):
        try:
            def f(x):
                v, g = ev(x)
                return (v, g) if use_jac else v
            results[i] = spo.minimize(f, starts[i], method='L-BFGS-B', jac=bool(use_jac), bounds=bounds,
                                      options=options or {})
        finally:
            ev.retire()
    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(starts))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return results


def _polish(cloud, value_only, value_and_grad):
    """L-BFGS-B from the first ten points of the cloud in centred/scaled coordinates (LMC.py:140-206).

    Overwrites those rows in place, as the reference does (its `thetaop` is a view of the cloud), and returns them."""
    d = cloud.shape[1]
    cen = np.mean(cloud, 0)
    sca = np.maximum(np.std(cloud, 0), 10 ** (-8) * np.std(cloud))
    z = (cloud - cen) / sca
    bounds = spo.Bounds(np.maximum(-10 * np.ones(d), np.min(z, 0)), np.minimum(10 * np.ones(d), np.max(z, 0)))
    nopt = min(10, cloud.shape[0])
    starts = z[:nopt].copy()

    def neg_value(X):
        return -np.reshape(value_only(cen + sca * X), -1), None
    if value_and_grad is None:
        res = _lockstep_minimise(starts, neg_value, bounds, False)
        for k in range(nopt):
            cloud[k] = cen + sca * res[k].x
        return cloud[:nopt], cen, sca

    def neg_both(X):
        v, g = value_and_grad(cen + sca * X)
        return -np.reshape(v, -1), -sca * g
    res = _lockstep_minimise(starts, neg_both, bounds, True)
    # the reference gives up on gradients once failures are frequent (LMC.py:186-196); which runs need the
    # short gradient-free fallback follows from the success flags alone
    fallback, trying, failures = [], True, 0
    for k in range(nopt):
        if trying and res[k].success:
            continue
        if trying:
            failures += 1
            a, b = failures + 0.25, k - failures + 1
            if a / (a + b) - 3 * np.sqrt(a * b / ((a + b + 1) * (a + b) ** 2)) > 0.25:
                trying = False
        fallback.append(k)
    if fallback:
        short = _lockstep_minimise(starts[fallback], neg_value, bounds, False, {'maxiter': 4, 'maxfun': 100})
        for k, r in zip(fallback, short):
            res[k] = r
    for k in range(nopt):
        cloud[k] = cen + sca * res[k].x
    return cloud[:nopt], cen, sca


def _ess(draws):
    """Per-coordinate effective sample size of (numchain, steps, d) draws from lag-1 autocorrelation (LMC.py:277-296)."""
    numchain, steps, _ = draws.shape
    mu_c = draws.mean(1)
    mu = mu_c.mean(0)
    dev = draws - mu_c[:, None, :]
    auto = np.mean(np.mean(dev[:, :-1] * dev[:, 1:], 1), 0)
    W = np.mean(np.mean(dev[:, :-1] ** 2, 1), 0)
    Bv = steps / (numchain - 1) * np.sum((mu_c - mu) ** 2, 0)
    varplus = W + Bv / steps
    if np.any(varplus < 10 ** (-10)):
        raise ValueError('Sampler failed to move at all.')
    rhohat = 1 - (W - auto) / varplus
    return 1 + numchain * steps * (1 - np.abs(rhohat))


def sampler(logpost_func, draw_func, numsamp=2000, theta0=None, numchain=100, maxiters=10, numsamppc=200,
            **lmc_options):
    """Langevin Monte Carlo; see the module docstring."""
    if theta0 is None:
        theta0 = draw_func(1000)
    theta0 = np.asarray(theta0, dtype=np.float64)
    d = theta0.shape[1]
    tarESS = np.max((150, 10 * d))

    probe = logpost_func(theta0[0:2, :])
    has_grad = isinstance(probe, tuple)
    if has_grad:
        if len(probe) != 2:
            raise ValueError('log density does not return 1 or 2 elements')
        if probe[1].shape[1] != d:
            raise ValueError('derivative appears to be the wrong shape')
        value_and_grad = logpost_func
        try:
            t = logpost_func(theta0[0:2, :], return_grad=False)
            if isinstance(t, tuple):
                raise ValueError('Cannot stop returning a grad')

            def value_only(th):
                return logpost_func(th, return_grad=False)
        except Exception:
            def value_only(th):
                return logpost_func(th)[0]
    else:
        value_and_grad = None
        value_only = logpost_func
    taracc = 0.60 if has_grad else 0.25

    # starting cloud: resample the distinct initial points with weights posterior^(1/4) (LMC.py:115-138)
    cloud = np.unique(theta0, axis=0)
    attempts = 0
    while True:
        lp = value_only(cloud) / 4
        top = np.max(lp)
        lp = lp - (top + np.log(np.sum(np.exp(lp - top))))
        w = np.exp(lp)
        w = w / np.sum(w)
        pick = cloud[np.random.choice(range(0, cloud.shape[0]), size=1000, p=w.reshape(cloud.shape[0])), :]
        if np.any(np.std(pick, 0) < 10 ** (-8) * np.min(np.std(cloud, 0))):
            best = cloud[np.argmax(lp), :]
            cloud = best + (cloud - best) / 2
            attempts += 1
        else:
            cloud = pick
            break
        if attempts > 10:
            raise ValueError('Could not find any points to vary.')

    polished, _cen, sca = _polish(cloud, value_only, value_and_grad)
    pool = np.vstack((cloud, polished))
    Lsave = value_only(pool)
    tau = -1
    rho = _rho(tau)
    cov0 = np.diag(sca)
    numsamppc = int(numsamppc)
    for it in range(maxiters):
        pool = pool[np.random.choice(np.arange(0, Lsave.shape[0]), size=Lsave.shape[0]), :]
        cov0 = 0.1 * cov0 + 0.9 * np.cov(pool.T)
        if cov0.ndim > 1:
            cov0 += 0.1 * np.diag(np.diag(cov0))
            Wc, Vc = np.linalg.eigh(cov0)
            hc = Vc @ np.diag(np.sqrt(Wc)) @ Vc.T
        else:
            hc = np.sqrt(cov0)
        cur = pool[np.random.choice(range(0, pool.shape[0]), size=numchain), :]
        if has_grad:
            fval, dfval = value_and_grad(cur)
        else:
            fval = value_only(cur)
        draws = np.zeros((numchain, numsamppc, d))
        Lsave = np.zeros((numchain, numsamppc))
        moved = 0
        for k in range(numsamppc):
            z = np.random.normal(0, 1, cur.shape)
            prop = cur + np.reshape(np.sqrt(2) * rho * (z @ hc), cur.shape)
            if has_grad:
                prop += rho ** 2 * (dfval @ cov0)
                fvalp, dfvalp = value_and_grad(prop)
                t1 = z / np.sqrt(2)
                t2 = (dfval + dfvalp) @ hc * rho / 2
                qadj = -(2 * np.sum(t1 * t2, 1) + np.sum(t2 ** 2, 1))
            else:
                fvalp = value_only(prop)
                qadj = np.zeros(fvalp.shape)
            logu = np.log(np.random.uniform(size=fval.shape[0]))
            take = np.where(np.squeeze(logu) < np.squeeze(fvalp - fval) + np.squeeze(qadj))[0]
            if take.shape[0] > 0:
                moved = moved + take.shape[0] / numchain
                cur[take, :] = prop[take, :]
                fval[take] = fvalp[take]
                if has_grad:
                    dfval[take, :] = dfvalp[take, :]
            if it < 1.5:                                # step-size adaptation during the first two rounds
                tau = tau + 1 / np.sqrt(1 + 100 / numchain * k) * (take.shape[0] / numchain - taracc)
                rho = _rho(tau)
            draws[:, k, :] = cur
            Lsave[:, k] = fval.reshape(len(fval))
        ESS = _ess(draws)
        pool = np.reshape(draws, (-1, d))
        accr = moved / numsamppc
        if it > 1.5 and taracc / 2 < accr < 1.5 * taracc and np.mean(ESS) > tarESS:
            break
        elif accr < taracc * 4 / 5 or accr > taracc * 5 / 4:
            tau = tau + 1 / (0.2 + 0.2 * it) * (accr - taracc)
            rho = _rho(tau)
        if taracc * 0.6 < accr < taracc * 1.5:
            numsamppc = np.ceil(numsamppc * np.min((1.5 * tarESS / np.mean(ESS), 4))).astype('int')
    theta = pool[np.random.choice(range(0, pool.shape[0]), size=numsamp), :]
    return {'theta': theta, 'logpost': Lsave}
