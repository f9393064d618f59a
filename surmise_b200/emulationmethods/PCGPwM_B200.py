"""PCGPwM_B200 -- surmise emulation method: PCGP with missingness on one B200 (fp64, sm_100a).

Drop-in sibling of surmise/emulationmethods/PCGPwM.py behind the same plugin contract
(`fit(fitinfo, x, theta, f, **args)`, `predict(predinfo, fitinfo, x, theta, **args)`,
surmise/emulation.py:150-153): `emulator(x, theta, f, method='PCGPwM_B200')` after
`surmise_b200.register()`.

What runs where
  host (numpy)  : standardisation / imputation / PCA (_pca.py), the share-or-own bookkeeping of
                  __fitGPs/__fitGP1d and scipy's L-BFGS-B driver -- the reference's own control flow;
  device (CUDA) : every covariance build, every likelihood+gradient evaluation, the full-n
                  factorisations, predict and back-projection (surmise_b200/csrc).

Differences from the reference that a user can observe (all documented in DESIGN.md):
  * one triangular factor Linv (Linv' Linv = R^-1) per distinct (hyper-parameters, gvar) pair is kept
    on the device instead of R, Rinv and Vh per PC (PCGPwM.py:817-846) -- 1/3 the memory or less;
  * the likelihood is evaluated through a Cholesky factorisation instead of eigh+abs (PCGPwM.py:858);
    where the reference's R is not positive definite (rounding noise in unscaled_pcstdvar) the
    evaluation reports breakdown and the optimiser treats the point as infeasible;
  * the full-n factorisation is done once after the three sweeps, batched over PCs, instead of after
    every __fitGP1d call (the intermediate ones are never read by the reference either).
"""
import itertools
import warnings
from pprint import pformat

import numpy as np
import scipy.optimize as spo

from .. import _pca

SIG2_OF_CONST = 0.01        # PCGPwM.py:705 / 719
_GENERATION = itertools.count(1)   # fit-generation token of a device state (caches downstream key on it)
_INFEASIBLE = 1e30          # value handed to L-BFGS-B when the factorisation breaks down


def _ops():
    from .. import ops
    return ops


# ------------------------------------------------------------------------------------------------
# fit
# ------------------------------------------------------------------------------------------------
def fit(fitinfo, x, theta, f, epsilonPC=0.001, epsilonImpute=10e-6, lognugmean=-10, lognugLB=-20,
        varconstant=None, dampalpha=0.3, eta=10, standardpcinfo=None, verbose=0, nhyptrain=None,
        fit_mode='reference', pca='device', distributed=True, merge_leaders=False, factors='owner', **kwargs):
    """Same arguments as PCGPwM.fit (PCGPwM.py:12-137) plus
    nhyptrain : None (reference rule min(20 d, n), PCGPwM.py:744), an int, or 'all'.
    fit_mode  : 'reference' -- the reference's sequential Gauss-Seidel sweeps (PCGPwM.py:691-721), one GPU;
                'parallel'  -- Jacobi sweeps: every PC sees the previous sweep's leaders, PCs are sharded over
                               the ranks of the initialised torch.distributed group (results do not depend on
                               the number of ranks); see surmise_b200/_dist.py.
    distributed : with fit_mode='parallel', shard the PCs over the torch.distributed group if one is initialised
                (every rank must then call fit with the same data); False keeps the fit on this process only.
    merge_leaders : with fit_mode='parallel', apply the reference's sharing criterion (PCGPwM.py:808-809) once more
                after the sweeps so that components adopt an earlier leader's hyper-parameters where that costs
                little likelihood: fewer full-data factors (C2: 20 -> 9), faster prediction / calibration, slightly
                larger hold-out error (0.065 -> 0.075 sd at C2).  Default False.
    factors   : with fit_mode='parallel' on more than one rank: 'owner' (default) -- every full-data factor Linv stays
                on the rank that computed it; predict then runs PC-sharded (each rank produces
                predvecs / predvars of the components whose factor it owns, one all-gather of B x k x (2+2d) doubles)
                and must be called by every rank with the same theta.  'replicated' -- owners broadcast their
                factors so that every rank holds all of them (needed for theta-sharded calibration, shard_theta).
    pca       : 'device' (default) -- standardisation / imputation / PCA on the GPU (_pca_device.py: Gram-matrix
                eigendecomposition and batched ridge solves); 'host' -- the numpy implementation (_pca.py), which
                takes the SVD exactly as the reference does.
    """
    ops = _ops()
    ops._torch()                                   # fail loudly without CUDA before any work
    clock = _PhaseClock(ops._torch())
    f = np.asarray(f, dtype=np.float64).T          # (n, m)
    theta = np.asarray(theta, dtype=np.float64)
    fitinfo['epsilonImpute'] = epsilonImpute
    fitinfo['epsilonPC'] = epsilonPC
    fitinfo['dampalpha'] = dampalpha
    fitinfo['eta'] = eta
    fitinfo['theta'] = theta
    fitinfo['f'] = f
    fitinfo['x'] = x

    if pca not in ('device', 'host'):
        raise ValueError("pca must be 'device' or 'host'")
    if factors not in ('owner', 'replicated'):
        raise ValueError("factors must be 'owner' or 'replicated'")
    if standardpcinfo is None and pca == 'device':
        from .. import _pca_device
        spc, mof, mofrows = _pca_device.standardize(f, epsilonPC, epsilonImpute)
        fitinfo.update(_pca_device.principal_components(spc, mof, mofrows, epsilonPC, epsilonImpute))
    else:
        if standardpcinfo is None:
            spc, mof, mofrows = _pca.standardize(f, epsilonPC, epsilonImpute)
        else:
            spc = _pca.complete_pcinfo(standardpcinfo, f)
            mof = None if np.all(np.isfinite(f)) else ~np.isfinite(f)
            mofrows = None if mof is None else np.where(mof.any(axis=1))[0]
        if 'U' in spc:
            U, S = spc['U'], spc['S']
        else:
            U, S, _vt = np.linalg.svd(spc['fs'].T, full_matrices=False)
        fitinfo.update(_pca.principal_components(spc['fs'], U, S, mof, mofrows, epsilonPC, epsilonImpute))
    fitinfo['mof'], fitinfo['mofrows'] = mof, mofrows
    fitinfo['standardpcinfo'] = spc
    clock.lap('pca')
    numpcs = fitinfo['pc'].shape[1]
    if verbose > 0:
        print(fitinfo.get('method', 'PCGPwM_B200'), 'considering ', numpcs, 'PCs')

    logvarc = np.log(varconstant) if varconstant is not None else None
    fitter = _GPFitter(fitinfo, theta, lognugmean, lognugLB, logvarc, nhyptrain)
    fitter.distributed = bool(distributed)
    fitter.merge_leaders = bool(merge_leaders)
    fitter.factors = factors
    fitter.clock = clock
    clock.lap('setup')
    if fit_mode not in ('reference', 'parallel'):
        raise ValueError("fit_mode must be 'reference' or 'parallel'")
    emulist = (fitter.run if fit_mode == 'reference' else fitter.run_parallel)(previous=fitinfo.get('emulist'))
    fitinfo['varc_status'] = 'fixed' if varconstant is not None else 'optimized'
    fitinfo['logvarc'] = np.array([e['hypvarconst'] for e in emulist])
    fitinfo['pcstdvar'] = np.exp(fitinfo['logvarc']) * fitinfo['unscaled_pcstdvar']
    fitinfo['emulist'] = emulist
    fitinfo['_b200'] = fitter.device_state(emulist, spc, fitinfo['pcti'])
    fitinfo['param_desc'] = _param_desc(fitinfo)
    clock.lap('state')
    fitinfo['_b200']['timing'] = clock.report()
    return


class _PhaseClock:
    """Wall-clock seconds per phase of a fit (device synchronised at every lap), kept in fitinfo['_b200']['timing']:
    'pca' (standardise / impute / PCA), 'setup', 'search' (the sweeps), 'factor' (full-data factorisations),
    'exchange' (collectives after the factorisations), 'weights', 'state'."""

    def __init__(self, torch):
        import time
        self._now, self._torch = time.perf_counter, torch
        self.t0 = self.last = self._now()
        self.phases = {}

    def lap(self, name):
        self._torch.cuda.synchronize()
        t = self._now()
        self.phases[name] = self.phases.get(name, 0.0) + (t - self.last)
        self.last = t

    def report(self):
        out = dict(self.phases)
        out['total'] = self._now() - self.t0
        return out


class _LbfgsbCore:
    """One bounded L-BFGS-B run driven step by step (reverse communication) through scipy's own core routine.

    Same routine, parameters and stopping rules as scipy.optimize.minimize(method='L-BFGS-B', jac=True,
    options={'gtol': gtol}) (scipy/optimize/_lbfgsb_py.py:_minimize_lbfgsb), so the iterates are the same; what is
    gone is one Python thread per optimiser: many cores can be advanced by one scheduler that batches their function
    requests into a single device launch.  `available()` is False when this scipy does not expose the routine with
    the expected signature -- callers then fall back to threads around the public API."""
    M, FTOL, MAXLS, MAXITER, MAXFUN = 10, 2.2204460492503131e-09, 20, 15000, 15000
    _ok = None

    @classmethod
    def available(cls):
        if cls._ok is None:
            try:
                core = cls(np.array([0.3, -0.2]), np.array([-1.0, -1.0]), np.array([1.0, 1.0]), 1e-8)
                n = 0
                while core.ask() is not None and n < 200:
                    core.tell(float(np.sum((core.x - 0.1) ** 2)), 2 * (core.x - 0.1))
                    n += 1
                cls._ok = bool(np.allclose(core.x, 0.1, atol=1e-6))
            except Exception:
                cls._ok = False
        return cls._ok

    def __init__(self, x0, lo, hi, gtol):
        from scipy.optimize import _lbfgsb
        self._setulb = _lbfgsb.setulb
        n = x0.shape[0]
        it = np.int32
        try:
            from scipy.optimize._lbfgsb_py import HAS_ILP64
            it = np.int64 if HAS_ILP64 else np.int32
        except ImportError:
            pass
        self.x = np.array(np.clip(x0, lo, hi), dtype=np.float64)
        self.lo, self.hi = np.array(lo, dtype=np.float64), np.array(hi, dtype=np.float64)
        self.nbd = np.full(n, 2, dtype=it)                      # both bounds finite
        self.f, self.g = np.array(0.0, dtype=np.float64), np.zeros(n, dtype=np.float64)
        m = self.M
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self.iwa = np.zeros(3 * n, dtype=it)
        self.task, self.ln_task = np.zeros(2, dtype=it), np.zeros(2, dtype=it)
        self.lsave, self.isave = np.zeros(4, dtype=it), np.zeros(44, dtype=it)
        self.dsave = np.zeros(29, dtype=np.float64)
        self.factr, self.pgtol = self.FTOL / np.finfo(float).eps, gtol
        self.nit = self.nfev = 0

    def ask(self):
        """Advance to the next point at which value and gradient are needed; None when the run has finished."""
        while True:
            self.g = self.g.astype(np.float64)
            self._setulb(self.M, self.x, self.lo, self.hi, self.nbd, self.f, self.g, self.factr, self.pgtol, self.wa,
                         self.iwa, self.task, self.lsave, self.isave, self.dsave, self.MAXLS, self.ln_task)
            if self.task[0] == 3:
                return self.x
            if self.task[0] == 1:
                self.nit += 1
                if self.nit >= self.MAXITER:
                    self.task[0], self.task[1] = 5, 504
                elif self.nfev > self.MAXFUN:
                    self.task[0], self.task[1] = 5, 502
            else:
                return None

    def tell(self, fval, grad):
        self.nfev += 1
        self.f, self.g = fval, np.asarray(grad, dtype=np.float64)


def minimize_many(fun_batch, x0s, lo, hi, gtol):
    """Bounded L-BFGS-B from every row of x0s, advanced together by one scheduler (requires _LbfgsbCore.available()).

    fun_batch(idx (r,), X (r, p)) -> (values (r,), gradients (r, p)) for the problems idx that are waiting.
    Returns (X (K, p), fun (K,), nfev (K,)) -- per problem what scipy.optimize.minimize(..., jac=True,
    options={'gtol': gtol}) returns."""
    cores = [_LbfgsbCore(np.asarray(x0, dtype=np.float64), lo, hi, gtol) for x0 in x0s]
    live = list(range(len(cores)))
    while live:
        asks = [(i, cores[i].ask()) for i in live]
        live = [i for i, xs in asks if xs is not None]
        if not live:
            break
        X = np.array([xs for _, xs in asks if xs is not None])
        vals, grads = fun_batch(np.array(live), X)
        for r, i in enumerate(live):
            cores[i].tell(float(vals[r]), grads[r])
    return (np.array([c.x for c in cores]), np.array([float(c.f) for c in cores]),
            np.array([c.nfev for c in cores]))


def hyper_prior(theta, lognugmean, lognugLB, logvarc):
    """Regulariser mean / box / std for [gamma_0..gamma_d, h_v, h_nu] (PCGPwM.py:728-742)."""
    d = theta.shape[1]
    base = 0.5 * np.log(d) + np.log(np.std(theta, 0))
    vmid, vlo, vhi = (4, -8, 8) if logvarc is None else (logvarc, logvarc - 0.5, logvarc + 0.5)
    mean = np.concatenate([base, [0, vmid, lognugmean]])
    lo = np.concatenate([base - 4, [-12, vlo, lognugLB]])
    hi = np.concatenate([base + 4, [2, vhi, -8]])
    std = (hi - lo) / 8
    std[-3] = 2
    std[-1] = 4
    return mean, lo, hi, std


class _GPFitter:
    """Hyper-parameter search for all PCs with device-side likelihoods (PCGPwM.py:678-847)."""

    def __init__(self, fitinfo, theta, lognugmean, lognugLB, logvarc, nhyptrain, variant=None):
        """variant: None for PCGPwM, or a dict for a sibling method that shares the search (PCSK_B200):
        prior (mean, lo, hi, std), nh_rule (d, n) -> int, share_bar (p) -> float, sweeps, gtol, skip_test,
        use_gvar, s0_noleader (sig2ofconst of a search that has no leaders yet)."""
        variant = variant or {}
        self.ops = _ops()
        torch = self.ops._torch()
        self.torch = torch
        self.theta = theta
        self.n, self.d = theta.shape
        self.pc = fitinfo['pc']
        # damped per-row variance of the imputed PCs: min(eta, v / (1-v)^alpha)   (PCGPwM.py:755)
        v = fitinfo['unscaled_pcstdvar']
        if variant.get('use_gvar', True):
            self.gvar = np.minimum(fitinfo['eta'], v / ((1 - v) ** fitinfo['dampalpha']))
        else:
            self.gvar = np.zeros_like(self.pc)         # sized by the scores: PCSK's simsd may be stale after update/remove
        if self.gvar.shape != self.pc.shape or self.pc.shape[0] != self.n:
            raise ValueError('PC scores (%s), their variances (%s) and theta (%d rows) do not match.'
                             % (self.pc.shape, self.gvar.shape, self.n))
        self.mean, self.lo, self.hi, self.std = variant.get('prior') or hyper_prior(theta, lognugmean, lognugLB, logvarc)
        self.p = self.mean.shape[0]
        self.share_bar = 1.25 * (self.p + 5 * np.sqrt(self.p))       # PCGPwM.py:776-777, 808-809
        if 'share_bar' in variant:
            self.share_bar = variant['share_bar'](self.p)
        self.sweeps = variant.get('sweeps', 3)
        self.gtol = variant.get('gtol', 0.1)
        self.skip_test = variant.get('skip_test', True)
        self.s0_noleader = variant.get('s0_noleader')
        if nhyptrain is None:
            self.nh = variant['nh_rule'](self.d, self.n) if 'nh_rule' in variant else min(20 * self.d, self.n)
        elif nhyptrain == 'all':
            self.nh = self.n
        else:
            self.nh = int(min(max(int(nhyptrain), 2), self.n))
        dev = torch.device('cuda', torch.cuda.current_device())
        self.theta_dev = torch.from_numpy(np.ascontiguousarray(theta)).to(dev)
        self.pc_dev = torch.from_numpy(np.ascontiguousarray(self.pc.T)).to(dev)        # (k, n)
        self.gvar_dev = torch.from_numpy(np.ascontiguousarray(self.gvar.T)).to(dev)    # (k, n)
        self.mean_dev = torch.from_numpy(self.mean).to(dev)
        self.std_dev = torch.from_numpy(self.std).to(dev)
        self.nevals = 0
        self._lockstep = None

    # -- one PC ---------------------------------------------------------------------------------
    def _draw_rows(self, j=None):
        """Random rows for the hyper-parameter search; consumes np.random like PCGPwM.py:744-750."""
        if self.n > self.nh:
            return np.random.choice(self.n, self.nh, replace=False)
        return np.arange(self.n)

    def _subsample(self, j, rows=None):
        if rows is None:
            rows = self._draw_rows()
        if self.n > self.nh:
            idx = self.torch.from_numpy(np.asarray(rows, dtype=np.int64)).to(self.theta_dev.device)
            data = self.ops.GPData(self.theta_dev[idx], self.pc_dev[j][idx], self.gvar_dev[j][idx])
        else:
            data = self.ops.GPData(self.theta_dev, self.pc_dev[j], self.gvar_dev[j])
        return rows, data

    def _nll(self, data, hyps, want_grad, s0=SIG2_OF_CONST):
        self.nevals += len(hyps)
        if self._lockstep is not None:
            val, grad, info = self._lockstep.evaluate(data, np.asarray(hyps), want_grad, s0)
        else:
            val, grad, info = self.ops.gp_nll_grad(data, np.asarray(hyps), self.mean_dev, self.std_dev,
                                                   s0, want_grad)
        bad = (info != 0) | ~np.isfinite(val)
        val = np.where(bad, _INFEASIBLE, val)
        if grad is not None:
            grad = np.where(bad[:, None], 0.0, grad)
        return val, grad

    def fit_one(self, j, cand, cand_ids, rows=None):
        """__fitGP1d without its full-n stage.  Returns (hyp, hypind or -1, subsample rows).

        Runs the fit_one_steps generator, answering each of its likelihood requests with a device call; when scipy's
        reverse-communication L-BFGS-B entry point is not usable (or inside the threaded rendezvous) the search goes
        through scipy.optimize.minimize instead -- the same iterates either way."""
        if self._lockstep is not None or not _LbfgsbCore.available():
            self.driver_used = 'scipy.optimize.minimize'
            return self._fit_one_scipy(j, cand, cand_ids, rows)
        self.driver_used = 'scipy _lbfgsb.setulb, reverse communication'
        steps = self.fit_one_steps(j, cand, cand_ids, rows)
        try:
            data, hyps, want_grad, s0 = next(steps)
            while True:
                data, hyps, want_grad, s0 = steps.send(self._nll(data, hyps, want_grad, s0))
        except StopIteration as stop:
            return stop.value

    def _fit_one_scipy(self, j, cand, cand_ids, rows=None):
        """fit_one with the L-BFGS-B run handed to scipy.optimize.minimize (PCGPwM.py:783-805 verbatim in structure)."""
        rows, data = self._subsample(j, rows)
        hyp = 1 * self.mean
        lead = -1
        s0 = self.s0_noleader if (cand is None and self.s0_noleader is not None) else SIG2_OF_CONST
        # candidate scan (PCGPwM.py:761-769): prior mean first, then every leader's hyper-parameters
        trial = [self.mean] if cand is None else [self.mean] + [c for c in cand]
        vals, _ = self._nll(data, trial, want_grad=False, s0=s0)
        L0 = vals[0]
        for c in range(1, len(trial)):
            if vals[c] < L0:
                hyp, L0, lead = trial[c], 1 * vals[c], cand_ids[c - 1]
        skip = False
        if self.skip_test and lead > -0.5 and cand.ndim > 1:        # PCGPwM.py:771-781
            _, dL = self._nll(data, [hyp], want_grad=True, s0=s0)
            nc = cand.shape[0]
            scal = np.std(cand, 0) * nc / (1 + nc) + self.std / (1 + nc)
            skip = np.sum((dL[0] * scal) ** 2) < self.share_bar
        if not skip:                                                # PCGPwM.py:783-805
            def scaled(v):
                val, grad = self._nll(data, [self.mean + v * self.std], want_grad=True, s0=s0)
                return float(val[0]), grad[0] * self.std
            res = spo.minimize(scaled, (hyp - self.mean) / self.std, method='L-BFGS-B', jac=True,
                               options={'gtol': self.gtol},
                               bounds=spo.Bounds((self.lo - self.mean) / self.std, (self.hi - self.mean) / self.std))
            hypn = self.mean + res.x * self.std
            likdiff = L0 - res.fun                                  # res.fun == nll(hypn): same kernel, same input
        else:
            likdiff = 0
        if lead > -0.5 and (2 * likdiff) < self.share_bar:           # PCGPwM.py:808-833
            return hyp, lead, rows
        return hypn, -1, rows

    def fit_one_steps(self, j, cand, cand_ids, rows=None):
        """fit_one as a generator: yields (data, hyps, want_grad, s0) whenever it needs likelihood values and receives
        (val, grad); the L-BFGS-B run goes through _LbfgsbCore.  Returns (hyp, hypind or -1, rows)."""
        rows, data = self._subsample(j, rows)
        hyp = 1 * self.mean
        lead = -1
        s0 = self.s0_noleader if (cand is None and self.s0_noleader is not None) else SIG2_OF_CONST
        trial = [self.mean] if cand is None else [self.mean] + [c for c in cand]
        vals, _ = yield (data, np.asarray(trial), False, s0)
        L0 = vals[0]
        for c in range(1, len(trial)):
            if vals[c] < L0:
                hyp, L0, lead = trial[c], 1 * vals[c], cand_ids[c - 1]
        skip = False
        if self.skip_test and lead > -0.5 and cand.ndim > 1:
            _, dL = yield (data, np.asarray([hyp]), True, s0)
            nc = cand.shape[0]
            scal = np.std(cand, 0) * nc / (1 + nc) + self.std / (1 + nc)
            skip = np.sum((dL[0] * scal) ** 2) < self.share_bar
        if not skip:
            core = _LbfgsbCore((hyp - self.mean) / self.std, (self.lo - self.mean) / self.std,
                               (self.hi - self.mean) / self.std, self.gtol)
            while True:
                xs = core.ask()
                if xs is None:
                    break
                val, grad = yield (data, np.asarray([self.mean + xs * self.std]), True, s0)
                core.tell(float(val[0]), grad[0] * self.std)
            hypn = self.mean + core.x * self.std
            likdiff = L0 - core.f
        else:
            likdiff = 0
        if lead > -0.5 and (2 * likdiff) < self.share_bar:
            return hyp, lead, rows
        return hypn, -1, rows

    def _run_block_coroutines(self, jobs):
        """Advance the fit_one_steps generators of several PCs together: every round gathers the likelihood requests
        of all live generators into ONE batched launch per (gradient flag, sig2ofconst).  One thread, no locks."""
        ops, torch = self.ops, self.torch
        gens = [job() for job in jobs]
        results, pending = [None] * len(gens), {}

        def advance(t, value=None):
            try:
                pending[t] = gens[t].send(value) if value is not None else next(gens[t])
            except StopIteration as stop:
                pending.pop(t, None)
                results[t] = stop.value
        for t in range(len(gens)):
            advance(t)
        while pending:
            order = sorted(pending)
            outs = {}
            for key in sorted({(pending[t][2], pending[t][3]) for t in order}):
                group = [t for t in order if (pending[t][2], pending[t][3]) == key]
                counts = [pending[t][1].shape[0] for t in group]
                datas = [pending[t][0] for t in group]
                theta = torch.cat([dd.theta.expand(c, -1, -1) if dd.theta.dim() == 2 else dd.theta
                                   for dd, c in zip(datas, counts)], 0)
                g = torch.cat([dd.g.expand(c, -1) for dd, c in zip(datas, counts)], 0)
                gv = torch.cat([dd.gvar.expand(c, -1) for dd, c in zip(datas, counts)], 0)
                hyps = np.concatenate([pending[t][1] for t in group], 0)
                self.nevals += hyps.shape[0]
                val, grad, info = ops.gp_nll_grad(ops.GPData(theta, g, gv), hyps, self.mean_dev, self.std_dev,
                                                  key[1], key[0])
                bad = (info != 0) | ~np.isfinite(val)
                val = np.where(bad, _INFEASIBLE, val)
                if grad is not None:
                    grad = np.where(bad[:, None], 0.0, grad)
                lo = 0
                for t, c in zip(group, counts):
                    outs[t] = (val[lo:lo + c], None if grad is None else grad[lo:lo + c])
                    lo += c
            for t in order:
                advance(t, outs[t])
        return results

    # -- all PCs --------------------------------------------------------------------------------
    def run(self, previous=None):
        """Three Gauss-Seidel sweeps (PCGPwM.py:678-722) followed by ONE batched full-n factorisation."""
        k = self.pc.shape[1]
        hypinds = -1 * np.ones(k)
        starts = None
        if previous is not None:                                     # warm start on refit (680-685)
            starts = np.zeros((k, previous[0]['hyp'].shape[0]))
            for j in range(min(k, len(previous))):
                starts[j] = previous[j]['hyp']
                hypinds[j] = previous[j]['hypind']
            if starts.shape[1] != self.p:
                starts, hypinds = None, -1 * np.ones(k)
        own = np.arange(k)
        hyps = [None] * k
        rows_used = [None] * k
        for _sweep in range(self.sweeps):
            for j in range(k):
                if np.sum(hypinds == own) > 0.5:
                    leaders = np.where(hypinds == own)[0]
                    hyp, hi, rows = self.fit_one(j, starts[leaders], leaders)
                else:
                    hyp, hi, rows = self.fit_one(j, None, None)
                    starts = np.zeros((k, self.p))
                hi = min(j, hi)
                starts[j] = hyp
                hypinds[j] = j if hi < -0.5 else hi
                hyps[j], rows_used[j] = hyp, rows
        if getattr(self, 'clock', None) is not None:
            self.clock.lap('search')
        return self._finalise(np.array(hyps), hypinds.astype(int), rows_used)

    def run_parallel(self, previous=None):
        """Jacobi sweeps with the PCs sharded over torch.distributed ranks (see _dist.jacobi_sweeps)."""
        from .. import _dist
        comm = _dist.Comm(enabled=getattr(self, 'distributed', True))
        k = self.pc.shape[1]
        start_hyps = start_inds = None
        if previous is not None and previous[0]['hyp'].shape[0] == self.p and len(previous) >= k:
            start_hyps = np.array([previous[j]['hyp'] for j in range(k)])
            start_inds = np.array([previous[j]['hypind'] for j in range(k)])

        def one(j, cand, cand_ids, rows):
            hyp, hi, _ = self.fit_one(j, cand, cand_ids, rows)
            return hyp, hi

        def one_steps(j, cand, cand_ids, rows):
            hyp, hi, _ = yield from self.fit_one_steps(j, cand, cand_ids, rows)
            return hyp, hi
        mode = self.lockstep
        if mode == 'coroutine' and not _LbfgsbCore.available():
            mode = 'threads'
        if mode == 'coroutine':
            self.driver_used = 'scipy _lbfgsb.setulb, reverse communication (lock-step generators)'
            fit_fn, run_block = one_steps, self._run_block_coroutines
        else:
            fit_fn, run_block = one, (self._run_block_lockstep if mode else None)
        hyps, hypinds, rows_used = _dist.jacobi_sweeps(fit_fn, k, self.p, comm, self._draw_rows,
                                                       start_hyps=start_hyps, start_inds=start_inds,
                                                       run_block=run_block)
        if self.merge_leaders:
            hyps, hypinds = self._merge_leaders(hyps, hypinds, rows_used)
        if getattr(self, 'clock', None) is not None:
            self.clock.lap('search')
        self._comm = comm                                     # the distinct factors are sharded over the ranks as well
        try:
            return self._finalise(hyps, hypinds, rows_used)
        finally:
            self._comm = None

    # Jacobi sweeps share hyper-parameters less often than the reference's Gauss-Seidel order (in the first sweep nobody
    # has a leader to adopt, afterwards everybody's own optimum is its best candidate), which costs one full-data
    # factor per component at fit time and in every prediction.  The merge pass applies the reference's own sharing
    # criterion (PCGPwM.py:808-809) once more after the sweeps, in component order.  Off by default: at C2 it takes the
    # distinct factors from 20 to 9 (the reference's order ends with 11) and the theta-log-likelihood cost with them, but
    # the hold-out RMSE from 0.065 to 0.075 sd; fit option merge_leaders=True turns it on.
    merge_leaders = False

    def _merge_leaders(self, hyps, hypinds, rows_used):
        """For each component in order: adopt the earlier leader whose hyper-parameters cost it the least, if
        2 (NLL_j(leader) - NLL_j(own)) < share_bar.  All k x k likelihoods come from k batched launches on the
        components' own sub-samples; every rank computes the same table, so the outcome is identical everywhere."""
        hyps = np.array(hyps, dtype=np.float64)
        k = hyps.shape[0]
        table = np.empty((k, k))
        for j in range(k):
            _rows, data = self._subsample(j, rows_used[j])
            table[j], _ = self._nll(data, hyps, want_grad=False)
        leaders = []
        out_inds = np.arange(k)
        for j in range(k):
            if leaders:
                cost = table[j, leaders]
                best = int(np.argmin(cost))
                if 2 * (cost[best] - table[j, j]) < self.share_bar:
                    out_inds[j] = leaders[best]
                    hyps[j] = hyps[leaders[best]]
                    continue
            leaders.append(j)
        return hyps, out_inds

    # how this rank's PCs are searched in parallel mode: 'coroutine' -- generators advanced by one scheduler that
    # batches their likelihood requests (default); 'threads' (or True) -- one scipy.optimize.minimize per Python thread
    # with a rendezvous (fallback when scipy's reverse-communication core is not usable); False -- one after the other
    lockstep = 'coroutine'

    def _run_block_lockstep(self, jobs):
        """Run fit_one for several PCs concurrently (one Python thread each); their likelihood requests are
        gathered and evaluated by ONE batched kernel launch whenever every live thread is waiting for a value.
        Per-problem arithmetic is unchanged (one CTA per problem), so results equal the sequential run."""
        import threading
        dev = self.theta_dev.device
        if len(jobs) <= 1:
            return [fn() for fn in jobs]
        self._lockstep = _LockstepEvaluator(self, len(jobs))
        results, errors = [None] * len(jobs), []

        def work(t, fn):
            try:
                self.torch.cuda.set_device(dev)
                results[t] = fn()
            except BaseException as exc:            # surface worker failures in the caller
                errors.append(exc)
            finally:
                self._lockstep.retire()
        threads = [threading.Thread(target=work, args=(t, fn)) for t, fn in enumerate(jobs)]
        try:
            for th in threads:
                th.start()
            for th in threads:
                th.join()
        finally:
            self._lockstep = None
        if errors:
            raise errors[0]
        return results

    def _finalise(self, hyps, hypinds, rows_used):
        """Full-data factorisations (PCGPwM.py:810-846), one per distinct (hyp, gvar), in one batch."""
        ops, torch = self.ops, self.torch
        k = hyps.shape[0]
        keys, fidx, members = {}, np.zeros(k, dtype=np.int32), []
        for j in range(k):
            key = (hyps[j].tobytes(), self.gvar[:, j].tobytes())
            if key not in keys:
                keys[key] = len(members)
                members.append(j)
            fidx[j] = keys[key]
        members = np.array(members)
        comm = getattr(self, '_comm', None)
        clock = getattr(self, 'clock', None)
        lap = clock.lap if clock is not None else (lambda name: None)
        dev = self.theta_dev.device
        self._shard = None
        if comm is not None and comm.world > 1 and len(members) > 1:
            # each rank factors a contiguous block of the distinct problems
            from .. import _dist
            sizes, starts = _dist.block_partition(len(members), comm.world)
            lo, hi = int(starts[comm.rank]), int(starts[comm.rank + 1])
            ldl = (self.n + 7) // 8 * 8
            local, flags = (torch.empty((0, self.n, ldl), dtype=torch.float64, device=dev), np.zeros(0, dtype=bool))
            if hi > lo:
                local, flags = self._factor_batch(members[lo:hi], hyps)
            lap('factor')
            clamped_f = comm.allgather_rows(flags.astype(np.float64).reshape(-1, 1), sizes).reshape(-1) > 0.5
            if getattr(self, 'factors', 'owner') == 'owner':
                # factors stay with their owner: the owner also produces pw / sig2 of the components that
                # use its factors; only those (k, n+1) doubles travel
                plan = _dist.owner_plan(fidx, starts, comm.rank)
                mine = plan['pcs_local']
                pw_l = torch.empty((len(mine), self.n), dtype=torch.float64, device=dev)
                s2_l = torch.empty(len(mine), dtype=torch.float64, device=dev)
                if len(mine):
                    idx = torch.from_numpy(mine).to(dev)
                    fl = torch.from_numpy((fidx[mine] - lo).astype(np.int32)).to(dev)
                    pw_l, s2_l = ops.gp_weights(local, fl, self.pc_dev[idx].contiguous(), SIG2_OF_CONST)
                lap('weights')
                table = comm.allgather_tensor(torch.cat([pw_l, s2_l[:, None]], 1), plan['counts'])
                table = table[torch.from_numpy(plan['inverse']).to(dev)]                 # gathered order -> PC order
                pw, sig2 = table[:, :self.n].contiguous(), table[:, self.n].contiguous()
                Linv = local
                self._shard = dict(plan, world=comm.world, rank=comm.rank, lo=lo, hi=hi)
                lap('exchange')
            else:
                Linv = torch.empty((len(members), self.n, ldl), dtype=torch.float64, device=dev)
                _dist.gather_blocks(comm, local if hi > lo else None, sizes, Linv)
                lap('exchange')
        else:
            Linv, clamped_f = self._factor_batch(members, hyps)
            lap('factor')
        if self._shard is None:
            # pw and sig2 depend on the scores g_j, which differ between PCs sharing a factor
            fidx_dev = torch.from_numpy(fidx).to(dev)
            pw, sig2 = ops.gp_weights(Linv, fidx_dev, self.pc_dev, SIG2_OF_CONST)
            lap('weights')
        nug = np.exp(hyps[:, -1]) / (1 + np.exp(hyps[:, -1]))        # PCGPwM.py:813
        pw_h, sig2_h = pw.cpu().numpy(), sig2.cpu().numpy()
        emulist = []
        for j in range(k):
            rows = rows_used[j]
            emulist.append({'hyp': hyps[j], 'hypcov': hyps[j][:-2], 'hypvarconst': hyps[j][-2], 'hypind': int(hypinds[j]),
                            'nug': nug[j], 'sig2': sig2_h[j], 'pw': pw_h[j], 'sig2ofconst': SIG2_OF_CONST,
                            'theta': self.theta[rows], 'g': self.pc[rows, j], 'gvar': self.gvar[rows, j],
                            'hypregmean': self.mean, 'hypregLB': self.lo, 'hypregUB': self.hi, 'hypregstd': self.std,
                            'gvar_clamped': bool(clamped_f[fidx[j]])})
        self._Linv, self._fidx, self._members, self._pw, self._sig2 = Linv, fidx, members, pw, sig2
        return emulist

    def _factor_batch(self, members, hyps):
        """Linv for each distinct problem; returns (Linv (nf,n,ldl), clamped flags (nf,))."""
        ops, torch = self.ops, self.torch
        idx = torch.from_numpy(members).to(self.theta_dev.device)
        gv = self.gvar_dev[idx]
        data = ops.GPData(self.theta_dev, self.pc_dev[idx], gv)
        Linv, _pw, _s2, _ld, info = ops.gp_factorize(data, hyps[members], SIG2_OF_CONST)
        clamped = np.zeros(len(members), dtype=bool)
        bad = np.where(info != 0)[0]
        if bad.size:
            # R not positive definite: inside the hyper-parameter box that only happens through rounding
            # noise (negative entries) in unscaled_pcstdvar; clamp those at their mathematical lower
            # bound 0 and refactor the affected problems.
            warnings.warn('PCGPwM_B200: covariance of %d component(s) not positive definite at the fitted '
                          'hyper-parameters; negative imputation variances clamped to 0.' % bad.size)
            bad_dev = torch.from_numpy(bad).to(idx.device)
            data_b = ops.GPData(self.theta_dev, self.pc_dev[idx[bad_dev]], torch.clamp(gv[bad_dev], min=0.0))
            Lb, _a, _b, _c, info_b = ops.gp_factorize(data_b, hyps[members[bad]], SIG2_OF_CONST)
            if np.any(info_b != 0):
                raise ValueError('PCGPwM_B200: covariance matrix not positive definite (info=%s)' % info_b)
            Linv[bad_dev] = Lb
            clamped[bad] = True
        return Linv, clamped

    def device_state(self, emulist, spc, pcti):
        """Everything predict needs, as device tensors (kept in fitinfo['_b200'])."""
        torch = self.torch
        dev = self.theta_dev.device
        hypcov = np.array([e['hypcov'] for e in emulist])

        def cov_sets(members):
            """distinct covariance hyper-parameter sets among these factors (cross-covariances are shared between factors)"""
            ukeys, fu, ulist = {}, np.zeros(len(members), dtype=np.int32), []
            for fi, j in enumerate(members):
                key = hypcov[j].tobytes()
                if key not in ukeys:
                    ukeys[key] = len(ulist)
                    ulist.append(hypcov[j])
                fu[fi] = ukeys[key]
            return np.array(ulist).reshape(len(ulist), hypcov.shape[1]), fu
        phi = pcti * spc['scale'][:, None]                           # PCGPwM.py:275
        t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
        nug = t(np.array([e['nug'] for e in emulist]))
        state = {'generation': next(_GENERATION), 'lbfgsb_driver': getattr(self, 'driver_used', None),
                 'theta': self.theta_dev, 'pw': self._pw, 'sig2': self._sig2, 'nug': nug,
                 'phi': t(phi), 'offset': t(spc['offset']), 'extravar': t(spc['extravar']),
                 'nevals': self.nevals, 'nf_total': int(len(self._members))}
        sh = getattr(self, '_shard', None)
        if sh is None:
            ulist, fu = cov_sets(self._members)
            state.update({'Linv': self._Linv, 'hypcov': t(ulist), 'fu': t(fu, torch.int32),
                          'fidx': t(self._fidx, torch.int32)})
            return state
        # factors stay with their owner: this rank predicts the components whose factor it holds ('local'), the
        # others arrive by all-gather (predict_pcspace); 'pw' / 'sig2' / 'nug' above cover all components
        lo, hi, pcs = sh['lo'], sh['hi'], sh['pcs_local']
        ulist, fu = cov_sets(self._members[lo:hi])
        sel = torch.from_numpy(pcs).to(dev)
        state['Linv'] = self._Linv                                   # (hi - lo, n, ldl): the factors this rank owns
        state['local'] = {'theta': self.theta_dev, 'Linv': self._Linv, 'pw': self._pw[sel].contiguous(),
                          'sig2': self._sig2[sel].contiguous(), 'nug': nug[sel].contiguous(), 'hypcov': t(ulist),
                          'fu': t(fu, torch.int32), 'fidx': t((self._fidx[pcs] - lo).astype(np.int32), torch.int32)}
        state['shard'] = {'world': sh['world'], 'rank': sh['rank'], 'counts': list(sh['counts']),
                          'inverse': torch.from_numpy(sh['inverse'].astype(np.int64)).to(dev),
                          'pcs_local': pcs, 'owner': sh['owner']}
        return state


class _LockstepEvaluator:
    """Rendezvous of the per-PC optimiser threads: the thread that completes the set of waiting requests
    launches them all at once (gp_nll_grad is batched over problems with per-problem data)."""

    def __init__(self, fitter, nthreads):
        import threading
        self.f = fitter
        self.cv = threading.Condition()
        self.live = nthreads
        self.pending = []

    def evaluate(self, data, hyps, want_grad, s0=SIG2_OF_CONST):
        req = {'data': data, 'hyps': hyps, 'grad': bool(want_grad), 's0': float(s0), 'out': None}
        with self.cv:
            self.pending.append(req)
            if len(self.pending) >= self.live:
                self._flush()
            else:
                while req['out'] is None:
                    self.cv.wait()
        out = req['out']
        if isinstance(out, BaseException):
            raise out
        return out

    def retire(self):
        with self.cv:
            self.live -= 1
            if self.live > 0 and len(self.pending) >= self.live:
                self._flush()

    def _flush(self):
        """Caller holds the lock.  One launch per (gradient flag, sig2ofconst) over all pending problems -- the same
        grouping as _GPFitter._run_block_coroutines."""
        f, torch = self.f, self.f.torch
        reqs, self.pending = self.pending, []
        try:
            for flag, s0 in sorted({(r['grad'], r['s0']) for r in reqs}):
                group = [r for r in reqs if r['grad'] == flag and r['s0'] == s0]
                counts = [r['hyps'].shape[0] for r in group]
                theta = torch.cat([r['data'].theta.expand(c, -1, -1) if r['data'].theta.dim() == 2
                                   else r['data'].theta for r, c in zip(group, counts)], 0)
                g = torch.cat([r['data'].g.expand(c, -1) for r, c in zip(group, counts)], 0)
                gv = torch.cat([r['data'].gvar.expand(c, -1) for r, c in zip(group, counts)], 0)
                data = f.ops.GPData(theta, g, gv)
                hyps = np.concatenate([r['hyps'] for r in group], 0)
                val, grad, info = f.ops.gp_nll_grad(data, hyps, f.mean_dev, f.std_dev, s0, flag)
                lo = 0
                for r, c in zip(group, counts):
                    r['out'] = (val[lo:lo + c], None if grad is None else grad[lo:lo + c], info[lo:lo + c])
                    lo += c
        except BaseException as exc:
            for r in reqs:
                if r['out'] is None:
                    r['out'] = exc
        self.cv.notify_all()


def import_state(theta, pc, gvar_damped, hyps, hypinds, pcti, spc):
    """Build the device state from given hyper-parameters (no optimisation): used to load a
    reference-fitted emulator (emulist[k]['hyp'], PCGPwM.py:817-846) for like-for-like prediction."""
    fi = {'pc': np.asarray(pc), 'unscaled_pcstdvar': np.zeros_like(pc), 'eta': 10, 'dampalpha': 0.3}
    fitter = _GPFitter(fi, np.asarray(theta, dtype=np.float64), -10, -20, None, None)
    fitter.gvar = np.asarray(gvar_damped, dtype=np.float64)
    fitter.gvar_dev = fitter.torch.from_numpy(np.ascontiguousarray(fitter.gvar.T)).to(fitter.theta_dev.device)
    rows = [np.arange(fitter.n)] * pc.shape[1]
    emulist = fitter._finalise(np.asarray(hyps, dtype=np.float64), np.asarray(hypinds).astype(int), rows)
    return emulist, fitter.device_state(emulist, spc, pcti)


def shard_state(fitinfo):
    """Turn a REPLICATED device state (every rank fitted or adopted the same emulator) into the owner layout of a
    fit with factors='owner': each rank of the initialised torch.distributed group keeps a contiguous block of the
    distinct factors and drops the rest; predict / calibration then run PC-sharded and collectively (SURVEY 8e,
    "by PC").  No-op on one rank or when the state is sharded already.  Returns fitinfo."""
    from .. import _dist
    ops = _ops()
    torch = ops._torch()
    st = fitinfo['_b200']
    comm = _dist.Comm()
    if comm.world == 1 or st.get('shard') is not None:
        return fitinfo
    nf = int(st['Linv'].shape[0])
    fidx = st['fidx'].cpu().numpy().astype(np.int64)
    fu = st['fu'].cpu().numpy().astype(np.int64)
    sizes, starts = _dist.block_partition(nf, comm.world)
    lo, hi = int(starts[comm.rank]), int(starts[comm.rank + 1])
    plan = _dist.owner_plan(fidx, starts, comm.rank)
    pcs = plan['pcs_local']
    dev = st['Linv'].device
    sets, fu_local = np.unique(fu[lo:hi], return_inverse=True) if hi > lo else (np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64))
    sel = torch.from_numpy(pcs).to(dev)
    local_L = st['Linv'][lo:hi].clone()
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
    new = dict(st)
    new['generation'] = next(_GENERATION)
    new['nf_total'] = nf
    new['Linv'] = local_L
    new['local'] = {'theta': st['theta'], 'Linv': local_L, 'pw': st['pw'][sel].contiguous(),
                    'sig2': st['sig2'][sel].contiguous(), 'nug': st['nug'][sel].contiguous(),
                    'hypcov': st['hypcov'][torch.from_numpy(sets).to(dev)].contiguous(),
                    'fu': t(fu_local, torch.int32), 'fidx': t(fidx[pcs] - lo, torch.int32)}
    new['shard'] = {'world': comm.world, 'rank': comm.rank, 'counts': list(plan['counts']),
                    'inverse': torch.from_numpy(plan['inverse'].astype(np.int64)).to(dev),
                    'pcs_local': pcs, 'owner': plan['owner']}
    for key in ('hypcov', 'fu', 'fidx'):
        new.pop(key, None)
    fitinfo['_b200'] = new
    del st
    torch.cuda.empty_cache()
    return fitinfo


def adopt(fitinfo):
    """Attach the device state to a `fitinfo` produced by the reference's own PCGPwM.fit (no optimisation).

    Factors R at the stored hyper-parameters of every component (emulist[k]['hyp'], 'hypind'; PCGPwM.py:810-846)
    and keeps the reference's PCA (pc, pcti, standardpcinfo).  After this, PCGPwM_B200.predict and
    directbayeswoodbury_B200 work on an emulator that was fitted on the CPU with method='PCGPwM':
        PCGPwM_B200.adopt(emu._info)
    directbayeswoodbury_B200 does it on first use."""
    need = ('emulist', 'pc', 'pcti', 'standardpcinfo', 'unscaled_pcstdvar', 'theta', 'eta', 'dampalpha')
    if not isinstance(fitinfo, dict) or any(k not in fitinfo for k in need):
        raise ValueError('adopt() needs the fitinfo dictionary of an emulator fitted with PCGPwM.')
    theta = np.asarray(fitinfo['theta'], dtype=np.float64)
    fitter = _GPFitter(fitinfo, theta, -10, -20, None, None)
    hyps = np.array([e['hyp'] for e in fitinfo['emulist']], dtype=np.float64)
    hypinds = np.array([e['hypind'] for e in fitinfo['emulist']]).astype(int)
    rows = [np.arange(fitter.n)] * hyps.shape[0]
    emulist = fitter._finalise(hyps, hypinds, rows)
    fitinfo['_b200'] = fitter.device_state(emulist, fitinfo['standardpcinfo'], fitinfo['pcti'])
    return fitinfo


# ------------------------------------------------------------------------------------------------
# predict
# ------------------------------------------------------------------------------------------------
def match_rows(x, xfit):
    """Rows of the fit-time x that each requested x row equals (PCGPwM.py:197-214).

    Returns (xind, xnewind): prediction row xnewind[i] is fit row xind[i]; unmatched rows are absent.
    """
    if x is None:
        idx = np.arange(xfit.shape[0])
        return idx, idx
    try:
        if x.shape == xfit.shape and (np.all(np.equal(x, xfit)) or np.allclose(x, xfit)):
            idx = np.arange(x.shape[0])
            return idx, idx
    except Exception:
        pass
    match = np.ones((x.shape[0], xfit.shape[0]), dtype=bool)
    for c in range(x.shape[1]):
        try:
            match &= np.isclose(x[:, c][:, None].astype(float), xfit[:, c].astype(float))
        except Exception:
            match &= np.equal(x[:, c][:, None], xfit[:, c])
    pairs = np.argwhere(match)
    return pairs[:, 1], pairs[:, 0]


def predict_pcspace(fitinfo, theta, return_grad=False, chunk=4096, population=None):
    """Device tensors predvecs, predvars (B,k) [+ gradients (B,k,d)] for theta (B,d) numpy/tensor.

    population: size of the whole population when theta is one rank's slice of it (see ops.gp_predict)."""
    ops = _ops()
    torch = ops._torch()
    st = fitinfo['_b200']
    tn = ops._f64(torch, np.atleast_2d(theta) if not isinstance(theta, torch.Tensor) else theta)
    B = tn.shape[0]
    population = B if population is None else population
    if st.get('shard') is not None:
        return _predict_pcspace_sharded(st, tn, return_grad, chunk, population)
    if B <= chunk:
        return ops.gp_predict(st, tn, return_grad, population=population)
    parts = [ops.gp_predict(st, tn[s:s + chunk], return_grad, population=population) for s in range(0, B, chunk)]
    return tuple(None if parts[0][i] is None else torch.cat([p[i] for p in parts], 0) for i in range(4))


def _predict_pcspace_sharded(st, tn, return_grad, chunk, population):
    """PC-sharded prediction for a fit whose factors stayed with their owners (fit option factors='owner').

    COLLECTIVE: every rank of the fit's process group calls with the same theta.  Each rank evaluates the components
    whose factor it owns; one all-gather of B x k x (2 + 2d) doubles gives every rank the full tables."""
    from .. import _dist
    ops = _ops()
    torch = ops._torch()
    sh, loc = st['shard'], st['local']
    comm = _dist.Comm()
    if comm.world != sh['world'] or comm.rank != sh['rank']:
        raise RuntimeError('this emulator was fitted with its factors sharded over %d ranks (this is rank %d of it); '
                           'predict must run in the same torch.distributed group (or refit with factors="replicated")'
                           % (sh['world'], sh['rank']))
    B, d = tn.shape
    kl = int(loc['pw'].shape[0])
    Q = 2 + (2 * d if return_grad else 0)
    packed = torch.empty((kl, B, Q), dtype=torch.float64, device=tn.device)
    if kl:
        for s0 in range(0, B, chunk):
            pv, pvar, dpv, dpvar = ops.gp_predict(loc, tn[s0:s0 + chunk], return_grad, population=population)
            blk = packed[:, s0:s0 + chunk]
            blk[:, :, 0], blk[:, :, 1] = pv.T, pvar.T
            if return_grad:
                blk[:, :, 2:2 + d] = dpv.permute(1, 0, 2)
                blk[:, :, 2 + d:] = dpvar.permute(1, 0, 2)
    full = comm.allgather_tensor(packed, sh['counts'])[st['shard']['inverse']]          # (k, B, Q), component order
    pv, pvar = full[:, :, 0].T.contiguous(), full[:, :, 1].T.contiguous()
    if not return_grad:
        return pv, pvar, None, None
    return pv, pvar, full[:, :, 2:2 + d].permute(1, 0, 2).contiguous(), full[:, :, 2 + d:].permute(1, 0, 2).contiguous()


def predict(predinfo, fitinfo, x, theta, **kwargs):
    """Same contract and output shapes as PCGPwM.predict (PCGPwM.py:140-328)."""
    ops = _ops()
    torch = ops._torch()
    return_grad = kwargs.get('return_grad', False) is True
    return_covx = not (kwargs.get('return_covx', True) is False)
    st = fitinfo['_b200']
    theta = np.asarray(theta, dtype=np.float64)
    if theta.ndim < 2:
        theta = theta.reshape(1, -1)
    B, d = theta.shape
    xind, xnewind = match_rows(x, fitinfo['x'])
    nx = fitinfo['x'].shape[0] if x is None else x.shape[0]

    pv, pvar, dpv, dpvar = predict_pcspace(fitinfo, theta, return_grad)
    sel = torch.from_numpy(np.asarray(xind, dtype=np.int64)).to(pv.device)
    phi = st['phi'][sel].contiguous()
    mean, var, cxh, mg, cg = ops.pc_backproject(phi, st['offset'][sel].contiguous(), st['extravar'][sel].contiguous(),
                                                pv, pvar, dpv, dpvar, want_covx=return_covx)
    K = pv.shape[1]

    in_order = xnewind.shape[0] == nx and np.array_equal(xnewind, np.arange(nx))

    def scatter(a, tail):
        if in_order:                                   # every requested x row matched, in order: nothing to fill
            return a
        out = np.full((nx,) + tail, np.nan)
        out[xnewind] = a
        return out

    want_cg = return_grad and return_covx
    if want_cg and not np.array_equal(xnewind, xind):
        raise ValueError('Check shuffling of x indices within predictions.')   # PCGPwM.py:318-320
    # one synchronisation for all results, copied through pinned host memory
    (mean_h, var_h, ev_h, pvar_h, pv_h, phi_h, cxh_h, mg_h, dpvar_h, dpv_h, cg_h) = ops.to_host(
        [mean, var, st['extravar'][sel], pvar, pv, phi, cxh if return_covx else None,
         mg if return_grad else None, dpvar if return_grad else None, dpv if return_grad else None,
         cg if want_cg else None])
    predinfo['mean'] = scatter(mean_h, (B,))
    predinfo['var'] = scatter(var_h, (B,))
    predinfo['extravar'] = ev_h
    predinfo['predvars'] = pvar_h
    predinfo['predvecs'] = pv_h
    predinfo['phi'] = phi_h
    predinfo['return_covx'] = return_covx
    predinfo['return_grad'] = return_grad
    if return_covx:
        predinfo['covxhalf'] = scatter(cxh_h, (B, K))
    if return_grad:
        predinfo['mean_gradtheta'] = scatter(mg_h, (B, d))
        predinfo['predvars_gradtheta'] = dpvar_h
        predinfo['predvecs_gradtheta'] = dpv_h
        if return_covx:
            predinfo['covxhalf_gradtheta'] = cg_h
    return


def predictlpdf(predinfo, f, addvar=0, **kwargs):
    """Log predictive density of simulator output f (m, B) at the predicted (x, theta); same contract as
    PCGPwM.predictlpdf (PCGPwM.py:331-364): (B,1), plus the gradient (B,d) when predict ran with return_grad.

    Host numpy on the small PC-space quantities of `predinfo` (k x k systems per theta, batched)."""
    totvar = addvar + predinfo['extravar']
    isd = 1 / np.sqrt(totvar)
    rf = (np.asarray(f, dtype=np.float64) - predinfo['mean']) * isd[:, None]        # (m, B)
    Gf = predinfo['phi'].T * isd                                                       # (k, m)
    Gfrf = (Gf @ rf).T                                                                 # (B, k)
    v = predinfo['predvars']                                                           # (B, k)
    M = Gf @ Gf.T + np.einsum('bk,kl->bkl', 1 / v, np.eye(v.shape[1]))                # diag(1/v_b) + Gf Gf'
    Minv = np.linalg.inv(M)
    t1 = np.einsum('bkl,bl->bk', Minv, Gfrf)
    _sign, logdet = np.linalg.slogdet(M)
    likv = np.sum(rf ** 2, 0) - np.sum(Gfrf * t1, 1) + np.sum(np.log(v), 1) + logdet
    if not predinfo['return_grad']:
        return (-likv / 2).reshape(-1, 1)
    rf2 = -predinfo['mean_gradtheta'] * isd[:, None, None]                            # (m, B, d)
    dv = predinfo['predvars_gradtheta']                                                # (B, k, d)
    grt = dv / v[:, :, None]
    dlikv = 2 * np.einsum('mbd,mb->bd', rf2, rf) + grt.sum(1) \
        - np.einsum('bk,bkd->bd', np.einsum('bkk->bk', Minv) / v, grt) \
        - 2 * np.einsum('km,mbd,bk->bd', Gf, rf2, t1) - np.einsum('bk,bkd->bd', (t1 / v) ** 2, dv)
    return (-likv / 2).reshape(-1, 1), -dlikv / 2


def supplementtheta(fitinfo, size, theta, thetachoices, choicecosts, cal, **kwargs):
    """Suggest `size` new parameters out of `thetachoices` (PCGPwM.py:367-530), as the reference behaves today.

    The reference's selection and obviation criteria are commented out (PCGPwM.py:468-478, 517-526): every
    criterion it compares is an array of zeros, so a pick is `argmax(0 / choicecosts)` over the remaining choices
    (the first one for positive costs), `info['crit']` is 0 / cost and no pending run is ever suggested for
    obviation.  The covariance matrices it still builds never reach the result and are not computed here; the
    zero criteria go through the same floating-point expressions so signed-zero / NaN costs select as they do there.
    kwargs: 'pending', 'costpending', 'includepending' as in the reference.  Returns (theta (size, d), info)."""
    if theta is None:
        raise ValueError('this method is designed to take in the theta values.')
    nfit, d = fitinfo['theta'].shape
    include = kwargs.get('includepending') is True
    costs = np.asarray(choicecosts)
    if include:
        base = kwargs['costpending'] if 'costpending' in kwargs else np.mean(costs)
        costpending = base * np.ones(nfit)
    poss = np.asarray(thetachoices)
    critsave = np.zeros(poss.shape[0])
    picks = np.zeros((size, d))
    with np.errstate(divide='ignore', invalid='ignore'):
        for j in range(size):
            if poss.shape[0] < 1.5:
                break
            crit = np.zeros(poss.shape[0])
            jstar = np.argmax(crit / costs)
            critsave[j] = crit[jstar] / costs[jstar]
            picks[j] = poss[jstar]
            poss = np.delete(poss, jstar, 0)
            costs = np.delete(costs, jstar)
        info = {'crit': critsave}
        if include:
            critpend = np.zeros(nfit)
            info['obviatesugg'] = np.where((np.mean(critsave[:size]) > critpend / costpending) > 0.5)[0]
    return picks, info


def _param_desc(fitinfo):
    """Human-readable summary used by emulator.__repr__ (PCGPwM.py:938-964)."""
    el = fitinfo['emulist']
    k = len(el)
    scales = 1 / (1 + np.exp(np.array([el[j]['hyp'][-2] for j in range(k)])))
    lengthscales = np.array([el[j]['hyp'][:-3] for j in range(k)])
    nuggets = np.array([el[j]['nug'] for j in range(k)])
    return ('\taverage emulation residual variance (from principal components):\t{:.3E}\n'
            '\tnumber of GP components:\t{:d}\n'
            '\tGP parameters, following Gramacy (ch.5, 2022) notations:\n'
            '\t\tscales (in log): \t{:s}\n'
            '\t\tlengthscales (in log):\n\t\t\t{:s}\n'
            '\t\tnuggets (in log):\t{:s}\n').format(
        fitinfo['standardpcinfo']['extravar'].mean(), k,
        pformat(['{:.3f}'.format(np.log(v)) for v in scales]),
        pformat(lengthscales).replace('\n', '\n\t\t\t'),
        pformat(['{:.3f}'.format(np.log(v)) for v in nuggets]))
