"""PTLMC_B200 -- parallel-tempering Langevin Monte Carlo sampler plugin, batched for the B200 path.

Same algorithm and options as surmise/utilitiesmethods/PTLMC.py (`sampler(logpostfunc, draw_func, theta0, numsamp,
numtemps, numchain, sampperchain, maxtemp)` -> {'theta', 'logpost'}, utilities.py:62-71), restructured so that the
log-posterior closure -- one fused device evaluation per call in directbayeswoodbury_B200 -- only ever sees large
batches:
  * the pre-optimiser (PTLMC.py:140-173: one L-BFGS-B run per chain, B = 1 calls, value and gradient separately:
    thousands of latency-bound calls) runs all chains in lock step, one batched call per optimiser step;
  * the temperature-exchange sweep (PTLMC.py:259-273, 8.8 ms per step as a Python loop) is a compiled host helper
    (sb200_host_tempexchange) fed with pre-drawn random numbers;
  * the gradient-free probe uses a 2-D theta, so `return_grad=False` is honoured instead of silently falling back to
    gradient evaluations (PTLMC.py:96-105).
Registered by surmise_b200.register() as surmise.utilitiesmethods.PTLMC_B200, so
`calibrator(..., args={'sampler': 'PTLMC_B200', ...})` is drop-in; usable without surmise as well.
"""
import ctypes
import threading

import numpy as np
import scipy.optimize as spo


def temperature_ladder(numtemps, numchain, maxtemp):
    """maxtemp ... maxtemp^(1/(numtemps+1)) geometrically, then numchain chains at T = 1 (PTLMC.py:75-82)."""
    hot = np.exp(np.linspace(np.log(maxtemp), np.log(maxtemp) / (numtemps + 1), numtemps))
    return np.concatenate((hot, np.ones(numchain)))[:, None]


def tempexchange(lpost, temps, rtv, logu):
    """Revised chain order after len(rtv)/B exchange passes (PTLMC.py:259-273); random draws supplied."""
    from .. import _lib
    lp = np.ascontiguousarray(np.squeeze(lpost), dtype=np.float64)
    tt = np.ascontiguousarray(np.squeeze(temps), dtype=np.float64)
    B = lp.shape[0]
    rtv = np.ascontiguousarray(rtv, dtype=np.int32)
    logu = np.ascontiguousarray(logu, dtype=np.float64)
    iters = rtv.size // B
    order = np.empty(B, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.lib.sb200_host_tempexchange(p(lp), p(tt), B, p(rtv), p(logu), iters, p(order))
    return order.astype(np.int64)


class _Lockstep:
    """Batch the evaluations requested by several optimiser threads into one closure call."""

    def __init__(self, fn, nthreads):
        self.fn, self.live, self.pending = fn, nthreads, []
        self.cv = threading.Condition()

    def __call__(self, theta_row):
        req = {'x': theta_row, 'out': None}
        with self.cv:
            self.pending.append(req)
            if len(self.pending) >= self.live:
                self._flush()
            else:
                while req['out'] is None:
                    self.cv.wait()
        if isinstance(req['out'], BaseException):
            raise req['out']
        return req['out']

    def retire(self):
        with self.cv:
            self.live -= 1
            if self.live > 0 and len(self.pending) >= self.live:
                self._flush()

    def _flush(self):
        reqs, self.pending = self.pending, []
        try:
            vals, grads = self.fn(np.array([r['x'] for r in reqs]))
            for i, r in enumerate(reqs):
                r['out'] = (float(vals[i]), None if grads is None else grads[i])
        except BaseException as exc:
            for r in reqs:
                r['out'] = exc
        self.cv.notify_all()


def _preoptimise(theta0, value_and_grad, has_grad):
    """One bounded L-BFGS-B run per chain from its starting point, then a random move off the optimum
    (PTLMC.py:122-173), all chains advancing in lock step."""
    numopt, d = theta0.shape
    cen = np.mean(theta0, 0)
    sca = np.maximum(np.std(theta0, 0), 1e-8 * np.std(theta0))
    lo = np.maximum(-10 * np.ones(d), np.min((theta0 - cen) / sca, 0))
    hi = np.minimum(10 * np.ones(d), np.max((theta0 - cen) / sca, 0))
    bounds = spo.Bounds(lo, hi)
    kicks = np.random.standard_normal(size=(numopt, 8, d))        # drawn up-front: results independent of scheduling
    out = np.array(theta0, dtype=np.float64)
    if has_grad:
        try:
            from ..emulationmethods.PCGPwM_B200 import _LbfgsbCore
            if _LbfgsbCore.available():
                return _preoptimise_steps(theta0, value_and_grad, cen, sca, lo, hi, kicks, out, _LbfgsbCore)
        except ImportError:
            pass

    def batch(thetas):
        v, g = value_and_grad(thetas)
        return np.squeeze(v, axis=1) if v.ndim > 1 else v, g
    ev = _Lockstep(batch, numopt)

    def neg(xp):                                                   # scaled coordinates, as the reference
        v, g = ev(cen + sca * xp)
        return (-v, -sca * g) if has_grad else -v

    def work(c):
        try:
            x0 = (theta0[c] - cen) / sca
            res = spo.minimize(neg, x0, method='L-BFGS-B', jac=bool(has_grad), bounds=bounds)
            best = cen + sca * res.x
            if c > 0:                                              # the first chain stays on its optimum (PTLMC.py:160)
                W, V = np.linalg.eigh(res.hess_inv @ np.eye(d))
                l0 = neg(res.x)[0] if has_grad else neg(res.x)
                step = 4.0
                for t in range(8):
                    r = (V.T * np.sqrt(np.abs(W))) @ (V @ kicks[c, t])
                    trial = neg(step * r + res.x)
                    trial = trial[0] if has_grad else trial
                    if trial - l0 < 3 * d:
                        best = cen + sca * (step * r + res.x)
                        break
                    step /= 2
                    if step < 1 / 16:
                        break
            out[c] = best
        finally:
            ev.retire()
    threads = [threading.Thread(target=work, args=(c,)) for c in range(numopt)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    return out


def _preoptimise_steps(theta0, value_and_grad, cen, sca, lo, hi, kicks, out, Core):
    """_preoptimise without threads: every chain is a generator that yields the points it needs; one scheduler
    evaluates all waiting points with a single batched closure call.  The L-BFGS-B runs go through scipy's own core
    in reverse-communication form with scipy.optimize.minimize's default tolerances (same iterates)."""
    numopt, d = theta0.shape

    def chain(c):
        core = Core((theta0[c] - cen) / sca, lo, hi, 1e-5)          # gtol: minimize's L-BFGS-B default
        while True:
            xs = core.ask()
            if xs is None:
                break
            v, g = yield cen + sca * xs
            core.tell(-v, -sca * g)
        best = cen + sca * core.x
        if c > 0:                                                  # the first chain stays on its optimum (PTLMC.py:160)
            m, n = core.M, d
            ncorr = int(min(core.isave[30], m))
            H = spo.LbfgsInvHessProduct(core.wa[0:m * n].reshape(m, n)[:ncorr],
                                        core.wa[m * n:2 * m * n].reshape(m, n)[:ncorr]) @ np.eye(d)
            W, V = np.linalg.eigh(H)
            v, _ = yield cen + sca * core.x
            l0 = -v
            step = 4.0
            for t in range(8):
                r = (V.T * np.sqrt(np.abs(W))) @ (V @ kicks[c, t])
                v, _ = yield cen + sca * (step * r + core.x)
                if -v - l0 < 3 * d:
                    best = cen + sca * (step * r + core.x)
                    break
                step /= 2
                if step < 1 / 16:
                    break
        return best
    gens = [chain(c) for c in range(numopt)]
    pending = {}

    def advance(c, value=None):
        try:
            pending[c] = gens[c].send(value) if value is not None else next(gens[c])
        except StopIteration as stop:
            pending.pop(c, None)
            out[c] = stop.value
    for c in range(numopt):
        advance(c)
    while pending:
        order = sorted(pending)
        vals, grads = value_and_grad(np.array([pending[c] for c in order]))
        vals = np.squeeze(vals, axis=1) if np.ndim(vals) > 1 else vals
        for r, c in enumerate(order):
            advance(c, (float(vals[r]), grads[r]))
    return out


def _device_chain(dl, thetac, fval, dfval, temps, cov0, hc, numtemps, numchain, tune, sampperchain, taracc,
                  steps_per_graph=25):
    """The main loop of the sampler (PTLMC.py:203-257) without leaving the GPU.

    dl: the closure's DeviceLogpost (prior box + fixed-buffer likelihood evaluation).  Per MCMC step the stream sees
    predict -> Woodbury -> ONE single-CTA kernel (accept, 5 B exchange proposals, adaptation or save, next Langevin
    proposal); `steps_per_graph` consecutive steps are captured once as a CUDA graph and replayed, so the host issues
    one launch per 25 steps and synchronises once at the end.  Returns the kept draws (numchain, sampperchain, d) and
    a dict of statistics, or None when the device loop cannot run (population too large for one CTA's shared memory)."""
    from .. import ops, _lib
    torch = ops._torch()
    B, d = thetac.shape
    lo, hi = dl.box
    ev = dl.evaluator(B, d)
    dev = ev.thetap.device
    if getattr(ev, 'peer', None) is not None:
        steps_per_graph -= steps_per_graph % 2                # the peer exchange alternates two buffers: whole pairs per replay
    t64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    seed = int(np.random.randint(0, 2 ** 31 - 1)) * 2 ** 31 + int(np.random.randint(0, 2 ** 31 - 1))
    tau0 = -1.0
    rho0 = 2 * (1 + np.tanh(tau0))
    st = {'B': B, 'd': d, 'numtemps': numtemps, 'tune': tune, 'nsave': sampperchain, 'seed': seed, 'taracc': taracc,
          'thetac': t64(thetac), 'fval': t64(np.squeeze(fval)), 'dfval': t64(dfval), 'thetap': ev.thetap,
          'z': torch.zeros((B, d), dtype=torch.float64, device=dev), 'adjrho': t64(np.squeeze(rho0 * temps ** (1 / 3))),
          'temps': t64(np.squeeze(temps)), 'lo': t64(lo), 'hi': t64(hi), 'hc': t64(hc), 'cov0': t64(cov0),
          'scal': t64(np.array([tau0, 0.0, 0.0, 0.0, rho0, 0.0, 0.0, 0.0])),
          'kstep': torch.zeros(1, dtype=torch.int64, device=dev),
          'save': torch.zeros((numchain, sampperchain, d), dtype=torch.float64, device=dev), 'valid': ev.valid}
    total = tune + sampperchain
    import time
    t_begin = time.perf_counter()

    def one_step():
        ev.launch()
        ops.ptlmc_step(st, 3, ev.ll, ev.dll)
    try:
        ops.ptlmc_step(st, 2)                                 # proposal of step 0
    except _lib.SurmiseB200Error:
        return None
    done, graph, used_graph = 0, None, False
    if total - 1 >= 3 * steps_per_graph:
        one_step()                                            # warm-up outside capture: workspaces, attributes, NCCL
        done = 1
        try:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(steps_per_graph):
                    one_step()                                # recorded, not executed: the replays below run them
        except Exception:                                     # capture unavailable (e.g. a collective that cannot be
            graph = None                                      # captured): same launches, issued eagerly
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    t_loop = time.perf_counter()
    done_before = done
    while total - 1 - done > 0:
        if graph is not None and total - 1 - done >= steps_per_graph:
            graph.replay()
            done += steps_per_graph
            used_graph = True
        else:
            one_step()
            done += 1
    ev.launch()                                               # likelihood of the last proposal
    ops.ptlmc_step(st, 1, ev.ll, ev.dll)
    save, scal = ops.to_host([st['save'], st['scal']])
    t_end = time.perf_counter()
    if hasattr(ev, 'check'):
        ev.check()
    stats = {'theta_evaluated_on_device': total * B, 'mcmc_steps': total, 'cuda_graph': used_graph,
             'exchange': ('none' if getattr(ev, 'segments', None) is None else
                          ('peer memory' if getattr(ev, 'peer', None) is not None else 'nccl all_gather')),
             'setup_and_capture_s': t_loop - t_begin, 'loop_s': t_end - t_loop,
             'ms_per_step_in_loop': 1e3 * (t_end - t_loop) / max(total - 1 - done_before, 1),
             'steps_per_graph': steps_per_graph if used_graph else 0, 'accept_rate': float(scal[2] / (total * B)),
             'swap_rate': float(scal[3] / (5 * B * total)), 'final_rho': float(scal[4]),
             'outside_box_rate': float(scal[5] / (total * B)), 'seed': seed}
    return save, stats


def sampler(logpostfunc, draw_func, theta0=None, numsamp=2000, numtemps=32, numchain=16, sampperchain=400,
            maxtemp=30, device_resident='auto', **ptlmc_options):
    """Parallel-tempering ensemble MCMC with Langevin proposals; returns {'theta': (numsamp, d), 'logpost': ...}.

    device_resident: 'auto' (default) -- when the closure comes from directbayeswoodbury_B200 and its prior declares a
    box (`thetaprior.box = (lo, hi)`: uniform inside, -inf outside), the main loop runs entirely on the GPU
    (_device_chain); otherwise, or with False, the host loop below calls the closure once per step (any Python prior).
    The two loops are the same algorithm with different random-number streams."""
    if theta0 is None or theta0.shape[0] < 10 * theta0.shape[1]:
        theta0 = draw_func(1000)
    theta0 = np.array(theta0, dtype=np.float64)
    d = theta0.shape[1]
    tune = int(np.ceil(sampperchain * 0.5))
    B = numtemps + numchain
    temps = temperature_ladder(numtemps, numchain, maxtemp)

    probe = logpostfunc(theta0[0:2, :])
    has_grad = isinstance(probe, tuple)
    if has_grad:
        if len(probe) != 2:
            raise ValueError('log density does not return 1 or 2 elements')
        if probe[1].shape[1] != d:
            raise ValueError('derivative appears to be the wrong shape')
        try:
            t = logpostfunc(theta0[0:2, :], return_grad=False)
            if isinstance(t, tuple):
                raise ValueError('Cannot stop returning a grad')

            def value_only(th):
                return logpostfunc(th, return_grad=False)
        except (TypeError, ValueError):
            def value_only(th):
                return logpostfunc(th)[0]

        def value_and_grad(th):
            return logpostfunc(th)
    else:
        def value_only(th):
            return logpostfunc(th)

        def value_and_grad(th):
            return logpostfunc(th), None
    taracc = 0.60 if has_grad else 0.25

    # starting points: best-ranked draws (with a random tie-breaker), each improved by a bounded optimiser
    import time
    t_start = time.perf_counter()
    rank = np.argsort(-np.squeeze(value_only(theta0)) + d * np.random.standard_normal(size=theta0.shape[0]) ** 2)
    thetac = _preoptimise(theta0[rank[:B]], value_and_grad, has_grad)
    t_preopt = time.perf_counter() - t_start

    if has_grad:
        fval, dfval = value_and_grad(thetac)
        fval, dfval = fval / temps, dfval / temps
    else:
        fval, dfval = value_only(thetac) / temps, None
    save = np.zeros((numchain, sampperchain, d))
    cov0 = np.atleast_2d(np.cov(thetac.T))
    if d > 1:
        cov0 = 0.9 * cov0 + 0.1 * np.diag(np.diag(cov0))
        W, V = np.linalg.eigh(cov0)
        hc = V @ np.diag(np.sqrt(W)) @ V.T
    else:
        hc = np.sqrt(cov0)
    dl = getattr(logpostfunc, 'b200', None)
    if (device_resident and has_grad and dl is not None and dl.box is not None and dl.box[0].shape[0] == d
            and np.all(np.isfinite(fval)) and np.all(np.isfinite(dfval))):
        # the box prior is constant inside: it cancels in every acceptance ratio, so the device chain carries the
        # log-likelihood alone (the constant is removed from the starting values)
        prior0 = np.asarray(dl.prior.lpdf(thetac), dtype=np.float64).reshape(-1, 1)
        if np.all(np.isfinite(prior0)):
            res = _device_chain(dl, thetac, fval - prior0 / temps, dfval, temps, cov0, hc, numtemps, numchain, tune,
                                sampperchain, taracc)
            if res is not None:
                save, stats = res
                stats['preoptimise_s'] = t_preopt
                dl.record(**stats)
                flat = save.reshape(-1, d)
                theta = flat[np.random.choice(flat.shape[0], size=numsamp)]
                return {'theta': theta, 'logpost': value_only(theta)}
    tau = -1.0
    rho = 2 * (1 + np.tanh(tau))
    adjrho = rho * temps ** (1 / 3)
    accepted = 0.0
    for k in range(tune + sampperchain):
        z = np.random.normal(0, 1, thetac.shape)
        thetap = thetac + np.sqrt(2) * adjrho * (z @ hc)
        if has_grad:
            thetap = thetap + (adjrho ** 2) * (dfval @ cov0)
            fvalp, dfvalp = value_and_grad(thetap)
            fvalp, dfvalp = fvalp / temps, dfvalp / temps
            t1 = z / np.sqrt(2)
            t2 = (adjrho / 2) * ((dfval + dfvalp) @ hc)
            qadj = -(2 * np.sum(t1 * t2, 1) + np.sum(t2 ** 2, 1))
        else:
            fvalp = value_only(thetap) / temps
            qadj = np.zeros(B)
        take = np.where(np.log(np.random.uniform(size=B)) < np.squeeze(fvalp - fval) + np.squeeze(qadj))[0]
        if take.size:
            accepted += take.size / B
            thetac[take] = thetap[take]
            fval[take] = fvalp[take]
            if has_grad:
                dfval[take] = dfvalp[take]
        # exchange along the temperature ladder (5 passes)
        fn = fval * temps
        rtv = np.random.choice(np.arange(1, B), size=5 * B)
        logu = np.log(np.random.uniform(size=5 * B))
        order = tempexchange(fn, temps, rtv, logu)
        fval = fn[order] / temps
        thetac = thetac[order]
        if has_grad:
            dfval = (temps * dfval)[order] / temps
        if k < tune and k % 10 == 0:
            tau = tau + (accepted / 10 - taracc) / np.sqrt(1 + k / 10)
            rho = 2 * (1 + np.tanh(tau))
            adjrho = rho * temps ** (1 / 3)
            accepted = 0.0
        elif k >= tune:
            save[:, k - tune, :] = thetac[numtemps:]
    flat = save.reshape(-1, d)
    theta = flat[np.random.choice(flat.shape[0], size=numsamp)]
    return {'theta': theta, 'logpost': value_only(theta)}
