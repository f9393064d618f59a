"""GPU: the device-resident PT-LMC sampler (surmise_b200/csrc/sampler.cu, utilitiesmethods/PTLMC_B200._device_chain)
against what surmise's PTLMC defines (utilitiesmethods/PTLMC.py:203-273).

The kernel draws its own random numbers (Philox), so draw-for-draw identity with the numpy sampler is not defined;
checked instead: the deterministic parts of a step exactly (Langevin proposal from the stored normals, box mask,
exchange as a permutation that conserves the multiset of log-posteriors, saved draws), the random parts statistically
(normal moments), and the whole chain distributionally on a target with known moments and through the calibrator."""
import numpy as np
import pytest

from gpu_util import report
from synth import make_simulator, make_observation

pytestmark = pytest.mark.gpu


def _state(B, d, numtemps, tune, nsave, seed=123):
    import torch
    from surmise_b200.utilitiesmethods.PTLMC_B200 import temperature_ladder
    dev = torch.device('cuda')
    rng = np.random.default_rng(seed)
    temps = temperature_ladder(numtemps, B - numtemps, 30.0)[:, 0]
    A = rng.normal(size=(d, d))
    cov0 = A @ A.T / d * 0.01 + 0.001 * np.eye(d)
    W, V = np.linalg.eigh(cov0)
    hc = V @ np.diag(np.sqrt(W)) @ V.T
    t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
    host = dict(thetac=rng.uniform(0.3, 0.7, size=(B, d)), fval=rng.normal(size=B), dfval=rng.normal(size=(B, d)),
                temps=temps, cov0=cov0, hc=hc, adjrho=0.5 * temps ** (1 / 3))
    st = {'B': B, 'd': d, 'numtemps': numtemps, 'tune': tune, 'nsave': nsave, 'seed': 987654321, 'taracc': 0.6,
          'thetac': t(host['thetac']), 'fval': t(host['fval']), 'dfval': t(host['dfval']),
          'thetap': torch.zeros((B, d), dtype=torch.float64, device=dev), 'z': torch.zeros((B, d), dtype=torch.float64, device=dev),
          'adjrho': t(host['adjrho']), 'temps': t(temps), 'lo': t(np.zeros(d)), 'hi': t(np.ones(d)), 'hc': t(hc),
          'cov0': t(cov0), 'scal': t(np.array([-1.0, 0, 0, 0, 0.5, 0, 0, 0])),
          'kstep': torch.zeros(1, dtype=torch.int64, device=dev),
          'save': torch.zeros((B - numtemps, max(nsave, 1), d), dtype=torch.float64, device=dev),
          'valid': torch.ones(B, dtype=torch.int32, device=dev)}
    return st, host


@pytest.mark.parametrize('B,d,numtemps', [(264, 8, 8), (24, 2, 6), (37, 1, 5), (48, 16, 32)])
def test_proposal_is_the_langevin_step_of_the_stored_normals(B, d, numtemps):
    from surmise_b200 import ops
    st, h = _state(B, d, numtemps, tune=10, nsave=4)
    ops.ptlmc_step(st, 2)
    z, thetap, valid = st['z'].cpu().numpy(), st['thetap'].cpu().numpy(), st['valid'].cpu().numpy()
    want = h['thetac'] + np.sqrt(2) * h['adjrho'][:, None] * (z @ h['hc']) + (h['adjrho'][:, None] ** 2) * (h['dfval'] @ h['cov0'])
    assert np.max(np.abs(thetap - want)) < 1e-13                                # PTLMC.py:205-211
    assert np.array_equal(valid.astype(bool), np.all((thetap >= 0) & (thetap <= 1), axis=1))
    assert int(st['kstep'].item()) == 0
    # a second draw (next step) uses fresh normals
    ops.ptlmc_step(st, 3, st['fval'] * 0 - 1e300, st['dfval'])                  # hopeless proposals: nothing is accepted
    z2 = st['z'].cpu().numpy()
    assert not np.array_equal(z, z2) and int(st['kstep'].item()) == 1


def test_normals_have_unit_moments():
    from surmise_b200 import ops
    st, h = _state(264, 8, 8, tune=10 ** 6, nsave=0)
    import torch
    zs = []
    ll = torch.full((264,), -1e300, dtype=torch.float64, device='cuda')
    ops.ptlmc_step(st, 2)
    for _ in range(200):
        zs.append(st['z'].clone())
        ops.ptlmc_step(st, 3, ll, st['dfval'])
    z = torch.stack(zs).cpu().numpy().reshape(-1)
    n = z.size                                                                   # 422 400 draws
    report('sampler_normals', mean=float(z.mean()), var=float(z.var()), kurt=float(np.mean(z ** 4)))
    assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.var() - 1) < 5 * np.sqrt(2 / n) and abs(np.mean(z ** 4) - 3) < 0.1
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 5 / np.sqrt(n)


def test_accept_all_then_exchange_permutes_and_saves():
    """With an overwhelming likelihood every inside proposal is accepted; the exchange sweep then only permutes:
    the multiset of un-tempered log-posteriors is conserved, theta rows travel with them, draws are saved."""
    import torch
    from surmise_b200 import ops
    B, d, nt = 40, 3, 8
    st, h = _state(B, d, nt, tune=0, nsave=3)
    ops.ptlmc_step(st, 2)
    thetap = st['thetap'].cpu().numpy()
    valid = st['valid'].cpu().numpy().astype(bool)
    rng = np.random.default_rng(5)
    ll = 1e6 + 50 * rng.normal(size=B)                                            # accepted wherever inside the box
    dll = rng.normal(size=(B, d))
    ops.ptlmc_step(st, 1, torch.from_numpy(ll).cuda(), torch.from_numpy(dll).cuda())
    thetac, fval, dfval = st['thetac'].cpu().numpy(), st['fval'].cpu().numpy(), st['dfval'].cpu().numpy()
    before_lp = np.where(valid, ll, h['fval'] * h['temps'])                       # fn = fval * temps  (PTLMC.py:240)
    before_th = np.where(valid[:, None], thetap, h['thetac'])
    before_g = np.where(valid[:, None], dll, h['dfval'] * h['temps'][:, None])
    after_lp = fval * h['temps']
    order = np.array([int(np.argmin(np.abs(before_lp - v))) for v in after_lp])
    assert sorted(order.tolist()) == list(range(B))                               # a permutation
    assert np.max(np.abs(after_lp - before_lp[order]) / np.abs(before_lp[order])) < 1e-13
    assert np.max(np.abs(thetac - before_th[order])) == 0.0
    assert np.max(np.abs(dfval * h['temps'][:, None] - before_g[order])) < 1e-9
    scal = st['scal'].cpu().numpy()
    assert scal[2] == valid.sum() and scal[5] == (~valid).sum() and scal[3] > 0
    assert np.array_equal(st['save'].cpu().numpy()[:, 0, :], thetac[nt:])         # k = 0 >= tune: first kept draw
    assert int(st['kstep'].item()) == 1


class _GaussianTarget:
    """Stands in for directbayeswoodbury_B200.DeviceLogpost: a correlated Gaussian inside a generous box, evaluated
    with torch ops into fixed buffers (so the CUDA-graph path is exercised too)."""

    def __init__(self, mu, cov, lo, hi):
        import torch
        self.box = (np.asarray(lo, float), np.asarray(hi, float))
        self.mu = torch.from_numpy(mu).cuda()
        self.P = torch.from_numpy(np.linalg.inv(cov)).cuda()
        self.stats = None

        class prior:
            @staticmethod
            def lpdf(th):
                th = np.atleast_2d(th)
                return np.where(np.all((th >= lo) & (th <= hi), axis=1), 0.0, -np.inf).reshape(-1, 1)
        self.prior = prior

    def evaluator(self, B, d):
        import torch
        tgt = self

        class Ev:
            thetap = torch.zeros((B, d), dtype=torch.float64, device='cuda')
            valid = torch.ones(B, dtype=torch.int32, device='cuda')
            ll = torch.zeros(B, dtype=torch.float64, device='cuda')
            dll = torch.zeros((B, d), dtype=torch.float64, device='cuda')

            def launch(self):
                r = self.thetap - tgt.mu
                g = -(r @ tgt.P)
                self.dll.copy_(g)
                self.ll.copy_(0.5 * (r * g).sum(1))
        return Ev()

    def record(self, **stats):
        self.stats = stats


def test_device_chain_samples_a_known_gaussian():
    from surmise_b200.utilitiesmethods import PTLMC_B200 as S
    mu = np.array([0.3, -0.2, 0.5])
    cov = np.array([[0.04, 0.018, 0.0], [0.018, 0.02, 0.004], [0.0, 0.004, 0.01]])
    tgt = _GaussianTarget(mu, cov, -5 * np.ones(3), 5 * np.ones(3))
    P = np.linalg.inv(cov)

    def logpost(theta, return_grad=True):
        theta = np.atleast_2d(theta)
        r = theta - mu
        lp = (-0.5 * np.sum((r @ P) * r, 1)).reshape(-1, 1)
        lp[np.any(np.abs(theta) > 5, axis=1)] = -np.inf
        return (lp, -(r @ P)) if return_grad else lp
    logpost.b200 = tgt
    np.random.seed(3)
    out = S.sampler(logpost, lambda n: np.random.uniform(-1, 1, size=(n, 3)), numsamp=20000, numtemps=8, numchain=64,
                    sampperchain=600)
    assert tgt.stats is not None and tgt.stats['theta_evaluated_on_device'] == 72 * 900
    th = out['theta']
    e_mu = np.abs(th.mean(0) - mu) / np.sqrt(np.diag(cov))
    e_cov = np.abs(np.cov(th.T) - cov) / np.sqrt(np.outer(np.diag(cov), np.diag(cov)))
    report('device_chain_gaussian', mean_err_sd=float(e_mu.max()), cov_err=float(e_cov.max()),
           accept=tgt.stats['accept_rate'], swap=tgt.stats['swap_rate'], graph=float(tgt.stats['cuda_graph']))
    assert e_mu.max() < 0.1 and e_cov.max() < 0.1
    assert 0.3 < tgt.stats['accept_rate'] < 0.9                                   # tuned towards 0.6 (PTLMC.py:119)
    assert tgt.stats['cuda_graph']
    # the host loop of the same plugin agrees (same algorithm, numpy random numbers)
    np.random.seed(3)
    host = S.sampler(logpost, lambda n: np.random.uniform(-1, 1, size=(n, 3)), numsamp=20000, numtemps=8, numchain=64,
                     sampperchain=600, device_resident=False)['theta']
    assert np.max(np.abs(host.mean(0) - th.mean(0)) / np.sqrt(np.diag(cov))) < 0.15


def test_calibration_runs_device_resident_when_the_prior_declares_its_box():
    """C5 in miniature through the calibration plugin: same posterior as the host loop, no host closure calls inside
    the main loop."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    x, theta, f, truth = make_simulator(150, 2, 25, 4, seed=21)
    np.random.seed(21)
    fi = {'method': 'PCGPwM_B200'}
    M.fit(fi, x, theta, f)
    theta_true, y, yvar = make_observation(truth, 2, seed=22)
    calls = {'rows': 0}

    class prior:
        box = (np.zeros(2), np.ones(2))

        @staticmethod
        def lpdf(th):
            th = np.atleast_2d(th)
            calls['rows'] += th.shape[0]
            return np.where(np.all((th >= 0) & (th <= 1), axis=1), 0.0, -np.inf).reshape(-1, 1)

        @staticmethod
        def lpdf_grad(th):
            return np.zeros_like(np.atleast_2d(th))

        @staticmethod
        def rnd(n):
            return np.random.uniform(size=(n, 2))

    class Emu:
        _info = fi
        _emulator__theta = theta
    post = {}
    for mode in ('auto', False):
        cal = {'thetaprior': prior, 'yvar': yvar}
        calls['rows'] = 0
        np.random.seed(5)
        C.fit(cal, Emu, x, y, sampler='PTLMC_B200', numsamp=4000, numtemps=6, numchain=26, sampperchain=400,
              device_resident=mode)
        post[mode] = cal['thetarnd']
        if mode == 'auto':
            st = cal['_b200sampler']
            assert st['theta_evaluated_on_device'] == 32 * 600 and st['cuda_graph']
            assert calls['rows'] < 32 * 600                     # starts, pre-optimiser and the final log-posteriors only:
                                                                # the main loop (19 200 proposals) never called the host prior
            rows_dev = calls['rows']
        else:
            assert '_b200sampler' not in cal and calls['rows'] > 32 * 600
    err = np.abs(post['auto'].mean(0) - theta_true[0])
    sd = post[False].std(0)
    gap = np.abs(post['auto'].mean(0) - post[False].mean(0)) / sd
    report('calibration_device_resident', err0=float(err[0]), err1=float(err[1]), gap_sd=float(gap.max()),
           sd_ratio=float(np.max(post['auto'].std(0) / sd)), host_prior_rows=rows_dev)
    assert np.all(err < 0.1) and gap.max() < 0.5
    assert np.all(np.abs(post['auto'].std(0) / sd - 1) < 0.5)
