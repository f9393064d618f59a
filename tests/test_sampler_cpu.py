"""CPU: the PTLMC_B200 sampler plugin (host logic + the compiled exchange helper) on analytic targets."""
import numpy as np
import pytest


def _py_tempexchange(lpostf, temps, rtv, logu):
    """Loop form of surmise's tempexchange (utilitiesmethods/PTLMC.py:259-273) with the draws made explicit."""
    B = lpostf.shape[0]
    order = np.arange(B)
    it = 0
    for rt in rtv:
        rhoh = 1 / temps[rt - 1] - 1 / temps[rt]
        if (lpostf[order[rt]] - lpostf[order[rt - 1]]) * rhoh > logu[it]:
            order[rt - 1], order[rt] = order[rt], order[rt - 1]
        it += 1
    return order


def test_tempexchange_helper_matches_loop_form():
    from surmise_b200.utilitiesmethods import PTLMC_B200 as S
    rng = np.random.default_rng(0)
    for B, nt in ((48, 32), (264, 8), (5, 3)):
        temps = S.temperature_ladder(nt, B - nt, 30.0)[:, 0]
        lp = rng.normal(size=B) * 20
        rtv = rng.integers(1, B, size=5 * B)
        logu = np.log(rng.uniform(size=5 * B))
        got = S.tempexchange(lp[:, None], temps[:, None], rtv, logu)
        assert np.array_equal(got, _py_tempexchange(lp, temps, rtv, logu))
        assert sorted(got.tolist()) == list(range(B))


def test_temperature_ladder_matches_reference_formula():
    from surmise_b200.utilitiesmethods import PTLMC_B200 as S
    t = S.temperature_ladder(8, 256, 30)
    assert t.shape == (264, 1) and np.all(t[8:] == 1)
    assert np.isclose(t[0, 0], 30) and np.isclose(t[7, 0], 30 ** (1 / 9))


@pytest.mark.parametrize('with_grad', [True, False])
def test_sampler_recovers_a_correlated_gaussian(with_grad):
    from surmise_b200.utilitiesmethods import PTLMC_B200 as S
    mu = np.array([0.3, -0.2, 0.5])
    C = np.array([[0.04, 0.018, 0.0], [0.018, 0.02, 0.004], [0.0, 0.004, 0.01]])
    Ci = np.linalg.inv(C)
    calls = {'n': 0, 'rows': 0}

    def logpost(theta, return_grad=True):
        theta = np.atleast_2d(theta)
        calls['n'] += 1
        calls['rows'] += theta.shape[0]
        r = theta - mu
        lp = -0.5 * np.einsum('bi,ij,bj->b', r, Ci, r).reshape(-1, 1)
        if with_grad and return_grad:
            return lp, -r @ Ci
        return lp
    np.random.seed(4)
    out = S.sampler(logpost, lambda n: np.random.uniform(-1, 1, size=(n, 3)), numsamp=4000, numtemps=8, numchain=24,
                    sampperchain=300, maxtemp=20)
    th = out['theta']
    assert th.shape == (4000, 3) and out['logpost'].shape[0] == 4000
    assert np.all(np.abs(th.mean(0) - mu) < 0.03)
    assert np.all(np.abs(np.cov(th.T) - C) < 0.012)
    assert calls['rows'] / calls['n'] > 16              # the closure sees batches, not single thetas


def test_dispatcher_contract():
    from surmise_b200.utilities import sampler
    with pytest.raises(ValueError):
        sampler(lambda t: t, lambda n: None, sampler='does_not_exist')


# ---- metropolis_hastings_B200 / LMC_B200 against chains drawn by the reference samplers (tests/golden/samplers.npz)
def _golden_samplers():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'samplers.npz'))


@pytest.mark.parametrize('step', ['normal', 'uniform'])
def test_mh_single_chain_reproduces_reference_chain(step):
    from surmise_b200.utilitiesmethods import metropolis_hastings_B200 as S
    from synth import boxed_logpost, box_draws
    gold = _golden_samplers()
    np.random.seed(11)
    out = S.sampler(boxed_logpost, box_draws, numsamp=400, stepType=step, burnSamples=150)
    assert np.array_equal(out['theta'], gold['mh_%s_theta' % step])
    assert out['acc_rate'] == float(gold['mh_%s_acc' % step])
    assert np.array_equal(out['lpostlist'], gold['mh_%s_lpostlist' % step])


def test_mh_population_batches_every_step_and_recovers_target():
    from surmise_b200.utilitiesmethods import metropolis_hastings_B200 as S
    from synth import boxed_logpost, box_draws, _MU
    rows = []

    def logpost(theta, return_grad=False):
        rows.append(theta.shape[0])
        return boxed_logpost(theta)
    np.random.seed(5)
    out = S.sampler(logpost, box_draws, numsamp=6000, numchain=64, burnSamples=400,
                    stepParam=np.array([0.15, 0.1, 0.08]))
    assert out['theta'].shape == (6000, 3) and set(rows) == {64}
    assert len(rows) == 400 + (-(-6000 // 64))
    assert 0.1 < out['acc_rate'] < 0.9
    assert np.all(np.abs(out['theta'].mean(0) - _MU) < 0.05)
    assert np.all(np.abs(out['theta'].std(0) - np.array([0.2, 0.1414, 0.1])) < 0.04)


@pytest.mark.parametrize('with_grad', [True, False])
def test_lmc_reproduces_reference_draws(with_grad):
    from surmise_b200.utilitiesmethods import LMC_B200 as S
    from synth import gaussian_logpost, box_draws
    gold = _golden_samplers()
    rows = []

    def logpost(theta, return_grad=True):
        rows.append(np.atleast_2d(theta).shape[0])
        if with_grad:
            return gaussian_logpost(theta, return_grad=return_grad)
        return gaussian_logpost(theta, return_grad=False)
    np.random.seed(12 if with_grad else 13)
    out = S.sampler(logpost, box_draws, numsamp=500)
    want = gold['lmc_grad_theta' if with_grad else 'lmc_nograd_theta']
    assert out['theta'].shape == want.shape
    assert np.allclose(out['theta'], want, rtol=0, atol=1e-12)
    # the polishing stage is batched: ten optimisers per call until the first of them converges
    assert 10 in rows and rows.count(1) < rows.count(10)


def test_ptlmc_preoptimiser_without_threads_equals_threaded_version(monkeypatch):
    """_preoptimise through the reverse-communication L-BFGS-B core (generators, one scheduler) gives the same
    starting points as one scipy.optimize.minimize per thread."""
    from surmise_b200.utilitiesmethods import PTLMC_B200 as S
    from surmise_b200.emulationmethods.PCGPwM_B200 import _LbfgsbCore
    from synth import gaussian_logpost
    if not _LbfgsbCore.available():
        pytest.skip('scipy core not usable')
    rows = []

    def vg(th):
        rows.append(th.shape[0])
        return gaussian_logpost(th)
    theta0 = np.random.default_rng(3).uniform(-1, 1, size=(24, 3))
    np.random.seed(9)
    a = S._preoptimise(theta0.copy(), vg, True)
    batched = list(rows)
    monkeypatch.setattr(_LbfgsbCore, 'available', classmethod(lambda cls: False))
    np.random.seed(9)
    b = S._preoptimise(theta0.copy(), vg, True)
    assert np.array_equal(a, b)
    assert max(batched) == 24 and np.mean(batched) > 8
