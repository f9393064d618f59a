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
