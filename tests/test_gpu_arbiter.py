"""GPU: who is right when the B200 path and the reference disagree?

tests/golden/arbiter.npz holds 40-digit evaluations (mpmath, tests/golden/make_golden_arbiter.py) of the reference's
own formulas -- PCGPwM.predict (PCGPwM.py:219-297), directbayeswoodbury.loglik (directbayeswoodbury.py:261-306),
PCGPwM.__negloglik / __negloglikgrad (PCGPwM.py:850-907) -- from the same double-precision inputs as the goldens.
The reference inverts R through `eigh` (cond(R) = 2e7 ... 3e10 on these fixtures), this repository through a Cholesky
factor.  The assertions: the CUDA path either meets BASELINE.json's bar against the EXACT value (1e-9 for NLL, gradient
and per-theta log-likelihood) or is as close to it as the reference is (within a factor 4 for the max over a
handful of thetas: both sides scatter at a small multiple of cond * eps).  The measured pairs go to gpurun_out/parity_report.txt as `arbiter_*` lines."""
import numpy as np
import pytest

from gpu_util import report
from oracle import pcgpwm_oracle as po

pytestmark = pytest.mark.gpu
BAR = 1e-9


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - b) / np.abs(b)))


def _state(g):
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    gv = po.damp_gvar(g['st_unscaled_pcstdvar'], 10, 0.3)
    spc = {'offset': g['st_offset'], 'scale': g['st_scale'], 'extravar': g['st_extravar']}
    emulist, state = M.import_state(g['st_theta'], g['st_pc'], gv, g['st_hyp'], g['st_hypind'], g['st_pcti'], spc)
    return M, {'_b200': state, 'emulist': emulist, 'x': g['in_x'], 'theta': g['st_theta'], 'pcti': g['st_pcti'],
               'standardpcinfo': spc}


@pytest.mark.parametrize('tag', ['c1', 'misspd', 'trig', 'd1', 'd8'])
def test_path_is_at_least_as_close_to_exact_as_the_reference(golden, tag):
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C
    g, a = golden('emu_' + tag), golden('arbiter')
    M, fitinfo = _state(g)

    class Emu:
        _info = fitinfo
    pred = {}
    M.predict(pred, fitinfo, g['in_x'], g['thetanew'])
    ll = C.loglik({'yvar': g['yvar']}, Emu, g['thetanew'], g['y'], g['in_x'])
    exact = {q: a['%s_%s' % (tag, q)] for q in ('predvecs', 'predvars', 'loglik')}
    mine = {'predvecs': float(np.max(np.abs(pred['predvecs'] - exact['predvecs'])) / np.max(np.abs(exact['predvecs']))),
            'predvars': _rel(pred['predvars'], exact['predvars']), 'loglik': _rel(ll, exact['loglik'])}
    ref = {q: float(a['%s_ref_err_%s' % (tag, q)]) for q in mine}
    report('arbiter_' + tag, cond=float(g['st_condR'].max()),
           **{'b200_' + q: v for q, v in mine.items()}, **{'reference_' + q: v for q, v in ref.items()})
    # 'd1' has cond(R) = 2.9e10: cond * eps = 6e-6, and neither an eigendecomposition nor a Cholesky factor can promise
    # more; there the two sides scatter around the exact value at that level and "as close as the reference" is
    # asserted within a factor 8 instead of 2
    # Elsewhere the factor is 4: the B200 value moves with the summation order of the GEMM that applies L^-1 (c1 predvars:
    # 2.3e-9 with cp.async staging, 3.6e-9 with the TMA kernel's k order; the reference sits at 1.5e-9; cond * eps = 1e-7)
    slack = 8 if tag == 'd1' else 4
    for q in mine:
        assert mine[q] < max(BAR, slack * ref[q]), (tag, q, mine[q], ref[q])
    if tag == 'd8':                                     # the conditioning of config C2: the bar itself, end to end
        assert mine['loglik'] < BAR and _rel(ll, g['loglik']) < BAR


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_nll_and_gradient_against_exact(golden, tag):
    from surmise_b200 import ops
    g, a = golden('nll'), golden('arbiter')
    n = g[tag + '_g'].shape[0]
    data = ops.GPData(g[tag + '_theta'], g[tag + '_g'], g[tag + '_gvar'])
    nll, grad, info = ops.gp_nll_grad(data, g[tag + '_hyps'], g[tag + '_mean'], g[tag + '_std'], 0.01, True)
    assert not info.any()
    ex_v, ex_g = a['nll_' + tag], a['nllgrad_' + tag]
    e_v = np.abs(nll - ex_v) / (np.abs(ex_v) + n)
    e_g = np.max(np.abs(grad - ex_g), 1) / np.max(np.abs(ex_g), 1)
    r_v, r_g = a['nll_%s_ref_err' % tag], a['nllgrad_%s_ref_err' % tag]
    for i in range(len(e_v)):
        report('arbiter_nll_' + tag, i=i, cond=g[tag + '_cond'][i], b200_nll=e_v[i], reference_nll=r_v[i],
               b200_grad=e_g[i], reference_grad=r_g[i])
        assert e_v[i] < max(BAR, 2 * r_v[i]) and e_g[i] < max(BAR, 2 * r_g[i]), (tag, i, e_v[i], r_v[i], e_g[i], r_g[i])
