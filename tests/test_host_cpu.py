"""CPU: host-side logic, the C-ABI surface, and the no-fallback guarantees (no compute calls)."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'surmise_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sb200_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    import ctypes
    import surmise_b200._lib as L
    names = _declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(L.SIGNATURES) == names            # the ctypes table binds exactly the header's surface
    assert L.lib.sb200_version() >= 100
    assert L.lib.sb200_launch_count() == 0


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: the header must compile as C99 (no C++ or torch types in the signatures) and a
    C program must be able to load the library and call it (a cgo / JNI / ctypes stub does nothing more)."""
    import shutil
    import subprocess
    import surmise_b200._lib as L
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('gcc not available')
    hdr = os.path.join(ROOT, 'include', 'surmise_b200.h')
    subprocess.check_call([gcc, '-std=c99', '-Wall', '-Werror', '-fsyntax-only', '-x', 'c', hdr])
    src = tmp_path / 'abi.c'
    src.write_text(
        '#include <stdio.h>\n#include <dlfcn.h>\n#include "surmise_b200.h"\n'
        'int main(int argc, char** argv) {\n'
        '  void* h = dlopen(argv[1], RTLD_NOW);\n'
        '  if (!h) { fprintf(stderr, "%s\\n", dlerror()); return 2; }\n'
        '  int (*ver)(void) = (int (*)(void))dlsym(h, "sb200_version");\n'
        '  const char* (*err)(void) = (const char* (*)(void))dlsym(h, "sb200_last_error");\n'
        '  if (!ver || !err) return 3;\n'
        '  printf("%d\\n", ver());\n'
        '  return 0;\n}\n')
    exe = tmp_path / 'abi'
    subprocess.check_call([gcc, '-std=c99', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe), '-ldl'])
    out = subprocess.check_output([str(exe), L.LIB_PATH], text=True)
    assert int(out.strip()) >= 100


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    import surmise_b200
    from surmise_b200._lib import SurmiseB200Error
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    with pytest.raises(SurmiseB200Error):
        surmise_b200.covmat(np.zeros((3, 2)), np.zeros((3, 2)), np.zeros(3))
    with pytest.raises(SurmiseB200Error):
        M.fit({}, np.arange(5.0)[:, None], np.random.rand(20, 2), np.random.rand(5, 20))


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'surmise_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M) or 'oracle/' in src:
                    bad.append(f)
    assert not bad, bad


def test_pca_matches_reference_internals(golden):
    """Batched imputation/PCA (_pca.py) vs the reference's per-row loops (PCGPwM.py:533-654).

    The 40 ill-conditioned ridge iterations amplify summation-order differences (the reference's own
    GEMM-vs-SYRK rounding moves fs by 2e-5), hence the tolerances."""
    from surmise_b200 import _pca
    g = golden('standardize')
    spc, mof, mofrows = _pca.standardize(g['f'].T, 0.001, 10e-6)
    assert np.allclose(spc['offset'], g['offset'], rtol=1e-13) and np.allclose(spc['scale'], g['scale'], rtol=1e-13)
    pcs = _pca.principal_components(spc['fs'], spc['U'], spc['S'], mof, mofrows, 0.001, 10e-6)
    assert pcs['pc'].shape == g['pc'].shape
    assert np.max(np.abs(spc['fs'] - g['fs'])) < 1e-3
    assert np.max(np.abs(spc['extravar'] - g['extravar'])) / np.max(g['extravar']) < 1e-3
    sign = np.sign(np.sum(pcs['pc'] * g['pc'], 0))               # singular-vector sign is arbitrary
    assert np.max(np.abs(pcs['pc'] * sign - g['pc'])) / np.max(np.abs(g['pc'])) < 1e-3
    assert np.max(np.abs(pcs['unscaled_pcstdvar'] - g['unscaled_pcstdvar'])) < 1e-2


def test_pca_complete_data_is_exact(golden):
    from surmise_b200 import _pca
    g = golden('emu_c1')
    spc, mof, mofrows = _pca.standardize(g['in_f'].T, 0.001, 10e-6)
    assert mof is None and mofrows is None
    pcs = _pca.principal_components(spc['fs'], spc['U'], spc['S'], None, None, 0.001, 10e-6)
    assert np.allclose(pcs['pc'], g['st_pc'], rtol=0, atol=1e-12)
    assert np.allclose(pcs['pcti'], g['st_pcti'], rtol=0, atol=1e-12)
    assert np.allclose(spc['extravar'], g['st_extravar'], rtol=1e-10, atol=1e-18)


def test_hyper_prior_and_row_matching(golden):
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    from oracle import pcgpwm_oracle as po
    g = golden('emu_c1')
    got = M.hyper_prior(g['st_theta'], -10, -20, None)
    want = po.hyper_regulariser(g['st_theta'], -10, -20)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    got = M.hyper_prior(g['st_theta'], -8, -16, np.log(0.5))
    want = po.hyper_regulariser(g['st_theta'], -8, -16, np.log(0.5))
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    x = g['in_x']
    assert np.array_equal(M.match_rows(x, x)[0], np.arange(x.shape[0]))
    assert np.array_equal(M.match_rows(None, x)[0], np.arange(x.shape[0]))
    xind, xnew = M.match_rows(np.vstack([x[[4, 2]], [[-1.0]]]), x)
    assert list(xind) == [4, 2] and list(xnew) == [0, 1]
    xo = np.array([['a', 1.0], ['b', 2.0], ['c', 3.0]], dtype=object)          # object x as in surmise's tests
    xind, xnew = M.match_rows(xo[[2, 0]], xo)
    assert list(xind) == [2, 0]


def test_register_into_surmise_if_installed():
    """With the real surmise importable (build container: PYTHONPATH=/tmp/ref/src) the plugin lookup works."""
    surmise = pytest.importorskip('surmise')
    import importlib
    import surmise_b200
    names = surmise_b200.register()
    assert 'PCGPwM_B200' in names
    m = importlib.import_module('surmise.emulationmethods.PCGPwM_B200')
    assert hasattr(m, 'fit') and hasattr(m, 'predict')
    smp = importlib.import_module('surmise.utilitiesmethods.PTLMC_B200')
    assert hasattr(smp, 'sampler') and 'PTLMC_B200' in names
    c = importlib.import_module('surmise.calibrationmethods.directbayeswoodbury_B200')
    for fn in ('fit', 'predict', 'thetarnd', 'thetalpdf', 'loglik', 'loglik_grad'):
        assert hasattr(c, fn)
    import torch
    if not torch.cuda.is_available():
        from surmise.emulation import emulator
        x = np.arange(6.0)[:, None]
        theta = np.random.default_rng(0).uniform(size=(30, 2))
        f = np.sin(x + theta.sum(1))
        with pytest.raises(Exception) as e:
            emulator(x, theta, f, method='PCGPwM_B200')
        assert 'CUDA' in str(e.value) or 'cuda' in str(e.value)


@pytest.mark.parametrize('tag', ['c1', 'misspd', 'trig', 'd1'])
def test_predictlpdf_matches_reference(golden, tag):
    """PCGPwM_B200.predictlpdf (host, batched) on the reference's own predict output vs the reference's value
    (PCGPwM.py:331-364); the oracle's loop form is checked against the same fixture."""
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    from oracle import pcgpwm_oracle as po
    g = golden('emu_' + tag)
    phi = g['st_pcti'] * g['st_scale'][:, None]
    pi = {'mean': g['mean'], 'extravar': g['st_extravar'], 'phi': phi, 'predvars': g['predvars'],
          'mean_gradtheta': g['mean_grad'], 'predvars_gradtheta': g['predvars_grad'], 'return_grad': True}
    for impl in (M.predictlpdf, po.predictlpdf):
        lp, dlp = impl(pi, g['lpdf_f'], addvar=0.003)
        assert lp.shape == g['lpdf'].shape and dlp.shape == g['dlpdf'].shape
        # the value is a difference of O(m) terms (sum rf^2, log-determinants): scale the bar accordingly
        assert np.max(np.abs(lp - g['lpdf']) / (np.abs(g['lpdf']) + g['mean'].shape[0])) < 1e-11
        assert np.max(np.abs(dlp - g['dlpdf'])) / np.max(np.abs(g['dlpdf'])) < 1e-9
    pi['return_grad'] = False
    assert np.allclose(M.predictlpdf(pi, g['lpdf_f'], addvar=0.003), g['lpdf'], rtol=1e-9, atol=1e-9)


def test_supplementtheta_matches_reference_outputs():
    """PCGPwM_B200.supplementtheta against the reference's function (tests/golden/supplement.npz)."""
    import os
    from surmise_b200.emulationmethods import PCGPwM_B200 as M
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'supplement.npz'))
    fitinfo = {'theta': g['theta']}
    for name, kw in (('plain', {}), ('costs', {}),
                     ('pending', {'pending': g['pending'], 'includepending': True, 'costpending': 0.3})):
        th, info = M.supplementtheta(fitinfo, 4, g['new'], g['choices'].copy(), g[name + '_costs'].copy(), None, **kw)
        assert np.array_equal(th, g[name + '_theta'])
        assert np.array_equal(info['crit'], g[name + '_crit'])
        assert np.array_equal(np.signbit(info['crit']), np.signbit(g[name + '_crit']))
        if name == 'pending':
            assert np.array_equal(info['obviatesugg'], g[name + '_obviate'])
    with pytest.raises(ValueError):
        M.supplementtheta(fitinfo, 4, None, g['choices'], np.ones(9), None)


def test_sibling_methods_host_stages_match_reference_fixtures():
    """Host-side stages of the sibling method modules against arrays produced by the reference (tests/golden):
    PCSK's simsd-driven component selection, PCGPwImpute's completion, indGP's hyper-parameter layout."""
    import os
    from surmise_b200.emulationmethods import PCSK_B200, PCGPwImpute_B200, indGP_B200
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    g = np.load(os.path.join(here, 'pcsk.npz'))
    fi = {'epsilonPC': 0.001, 'numpcs': -1, 'standardpcinfo': PCSK_B200.standardize(g['f'].T)}
    PCSK_B200.principal_components(fi, g['simsd'])
    assert fi['pc'].shape == g['pc'].shape
    sign = np.sign(np.sum(fi['pc'] * g['pc'], 0))
    assert np.allclose(fi['pc'] * sign, g['pc'], rtol=0, atol=1e-10 * np.abs(g['pc']).max())
    assert np.allclose(fi['unscaled_pcstdvar'], g['pcstdvar'], rtol=1e-9, atol=0)
    mean, lo, hi, std = PCSK_B200.hyper_prior(g['theta'])
    assert mean.shape == (g['theta'].shape[1] + 3,) and mean[-1] == -17 and np.all(std == (hi - lo) / 4)

    gi = np.load(os.path.join(here, 'impute.npz'))
    for tag, method in (('knn', 'KNN'), ('br', 'BayesianRidge')):
        fin = gi[tag + '_f'].T.copy()
        assert np.isnan(fin).any()
        done = PCGPwImpute_B200._complete(fin, method)
        assert np.allclose(done, gi[tag + '_fcompleted'], rtol=1e-10, atol=1e-12)
    with pytest.raises(ValueError):
        PCGPwImpute_B200._complete(fin, 'nope')

    mean, lo, hi, std = indGP_B200.hyper_prior(g['theta'], -10, -20)
    assert mean.shape == (g['theta'].shape[1] + 2,) and std[-2] == 2 and std[-1] == 4 and mean[-1] == -10
    wide = indGP_B200._widen(np.array([[1.0, 2.0, 3.0, 4.0]]), -1e-8)
    assert np.array_equal(wide, [[1.0, 2.0, 3.0, -1e-8, 4.0]])


def test_reverse_communication_lbfgsb_equals_scipy_minimize():
    """_LbfgsbCore / minimize_many (the parallel fit's optimiser driver) against scipy.optimize.minimize: same
    iterates, values and evaluation counts, problem by problem, also when the problems advance together."""
    import scipy.optimize as spo
    from surmise_b200.emulationmethods.PCGPwM_B200 import _LbfgsbCore, minimize_many
    if not _LbfgsbCore.available():
        pytest.skip("this scipy does not expose the L-BFGS-B core with the expected signature")
    rng = np.random.default_rng(0)
    centres = rng.uniform(-0.5, 0.5, size=(7, 4))

    def fg(c, x):
        f = 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2 + 0.5 * np.sum((x[2:] - c[2:]) ** 4) + np.sum(c[:2] * x[:2])
        g = np.zeros_like(x)
        g[0] = -400 * x[0] * (x[1] - x[0] ** 2) - 2 * (1 - x[0]) + c[0]
        g[1] = 200 * (x[1] - x[0] ** 2) + c[1]
        g[2:] = 2 * (x[2:] - c[2:]) ** 3
        return f, g
    lo, hi = -2 * np.ones(4), np.array([0.8, 2.0, 2.0, 2.0])
    x0s = rng.uniform(-1.5, 0.7, size=(7, 4))
    calls = []

    def batch(idx, X):
        calls.append(len(idx))
        out = [fg(centres[i], x) for i, x in zip(idx, X)]
        return np.array([o[0] for o in out]), np.array([o[1] for o in out])
    for gtol in (0.1, 1e-6):
        X, F, nfev = minimize_many(batch, x0s, lo, hi, gtol)
        for i in range(7):
            r = spo.minimize(lambda x: fg(centres[i], x), x0s[i], method='L-BFGS-B', jac=True,
                             bounds=spo.Bounds(lo, hi), options={'gtol': gtol})
            assert np.array_equal(r.x, X[i]) and r.fun == F[i] and r.nfev == nfev[i]
    assert max(calls) == 7 and min(calls) >= 1


def test_lazy_standardpcinfo_dict(monkeypatch):
    """LazyHostDict (the standardpcinfo of the device PCA): deferred entries are copied to the host on first read, by
    every way of reading a mapping, and the pickled form is a plain dict of numpy arrays."""
    import copy
    import pickle
    import types
    import surmise_b200
    from surmise_b200 import _pca_device as P
    calls = []
    fake = types.SimpleNamespace(to_host=lambda ts: [calls.append(len(ts)) or np.asarray(t) for t in ts])
    monkeypatch.setitem(sys.modules, 'surmise_b200.ops', fake)
    monkeypatch.setattr(surmise_b200, 'ops', fake, raising=False)

    def fresh():
        d = P.LazyHostDict({'offset': np.zeros(2)})
        d.defer('fs', [[1.0, 2.0]])
        return d
    d = fresh()
    assert 'fs' in d and len(d) == 2 and not calls and d.device('fs') is not None
    assert isinstance(d['fs'], np.ndarray) and calls == [1] and d.device('fs') is None
    assert d['fs'] is d['fs'] and calls == [1]                       # copied once
    for read in (lambda m: dict(m)['fs'], lambda m: {**m}['fs'], lambda m: m.get('fs'), lambda m: dict(m.items())['fs'],
                 lambda m: list(m.values())[1], lambda m: m.copy()['fs'], lambda m: copy.deepcopy(m)['fs'],
                 lambda m: pickle.loads(pickle.dumps(m))['fs'], lambda m: m.pop('fs')):
        assert np.array_equal(read(fresh()), [[1.0, 2.0]])
    assert type(pickle.loads(pickle.dumps(fresh()))) is dict
    d = fresh()
    d['fs'] = 3
    assert d['fs'] == 3 and d.device('fs') is None


def test_calibrator_chooses_its_path_by_the_emulators_state():
    """directbayeswoodbury_B200: the fused PC-space path for an emulator with device factors or one that can be adopted
    (a reference PCGPwM fitinfo), the generic path (that emulator's own predict, likelihood of dbw.py:261-306 on the
    device, no gradient) for anything else -- decided from emu._info alone, no CUDA needed."""
    from surmise_b200.calibrationmethods import directbayeswoodbury_B200 as C

    class Emu:
        def __init__(self, info):
            self._info = info
    assert C._pc_family(Emu({'_b200': {'phi': object()}}))
    assert C._pc_family(Emu({'emulist': [], 'unscaled_pcstdvar': np.zeros((2, 1))}))          # reference PCGPwM: adopt()
    assert not C._pc_family(Emu({'_b200': {'theta': object()}}))                               # e.g. indGP_B200
    assert not C._pc_family(Emu({'method': 'PCGP', 'emulist': []}))
    assert not C._pc_family(Emu(None)) and not C._pc_family(object())
    # the gradient form is refused for the generic path before anything touches the device
    with pytest.raises(ValueError, match='gradient'):
        C.loglik_grad({'yvar': np.ones(3)}, Emu({'method': 'PCGP'}), np.zeros((2, 2)), np.zeros(3), None)

