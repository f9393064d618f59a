"""Extended-precision arbiter for the fitted-emulator goldens (40 significant digits, mpmath).

Why: through the whole path (factorise -> predict -> Woodbury) the B200 results and the reference's differ by
3e-9 ... 3e-6 on the golden emulators, whose covariance matrices have condition numbers 2e7 ... 3e10.  The reference
inverts R through `eigh`, this repository through a Cholesky factor; both are backward stable, neither is "the truth".
This script evaluates the SAME formulas (PCGPwM.py:219-297 predict, directbayeswoodbury.py:261-306 loglik) in 40-digit
arithmetic from the SAME double-precision inputs (design, scores, fitted hyper-parameters, PCA maps, y, yvar) stored
in emu_<tag>.npz, so that a test can say which of the two is closer to the exact value:

    python tests/golden/make_golden_arbiter.py [tags...]      ->  tests/golden/arbiter.npz

Needs only numpy + mpmath and the emu_<tag>.npz files (not the reference).  Keys per tag: <tag>_predvecs,
<tag>_predvars, <tag>_loglik (the exact values rounded to double), <tag>_ref_err_{predvecs,predvars,loglik} (max
relative error of the REFERENCE's stored outputs against them), <tag>_fired.
"""
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
mp.mp.dps = 40
S0 = mp.mpf('0.01')                      # sig2ofconst, PCGPwM.py:705


def M(x):
    return mp.mpf(float(x))


def corr_row(t, theta, gam):
    """Separable Matern-3/2 correlation of one point with every design point (matern_covmat.pyx:27-64), exact."""
    d = len(t)
    ell = [mp.exp(M(gam[j])) for j in range(d)]
    c = 1 / (1 + mp.exp(M(gam[d])))
    out = []
    for i in range(theta.shape[0]):
        prod, ssum = mp.mpf(1), mp.mpf(0)
        for j in range(d):
            s = abs(M(t[j]) / ell[j] - M(theta[i, j]) / ell[j])
            prod *= (1 + s)
            ssum += s
        out.append(c * prod * mp.exp(-ssum) + (1 - c))
    return out


def forward(L, b):
    """Solve L y = b for lower-triangular L (list of mp rows)."""
    n = len(b)
    y = [mp.mpf(0)] * n
    for i in range(n):
        y[i] = (b[i] - mp.fdot(L[i][:i], y[:i])) / L[i][i]
    return y


def backward(L, y):
    """Solve L' x = y."""
    n = len(y)
    x = [mp.mpf(0)] * n
    for i in range(n - 1, -1, -1):
        x[i] = (y[i] - mp.fdot([L[r][i] for r in range(i + 1, n)], x[i + 1:])) / L[i][i]
    return x


def factor(theta, hyp, gvar):
    """Rows of the Cholesky factor of R = (1-nu) R0 + nu I + exp(h_v) diag(gvar)  (PCGPwM.py:810-815, 850-858)."""
    n, d = theta.shape
    nu = mp.exp(M(hyp[-1])) / (1 + mp.exp(M(hyp[-1])))
    ev = mp.exp(M(hyp[-2]))
    R = mp.matrix(n, n)
    for i in range(n):
        row = corr_row(theta[i], theta[:i + 1], hyp[:d + 1])
        for j in range(i + 1):
            R[i, j] = R[j, i] = (1 - nu) * row[j]
        R[i, i] = (1 - nu) * row[i] + nu + ev * M(gvar[i])
    Lm = mp.cholesky(R)
    return [[Lm[i, j] for j in range(i + 1)] for i in range(n)], nu


def arbiter(tag):
    g = dict(np.load(os.path.join(HERE, 'emu_%s.npz' % tag)))
    theta, hyps, pc = g['st_theta'], g['st_hyp'], g['st_pc']
    n, d = theta.shape
    k = hyps.shape[0]
    v = g['st_unscaled_pcstdvar']
    gvar = np.minimum(10, v / ((1 - v) ** 0.3))                   # PCGPwM.py:755 (double, an input to R on both sides)
    tn = g['thetanew']
    B = tn.shape[0]
    factors = {}
    pv = [[None] * k for _ in range(B)]
    pvar = [[None] * k for _ in range(B)]
    for j in range(k):
        key = (hyps[j].tobytes(), gvar[:, j].tobytes())
        if key not in factors:
            L, nu = factor(theta, hyps[j], gvar[:, j])
            rows = [[(1 - nu) * r for r in corr_row(tn[b], theta, hyps[j][:d + 1])] for b in range(B)]
            ts = [forward(L, rows[b]) for b in range(B)]
            quad = [mp.fdot(t, t) for t in ts]
            factors[key] = (L, rows, quad)
            print(tag, 'factor', len(factors), 'of component', j, flush=True)
        L, rows, quad = factors[key]
        gj = [M(a) for a in pc[:, j]]
        y = forward(L, gj)
        pw = backward(L, y)
        sig2 = (mp.fdot(y, y) + S0) / (n + S0)                    # PCGPwM.py:838-846
        for b in range(B):
            pv[b][j] = mp.fdot(rows[b], pw)                       # PCGPwM.py:236
            pvar[b][j] = sig2 * abs(1 - quad[b])                  # PCGPwM.py:270
    # back-projection and Woodbury log-likelihood (PCGPwM.py:273-297, directbayeswoodbury.py:261-306)
    m = g['st_pcti'].shape[0]
    phi = [[M(g['st_pcti'][i, j]) * M(g['st_scale'][i]) for j in range(k)] for i in range(m)]
    extrav = [M(a) for a in g['st_extravar']]
    fired = False
    for b in range(B):
        for i in range(m):
            s = mp.fsum(pvar[b][j] * phi[i][j] ** 2 for j in range(k))
            if abs((extrav[i] + s) / (mp.mpf('1e-4') + (1 + mp.mpf('1e-4')) * s)) > 1:
                fired = True
    obsvar = [M(a) for a in np.broadcast_to(np.squeeze(g['yvar']), (m,))]
    if fired:
        obsvar = [o + abs(e) for o, e in zip(obsvar, extrav)]      # dbw.py:280-284: mean over theta of a constant
    ll = []
    for b in range(B):
        mean = [mp.fdot(phi[i], pv[b]) + M(g['st_offset'][i]) for i in range(m)]
        sr = [(M(g['y'][i]) - mean[i]) / mp.sqrt(obsvar[i]) for i in range(m)]
        J = mp.matrix(m, k)
        for i in range(m):
            for j in range(k):
                J[i, j] = phi[i][j] * mp.sqrt(pvar[b][j]) / mp.sqrt(obsvar[i])
        A = mp.eye(k) + J.T * J
        J2 = J.T * mp.matrix(sr)
        J3 = mp.lu_solve(A, J2)
        resid = mp.fdot(sr, sr) - mp.fdot(list(J2), list(J3))
        ll.append(-resid / 2 - mp.log(mp.det(A)) / 2)
    out = {'predvecs': np.array([[float(a) for a in row] for row in pv]),
           'predvars': np.array([[float(a) for a in row] for row in pvar]),
           'loglik': np.array([float(a) for a in ll]).reshape(-1, 1), 'fired': np.array(fired)}
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.abs(b)))
    out['ref_err_predvecs'] = np.array(float(np.max(np.abs(g['predvecs'] - out['predvecs'])) / np.max(np.abs(out['predvecs']))))
    out['ref_err_predvars'] = np.array(rel(g['predvars'], out['predvars']))
    out['ref_err_loglik'] = np.array(rel(g['loglik'], out['loglik']))
    print(tag, 'reference vs exact: predvecs %.2e predvars %.2e loglik %.2e fired %s (stored %s) cond %.1e'
          % (out['ref_err_predvecs'], out['ref_err_predvars'], out['ref_err_loglik'], fired, bool(g['fired']),
             g['st_condR'].max()), flush=True)
    return out


def nll_arbiter():
    """Exact penalised NLL and gradient (PCGPwM.py:850-907) for every hyper-parameter point of nll.npz."""
    g = dict(np.load(os.path.join(HERE, 'nll.npz')))
    out = {}
    for tag in ('a', 'b', 'c'):
        theta, gg, gvar = g[tag + '_theta'], g[tag + '_g'], g[tag + '_gvar']
        mu, sd, hyps = g[tag + '_mean'], g[tag + '_std'], g[tag + '_hyps']
        n, d = theta.shape
        p = d + 3
        vals, grads = [], []
        for hyp in hyps:
            nu = mp.exp(M(hyp[-1])) / (1 + mp.exp(M(hyp[-1])))
            ev = mp.exp(M(hyp[-2]))
            ell = [mp.exp(M(hyp[j])) for j in range(d)]
            c = 1 / (1 + mp.exp(M(hyp[d])))
            R = mp.matrix(n, n)
            dR = [mp.matrix(n, n) for _ in range(p)]
            for a in range(n):
                for b in range(a + 1):
                    S = [abs(M(theta[a, j]) / ell[j] - M(theta[b, j]) / ell[j]) for j in range(d)]
                    pre = c * mp.fprod(1 + s for s in S) * mp.exp(-mp.fsum(S))
                    r0 = pre + (1 - c)
                    R[a, b] = R[b, a] = (1 - nu) * r0 + (nu + ev * M(gvar[a]) if a == b else 0)
                    for j in range(d):
                        dR[j][a, b] = dR[j][b, a] = (1 - nu) * pre * S[j] ** 2 / (1 + S[j])
                    dR[d][a, b] = dR[d][b, a] = (1 - nu) * (1 - c) * (c - pre)
                    dR[d + 1][a, b] = dR[d + 1][b, a] = ev * M(gvar[a]) if a == b else mp.mpf(0)
                    dR[d + 2][a, b] = dR[d + 2][b, a] = nu * (1 - nu) * ((1 if a == b else 0) - r0)
            Lm = mp.cholesky(R)
            L = [[Lm[i, j] for j in range(i + 1)] for i in range(n)]
            y = forward(L, [M(v) for v in gg])
            alpha = backward(L, y)
            sig2 = (mp.fdot(y, y) + S0) / (n + S0)
            pen = [(mp.mpf('1e-8') + M(hyp[j]) - M(mu[j])) / M(sd[j]) for j in range(p)]
            nll = mp.fsum(mp.log(L[i][i]) for i in range(n)) + n * mp.log(sig2) / 2 + mp.fsum(q * q for q in pen) / 2
            Rinv = mp.matrix(n, n)
            for i in range(n):
                e = [mp.mpf(1 if r == i else 0) for r in range(n)]
                col = backward(L, forward(L, e))
                for r in range(n):
                    Rinv[r, i] = col[r]
            grad = []
            av = mp.matrix(alpha)
            for j in range(p):
                tr = mp.fsum(Rinv[a, b] * dR[j][a, b] for a in range(n) for b in range(n))
                quad = (av.T * dR[j] * av)[0, 0]
                grad.append(tr / 2 - n * quad / (2 * (n + S0) * sig2) + pen[j] / M(sd[j]))
            vals.append(float(nll))
            grads.append([float(q) for q in grad])
            print('nll', tag, len(vals), flush=True)
        vals, grads = np.array(vals), np.array(grads)
        out['nll_%s' % tag], out['nllgrad_%s' % tag] = vals, grads
        out['nll_%s_ref_err' % tag] = np.abs(g[tag + '_nll'] - vals) / (np.abs(vals) + n)
        out['nllgrad_%s_ref_err' % tag] = np.max(np.abs(g[tag + '_grad'] - grads), 1) / np.max(np.abs(grads), 1)
        print('nll', tag, 'reference vs exact: nll', out['nll_%s_ref_err' % tag], 'grad', out['nllgrad_%s_ref_err' % tag])
    return out


if __name__ == '__main__':
    if sys.argv[1:] == ['nll']:
        path = os.path.join(HERE, 'arbiter.npz')
        store = dict(np.load(path)) if os.path.exists(path) else {}
        store.update(nll_arbiter())
        np.savez_compressed(path, **store)
        sys.exit(0)
    tags = sys.argv[1:] or ['d1', 'trig', 'misspd', 'c1']
    path = os.path.join(HERE, 'arbiter.npz')
    store = dict(np.load(path)) if os.path.exists(path) else {}
    for tag in tags:
        for key, val in arbiter(tag).items():
            store['%s_%s' % (tag, key)] = val
        np.savez_compressed(path, **store)
