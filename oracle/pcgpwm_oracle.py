"""CPU oracle for the PCGPwM hot path (TEST INFRASTRUCTURE -- never imported by the product).

This is a numpy restatement of the algorithm that bandframework/surmise implements in
  * src/surmise/emulationsupport/matern_covmat.pyx   (Matern-3/2 correlation + gradients)
  * src/surmise/emulationmethods/PCGPwM.py           (standardise, PCA, per-PC GP fit, predict)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it, and only as the checker or the timed CPU arm.

Parity pinning: the reference's own test-suite holds no golden vectors for this path
(SURVEY.md section 4), so this oracle is pinned against outputs of the reference itself, run
in the build container by tests/golden/make_golden.py; the resulting fixtures are committed
under tests/golden/ and checked by tests/test_oracle_golden.py.

Two formulations of the likelihood are provided on purpose:
  * `*_eigh`  follows the reference operation by operation (symmetric eigendecomposition) and
               is what the CPU baseline times;
  * `*_chol`  is the Cholesky / explicit-inverse form that the CUDA kernels implement; the
               tests check chol == eigh == reference so that a GPU mismatch can be localised.
"""
import numpy as np
import scipy.linalg as sla
import scipy.optimize as sopt

SIG2_OF_CONST = 0.01          # PCGPwM.py:705,719  (sig2ofconst passed by __fitGPs)
PRIOR_SHIFT = 1e-8            # PCGPwM.py:866,905  (10 ** (-8) added inside the regulariser)


# --------------------------------------------------------------------------------------
# Matern covariance                                   matern_covmat.pyx:27-64 (covmat), 13-24 (map4)
# --------------------------------------------------------------------------------------
def matern_covmat(x1, x2, gammav, grad=None):
    """Separable Matern-3/2 correlation with scale mixing.

    R = c * prod_j (1 + S_j) * exp(-sum_j S_j) + (1 - c),  S_j = |x1_j - x2_j| / exp(gamma_j),
    c = 1 / (1 + exp(gamma_d)).                              (matern_covmat.pyx:29-48, 60)
    grad='hyp' -> dR/dgamma, shape (n1, n2, d+1)             (matern_covmat.pyx:49-57)
    grad='x1'  -> dR/dx1,    shape (n1, n2, d)               (matern_covmat.pyx:13-24, 58-59)
    """
    gammav = np.asarray(gammav, dtype=np.float64)
    d = gammav.shape[0] - 1
    ell = np.exp(gammav[:d])
    a = (np.reshape(x1, (1, d)) if np.ndim(x1) < 2 else np.asarray(x1)) / ell
    b = (np.reshape(x2, (1, d)) if np.ndim(x2) < 2 else np.asarray(x2)) / ell
    c = 1.0 / (1.0 + np.exp(gammav[d]))
    diff = a[:, None, :] - b[None, :, :]                      # (n1, n2, d) scaled differences
    S = np.abs(diff)
    # product accumulated dimension by dimension in the reference's order (pyx:43-48)
    prod = np.full(S.shape[:2], c)
    ssum = np.zeros(S.shape[:2])
    for j in range(d):
        prod = prod * (1.0 + S[:, :, j])
        ssum = ssum - S[:, :, j]
    Rpre = prod * np.exp(ssum)
    R = Rpre + np.exp(gammav[d]) / (1.0 + np.exp(gammav[d]))
    if grad is None:
        return R
    if grad == 'hyp':
        dR = np.empty(S.shape[:2] + (d + 1,))
        dR[:, :, :d] = (S ** 2 / (1.0 + S)) * Rpre[:, :, None]
        dR[:, :, d] = np.exp(gammav[d]) / (1.0 + np.exp(gammav[d])) * (c - Rpre)
        return R, dR
    if grad == 'x1':
        ginv = np.exp(-gammav[:d])
        dR = (-ginv * diff / (1.0 + S)) * Rpre[:, :, None]
        return R, dR
    raise ValueError('grad must be None, "hyp" or "x1"')


# --------------------------------------------------------------------------------------
# per-PC penalised negative log-likelihood                     PCGPwM.py:850-907
# --------------------------------------------------------------------------------------
def _logistic(h):
    return np.exp(h) / (1.0 + np.exp(h))


def assemble_gp_matrix(theta, hyp, gvar):
    """R = (1-nu) R0 + nu I + exp(h_v) diag(gvar)                (PCGPwM.py:852-857)."""
    R0 = matern_covmat(theta, theta, hyp[:-2])
    nu = _logistic(hyp[-1])
    R = (1.0 - nu) * R0 + nu * np.eye(theta.shape[0])
    if gvar is not None:
        R = R + np.exp(hyp[-2]) * np.diag(gvar)
    return R, R0, nu


def negloglik_eigh(hyp, info):
    """Reference form (PCGPwM.py:850-868)."""
    R, _, _ = assemble_gp_matrix(info['theta'], hyp, info['gvar'])
    W, V = np.linalg.eigh(R)
    fc = (V / np.sqrt(np.abs(W))).T @ info['g']
    n = info['g'].shape[0]
    s0 = info['sig2ofconst']
    sig2hat = (n * np.mean(fc ** 2) + s0) / (n + s0)
    val = 0.5 * np.sum(np.log(np.abs(W))) + 0.5 * n * np.log(sig2hat)
    val += 0.5 * np.sum(((PRIOR_SHIFT + hyp - info['hypregmean']) / info['hypregstd']) ** 2)
    return val


def _hyper_derivative_stack(theta, hyp, gvar):
    """dR/dhyp_j for the p = d+3 hyper-parameters            (PCGPwM.py:874-887)."""
    R0, dR0 = matern_covmat(theta, theta, hyp[:-2], grad='hyp')
    n = theta.shape[0]
    nu = _logistic(hyp[-1])
    R = (1.0 - nu) * R0 + nu * np.eye(n)
    parts = [(1.0 - nu) * dR0]
    if gvar is not None:
        R = R + np.exp(hyp[-2]) * np.diag(gvar)
        parts.append((np.exp(hyp[-2]) * np.diag(gvar))[:, :, None])
    else:
        parts.append(np.zeros((n, n, 1)))
    parts.append((nu / (1.0 + np.exp(hyp[-1])) * (np.eye(n) - R0))[:, :, None])
    return R, np.concatenate(parts, axis=2)


def negloglikgrad_eigh(hyp, info):
    """Reference form (PCGPwM.py:871-907), including its per-parameter n^3 products."""
    R, dR = _hyper_derivative_stack(info['theta'], hyp, info['gvar'])
    W, V = np.linalg.eigh(R)
    Vh = V / np.sqrt(np.abs(W))
    fc = Vh.T @ info['g']
    n = info['g'].shape[0]
    s0 = info['sig2ofconst']
    sig2hat = (n * np.mean(fc ** 2) + s0) / (n + s0)
    Rinv = Vh @ Vh.T
    out = np.zeros(dR.shape[2])
    for j in range(dR.shape[2]):
        quad = np.sum((Vh @ np.multiply.outer(fc, fc) @ Vh.T) * dR[:, :, j])
        out[j] = 0.5 * n * (-quad / (n + s0)) / sig2hat + 0.5 * np.sum(Rinv * dR[:, :, j])
    out += (PRIOR_SHIFT + hyp - info['hypregmean']) / info['hypregstd'] ** 2
    return out


def negloglik_and_grad_chol(hyp, info, want_grad=True):
    """Cholesky restatement -- the formulation the CUDA path implements.

    0.5*sum log|W| = sum log L_ii ; n*mean(fc^2) = ||L^-1 g||^2 ; Rinv = L^-T L^-1 ; alpha = Rinv g.
    grad_j = 0.5 tr(Rinv dR_j) - 0.5 n alpha' dR_j alpha / ((n+s0) sig2hat) + prior'.
    Returns (nll, grad or None, info_code) with info_code > 0 when the factorisation broke down.
    """
    theta, g, gvar = info['theta'], info['g'], info['gvar']
    n = g.shape[0]
    s0 = info['sig2ofconst']
    if want_grad:
        R, dR = _hyper_derivative_stack(theta, hyp, gvar)
    else:
        R, _, _ = assemble_gp_matrix(theta, hyp, gvar)
    try:
        L = np.linalg.cholesky(R)
    except np.linalg.LinAlgError:
        return np.inf, None, 1
    y = sla.solve_triangular(L, g, lower=True)
    sig2hat = (y @ y + s0) / (n + s0)
    prior = (PRIOR_SHIFT + hyp - info['hypregmean']) / info['hypregstd']
    nll = np.sum(np.log(np.diag(L))) + 0.5 * n * np.log(sig2hat) + 0.5 * np.sum(prior ** 2)
    if not want_grad:
        return nll, None, 0
    Linv = sla.solve_triangular(L, np.eye(n), lower=True)
    Rinv = Linv.T @ Linv
    alpha = Linv.T @ y
    Wmat = 0.5 * Rinv - (0.5 * n / ((n + s0) * sig2hat)) * np.outer(alpha, alpha)
    grad = np.einsum('ab,abj->j', Wmat, dR) + prior / info['hypregstd']
    return nll, grad, 0


# --------------------------------------------------------------------------------------
# standardisation / imputation / PCA                            PCGPwM.py:533-654
# --------------------------------------------------------------------------------------
def standardize_f(fitinfo):
    """offset/scale, iterative low-rank imputation of NaNs, extravar   (PCGPwM.py:533-605)."""
    f = fitinfo['f']                     # (n, m)
    mof, mofrows = fitinfo['mof'], fitinfo['mofrows']
    epsPC = fitinfo['epsilonPC']
    n, m = f.shape
    offset = np.array([np.nanmean(f[:, k]) for k in range(m)])
    scale = np.array([np.nanstd(f[:, k]) / np.sqrt(1 - np.isnan(f[:, k]).mean()) for k in range(m)])
    if np.any(scale == 0):
        raise ValueError("You have a row that is non-varying.")
    if mof is None:
        fs = (f - offset) / scale
    else:
        epsImp = fitinfo['epsilonImpute']
        fs = np.zeros(f.shape)
        for k in range(m):                                   # seed missing cells with +-2 (565-571)
            fs[:, k] = (f[:, k] - offset[k]) / scale[k]
            nmiss = int(np.sum(mof[:, k]))
            if nmiss > 0:
                seed = np.empty(nmiss)
                seed[::2] = 2
                seed[1::2] = -2
                fs[np.where(mof[:, k])[0], k] = seed - np.mean(seed)
        for _ in range(40):                                  # 573-587
            U, S, _vt = np.linalg.svd(fs.T, full_matrices=False)
            Sp = S ** 2 - epsPC
            Up = U[:, Sp > 0]
            Sp = np.sqrt(Sp[Sp > 0])
            for rv in mofrows:
                miss = np.where(mof[rv, :] > 0.5)[0]
                obs = np.where(mof[rv, :] < 0.5)[0]
                Uo = Up[obs, :]
                # NB two separate copies: numpy turns X.T @ X on one object into SYRK, whose rounding
                # differs from the reference's GEMM and is amplified by the ill-conditioned solve below
                H = Up[obs, :].T @ Uo
                Amat = epsImp * np.diag(1 / Sp ** 2) + H
                J = Uo.T @ fs[rv, obs]
                fs[rv, miss] = (Up[miss, :] * ((Sp / np.sqrt(epsImp)) ** 2)) @ \
                    (J - H @ sla.solve(Amat, J, assume_a='pos'))
    U, S, _vt = np.linalg.svd(fs.T, full_matrices=False)
    Sp = S ** 2 - epsPC
    Up = U[:, Sp > 0]
    extravar = np.nanmean((fs - fs @ Up @ Up.T) ** 2, 0) * scale ** 2
    fitinfo['standardpcinfo'] = {'offset': offset, 'scale': scale, 'fs': fs,
                                 'U': U, 'S': S, 'extravar': extravar}


def principal_components(fitinfo):
    """PC basis, scores and per-row imputation variance            (PCGPwM.py:608-654)."""
    f, mof, mofrows = fitinfo['f'], fitinfo['mof'], fitinfo['mofrows']
    epsPC = fitinfo['epsilonPC']
    spc = fitinfo['standardpcinfo']
    fs = spc['fs']
    if 'U' in spc:
        U, S = spc['U'], spc['S']
    else:
        U, S, _vt = np.linalg.svd(fs.T, full_matrices=False)
    Sp = S ** 2 - epsPC
    pct = U[:, Sp > 0]
    pcw = np.sqrt(Sp[Sp > 0])
    pc = fs @ pct
    pcstdvar = np.zeros((f.shape[0], pct.shape[1]))
    if mof is not None:
        epsImp = fitinfo['epsilonImpute']
        for rv in mofrows:
            obs = np.where(mof[rv, :] < 0.5)[0]
            Po = pct[obs, :]
            H = pct[obs, :].T @ Po                               # separate copies: GEMM, not SYRK
            Q = np.diag(epsImp / pcw ** 2) + H
            J = Po.T @ fs[rv, obs]
            pc[rv, :] = (pcw ** 2 / epsImp + 1) * (J - H @ np.linalg.solve(Q, J))
            fs[rv, :] = pc[rv, :] @ pct.T
            term3 = np.diag(H) - np.sum(H * sla.solve(Q, H, assume_a='pos'), 0)
            pcstdvar[rv, :] = 1 - (pcw ** 2 / epsImp + 1) * term3
    effn = np.sum(np.clip(1 - pcstdvar, 0, 1))
    fitinfo['pcw'] = pcw
    fitinfo['pcto'] = 1 * pct
    fitinfo['pct'] = pct * pcw / np.sqrt(effn)
    fitinfo['pcti'] = pct * (np.sqrt(effn) / pcw)
    fitinfo['pc'] = fs @ fitinfo['pct']
    fitinfo['unscaled_pcstdvar'] = pcstdvar


# --------------------------------------------------------------------------------------
# per-PC GP fit                                                  PCGPwM.py:678-847
# --------------------------------------------------------------------------------------
def hyper_regulariser(theta, lognugmean, lognugLB, logvarconst=None):
    """Prior mean / bounds / std of the p=d+3 hyper-parameters     (PCGPwM.py:728-742)."""
    vmean = 4 if logvarconst is None else logvarconst
    vLB = -8 if logvarconst is None else logvarconst - 0.5
    vUB = 8 if logvarconst is None else logvarconst + 0.5
    base = 0.5 * np.log(theta.shape[1]) + np.log(np.std(theta, 0))
    mean = np.append(base, (0, vmean, lognugmean))
    LB = np.append(base - 4, (-12, vLB, lognugLB))
    UB = np.append(base + 4, (2, vUB, -8))
    std = (UB - LB) / 8
    std[-3] = 2
    std[-1] = 4
    return mean, LB, UB, std


def damp_gvar(gvar, eta, dampalpha):
    """gvar <- min(eta, gvar / (1-gvar)^alpha)                      (PCGPwM.py:755)."""
    return np.minimum(eta, gvar / ((1 - gvar) ** dampalpha))


def share_threshold(p):
    """1.25 (p + 5 sqrt p)                                          (PCGPwM.py:776-777, 808-809)."""
    return 1.25 * (p + 5 * np.sqrt(p))


def final_factorisation(theta, g, gvar_damped, hyp, s0=SIG2_OF_CONST, shared=False):
    """Full-n stage of __fitGP1d: nug, sig2, pw (+ R, Vh, Rinv)     (PCGPwM.py:810-846).

    shared=True follows the share-branch, which forms Rinv = V diag(1/W) V' (PCGPwM.py:827);
    the own-branch forms Vh Vh' (PCGPwM.py:844).  They differ only by rounding when R is SPD.
    """
    out = {}
    out['hypcov'] = hyp[:-2]
    out['hypvarconst'] = hyp[-2]
    out['nug'] = _logistic(hyp[-1])
    R = matern_covmat(theta, theta, out['hypcov'])
    R = (1 - out['nug']) * R + out['nug'] * np.eye(R.shape[0])
    if gvar_damped is not None:
        R = R + np.exp(out['hypvarconst']) * np.diag(gvar_damped)
    W, V = np.linalg.eigh(R)
    Vh = V / np.sqrt(np.abs(W))
    fc = Vh.T @ g
    n = R.shape[0]
    out['R'] = R
    out['Vh'] = Vh
    out['sig2'] = (np.mean(fc ** 2) * n + s0) / (n + s0)
    out['Rinv'] = (V @ np.diag(1 / W) @ V.T) if shared else (Vh @ Vh.T)
    out['pw'] = out['Rinv'] @ g
    return out


def fit_gp_1d(theta, g, hyp1, hyp2, hypvarconst, gvar, dampalpha, eta,
              hypstarts=None, hypinds=None, sig2ofconst=SIG2_OF_CONST,
              nll=negloglik_eigh, nllgrad=negloglikgrad_eigh):
    """One principal component's GP fit                           (PCGPwM.py:725-847).

    Consumes the global numpy RNG exactly as the reference does (one np.random.choice
    when n > 20 d, PCGPwM.py:744-748) so seeded runs are comparable.
    """
    mean, LB, UB, std = hyper_regulariser(theta, hyp1, hyp2, hypvarconst)
    sub = {'hypregmean': mean, 'hypregLB': LB, 'hypregUB': UB, 'hypregstd': std,
           'hyp': 1 * mean, 'sig2ofconst': sig2ofconst}
    n, d = theta.shape
    nh = min(20 * d, n)
    rows = np.random.choice(n, nh, replace=False) if n > nh else range(0, n)
    gvar = damp_gvar(gvar, eta, dampalpha)
    sub['theta'] = theta[rows, :]
    sub['g'] = g[rows]
    sub['gvar'] = gvar[rows]
    p = mean.shape[0]

    lead = -1
    L0 = nll(sub['hyp'], sub)
    if hypstarts is not None:
        for c in range(hypstarts.shape[0]):                      # candidate scan (761-769)
            L1 = nll(hypstarts[c, :], sub)
            if L1 < L0:
                sub['hyp'] = hypstarts[c, :]
                L0 = 1 * L1
                lead = hypinds[c]

    skip = False
    if lead > -0.5 and hypstarts.ndim > 1:                       # skip test (771-781)
        dL = nllgrad(sub['hyp'], sub)
        nc = hypstarts.shape[0]
        scal = np.std(hypstarts, 0) * nc / (1 + nc) + std / (1 + nc)
        skip = np.sum((dL * scal) ** 2) < share_threshold(p)

    if not skip:                                                 # L-BFGS-B in scaled coords (783-805)
        res = sopt.minimize(lambda v: nll(mean + v * std, sub),
                            (sub['hyp'] - mean) / std,
                            method='L-BFGS-B', options={'gtol': 0.1},
                            jac=lambda v: nllgrad(mean + v * std, sub) * std,
                            bounds=sopt.Bounds((LB - mean) / std, (UB - mean) / std))
        hypn = mean + res.x * std
        likdiff = L0 - nll(hypn, sub)
    else:
        likdiff = 0
    shared = bool(lead > -0.5 and (2 * likdiff) < share_threshold(p))
    if shared:                                                   # share-or-own (808-833)
        sub['hypind'] = lead
    else:
        sub['hyp'] = hypn
        sub['hypind'] = -1
    sub['gvar'] = gvar[rows]
    sub.update(final_factorisation(theta, g, gvar, sub['hyp'], sig2ofconst, shared=shared))
    return sub


def fit_gps(fitinfo, theta, numpcs, hyp1, hyp2, varconstant, **fit1d_kwargs):
    """Three Gauss-Seidel sweeps over the PCs                      (PCGPwM.py:678-722)."""
    hypinds = -1 * np.ones(numpcs)
    hypstarts = None
    if 'emulist' in fitinfo:
        hypstarts = np.zeros((numpcs, fitinfo['emulist'][0]['hyp'].shape[0]))
        for j in range(min(numpcs, len(fitinfo['emulist']))):
            hypstarts[j, :] = fitinfo['emulist'][j]['hyp']
            hypinds[j] = fitinfo['emulist'][j]['hypind']
    emulist = [dict() for _ in range(numpcs)]
    own = np.arange(numpcs)
    for _sweep in range(3):
        for j in range(numpcs):
            common = dict(theta=theta, g=fitinfo['pc'][:, j], hyp1=hyp1, hyp2=hyp2,
                          hypvarconst=varconstant, gvar=fitinfo['unscaled_pcstdvar'][:, j],
                          dampalpha=fitinfo['dampalpha'], eta=fitinfo['eta'],
                          sig2ofconst=SIG2_OF_CONST, **fit1d_kwargs)
            if np.sum(hypinds == own) > 0.5:
                leaders = np.where(hypinds == own)[0]
                emulist[j] = fit_gp_1d(hypstarts=hypstarts[leaders, :], hypinds=leaders, **common)
            else:
                emulist[j] = fit_gp_1d(**common)
                hypstarts = np.zeros((numpcs, emulist[j]['hyp'].shape[0]))
            emulist[j]['hypind'] = min(j, emulist[j]['hypind'])
            hypstarts[j, :] = emulist[j]['hyp']
            if emulist[j]['hypind'] < -0.5:
                emulist[j]['hypind'] = 1 * j
            hypinds[j] = 1 * emulist[j]['hypind']
    return emulist


def fit(fitinfo, x, theta, f, epsilonPC=0.001, epsilonImpute=10e-6, lognugmean=-10, lognugLB=-20,
        varconstant=None, dampalpha=0.3, eta=10, standardpcinfo=None, **kwargs):
    """Plugin entry: f is (m, n) as the facade passes it           (PCGPwM.py:12-137)."""
    f = f.T
    if not np.all(np.isfinite(f)):
        fitinfo['mof'] = np.logical_not(np.isfinite(f))
        fitinfo['mofrows'] = np.where(np.any(fitinfo['mof'] > 0.5, 1))[0]
    else:
        fitinfo['mof'] = None
        fitinfo['mofrows'] = None
    fitinfo.update(epsilonImpute=epsilonImpute, epsilonPC=epsilonPC, dampalpha=dampalpha,
                   eta=eta, theta=theta, f=f, x=x)
    if standardpcinfo is None:
        standardize_f(fitinfo)
    else:
        fitinfo['standardpcinfo'] = standardpcinfo
    principal_components(fitinfo)
    k = fitinfo['pc'].shape[1]
    logvarc = np.log(varconstant) if varconstant is not None else None
    emulist = fit_gps(fitinfo, theta, k, lognugmean, lognugLB, logvarc)
    fitinfo['logvarc'] = np.array([e['hypvarconst'] for e in emulist])
    fitinfo['pcstdvar'] = np.exp(fitinfo['logvarc']) * fitinfo['unscaled_pcstdvar']
    fitinfo['emulist'] = emulist


# --------------------------------------------------------------------------------------
# predict                                                        PCGPwM.py:140-328
# --------------------------------------------------------------------------------------
def predict_pcspace(fitinfo, theta, return_grad=False):
    """predvecs, predvars (B,k) [+ gradients (B,k,d)]              (PCGPwM.py:219-270)."""
    infos = fitinfo['emulist']
    theta = np.atleast_2d(theta)
    B, d = theta.shape
    k = len(infos)
    pv = np.zeros((B, k))
    pvar = np.zeros((B, k))
    dpv = np.zeros((B, k, d))
    dpvar = np.zeros((B, k, d))
    cache = {}
    for j in range(k):
        lead = int(infos[j]['hypind'])
        if lead == j:
            if return_grad:
                cache[j] = matern_covmat(theta, fitinfo['theta'], infos[j]['hypcov'], grad='x1')
            else:
                cache[j] = (matern_covmat(theta, fitinfo['theta'], infos[j]['hypcov']), None)
        r = (1 - infos[j]['nug']) * cache[lead][0]
        rVh = r @ infos[j]['Vh']
        pv[:, j] = r @ infos[j]['pw']
        pvar[:, j] = infos[j]['sig2'] * np.abs(1 - np.sum(rVh ** 2, 1))
        if return_grad:
            dr = (1 - infos[j]['nug']) * cache[lead][1]          # (B, n, d)
            rRinv = rVh @ infos[j]['Vh'].T
            # reference quirk kept for parity: dr already carries (1-nug) and PCGPwM.py:268 multiplies
            # by (1-nug) again; only the B=1,d=1 branch (PCGPwM.py:257-259) applies it once
            extra = 1.0 if (B == 1 and d == 1) else (1 - infos[j]['nug'])
            dpv[:, j, :] = extra * np.einsum('bnd,n->bd', dr, infos[j]['pw'])
            dpvar[:, j, :] = -2 * infos[j]['sig2'] * np.einsum('bn,bnd->bd', rRinv, dr)
    if return_grad:
        return pv, pvar, dpv, dpvar
    return pv, pvar


def factor_state_chol(theta, g, gvar_damped, hyp, s0=SIG2_OF_CONST):
    """Cholesky restatement of the full-n stage: what the CUDA path stores per PC.

    Linv = L^-1 (lower), pw = Linv' (Linv g), sig2 = (||Linv g||^2 + s0) / (n + s0).
    """
    nug = _logistic(hyp[-1])
    R = (1 - nug) * matern_covmat(theta, theta, hyp[:-2]) + nug * np.eye(theta.shape[0])
    if gvar_damped is not None:
        R = R + np.exp(hyp[-2]) * np.diag(gvar_damped)
    L = np.linalg.cholesky(R)
    Linv = sla.solve_triangular(L, np.eye(R.shape[0]), lower=True)
    yv = Linv @ g
    n = R.shape[0]
    return {'Linv': Linv, 'pw': Linv.T @ yv, 'sig2': (yv @ yv + s0) / (n + s0), 'nug': nug,
            'hypcov': hyp[:-2]}


def predict_pcspace_chol(theta_fit, states, hypind, theta, return_grad=False):
    """PC-space prediction in the form the CUDA path uses (T1 = Linv r', T2 = Linv' T1)."""
    theta = np.atleast_2d(theta)
    B, d = theta.shape
    k = len(states)
    pv, pvar = np.zeros((B, k)), np.zeros((B, k))
    dpv, dpvar = np.zeros((B, k, d)), np.zeros((B, k, d))
    for j in range(k):
        st = states[j]
        R0, dR0 = matern_covmat(theta, theta_fit, states[int(hypind[j])]['hypcov'], grad='x1')
        r = (1 - st['nug']) * R0
        T1 = st['Linv'] @ r.T                                    # (n, B)
        pv[:, j] = r @ st['pw']
        pvar[:, j] = st['sig2'] * np.abs(1 - np.sum(T1 ** 2, 0))
        if return_grad:
            T2 = st['Linv'].T @ T1                               # (n, B) = Rinv r'
            dr = (1 - st['nug']) * dR0
            extra = 1.0 if (B == 1 and d == 1) else (1 - st['nug'])
            dpv[:, j, :] = extra * np.einsum('bnd,n->bd', dr, st['pw'])
            dpvar[:, j, :] = -2 * st['sig2'] * np.einsum('nb,bnd->bd', T2, dr)
    if return_grad:
        return pv, pvar, dpv, dpvar
    return pv, pvar


def predict(predinfo, fitinfo, x, theta, return_grad=False, return_covx=True, xind=None):
    """Back-projection through Phi = pcti * scale                  (PCGPwM.py:273-327).

    x-row matching (PCGPwM.py:197-214) is reduced to an explicit `xind` (rows of the fit x
    to predict at); None = all rows in fit order.
    """
    theta = np.atleast_2d(theta)
    out = predict_pcspace(fitinfo, theta, return_grad)
    pv, pvar = out[0], out[1]
    spc = fitinfo['standardpcinfo']
    phi = fitinfo['pcti'] * spc['scale'][:, None]
    if xind is None:
        xind = np.arange(phi.shape[0])
    phi = phi[xind, :]
    predinfo['mean'] = (pv @ phi.T + spc['offset'][xind]).T
    predinfo['var'] = (spc['extravar'][xind] + pvar @ (phi ** 2).T).T
    predinfo['extravar'] = 1 * spc['extravar'][xind]
    predinfo['predvars'] = pvar
    predinfo['predvecs'] = pv
    predinfo['phi'] = phi
    predinfo['return_grad'] = return_grad
    predinfo['return_covx'] = return_covx
    if return_covx:
        predinfo['covxhalf'] = (np.sqrt(pvar)[:, :, None] * phi.T[None, :, :]).transpose(2, 0, 1)
    if return_grad:
        dpv, dpvar = out[2], out[3]
        predinfo['mean_gradtheta'] = np.einsum('bkd,xk->xbd', dpv, phi)
        predinfo['predvecs_gradtheta'] = dpv
        predinfo['predvars_gradtheta'] = dpvar
        if return_covx:
            dsq = 0.5 * dpvar / np.sqrt(pvar)[:, :, None]        # (B,k,d)
            predinfo['covxhalf_gradtheta'] = np.einsum('bkd,xk->xbkd', dsq, phi)


def predictlpdf(predinfo, f, addvar=0):
    """Log predictive density of simulator output f (m, B) given a predict() result   (PCGPwM.py:331-364)."""
    totvar = addvar + predinfo['extravar']
    isd = 1 / np.sqrt(totvar)
    rf = (f - predinfo['mean']) * isd[:, None]
    Gf = predinfo['phi'].T * isd
    Gfrf = Gf @ rf
    Gf2 = Gf @ Gf.T
    likv = np.sum(rf ** 2, 0)
    grad = predinfo['return_grad']
    B = predinfo['predvars'].shape[0]
    if grad:
        rf2 = -predinfo['mean_gradtheta'] * isd[:, None, None]                 # (m, B, d)
        Gfrf2 = np.einsum('km,mbd->kbd', Gf, rf2)
        dlikv = 2 * np.einsum('mbd,mb->bd', rf2, rf)
    for c in range(B):
        v = predinfo['predvars'][c]
        w, V = np.linalg.eig(np.diag(1 / v) + Gf2)
        t1 = (V * (1 / w)) @ (V.T @ Gfrf[:, c])
        likv[c] += -Gfrf[:, c] @ t1 + np.sum(np.log(v)) + np.sum(np.log(w))
        if grad:
            Si = (V * (1 / w)) @ V.T
            dv = predinfo['predvars_gradtheta'][c]                              # (k, d)
            grt = dv / v[:, None]
            dlikv[c] += np.sum(grt, 0) + np.diag(Si) @ (-grt / v[:, None]) - 2 * Gfrf2[:, c, :].T @ t1 \
                - ((t1 / v) ** 2) @ dv
    if grad:
        return (-likv / 2).reshape(-1, 1), -dlikv / 2
    return (-likv / 2).reshape(-1, 1)
