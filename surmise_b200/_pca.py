"""Host-side standardisation, missing-value imputation and principal components for PCGPwM_B200.

Same mathematics as surmise's PCGPwM.__standardizef / __PCs (emulationmethods/PCGPwM.py:533-654),
restructured so that the per-row k x k ridge systems of all rows with missing outputs are formed
and solved as one batch (masked Gram matrices via a single GEMM, batched LAPACK solve) instead of
a Python loop per row.  Runs in numpy: moving it to the device is listed as "next" in DESIGN.md.
"""
import numpy as np


def _alternating_seed(count):
    """+2, -2, +2, ... minus its mean: the reference's initial fill of missing cells (PCGPwM.py:565-571)."""
    a = np.empty(count)
    a[::2] = 2
    a[1::2] = -2
    return a - a.mean()


def _masked_grams(basis, obs):
    """H_r = basis' diag(obs_r) basis for every row r: (nr, k, k) from one (nr, m) @ (m, k*k) product."""
    m, k = basis.shape
    outer = (basis[:, :, None] * basis[:, None, :]).reshape(m, k * k)
    return (obs @ outer).reshape(obs.shape[0], k, k)


def standardize(f, eps_pc, eps_impute, n_iter=40):
    """f: (n, m) with NaN/inf for missing.  Returns (standardpcinfo dict, mof, mofrows)."""
    f = np.asarray(f, dtype=np.float64)
    mof = ~np.isfinite(f)
    any_missing = bool(mof.any())
    offset = np.nanmean(f, axis=0)
    scale = np.nanstd(f, axis=0) / np.sqrt(1 - np.isnan(f).mean(axis=0))
    if np.any(scale == 0):
        raise ValueError("You have a row that is non-varying.")
    fs = (f - offset) / scale
    mofrows = None
    if any_missing:
        mofrows = np.where(mof.any(axis=1))[0]
        for col in np.where(mof.any(axis=0))[0]:
            idx = np.where(mof[:, col])[0]
            fs[idx, col] = _alternating_seed(idx.shape[0])
        miss = mof[mofrows]
        obs = (~miss).astype(np.float64)
        for _ in range(n_iter):                                   # PCGPwM.py:573-587
            U, S, _vt = np.linalg.svd(fs.T, full_matrices=False)
            lam = S ** 2 - eps_pc
            Up = U[:, lam > 0]
            sp2 = lam[lam > 0]
            rows = fs[mofrows]
            H = _masked_grams(Up, obs)
            J = (obs * rows) @ Up
            A = H + eps_impute * np.diag(1 / sp2)
            sol = np.linalg.solve(A, J[:, :, None])[:, :, 0]
            vec = J - np.einsum('rab,rb->ra', H, sol)
            filled = vec @ (Up * (sp2 / eps_impute)).T
            fs[mofrows] = np.where(miss, filled, rows)
    U, S, _vt = np.linalg.svd(fs.T, full_matrices=False)
    Up = U[:, (S ** 2 - eps_pc) > 0]
    extravar = np.nanmean((fs - fs @ Up @ Up.T) ** 2, 0) * scale ** 2     # PCGPwM.py:594
    info = {'offset': offset, 'scale': scale, 'fs': fs, 'U': U, 'S': S, 'extravar': extravar}
    return info, (mof if any_missing else None), mofrows


def principal_components(fs, U, S, mof, mofrows, eps_pc, eps_impute):
    """PC loadings/scores and the per-row imputation variance (PCGPwM.py:608-654).

    Returns dict(pct, pcti, pcto, pcw, pc, unscaled_pcstdvar); modifies fs rows with missing outputs
    exactly like the reference (they are replaced by their rank-k reconstruction).
    """
    lam = S ** 2 - eps_pc
    basis = U[:, lam > 0]
    pcw = np.sqrt(lam[lam > 0])
    pcstdvar = np.zeros((fs.shape[0], basis.shape[1]))
    if mof is not None:
        obs = (~mof[mofrows]).astype(np.float64)
        H = _masked_grams(basis, obs)
        Q = H + np.diag(eps_impute / pcw ** 2)
        J = (obs * fs[mofrows]) @ basis
        gain = pcw ** 2 / eps_impute + 1
        sol = np.linalg.solve(Q, J[:, :, None])[:, :, 0]
        pcrow = gain * (J - np.einsum('rab,rb->ra', H, sol))
        fs[mofrows] = pcrow @ basis.T
        QiH = np.linalg.solve(Q, H)
        term3 = np.einsum('raa->ra', H) - np.einsum('rab,rab->rb', H, QiH)
        pcstdvar[mofrows] = 1 - gain * term3
    effn = np.sum(np.clip(1 - pcstdvar, 0, 1))
    pct = basis * pcw / np.sqrt(effn)
    return {'pcw': pcw, 'pcto': 1 * basis, 'pct': pct, 'pcti': basis * (np.sqrt(effn) / pcw),
            'pc': fs @ pct, 'unscaled_pcstdvar': pcstdvar}


def complete_pcinfo(pcinfo, f):
    """User-supplied basis only (PCGPwM.py:910-935): fill offset/scale/S/fs/extravar."""
    if 'U' not in pcinfo:
        raise AttributeError("'U', the basis vectors must be provided.")
    if len(pcinfo) == 1:
        Phi = pcinfo['U']
        offset = f.mean(0)
        fs = f - offset
        pcinfo.update(offset=offset, scale=np.ones(f.shape[1]), S=np.ones(Phi.shape[1]), fs=fs,
                      extravar=np.mean((fs.T - Phi @ Phi.T @ fs.T) ** 2, 1))
    return pcinfo
