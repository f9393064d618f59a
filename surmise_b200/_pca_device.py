"""Device (torch CUDA) version of the standardisation / imputation / PCA stage of PCGPwM.

Same mathematics as surmise's PCGPwM.__standardizef / __PCs (emulationmethods/PCGPwM.py:533-654) and as the
host implementation in _pca.py; SURVEY.md section 8(f) row 2.  This stage is not one of the hand-written
kernels of the hot path: it is expressed with torch.linalg on the device (library eigensolver / batched solve),
because with missing data it otherwise dominates the whole fit (config C3: 287 s on the host, 41 SVDs of a
1000 x 4000 matrix plus 160 k ridge systems).

The reference takes the SVD of fs' (m x n) and only ever uses U and S.  Here U and S^2 come from the symmetric
eigendecomposition of the m x m Gram matrix fs' fs, which is much cheaper and accurate for every component the
method keeps (S^2 > epsilonPC; relative eigenvalue error ~ eps * S_max^2 / S_k^2 <= 1e-9).  Singular-vector signs
are arbitrary in both; nothing downstream depends on them.
"""
import numpy as np


def _torch():
    from . import _lib
    return _lib.require_cuda()


def _left_singular(fs, torch):
    """U (m, m) and S (m,) of fs' (descending), from eigh of the Gram matrix."""
    lam, U = torch.linalg.eigh(fs.T @ fs)
    lam = torch.flip(lam, [0])
    U = torch.flip(U, [1])
    return U, torch.sqrt(torch.clamp(lam, min=0.0)), lam


def _leading_eig(G, V0, eps_pc, torch, steps=2, tol=1e-11):
    """Eigenpairs of the PSD Gram matrix G above eps_pc from a warm start, or None if that cannot be certified.

    Once the imputation sweeps have settled, only k << m components exceed eps_pc and the basis barely moves from
    one sweep to the next, so a block power iteration from the previous basis V0 (m, k + pad) plus Rayleigh-Ritz
    replaces the full m x m eigendecomposition (C3: 13 ms -> 2 ms per sweep).  Two checks make it safe:
      * trace(G) - sum(Ritz values) bounds the sum of ALL remaining eigenvalues (Ritz values never exceed the true
        ones); if that is below eps_pc, no component outside the block can pass the threshold;
      * the residual |G X - X diag(w)| of the kept pairs must be at rounding level.
    Returns (X_keep (m, k) in descending order, w_keep (k,), X_all for the next warm start) or None."""
    V = V0
    for _ in range(steps):
        V, _r = torch.linalg.qr(G @ V)
    GV = G @ V
    w, Q = torch.linalg.eigh(V.T @ GV)
    w, Q = torch.flip(w, [0]), torch.flip(Q, [1])
    X = V @ Q
    if float(torch.trace(G) - w.sum()) >= eps_pc:
        return None
    keep = (w - eps_pc) > 0
    if bool(keep.all()):                                # no Ritz value below the threshold inside the block: widen it
        return None
    Xk, wk = X[:, keep], w[keep]
    resid = torch.linalg.matrix_norm(G @ Xk - Xk * wk) / torch.linalg.matrix_norm(G)
    if float(resid) > tol:
        return None
    return Xk, wk, X


_CHUNK_BYTES = 8 << 30      # bound on the temporaries of one row chunk (early imputation sweeps keep k ~ m PCs)


def _row_chunks(nrows, m, k):
    per_row = 8 * k * max(m, 3 * k)
    step = max(1, min(nrows, _CHUNK_BYTES // max(per_row, 1)))
    return [(s, min(s + step, nrows)) for s in range(0, nrows, step)]


def _ridge_rows(basis, obs, cur, diag_add, torch, want_qih=False):
    """Per row r: H = basis' diag(obs_r) basis, J = basis' (obs_r * cur_r), sol = (H + diag_add)^-1 J.

    Returns vec = J - H sol  (nr, k) and, if want_qih, term3 = diag(H) - sum_a H_ab [(H + diag_add)^-1 H]_ab.
    Rows are processed in chunks so the (chunk, k, k) temporaries stay below _CHUNK_BYTES."""
    nr, m = obs.shape
    k = basis.shape[1]
    vec = torch.empty((nr, k), dtype=torch.float64, device=basis.device)
    term3 = torch.empty((nr, k), dtype=torch.float64, device=basis.device) if want_qih else None
    D = torch.diag(diag_add)
    for lo, hi in _row_chunks(nr, m, k):
        o = obs[lo:hi]
        H = torch.matmul((basis.unsqueeze(0) * o.unsqueeze(2)).transpose(1, 2), basis)    # (c, k, k)
        J = (o * cur[lo:hi]) @ basis
        A = H + D
        sol = torch.linalg.solve(A, J.unsqueeze(2)).squeeze(2)
        vec[lo:hi] = J - torch.einsum('rab,rb->ra', H, sol)
        if want_qih:
            QiH = torch.linalg.solve(A, H)
            term3[lo:hi] = torch.diagonal(H, dim1=1, dim2=2) - torch.einsum('rab,rab->rb', H, QiH)
    return vec, term3


def _missing_index(miss, torch):
    """Padded column indices of the missing entries of each row: idx (nr, qmax) and valid mask (nr, qmax)."""
    q = miss.sum(1)
    qmax = int(q.max().item())
    order = torch.argsort((~miss).to(torch.int8), dim=1, stable=True)      # missing columns first, in column order
    idx = order[:, :qmax]
    valid = torch.arange(qmax, device=miss.device)[None, :] < q[:, None]
    return idx, valid


def _impute_schur(Up, sp2, eps_impute, cur, obs, idx, valid, torch):
    """The ridge update of the missing entries (PCGPwM.py:576-586) through the |missing| x |missing| system.

    The reference solves, per row, the k x k system (H + D) sol = J with H = Up' diag(obs) Up, D = diag(eps / sp2),
    J = Up' (obs * cur) and fills the missing entries with Up sol.  Up has orthonormal columns, so
    H + D = Lam - M' M with Lam = I + D and M = Up[missing rows of this output row], and by the Woodbury identity the
    filled values are  w = (I - P_mm)^-1 P_mo cur_o  with P = Up Lam^-1 Up' (m x m, once per sweep): one GEMM for
    all rows plus a batched Cholesky of size |missing| instead of size k.  Early sweeps keep k ~ m components
    (k^3 per row); |missing| is the missing fraction of m.  Returns the filled values, (nr, qmax), 0 where padded."""
    lam_inv = sp2 / (sp2 + eps_impute)
    P = (Up * lam_inv) @ Up.T
    Z = (obs * cur) @ P                                            # (nr, m): P_{.,o} cur_o for every column
    z = torch.gather(Z, 1, idx) * valid
    nr, qmax = idx.shape
    w = torch.zeros_like(z)
    eye = torch.eye(qmax, dtype=Up.dtype, device=Up.device)
    per_row = 8 * qmax * qmax * 3
    step = max(1, min(nr, _CHUNK_BYTES // max(per_row, 1)))
    from . import ops
    small = qmax <= ops.small_max_n()                 # the register-resident sweep kernel solves these directly
    for lo in range(0, nr, step):
        hi = min(lo + step, nr)
        ii = idx[lo:hi]
        vv = valid[lo:hi].to(Up.dtype)
        S = P[ii[:, :, None], ii[:, None, :]] * (vv[:, :, None] * vv[:, None, :])
        S = eye - S                                                # padded rows / columns: identity
        if small:
            lda = (qmax + 7) // 8 * 8
            A = torch.zeros((hi - lo, qmax + 1, lda), dtype=Up.dtype, device=Up.device)
            A[:, :qmax, :qmax] = S
            A[:, qmax, :qmax] = z[lo:hi]
            info = ops.spd_solve_small(A, qmax)
            if bool((info != 0).any()):
                raise ValueError('imputation: a Schur system of the missing entries is not positive definite')
            w[lo:hi] = A[:, qmax, :qmax]
        else:
            L = torch.linalg.cholesky(S)
            w[lo:hi] = torch.cholesky_solve(z[lo:hi, :, None], L).squeeze(2)
    return w * valid


def standardize(f, eps_pc, eps_impute, n_iter=40):
    """f: (n, m) numpy with NaN/inf for missing.  Returns (standardpcinfo of numpy arrays, mof, mofrows)."""
    torch = _torch()
    dev = torch.device('cuda', torch.cuda.current_device())
    f_np = np.asarray(f, dtype=np.float64)
    mof_np = ~np.isfinite(f_np)
    any_missing = bool(mof_np.any())
    # offset / scale exactly as the reference computes them (PCGPwM.py:553-557), on the host: O(n m)
    if any_missing:
        offset_np = np.nanmean(f_np, axis=0)
        scale_np = np.nanstd(f_np, axis=0) / np.sqrt(1 - np.isnan(f_np).mean(axis=0))
    else:                           # same bits as the nan-versions on complete data, three times faster (C4: 0.2 s)
        offset_np = np.mean(f_np, axis=0)
        scale_np = np.std(f_np, axis=0)
    if np.any(scale_np == 0):
        raise ValueError("You have a row that is non-varying.")
    fs_np = (f_np - offset_np) / scale_np
    mofrows_np = None
    if any_missing:
        mofrows_np = np.where(mof_np.any(axis=1))[0]
        # initial fill +2, -2, +2, ... minus its mean, in row order within each column (PCGPwM.py:565-571)
        rank = np.cumsum(mof_np, axis=0) - 1
        alt = np.where(rank % 2 == 0, 2.0, -2.0)
        cnt = mof_np.sum(axis=0)
        col_mean = np.where(cnt > 0, (alt * mof_np).sum(axis=0) / np.maximum(cnt, 1), 0.0)
        fs_np = np.where(mof_np, alt - col_mean, fs_np)
    fs = torch.from_numpy(np.ascontiguousarray(fs_np)).to(dev)
    if any_missing:
        rows = torch.from_numpy(mofrows_np).to(dev)
        miss = torch.from_numpy(mof_np[mofrows_np]).to(dev)
        obs = (~miss).to(torch.float64)
        idx, valid = _missing_index(miss, torch)
        warm, pad = None, 8
        for _ in range(n_iter):                                   # PCGPwM.py:573-587
            lead = None
            if warm is not None and warm.shape[1] <= fs.shape[1] // 4:
                lead = _leading_eig(fs.T @ fs, warm, eps_pc, torch)
            if lead is not None:
                Up, lam_keep, warm = lead
                sp2 = lam_keep - eps_pc
            else:
                U, _S, lam = _left_singular(fs, torch)
                keep = (lam - eps_pc) > 0
                Up = U[:, keep]
                sp2 = lam[keep] - eps_pc
                nk = int(keep.sum().item())
                warm = U[:, :min(nk + pad, U.shape[1])]
            cur = fs[rows]
            if Up.shape[1] > idx.shape[1]:                        # more components than missing entries per row
                w = _impute_schur(Up, sp2, eps_impute, cur, obs, idx, valid, torch)
                filled = cur.scatter(1, idx, torch.where(valid, w, torch.gather(cur, 1, idx)))
                fs[rows] = filled
            else:
                vec, _ = _ridge_rows(Up, obs, cur, eps_impute / sp2, torch)
                filled = vec @ (Up * (sp2 / eps_impute)).T
                fs[rows] = torch.where(miss, filled, cur)
    U, S, lam = _left_singular(fs, torch)
    Up = U[:, (lam - eps_pc) > 0]
    resid = fs - (fs @ Up) @ Up.T
    extravar = (resid ** 2).mean(0).cpu().numpy() * scale_np ** 2   # PCGPwM.py:594
    info = {'offset': offset_np, 'scale': scale_np, 'fs': fs, 'U': U, 'S': S, 'extravar': extravar, '_lam': lam}
    return info, (mof_np if any_missing else None), mofrows_np


def principal_components(spc, mof, mofrows, eps_pc, eps_impute):
    """PC loadings / scores / per-row imputation variance (PCGPwM.py:608-654) from a device standardpcinfo.

    Returns a dict of numpy arrays and converts spc's device tensors to numpy in place."""
    torch = _torch()
    fs, U, lam = spc['fs'], spc['U'], spc.pop('_lam')
    keep = (lam - eps_pc) > 0
    basis = U[:, keep]
    pcw = torch.sqrt(lam[keep] - eps_pc)
    pcstdvar = torch.zeros((fs.shape[0], basis.shape[1]), dtype=torch.float64, device=fs.device)
    if mof is not None:
        rows = torch.from_numpy(mofrows).to(fs.device)
        obs = torch.from_numpy(~mof[mofrows]).to(fs.device).to(torch.float64)
        gain = pcw ** 2 / eps_impute + 1
        vec, term3 = _ridge_rows(basis, obs, fs[rows], eps_impute / pcw ** 2, torch, want_qih=True)
        fs[rows] = (gain * vec) @ basis.T
        pcstdvar[rows] = 1 - gain * term3
    effn = torch.sum(torch.clamp(1 - pcstdvar, 0, 1))
    pct = basis * pcw / torch.sqrt(effn)
    out = {'pcw': pcw, 'pcto': basis.clone(), 'pct': pct, 'pcti': basis * (torch.sqrt(effn) / pcw),
           'pc': fs @ pct, 'unscaled_pcstdvar': pcstdvar}
    from . import ops
    keys = list(out)
    host = ops.to_host([out[k] for k in keys] + [spc[k] for k in ('fs', 'U', 'S')])   # pinned copies, one sync
    out = dict(zip(keys, host[:len(keys)]))
    for k, v in zip(('fs', 'U', 'S'), host[len(keys):]):
        spc[k] = v
    return out
