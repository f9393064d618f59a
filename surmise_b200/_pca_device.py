"""Device version of the standardisation / imputation / PCA stage of PCGPwM.

Same mathematics as surmise's PCGPwM.__standardizef / __PCs (emulationmethods/PCGPwM.py:533-654) and as the
host implementation in _pca.py; SURVEY.md section 8(f) row 2.  Everything numerical runs in this package's own kernels
through the C ABI (include/surmise_b200.h): the column statistics and the scaling (sb200_pca_col_stats /
sb200_pca_standardize), every dense product (sb200_dgemm, FP64 DMMA), the eigen-decomposition (block subspace
iteration on sb200_dgemm with the Jacobi kernel sb200_syevj_batched for orthonormalisation and Rayleigh-Ritz) and the
per-row ridge / Schur systems of the imputation (sb200_spd_solve_small, or the blocked sb200_potrf_batched /
sb200_trtri_batched beyond its size); torch supplies memory, indexing and elementwise glue.

The reference takes the SVD of fs' (m x n) and only ever uses the left singular vectors whose S^2 exceeds epsilonPC
(PCGPwM.py:575-577, 589-591, 620-622).  Here those come from the m x m Gram matrix G = fs' fs as a CERTIFIED leading
block: Ritz pairs (X, w) of a p-dimensional subspace with ||G - X diag(w) X'||_F < epsilonPC, which bounds every
eigenvalue outside the block (Weyl), so the set of components above the threshold is the reference's.  U and S in
standardpcinfo are that block (m x p, p), not the full m x m basis -- nothing downstream reads more.  If no block of
at most sb200_syevj_max_p() columns can be certified (a spectrum that does not drop below epsilonPC within 112
components, e.g. the first imputation sweeps) the full symmetric eigensolver of the CUDA library (torch.linalg.eigh)
is used for that decomposition and the fact is recorded in standardpcinfo['_eig_path'].
Singular-vector signs are arbitrary in both implementations; nothing downstream depends on them.
"""
import numpy as np


def _torch():
    from . import _lib
    return _lib.require_cuda()


class LazyHostDict(dict):
    """standardpcinfo whose large entries (the standardised n x m matrix fs) stay on the device until someone reads
    them: the reference only reads fs inside the fit itself (PCGPwM.py:616), and copying 128 MB to fresh host memory was
    a quarter of a C4 fit.  Reading through [], get, items, values, copying or pickling materialises numpy arrays;
    the mapping then behaves like the plain dict the reference stores."""
    __slots__ = ('_pending',)

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._pending = {}

    def defer(self, key, tensor):
        self._pending[key] = tensor
        dict.__setitem__(self, key, None)

    def _materialise(self, key=None):
        keys = [key] if key is not None else list(self._pending)
        keys = [k for k in keys if k in self._pending]
        if keys:
            from . import ops
            for k, v in zip(keys, ops.to_host([self._pending[k] for k in keys])):
                dict.__setitem__(self, k, v)
                del self._pending[k]

    def __getitem__(self, key):
        if key in self._pending:
            self._materialise(key)
        return dict.__getitem__(self, key)

    def __setitem__(self, key, value):
        self._pending.pop(key, None)
        dict.__setitem__(self, key, value)

    def __delitem__(self, key):
        self._pending.pop(key, None)
        dict.__delitem__(self, key)

    def __iter__(self):                      # (overridden so that dict(d) / {**d} go through keys() and [])
        return dict.__iter__(self)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def pop(self, key, *default):
        self._materialise(key)
        return dict.pop(self, key, *default)

    def items(self):
        self._materialise()
        return dict.items(self)

    def values(self):
        self._materialise()
        return dict.values(self)

    def copy(self):
        self._materialise()
        return dict(dict.items(self))

    def device(self, key):
        """The device tensor of a deferred entry (None once it has been materialised)."""
        return self._pending.get(key)

    def __reduce__(self):
        self._materialise()
        return (dict, (dict(dict.items(self)),))


def _ops():
    from . import ops
    return ops


# ------------------------------------------------------------------------------------------------------------------
# eigen-decomposition of the Gram matrix
def _full_eig(G, torch):
    """All eigenpairs, descending (library fallback, see the module docstring)."""
    lam, U = torch.linalg.eigh(G)
    return torch.flip(U, [1]).contiguous(), torch.flip(lam, [0]).contiguous()


def _orthonormalise(Y, ops, torch):
    """Columns of Y -> an orthonormal basis of their span (SVQB: Y Z Lam^-1/2 from the Jacobi eigen-decomposition of
    the column-scaled Gram matrix; a second pass when the first met a badly conditioned block)."""
    for _ in range(2):
        Y = Y / torch.sqrt((Y * Y).sum(0)).clamp_min(1e-300)
        lam, Z, _s = ops.syevj(ops.mm(Y, Y, ta=True))
        top = lam[0]
        Y = ops.mm(Y, Z / torch.sqrt(lam.clamp_min(top * 1e-28)))
        if float(lam[-1]) > 1e-3 * float(top):
            break
    return Y


def _leading_block(G, p, warm, eps_pc, ops, torch, max_iter=60):
    """Ritz pairs of a p-dimensional invariant subspace of the PSD matrix G by block subspace iteration (two products
    with G per step) with Rayleigh-Ritz.  Returns (X (m,p), w (p) descending, certified) where certified means
    ||G - X diag(w) X'||_F < eps_pc and the pairs above eps_pc have residuals at rounding level."""
    m = G.shape[0]
    gen = torch.Generator(device=G.device)
    gen.manual_seed(1234567 + p)
    X = torch.randn((m, p), dtype=torch.float64, device=G.device, generator=gen)
    if warm is not None:
        c = min(p, warm.shape[1])
        X[:, :c] = warm[:, :c]
    best = None
    prev = np.inf
    for it in range(max_iter):
        Q = _orthonormalise(ops.mm(G, ops.mm(G, X)) if it else ops.mm(G, X), ops, torch)
        GQ = ops.mm(G, Q)
        w, Z, _s = ops.syevj(ops.mm(Q, GQ, ta=True))
        X = ops.mm(Q, Z)
        R = ops.mm(GQ, Z) - X * w
        rn = torch.sqrt((R * R).sum(0))
        keep = w > eps_pc
        if p < m and bool(keep[-1]):                   # Ritz values never exceed the eigenvalues they approximate: the
            return X, w, False                         # p-th eigenvalue is above the threshold, the block is too small
        top = float(w[0])
        worst = float((rn * keep).max())
        best = (X, w)
        if worst <= 1e-13 * top:
            break
        if worst <= 1e-10 * top and worst > 0.5 * prev:       # stagnated at rounding level
            break
        prev = worst
    else:
        return X, w, False
    E = ops.mm(X * w, X, tb=True, alpha=-1.0, beta=1.0, out=G.clone())
    return best[0], best[1], bool(float(ops.sumsq(E)) < eps_pc * eps_pc)


def _eig_above(G, eps_pc, warm, ops, torch, note):
    """(U, lam) descending such that every eigenvalue of G above eps_pc is among them (see module docstring)."""
    m = G.shape[0]
    cap = min(ops.syevj_max_p(), m)
    p = min(cap, max(16, warm.shape[1] if warm is not None else 32))
    while cap >= 2:
        X, w, ok = _leading_block(G, p, warm, eps_pc, ops, torch)
        if ok:
            note['subspace'] = note.get('subspace', 0) + 1
            return X, w
        if p >= cap:
            break
        warm, p = X, min(cap, (p + max(16, p // 2) + 7) // 8 * 8)
    note['library_eigh'] = note.get('library_eigh', 0) + 1
    return _full_eig(G, torch)


# ------------------------------------------------------------------------------------------------------------------
# batched small SPD systems of the imputation
_CHUNK_BYTES = 8 << 30      # bound on the temporaries of one row chunk (early imputation sweeps keep k ~ m PCs)


def _spd_inverse(A, ops, torch):
    """Inverses of a batch (c, q, q) of symmetric positive definite matrices (full symmetric), on our own kernels."""
    c, q, _ = A.shape
    if q <= ops.small_max_n():                       # register-resident sweep kernel
        lda = (q + 7) // 8 * 8
        W = torch.zeros((c, q + 1, lda), dtype=A.dtype, device=A.device)
        W[:, :q, :q] = A
        info = ops.spd_solve_small(W, q)
        low = W[:, :q, :q]
    else:                                            # blocked Cholesky / triangular inverse / L^-T L^-1 on the DMMA GEMM
        lda = (q + 7) // 8 * 8
        W = torch.zeros((c, q, lda), dtype=A.dtype, device=A.device)
        W[:, :, :q] = A
        Linv, info = ops.potrf_batched(W, q, q)
        ops.trtri_batched(W, Linv, q)
        low = ops.lauum_batched(Linv, q)[:, :, :q]
    if bool((info != 0).any()):
        raise ValueError('imputation: a ridge system of the missing entries is not positive definite')
    low = torch.tril(low)
    return low + torch.tril(low, -1).transpose(1, 2)


def _batch_cap(q, ops):
    """Largest batch one call of the batched solvers takes (grid limits of the blocked path beyond the sweep kernel)."""
    return 60000 if q <= ops.small_max_n() else 8000


def _row_chunks(nrows, m, k, ops):
    per_row = 8 * k * max(m, 4 * k)
    step = max(1, min(nrows, _batch_cap(k, ops), _CHUNK_BYTES // max(per_row, 1)))
    return [(s, min(s + step, nrows)) for s in range(0, nrows, step)]


def _ridge_rows(basis, obs, cur, diag_add, ops, torch, want_qih=False):
    """Per row r: H = basis' diag(obs_r) basis, J = basis' (obs_r * cur_r), sol = (H + diag_add)^-1 J.

    Returns vec = J - H sol  (nr, k) and, if want_qih, term3 = diag(H) - sum_a H_ab [(H + diag_add)^-1 H]_ab.
    The masked Gram matrices of all rows of a chunk are ONE product obs (c, m) x [basis_a basis_b] (m, k^2)."""
    nr, m = obs.shape
    k = basis.shape[1]
    vec = torch.empty((nr, k), dtype=torch.float64, device=basis.device)
    term3 = torch.empty((nr, k), dtype=torch.float64, device=basis.device) if want_qih else None
    D = torch.diag(diag_add)
    outer = (basis[:, :, None] * basis[:, None, :]).reshape(m, k * k).contiguous()
    for lo, hi in _row_chunks(nr, m, k, ops):
        o = obs[lo:hi].contiguous()
        c = hi - lo
        H = ops.mm(o, outer).view(c, k, k)
        J = ops.mm((o * cur[lo:hi]).contiguous(), basis)
        Ainv = _spd_inverse(H + D, ops, torch)
        sol = (Ainv * J[:, None, :]).sum(2)
        vec[lo:hi] = J - (H * sol[:, None, :]).sum(2)
        if want_qih:
            QiH = torch.empty_like(H)
            ops.dgemm(Ainv, H, QiH, True, False, k, k, k, batch=c, strides=(k * k, k * k, k * k), lds=(k, k, k))
            term3[lo:hi] = torch.diagonal(H, dim1=1, dim2=2) - (H * QiH).sum(1)
    return vec, term3


def _missing_index(miss, torch):
    """Padded column indices of the missing entries of each row: idx (nr, qmax) and valid mask (nr, qmax)."""
    q = miss.sum(1)
    qmax = int(q.max().item())
    order = torch.argsort((~miss).to(torch.int8), dim=1, stable=True)      # missing columns first, in column order
    idx = order[:, :qmax]
    valid = torch.arange(qmax, device=miss.device)[None, :] < q[:, None]
    return idx, valid


def _all_kept_projector(G, eps_pc, eps_impute, ops, torch):
    """P = U diag((lam - eps_pc) / (lam - eps_pc + eps_impute)) U' WITHOUT an eigen-decomposition, for a sweep in which
    every one of the m components passes the threshold (the first imputation sweeps: the +-2 seed pollutes every
    direction).  Then P is a rational function of G alone,
        P = I - eps_impute (G - (eps_pc - eps_impute) I)^-1,
    i.e. one SPD inverse on the blocked Cholesky / trtri / lauum kernels.  Returns None unless G - eps_pc I is positive
    definite (checked by a Cholesky factorisation: exactly the reference's condition S^2 - eps_pc > 0 for all S)."""
    m = G.shape[0]
    lda = (m + 7) // 8 * 8
    eye = torch.eye(m, dtype=G.dtype, device=G.device)
    W = torch.zeros((2, m, lda), dtype=G.dtype, device=G.device)
    W[0, :, :m] = G - eps_pc * eye
    W[1, :, :m] = G - (eps_pc - eps_impute) * eye
    if m <= ops.small_max_n():
        return None                                   # tiny problems: the eigen path is cheap enough
    Linv, info = ops.potrf_batched(W, m, m)
    if bool((info != 0).any()):
        return None
    ops.trtri_batched(W, Linv, m)
    low = torch.tril(ops.lauum_batched(Linv[1:2], m)[0, :, :m])
    Ainv = low + torch.tril(low, -1).T
    return eye - eps_impute * Ainv


def _impute_schur(Up, sp2, eps_impute, cur, obs, idx, valid, ops, torch, P=None):
    """The ridge update of the missing entries (PCGPwM.py:576-586) through the |missing| x |missing| system.

    The reference solves, per row, the k x k system (H + D) sol = J with H = Up' diag(obs) Up, D = diag(eps / sp2),
    J = Up' (obs * cur) and fills the missing entries with Up sol.  Up has orthonormal columns, so
    H + D = Lam - M' M with Lam = I + D and M = Up[missing rows of this output row], and by the Woodbury identity the
    filled values are  w = (I - P_mm)^-1 P_mo cur_o  with P = Up Lam^-1 Up' (m x m, once per sweep): one GEMM for
    all rows plus a batched Cholesky of size |missing| instead of size k.  Early sweeps keep k ~ m components
    (k^3 per row); |missing| is the missing fraction of m.  P may be handed in (see _all_kept_projector).
    Returns the filled values, (nr, qmax), 0 where padded."""
    if P is None:
        lam_inv = sp2 / (sp2 + eps_impute)
        P = ops.mm((Up * lam_inv).contiguous(), Up.contiguous(), tb=True)
    Z = ops.mm((obs * cur).contiguous(), P)                        # (nr, m): P_{.,o} cur_o for every column
    z = torch.gather(Z, 1, idx) * valid
    nr, qmax = idx.shape
    w = torch.zeros_like(z)
    eye = torch.eye(qmax, dtype=cur.dtype, device=cur.device)
    per_row = 8 * qmax * qmax * 4
    step = max(1, min(nr, _batch_cap(qmax, ops), _CHUNK_BYTES // max(per_row, 1)))
    small = qmax <= ops.small_max_n()                 # the register-resident sweep kernel solves these directly
    for lo in range(0, nr, step):
        hi = min(lo + step, nr)
        ii = idx[lo:hi]
        vv = valid[lo:hi].to(cur.dtype)
        S = P[ii[:, :, None], ii[:, None, :]] * (vv[:, :, None] * vv[:, None, :])
        S = eye - S                                                # padded rows / columns: identity
        if small:
            lda = (qmax + 7) // 8 * 8
            A = torch.zeros((hi - lo, qmax + 1, lda), dtype=cur.dtype, device=cur.device)
            A[:, :qmax, :qmax] = S
            A[:, qmax, :qmax] = z[lo:hi]
            info = ops.spd_solve_small(A, qmax)
            if bool((info != 0).any()):
                raise ValueError('imputation: a Schur system of the missing entries is not positive definite')
            w[lo:hi] = A[:, qmax, :qmax]
        else:
            w[lo:hi] = (_spd_inverse(S, ops, torch) * z[lo:hi, None, :]).sum(2)
    return w * valid


# ------------------------------------------------------------------------------------------------------------------
def _upload(f_np, torch, dev):
    """(n, m) host matrix -> contiguous device tensor; a transposed view of a C-contiguous (m, n) array (what
    PCGPwM_B200.fit holds) is copied as it lies in memory and transposed on the device."""
    if f_np.flags.c_contiguous:
        return torch.from_numpy(f_np).to(dev)
    if f_np.T.flags.c_contiguous:
        return torch.from_numpy(f_np.T).to(dev).T.contiguous()
    return torch.from_numpy(np.ascontiguousarray(f_np)).to(dev)


def standardize(f, eps_pc, eps_impute, n_iter=40):
    """f: (n, m) numpy with NaN/inf for missing.  Returns (standardpcinfo, mof, mofrows); fs / U / S / _lam are device
    tensors until principal_components() converts them."""
    torch = _torch()
    ops = _ops()
    dev = torch.device('cuda', torch.cuda.current_device())
    f_np = np.asarray(f, dtype=np.float64)
    fd = _upload(f_np, torch, dev)
    offset, scale, nmiss = ops.col_stats(fd)                       # PCGPwM.py:553-557
    any_missing = bool(nmiss.sum().item() > 0)
    if bool((scale == 0).any()):
        raise ValueError("You have a row that is non-varying.")
    fs, miss_all = ops.standardize(fd, offset, scale, nmiss, want_mask=any_missing)    # + seed, PCGPwM.py:565-571
    del fd
    note = {}
    mof_np = mofrows_np = None
    if any_missing:
        rows = torch.nonzero(miss_all.any(dim=1)).squeeze(1)
        miss = miss_all[rows].bool()
        obs = (~miss).to(torch.float64)
        idx, valid = _missing_index(miss, torch)
        warm, pad, retry_block = None, 8, False
        cap = ops.syevj_max_p()
        all_kept = True                                           # until a sweep says otherwise (it starts that way)
        for sweep in range(n_iter):                               # PCGPwM.py:573-587
            G = ops.mm(fs, fs, ta=True)
            cur = fs[rows]
            if all_kept and G.shape[0] > idx.shape[1]:
                P = _all_kept_projector(G, eps_pc, eps_impute, ops, torch)
                if P is not None:                                 # every component kept: no decomposition needed
                    note['all_kept_sweeps'] = note.get('all_kept_sweeps', 0) + 1
                    note.setdefault('kept_per_sweep', []).append(int(G.shape[0]))
                    w = _impute_schur(None, None, eps_impute, cur, obs, idx, valid, ops, torch, P=P)
                    fs[rows] = cur.scatter(1, idx, torch.where(valid, w, torch.gather(cur, 1, idx)))
                    continue
                all_kept = False
            if warm is not None or sweep == 0 or retry_block:
                U, lam = _eig_above(G, eps_pc, warm, ops, torch, note)
            else:                                                 # the last sweep kept more components than a block holds
                U, lam = _full_eig(G, torch)
                note['library_eigh'] = note.get('library_eigh', 0) + 1
            keep = (lam - eps_pc) > 0
            Up = U[:, keep].contiguous()
            sp2 = lam[keep] - eps_pc
            nk = int(Up.shape[1])
            note.setdefault('kept_per_sweep', []).append(nk)
            warm = U[:, :min(nk + pad, U.shape[1])] if nk + pad <= cap else None
            # the count collapses within a few sweeps (C3: 1000, 1000, 1000, 675, 261, 35, 33, ...): once it has at least
            # halved towards the block size, try the block iteration (two cheap steps tell if the block is too small)
            retry_block = warm is None and nk <= 4 * cap
            if nk > idx.shape[1]:                                 # more components than missing entries per row
                w = _impute_schur(Up, sp2, eps_impute, cur, obs, idx, valid, ops, torch)
                filled = cur.scatter(1, idx, torch.where(valid, w, torch.gather(cur, 1, idx)))
                fs[rows] = filled
            else:
                vec, _ = _ridge_rows(Up, obs, cur, eps_impute / sp2, ops, torch)
                filled = ops.mm(vec, (Up * (sp2 / eps_impute)).contiguous(), tb=True)
                fs[rows] = torch.where(miss, filled, cur)
        mof_np = miss_all.cpu().numpy().astype(bool)
        mofrows_np = rows.cpu().numpy()
        U, lam = _eig_above(ops.mm(fs, fs, ta=True), eps_pc, warm, ops, torch, note)
    else:
        U, lam = _eig_above(ops.mm(fs, fs, ta=True), eps_pc, None, ops, torch, note)
    Up = U[:, (lam - eps_pc) > 0].contiguous()
    resid = fs - ops.mm(ops.mm(fs, Up), Up, tb=True)
    offset_np, scale_np = offset.cpu().numpy(), scale.cpu().numpy()
    extravar = (resid ** 2).mean(0).cpu().numpy() * scale_np ** 2   # PCGPwM.py:594
    info = LazyHostDict({'offset': offset_np, 'scale': scale_np, 'fs': fs, 'U': U,
                         'S': torch.sqrt(torch.clamp(lam, min=0.0)), 'extravar': extravar, '_lam': lam, '_eig_path': note})
    return info, mof_np, mofrows_np


def principal_components(spc, mof, mofrows, eps_pc, eps_impute):
    """PC loadings / scores / per-row imputation variance (PCGPwM.py:608-654) from a device standardpcinfo.

    Returns a dict of numpy arrays and converts spc's small device tensors to numpy in place; fs is deferred."""
    torch = _torch()
    ops = _ops()
    fs, U, lam = spc['fs'], spc['U'], spc.pop('_lam')
    keep = (lam - eps_pc) > 0
    basis = U[:, keep].contiguous()
    pcw = torch.sqrt(lam[keep] - eps_pc)
    pcstdvar = torch.zeros((fs.shape[0], basis.shape[1]), dtype=torch.float64, device=fs.device)
    if mof is not None:
        rows = torch.from_numpy(mofrows).to(fs.device)
        obs = torch.from_numpy(~mof[mofrows]).to(fs.device).to(torch.float64)
        gain = pcw ** 2 / eps_impute + 1
        vec, term3 = _ridge_rows(basis, obs, fs[rows], eps_impute / pcw ** 2, ops, torch, want_qih=True)
        fs[rows] = ops.mm((gain * vec).contiguous(), basis, tb=True)
        pcstdvar[rows] = 1 - gain * term3
    effn = torch.sum(torch.clamp(1 - pcstdvar, 0, 1))
    pct = (basis * pcw / torch.sqrt(effn)).contiguous()
    out = {'pcw': pcw, 'pcto': basis.clone(), 'pct': pct, 'pcti': basis * (torch.sqrt(effn) / pcw),
           'pc': ops.mm(fs, pct), 'unscaled_pcstdvar': pcstdvar}
    keys = list(out)
    host = ops.to_host([out[k] for k in keys] + [spc[k] for k in ('U', 'S')])        # pinned copies, one sync
    out = dict(zip(keys, host[:len(keys)]))
    for k, v in zip(('U', 'S'), host[len(keys):]):
        spc[k] = v
    spc.defer('fs', fs)                                   # n x m: copied to the host only if somebody reads it
    return out
