"""Device-tensor level wrappers over the C ABI (torch supplies memory and streams only).

All tensors are float64 CUDA tensors (int32 where noted), contiguous.  Every function checks that
it is handed CUDA memory: there is no CPU code path.
"""
import numpy as np

from . import _lib
from ._lib import lib, check

_workspaces = {}


def _torch():
    return _lib.require_cuda()


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream(torch):
    """Raw handle of torch's current stream (every launch passes it through the C ABI)."""
    try:                                            # the Python Stream object costs ~13 us to build; this is ~0.3 us
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:
        return torch.cuda.current_stream().cuda_stream


_PIN_LIMIT = 4 << 30        # results above this size are copied through pageable memory


def to_host(tensors):
    """Device tensors (or None) -> numpy arrays with ONE synchronisation.

    Each result is copied asynchronously into pinned host memory from torch's caching host allocator (a pageable
    .cpu() copy runs at ~2.5 GB/s here, a pinned one at PCIe speed); the numpy array owns that block, which goes
    back to the allocator's cache when the caller drops the array."""
    torch = _torch()
    staged = []
    for t in tensors:
        if t is None:
            staged.append(None)
            continue
        t = t.contiguous()
        h = None
        if t.is_cuda and 0 < t.numel() * t.element_size() <= _PIN_LIMIT:
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
        staged.append((t, h))
    torch.cuda.current_stream().synchronize()
    return [None if s is None else (s[1].numpy() if s[1] is not None else s[0].cpu().numpy()) for s in staged]


def _f64(torch, a, device=None):
    """numpy / tensor -> contiguous float64 CUDA tensor."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    t = t.to(device=device or torch.device('cuda', torch.cuda.current_device()), dtype=torch.float64)
    return t.contiguous()


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.SurmiseB200Error('surmise_b200 ops need CUDA tensors (no CPU fallback)')


def workspace(nbytes):
    """Grow-only per-device scratch buffer (the library itself never allocates)."""
    torch = _torch()
    dev = torch.cuda.current_device()
    ws = _workspaces.get(dev)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _workspaces.pop(dev, None)
        ws = torch.empty(int(nbytes * 1.1) + 1024, dtype=torch.uint8, device=torch.device('cuda', dev))
        _workspaces[dev] = ws
    return ws


def release_workspaces():
    _workspaces.clear()
    _wb_scratch.clear()


# ----------------------------------------------------------------------------------------------
def covmat_device(x1, x2, gammav, grad_mode=0):
    """R (n1,n2) [, dR planes (D,n1,n2)] on device.  Replaces matern_covmat.covmat (pyx:27-64)."""
    torch = _torch()
    _need_cuda(x1, x2, gammav)
    n1, d = x1.shape
    n2 = x2.shape[0]
    R = torch.empty((n1, n2), dtype=torch.float64, device=x1.device)
    dR = None
    if grad_mode:
        planes = d + 1 if grad_mode == 1 else d
        dR = torch.empty((planes, n1, n2), dtype=torch.float64, device=x1.device)
    check(lib.sb200_matern_covmat(_ptr(x1), n1, _ptr(x2), n2, d, _ptr(gammav), _ptr(R), n2, _ptr(dR), grad_mode,
                                  _stream(torch)), 'sb200_matern_covmat')
    return R, dR


def covmat(x1, x2, gammav, return_gradhyp=False, return_gradx1=False):
    """numpy-in / numpy-out drop-in for surmise.emulationsupport.matern_covmat.covmat."""
    torch = _torch()
    gam = np.asarray(gammav, dtype=np.float64)
    d = gam.shape[0] - 1
    a = np.asarray(x1, dtype=np.float64)
    b = np.asarray(x2, dtype=np.float64)
    a = a.reshape(1, d) if a.ndim < 2 else a          # pyx:29-32
    b = b.reshape(1, d) if b.ndim < 2 else b
    mode = 1 if return_gradhyp else (2 if return_gradx1 else 0)
    R, dR = covmat_device(_f64(torch, a), _f64(torch, b), _f64(torch, gam), mode)
    R_h, dR_h = to_host([R, dR if mode else None])
    if mode:
        return R_h, dR_h.transpose(1, 2, 0)           # a transposed view of the (., n1, n2) buffer, as pyx:52
    return R_h


# ----------------------------------------------------------------------------------------------
class GPData:
    """Device-resident (theta, g, gvar) of one or several GP problems sharing n and d.

    theta: (n,d) shared or (nb,n,d); g: (n) or (nb,n); gvar: same shape as g or None.
    """

    def __init__(self, theta, g, gvar=None):
        torch = _torch()
        self.theta = _f64(torch, theta)
        self.g = _f64(torch, g)
        self.gvar = None if gvar is None else _f64(torch, gvar)
        self.n, self.d = self.theta.shape[-2], self.theta.shape[-1]
        self.stride_theta = 0 if self.theta.dim() == 2 else self.n * self.d
        self.stride_g = 0 if self.g.dim() == 1 else self.n
        self.stride_gvar = 0 if (self.gvar is None or self.gvar.dim() == 1) else self.n
        self.nb_fixed = None
        if self.theta.dim() == 3:
            self.nb_fixed = self.theta.shape[0]
        if self.g.dim() == 2:
            self.nb_fixed = self.g.shape[0]


_io_cache = {}


def _io_buffers(torch, dev, nb, p):
    """Per-(device, nb, p) staging: pinned host hyp buffer, one flat device result buffer and its pinned mirror.

    Layout of the result buffer (doubles): nll[nb] | grad[nb*p] | info[nb] (int32 in the low half of each slot)."""
    key = (dev.index, nb, p)
    buf = _io_cache.get(key)
    if buf is None:
        if len(_io_cache) > 64:
            _io_cache.clear()
        total = nb * (p + 2)
        buf = {'hyp_h': torch.empty((nb, p), dtype=torch.float64).pin_memory(),
               'hyp_d': torch.empty((nb, p), dtype=torch.float64, device=dev),
               'res_d': torch.empty(total, dtype=torch.float64, device=dev),
               'res_h': torch.empty(total, dtype=torch.float64).pin_memory()}
        _io_cache[key] = buf
    return buf


def gp_nll_grad(data, hyp, regmean, regstd, sig2ofconst=0.01, want_grad=True, return_device=False):
    """Batched PCGPwM.__negloglik (+ __negloglikgrad) (PCGPwM.py:850-907).

    hyp: (nb, d+3) numpy or tensor.  Returns numpy (nll (nb,), grad (nb,p) or None, info (nb,)), or with
    return_device=True one device tensor (nb, p+2) = [nll | grad | info] without synchronising.
    The numpy path stages hyp through pinned memory and reads everything back with ONE device-to-host copy.
    """
    torch = _torch()
    n, d = data.n, data.d
    p = d + 3
    dev = data.theta.device
    if isinstance(hyp, torch.Tensor):
        hyp_t = _f64(torch, hyp)
        nb = hyp_t.shape[0]
        io = None
    else:
        hyp_np = np.atleast_2d(np.asarray(hyp, dtype=np.float64))
        nb = hyp_np.shape[0]
        io = _io_buffers(torch, dev, nb, p)
        io['hyp_h'].numpy()[...] = hyp_np
        io['hyp_d'].copy_(io['hyp_h'], non_blocking=True)
        hyp_t = io['hyp_d']
    if hyp_t.shape[1] != p:
        raise ValueError('hyp must have d+3 = %d columns' % p)
    if data.nb_fixed is not None and data.nb_fixed != nb:
        raise ValueError('batch mismatch between data (%d) and hyp (%d)' % (data.nb_fixed, nb))
    rm = regmean if (isinstance(regmean, torch.Tensor) and regmean.is_cuda) else _f64(torch, regmean)
    rs = regstd if (isinstance(regstd, torch.Tensor) and regstd.is_cuda) else _f64(torch, regstd)
    if io is None:
        io = _io_buffers(torch, dev, nb, p)
    res = io['res_d']
    base = res.data_ptr()
    nll_ptr, grad_ptr, info_ptr = base, base + 8 * nb, base + 8 * nb * (p + 1)
    nbytes = lib.sb200_gp_nll_grad_workspace(nb, n, d, int(want_grad))
    ws = workspace(nbytes)
    check(lib.sb200_gp_nll_grad(nb, n, d, _ptr(data.theta), data.stride_theta, _ptr(data.g), data.stride_g,
                                _ptr(data.gvar), data.stride_gvar, _ptr(hyp_t),
                                _ptr(rm), _ptr(rs), float(sig2ofconst), int(want_grad), nll_ptr,
                                grad_ptr if want_grad else 0, info_ptr, _ptr(ws), ws.numel(), _stream(torch)),
          'sb200_gp_nll_grad')
    if return_device:
        info = res[nb * (p + 1):].view(torch.int32)[:nb].to(torch.float64)
        cols = [res[:nb, None]] + ([res[nb:nb * (p + 1)].view(nb, p)] if want_grad else []) + [info[:, None]]
        return torch.cat(cols, dim=1)
    io['res_h'].copy_(res, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    host = io['res_h'].numpy()
    nll = host[:nb].copy()
    info = host[nb * (p + 1):].view(np.int32)[:nb].astype(np.int64)
    grad = host[nb:nb * (p + 1)].reshape(nb, p).copy() if want_grad else None
    return nll, grad, info


def spd_solve_small(A, n):
    """In-place batched SPD solve on the sweep kernel: A (batch, n+1, lda) = [lower(S); z'] -> [lower(S^-1); (S^-1 z)'].

    Returns info (device int32, batch)."""
    torch = _torch()
    _need_cuda(A)
    batch, rows, lda = A.shape
    if rows != n + 1 or not A.is_contiguous():
        raise ValueError('spd_solve_small: A must be a contiguous (batch, n+1, lda) tensor')
    info = torch.empty(batch, dtype=torch.int32, device=A.device)
    check(lib.sb200_spd_solve_small(batch, n, _ptr(A), lda, rows * lda, 0, _ptr(info), _stream(torch)),
          'sb200_spd_solve_small')
    return info


def force_blocked_path(on):
    """Testing hook: send n <= small_max_n() problems through the blocked multi-launch path too."""
    lib.sb200_gp_force_blocked_path(int(bool(on)))


def small_max_n():
    return int(lib.sb200_gp_nll_small_max_n())


def gp_factorize(data, hyp, sig2ofconst=0.01):
    """Full-data stage of __fitGP1d for a batch (PCGPwM.py:810-846).

    Returns device tensors Linv (nb,n,ldl), pw (nb,n), sig2 (nb,), logdet_half (nb,) and numpy info.
    """
    torch = _torch()
    hyp_t = _f64(torch, np.atleast_2d(hyp) if not isinstance(hyp, torch.Tensor) else hyp)
    nb = hyp_t.shape[0]
    n, d = data.n, data.d
    dev = hyp_t.device
    ldl = (n + 7) // 8 * 8
    Linv = torch.empty((nb, n, ldl), dtype=torch.float64, device=dev)
    pw = torch.empty((nb, n), dtype=torch.float64, device=dev)
    sig2 = torch.empty(nb, dtype=torch.float64, device=dev)
    ldh = torch.empty(nb, dtype=torch.float64, device=dev)
    info = torch.empty(nb, dtype=torch.int32, device=dev)
    nbytes = lib.sb200_gp_factorize_workspace(nb, n)
    ws = workspace(nbytes)
    check(lib.sb200_gp_factorize(nb, n, d, _ptr(data.theta), data.stride_theta, _ptr(data.g), data.stride_g,
                                 _ptr(data.gvar), data.stride_gvar, _ptr(hyp_t),
                                 float(sig2ofconst), _ptr(Linv), ldl, n * ldl, _ptr(pw), _ptr(sig2), _ptr(ldh),
                                 _ptr(info), _ptr(ws), ws.numel(), _stream(torch)), 'sb200_gp_factorize')
    return Linv, pw, sig2, ldh, info.cpu().numpy()


def gp_weights(Linv, fidx, g, sig2ofconst=0.01):
    """pw (k,n), sig2 (k) for PC scores g (k,n) through the shared factors Linv[fidx[j]] (PCGPwM.py:823-846)."""
    torch = _torch()
    _need_cuda(Linv, fidx, g)
    K, n = g.shape
    ldl = Linv.shape[2]
    pw = torch.empty((K, n), dtype=torch.float64, device=g.device)
    sig2 = torch.empty(K, dtype=torch.float64, device=g.device)
    ytmp = torch.empty((K, n), dtype=torch.float64, device=g.device)
    check(lib.sb200_gp_weights(K, n, _ptr(Linv), ldl, n * ldl, _ptr(fidx), _ptr(g.contiguous()), float(sig2ofconst),
                               _ptr(pw), _ptr(sig2), _ptr(ytmp), _stream(torch)), 'sb200_gp_weights')
    return pw, sig2


def gp_predict(state, thetanew, want_grad=False, out=None, population=None, out_ld=0):
    """PC-space prediction (PCGPwM.py:219-270) from a device state dict (see emulation module).

    Returns device tensors predvecs, predvars (B,k) and, if want_grad, gradients (B,k,d).
    out: optional (pv, pvar, dpv, dpvar) to write into (callers that evaluate the same shape repeatedly).
    population: size of the caller's WHOLE theta population when thetanew is a chunk or a rank's slice of it (default:
    thetanew itself).  Only a population of one theta with d == 1 takes the reference's single-(1-nug) gradient
    (PCGPwM.py:236, 262-268), so results do not depend on chunking or on the number of ranks.
    out_ld: with `out`, the row length (in components) of the output buffers when they are wider than this state's k
    (PC-sharded prediction writes each rank's components into a padded (B, kmax[, d]) block).
    """
    torch = _torch()
    tn = _f64(torch, thetanew)
    B, d = tn.shape
    K = state['pw'].shape[0]
    nf, n, ldl = state['Linv'].shape
    nu = state['hypcov'].shape[0]
    dev = tn.device
    if out is not None:
        pv, pvar, dpv, dpvar = out
    else:
        pv = torch.empty((B, K), dtype=torch.float64, device=dev)
        pvar = torch.empty((B, K), dtype=torch.float64, device=dev)
        dpv = torch.empty((B, K, d), dtype=torch.float64, device=dev) if want_grad else None
        dpvar = torch.empty((B, K, d), dtype=torch.float64, device=dev) if want_grad else None
    nbytes = lib.sb200_gp_predict_workspace(K, nf, nu, n, d, B, int(want_grad))
    ws = workspace(nbytes)
    total = B if population is None else int(population)
    grad_mode = (2 if (total == 1 and d == 1) else 1) if want_grad else 0
    check(lib.sb200_gp_predict(K, nf, nu, n, d, B, _ptr(state['theta']), _ptr(tn), _ptr(state['hypcov']),
                               _ptr(state['fu']), _ptr(state['fidx']), _ptr(state['nug']), _ptr(state['sig2']),
                               _ptr(state['Linv']), ldl, n * ldl, _ptr(state['pw']), grad_mode, int(out_ld), _ptr(pv),
                               _ptr(pvar), _ptr(dpv), _ptr(dpvar), _ptr(ws), ws.numel(), _stream(torch)),
          'sb200_gp_predict')
    return pv, pvar, dpv, dpvar


def pc_backproject(phi, offset, extravar, pv, pvar, dpv=None, dpvar=None, want_covx=True):
    """mean, var (m,B) [, covxhalf (m,B,k), mean_grad (m,B,d), covxhalf_grad (m,B,k,d)] (PCGPwM.py:273-327)."""
    torch = _torch()
    _need_cuda(phi, offset, extravar, pv, pvar)
    m, K = phi.shape
    B = pv.shape[0]
    d = dpv.shape[2] if dpv is not None else 1
    dev = pv.device
    mean = torch.empty((m, B), dtype=torch.float64, device=dev)
    var = torch.empty((m, B), dtype=torch.float64, device=dev)
    cxh = torch.empty((m, B, K), dtype=torch.float64, device=dev) if want_covx else None
    mg = torch.empty((m, B, d), dtype=torch.float64, device=dev) if dpv is not None else None
    cg = torch.empty((m, B, K, d), dtype=torch.float64, device=dev) if (dpvar is not None and want_covx) else None
    check(lib.sb200_pc_backproject(m, K, B, d, _ptr(phi), _ptr(offset), _ptr(extravar), _ptr(pv), _ptr(pvar),
                                   _ptr(dpv), _ptr(dpvar), _ptr(mean), _ptr(var), _ptr(cxh), _ptr(mg), _ptr(cg),
                                   _stream(torch)), 'sb200_pc_backproject')
    return mean, var, cxh, mg, cg


def woodbury_trigger(consts, pvar, row_valid=None):
    """Device int32 flag: does the unmodelled-variance test (dbw.py:277-279) fire for this batch?"""
    torch = _torch()
    _need_cuda(pvar, consts['phiT'])
    B, K = pvar.shape
    fired = torch.empty(1, dtype=torch.int32, device=pvar.device)
    check(lib.sb200_woodbury_trigger(B, K, consts['phiT'].shape[1], _ptr(pvar), _ptr(consts['phiT']),
                                     _ptr(consts['extravar']), _ptr(fired), _ptr(row_valid), 0, 0, 0, _stream(torch)),
          'sb200_woodbury_trigger')
    return fired


def woodbury_loglik(consts, pv, pvar, dpv=None, dpvar=None, variant=-1, out=None, row_valid=None, segments=None,
                    seg_pos=None):
    """Batched directbayeswoodbury log-likelihood (+ gradient) (dbw.py:261-383).

    consts: dict of device tensors phiT (k,m), resid0 (m), extravar (m), sinv (2,m), G (2,k,k).
    row_valid: optional int32 (B,) mask of the rows that take part in the batch-wide variance test.
    segments: None for dense (B,k) inputs, or (seg_k, seg_stride) when pv / pvar / dpv / dpvar are the FIRST blocks
    (B, seg_k[, d]) of an all-gathered PC-sharded prediction whose blocks lie seg_stride doubles apart; k is then
    consts['phiT'].shape[0] (see include/surmise_b200.h).  seg_pos: optional int32 device tensor (k,) with the padded slot
    of every real component; consts are then the ordinary unpadded arrays.
    Returns device tensors ll (B,), dll (B,d) or None, fired (int32 scalar tensor).
    """
    torch = _torch()
    _need_cuda(pv, pvar, consts['phiT'])
    B, K = pv.shape
    seg_k, seg_stride = (0, 0) if segments is None else segments
    if segments is not None:
        K = consts['phiT'].shape[0]
    m = consts['phiT'].shape[1]
    want_grad = dpv is not None
    d = dpv.shape[2] if want_grad else 1
    dev = pv.device
    if out is not None:                              # (ll, dll, fired) preallocated by the caller
        ll, dll, fired = out
    else:
        ll = torch.empty(B, dtype=torch.float64, device=dev)
        dll = torch.empty((B, d), dtype=torch.float64, device=dev) if want_grad else None
        fired = torch.empty(1, dtype=torch.int32, device=dev)
    # k x k systems beyond shared memory (k > 112 at m <= 3000) work in a global scratch buffer of their own
    nbytes = lib.sb200_woodbury_workspace(B, K, m)
    ws = _woodbury_scratch(torch, dev, nbytes) if nbytes else None
    check(lib.sb200_woodbury_loglik_ws(B, K, m, d, _ptr(pv), _ptr(pvar), _ptr(dpv), _ptr(dpvar), _ptr(consts['phiT']),
                                       _ptr(consts['resid0']), _ptr(consts['extravar']), _ptr(consts['sinv']),
                                       _ptr(consts['G']), int(want_grad), int(variant), _ptr(ll), _ptr(dll), _ptr(fired),
                                       _ptr(row_valid), int(seg_k), int(seg_stride), _ptr(seg_pos), _ptr(ws),
                                       ws.numel() if ws is not None else 0, _stream(torch)),
          'sb200_woodbury_loglik_ws')
    return ll, dll, fired


def woodbury_loglik_mspace(meanT, varT, cxhT, y, obsvar=None):
    """directbayeswoodbury.loglik as written (dbw.py:261-306) for an emulator outside the PCGPwM family.

    meanT (B,m), varT (B,m), cxhT (B,m,k): that emulator's predictive mean / variance / covxhalf at the proposals, on the
    device, proposal-major.  obsvar None: returns the device int32 flag of the batch-wide unmodelled-variance test
    (dbw.py:277-279); obsvar (m,) device tensor: returns ll (B,) for that observation variance."""
    torch = _torch()
    _need_cuda(meanT, varT, cxhT)
    B, m, K = cxhT.shape
    dev = cxhT.device
    if obsvar is None:
        fired = torch.empty(1, dtype=torch.int32, device=dev)
        check(lib.sb200_woodbury_loglik_mspace(B, K, m, _ptr(meanT), _ptr(varT), _ptr(cxhT), None, None, 1, None,
                                               _ptr(fired), _stream(torch)), 'sb200_woodbury_loglik_mspace')
        return fired
    sinv = (1.0 / obsvar).contiguous()
    ll = torch.empty(B, dtype=torch.float64, device=dev)
    check(lib.sb200_woodbury_loglik_mspace(B, K, m, _ptr(meanT), _ptr(varT), _ptr(cxhT), _ptr(y), _ptr(sinv), 0, _ptr(ll),
                                           None, _stream(torch)), 'sb200_woodbury_loglik_mspace')
    return ll


_wb_scratch = {}


def _woodbury_scratch(torch, dev, nbytes):
    """Grow-only per-device buffer for the Woodbury kernel's global-scratch path (separate from `workspace`, whose
    contents a prediction running ahead on the same stream may still be using)."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    ws = _wb_scratch.get(key)
    if ws is None or ws.numel() < nbytes:
        _wb_scratch.pop(key, None)
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        _wb_scratch[key] = ws
    return ws


def woodbury_force_global(on):
    """Testing switch: k x k systems in global scratch at any k (the path large k takes)."""
    lib.sb200_woodbury_force_global(int(bool(on)))


def ptlmc_step(state, mode, ll=None, dll=None):
    """One launch of the device-resident PT-LMC step (PTLMC.py:203-273; see include/surmise_b200.h).

    state: dict built by utilitiesmethods.PTLMC_B200 -- device tensors thetac, fval, dfval, thetap, z, adjrho, temps, lo,
    hi, hc, cov0, scal, kstep, save, valid and the ints B, d, numtemps, tune, nsave, seed, plus taracc."""
    torch = _torch()
    s = state
    check(lib.sb200_ptlmc_step(s['B'], s['d'], s['numtemps'], s['tune'], s['nsave'], int(mode), int(s['seed']),
                               float(s['taracc']), _ptr(s['thetac']), _ptr(s['fval']), _ptr(s['dfval']), _ptr(s['thetap']),
                               _ptr(s['z']), _ptr(s['adjrho']), _ptr(ll), _ptr(dll), _ptr(s['temps']), _ptr(s['lo']),
                               _ptr(s['hi']), _ptr(s['hc']), _ptr(s['cov0']), _ptr(s['scal']), _ptr(s['kstep']),
                               _ptr(s['save']), _ptr(s['valid']), _stream(torch)), 'sb200_ptlmc_step')


# ----------------------------------------------------------------------------------------------
# building blocks (exported by the C ABI for testing / reuse)
def dgemm(A, B, C, a_kc, b_kc, M, N, K, alpha=1.0, beta=0.0, flags=0, tile=-1, batch=1, strides=(0, 0, 0),
          lds=None):
    torch = _torch()
    _need_cuda(A, B, C)
    lda, ldb, ldc = lds
    check(lib.sb200_dgemm(int(a_kc), int(b_kc), M, N, K, float(alpha), _ptr(A), lda, strides[0], _ptr(B), ldb,
                          strides[1], float(beta), _ptr(C), ldc, strides[2], batch, flags, tile, _stream(torch)),
          'sb200_dgemm')


def potrf_batched(A, n, mrows):
    """A: (batch, mrows, lda) in place; returns (Linv with diagonal blocks only, info numpy)."""
    torch = _torch()
    _need_cuda(A)
    batch, _, lda = A.shape
    Linv = torch.empty((batch, n, lda), dtype=torch.float64, device=A.device)
    info = torch.empty(lib.sb200_potrf_info_ints(batch, mrows), dtype=torch.int32, device=A.device)   # [info | scratch]
    check(lib.sb200_potrf_batched(batch, n, mrows, _ptr(A), lda, A.stride(0), _ptr(Linv), lda, n * lda, _ptr(info),
                                  _stream(torch)), 'sb200_potrf_batched')
    return Linv, info[:batch]


def trtri_batched(L, Linv, n):
    torch = _torch()
    batch = L.shape[0]
    te = lib.sb200_trtri_tmp_elems(n)
    tmp = torch.empty((batch, max(te, 1)), dtype=torch.float64, device=L.device)
    check(lib.sb200_trtri_batched(batch, n, _ptr(L), L.shape[2], L.stride(0), _ptr(Linv), Linv.shape[2],
                                  Linv.stride(0), _ptr(tmp), te, _stream(torch)), 'sb200_trtri_batched')


def lauum_batched(Linv, n):
    torch = _torch()
    batch, _, ldl = Linv.shape
    Cm = torch.zeros((batch, n, ldl), dtype=torch.float64, device=Linv.device)
    check(lib.sb200_lauum_batched(batch, n, _ptr(Linv), ldl, Linv.stride(0), _ptr(Cm), ldl, Cm.stride(0),
                                  _stream(torch)), 'sb200_lauum_batched')
    return Cm


# ----------------------------------------------------------------------------------------------
# standardisation / PCA stage (PCGPwM.py:533-654)
def mm(A, B, ta=False, tb=False, alpha=1.0, beta=0.0, out=None, flags=0):
    """op(A) @ op(B) for contiguous 2-D float64 CUDA tensors on the DMMA GEMM (sb200_dgemm); ta / tb transpose."""
    torch = _torch()
    _need_cuda(A, B)
    M, K = (A.shape[1], A.shape[0]) if ta else (A.shape[0], A.shape[1])
    K2, N = (B.shape[1], B.shape[0]) if tb else (B.shape[0], B.shape[1])
    if K != K2:
        raise ValueError('mm: inner dimensions %d and %d differ' % (K, K2))
    if not (A.is_contiguous() and B.is_contiguous()):
        raise ValueError('mm: operands must be contiguous')
    if out is None:
        out = torch.empty((M, N), dtype=torch.float64, device=A.device)
    if M == 0 or N == 0:
        return out
    if K == 0:
        return out.zero_() if beta == 0.0 else out.mul_(beta)
    check(lib.sb200_dgemm(int(not ta), int(tb), M, N, K, float(alpha), _ptr(A), A.shape[1], 0, _ptr(B), B.shape[1], 0,
                          float(beta), _ptr(out), out.stride(0), 0, 1, int(flags), -1, _stream(torch)), 'sb200_dgemm')
    return out


def col_stats(f):
    """offset (m), scale (m), nmiss (m, int32) of a (n, m) device matrix with non-finite cells missing."""
    torch = _torch()
    _need_cuda(f)
    n, m = f.shape
    offset = torch.empty(m, dtype=torch.float64, device=f.device)
    scale = torch.empty(m, dtype=torch.float64, device=f.device)
    nmiss = torch.empty(m, dtype=torch.int32, device=f.device)
    nbytes = lib.sb200_pca_col_stats_workspace(n, m)
    ws = workspace(nbytes)
    check(lib.sb200_pca_col_stats(n, m, _ptr(f), f.stride(0), _ptr(offset), _ptr(scale), _ptr(nmiss), _ptr(ws), nbytes,
                                  _stream(torch)), 'sb200_pca_col_stats')
    return offset, scale, nmiss


def standardize(f, offset, scale, nmiss=None, want_mask=False):
    """fs = (f - offset) / scale (n, m); want_mask: also the uint8 mask of the missing cells, which are seeded."""
    torch = _torch()
    _need_cuda(f, offset, scale)
    n, m = f.shape
    fs = torch.empty((n, m), dtype=torch.float64, device=f.device)
    miss = torch.empty((n, m), dtype=torch.uint8, device=f.device) if want_mask else None
    check(lib.sb200_pca_standardize(n, m, _ptr(f), f.stride(0), _ptr(offset), _ptr(scale), _ptr(fs), m, _ptr(miss),
                                    _ptr(nmiss), _stream(torch)), 'sb200_pca_standardize')
    return fs, miss


def sumsq(x):
    """Device scalar tensor: sum of squares of a contiguous float64 tensor (deterministic)."""
    torch = _torch()
    _need_cuda(x)
    if not x.is_contiguous():
        x = x.contiguous()
    out = torch.empty(1, dtype=torch.float64, device=x.device)
    scratch = torch.empty(1024, dtype=torch.float64, device=x.device)
    check(lib.sb200_pca_sumsq(x.numel(), _ptr(x), _ptr(out), _ptr(scratch), _stream(torch)), 'sb200_pca_sumsq')
    return out


def syevj_max_p():
    return int(lib.sb200_syevj_max_p())


def syevj(H, max_sweeps=40):
    """Eigen-decomposition of a symmetric (p, p) or batch of (b, p, p) device matrices, p <= syevj_max_p():
    W descending (.., p), Z (.., p, p) with the eigenvectors in its columns, sweeps (b) int32."""
    torch = _torch()
    _need_cuda(H)
    Hb = H if H.dim() == 3 else H.unsqueeze(0)
    Hb = Hb.contiguous()
    b, p, _ = Hb.shape
    W = torch.empty((b, p), dtype=torch.float64, device=H.device)
    Z = torch.empty((b, p, p), dtype=torch.float64, device=H.device)
    sweeps = torch.empty(b, dtype=torch.int32, device=H.device)
    check(lib.sb200_syevj_batched(b, p, _ptr(Hb), p, p * p, _ptr(W), p, _ptr(Z), p, p * p, int(max_sweeps), _ptr(sweeps),
                                  _stream(torch)), 'sb200_syevj_batched')
    if H.dim() == 3:
        return W, Z, sweeps
    return W[0], Z[0], sweeps
