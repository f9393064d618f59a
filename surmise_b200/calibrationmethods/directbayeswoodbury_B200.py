"""directbayeswoodbury_B200 -- surmise calibration method with the Woodbury log-likelihood on a B200.

Drop-in sibling of surmise/calibrationmethods/directbayeswoodbury.py behind the same plugin contract
(`fit(fitinfo, emu, x, y, **args)`, optional `predict`, `thetarnd`, `thetalpdf`;
surmise/calibration.py:196-234, 271-272, 506-577):
`calibrator(emu, y, x, thetaprior, method='directbayeswoodbury_B200', yvar=..., args={'sampler': ...})`.

The sampler plugins of surmise are used unmodified: they call the log-posterior closure once per step
with the whole population as one (B, d) array; each such call is ONE fused device evaluation
(cross-covariance -> two triangular DMMA products -> PC-space reduction -> batched Woodbury kernel).
Nothing of size m x B x k x d is ever formed (the reference builds covxhalf_gradtheta of that size,
directbayeswoodbury.py:319-324).

The fused path needs an emulator fitted with method='PCGPwM_B200' (its device state is read from emu._info) or the
reference's 'PCGPwM' (adopted onto the device).  Any other emulator (PCGP, indGP, ...) is evaluated through its own
predict and the likelihood as the reference writes it (dbw.py:261-306) on the device, without gradients.
"""
import copy

import os

import numpy as np

from ..emulationmethods import PCGPwM_B200 as _emu_method


def _ops():
    from .. import ops
    return ops


def _device_state(emu):
    info = getattr(emu, '_info', None)
    if isinstance(info, dict) and '_b200' not in info and 'emulist' in info and 'unscaled_pcstdvar' in info:
        _emu_method.adopt(info)                      # an emulator fitted by the reference's PCGPwM: factor on the device
    if not isinstance(info, dict) or 'phi' not in info.get('_b200', {}):        # e.g. an indGP_B200 emulator
        raise ValueError("directbayeswoodbury_B200 needs an emulator fitted with method='PCGPwM_B200' "
                         "(or the reference's 'PCGPwM').")
    return info


def _pc_family(emu):
    """True for an emulator that carries (or can be given, PCGPwM_B200.adopt) PC-space factors on the device."""
    info = getattr(emu, '_info', None)
    if not isinstance(info, dict):
        return False
    if 'phi' in info.get('_b200', {}):
        return True
    return '_b200' not in info and 'emulist' in info and 'unscaled_pcstdvar' in info


def _loglik_any_emulator(fitinfo, emu, theta, y, x):
    """directbayeswoodbury.loglik as written (dbw.py:261-306) for an emulator OUTSIDE the PCGPwM family (PCGP, indGP,
    ...): that emulator's own predict supplies mean / var / covxhalf at the proposals, the batch-wide variance test and
    the Woodbury-form Gaussian density run on the device (sb200_woodbury_loglik_mspace).  No gradient in this form."""
    if 'yvar' not in fitinfo:
        raise ValueError('Must provide yvar in this software.')            # dbw.py:266-269
    ops = _ops()
    torch = ops._torch()
    dev = torch.device('cuda', torch.cuda.current_device())
    pred = emu.predict(x, theta)
    mean, var, cxh = (np.asarray(a, dtype=np.float64) for a in (pred.mean(), pred.var(), pred.covxhalf()))
    m, B = mean.shape
    if cxh.ndim == 2:
        cxh = cxh[:, :, None]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    meanT, varT, cxhT = t(mean.T), t(var.T), t(np.transpose(cxh, (1, 0, 2)))
    obsvar = 1 * np.squeeze(np.asarray(fitinfo['yvar'], dtype=np.float64)) * np.ones(m)
    if int(ops.woodbury_loglik_mspace(meanT, varT, cxhT, None).item()):       # dbw.py:277-284
        old = emu.predict(x)
        obsvar = obsvar + np.mean(np.abs(old.var() - np.sum(np.square(old.covxhalf()), 2)), 1)
    ll = ops.woodbury_loglik_mspace(meanT, varT, cxhT, t(np.squeeze(np.asarray(y, dtype=np.float64)) * np.ones(m)),
                                    t(obsvar))
    return ll.cpu().numpy().reshape(B, 1)


def woodbury_constants(fitinfo, emu, x, y):
    """theta-independent device constants for (x, y, yvar); cached in fitinfo['_b200cal']."""
    if 'yvar' not in fitinfo:
        raise ValueError('Must provide yvar in this software.')            # dbw.py:266-269
    cache = fitinfo.get('_b200cal')
    info = _device_state(emu)
    st = info['_b200']
    # fast path for the sampler loop: the same x / y / yvar objects (with unchanged sums, in case they were edited in
    # place) as last time -> skip row matching and the byte-wise comparison.  The emulator's device state carries a
    # process-wide fit-generation token (a refit / update / adopt gets a new one), and the cache keeps a strong
    # reference to the state it was built from, so a recycled id() can never alias a previous fit.
    ident = (id(x), id(y), id(fitinfo['yvar']), st.get('generation'), float(np.sum(y)), float(np.sum(fitinfo['yvar'])))
    if (cache is not None and cache.get('ident') == ident and cache['refs'][0] is x and cache['refs'][1] is y
            and cache.get('state') is st):
        return cache
    ops = _ops()
    torch = ops._torch()
    y_in = y
    y = np.squeeze(np.asarray(y, dtype=np.float64))
    yvar = 1 * np.squeeze(np.asarray(fitinfo['yvar'], dtype=np.float64))
    xind, xnewind = _emu_method.match_rows(x, info['x'])
    nx = info['x'].shape[0] if x is None else x.shape[0]
    if xind.shape[0] != nx or not np.array_equal(xnewind, np.arange(nx)):
        raise ValueError('Every calibration x must match exactly one row of the emulator x.')
    key = (xind.tobytes(), y.tobytes(), yvar.tobytes(), st.get('generation'))
    if cache is not None and cache['key'] == key and cache.get('state') is st:
        cache['ident'], cache['refs'] = ident, (x, y_in, fitinfo['yvar'])
        return cache
    dev = st['phi'].device
    sel = torch.from_numpy(xind.astype(np.int64)).to(dev)
    phi = st['phi'][sel].cpu().numpy()
    offset = st['offset'][sel].cpu().numpy()
    extravar = st['extravar'][sel].cpu().numpy()
    yv = yvar * np.ones(phi.shape[0])
    sinv = np.stack([1 / yv, 1 / (yv + np.abs(extravar))])                   # dbw.py:280-284 when triggered
    G = np.stack([phi.T @ (phi * s[:, None]) for s in sinv])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cache = {'key': key, 'ident': ident, 'refs': (x, y_in, fitinfo['yvar']), 'state': st, 'phiT': t(phi.T), 'resid0': t(y - offset),
             'extravar': t(extravar), 'sinv': t(sinv), 'G': t(G), 'xind': xind}
    fitinfo['_b200cal'] = cache
    return cache


def loglik_device(fitinfo, emu, theta, y, x, return_grad=False, chunk=4096):
    """Device tensors (ll (B,), dll (B,d) or None, fired) -- the fused evaluation."""
    ops = _ops()
    torch = ops._torch()
    consts = woodbury_constants(fitinfo, emu, x, y)
    info = _device_state(emu)
    tn = ops._f64(torch, np.atleast_2d(theta) if not isinstance(theta, torch.Tensor) else theta)
    pv, pvar, dpv, dpvar = _emu_method.predict_pcspace(info, tn, return_grad, chunk=chunk)
    return ops.woodbury_loglik(consts, pv, pvar, dpv, dpvar)


_closure_buffers = {}


def _buffers(torch, dev, B, d, K, return_grad):
    """Per-shape staging of the sampler closure: pinned theta, device theta / predictions / results, pinned results.

    A sampler evaluates the same population size thousands of times; allocating eight small tensors per call cost
    more host time than launching the kernels."""
    key = (dev.index, B, d, K, bool(return_grad))
    buf = _closure_buffers.get(key)
    if buf is None:
        if len(_closure_buffers) > 32:
            _closure_buffers.clear()
        f64 = dict(dtype=torch.float64, device=dev)
        width = 1 + (d if return_grad else 0)
        res = torch.empty(B * width, **f64)
        buf = {'theta_h': torch.empty((B, d), dtype=torch.float64).pin_memory(), 'theta_d': torch.empty((B, d), **f64),
               'pred': (torch.empty((B, K), **f64), torch.empty((B, K), **f64),
                        torch.empty((B, K, d), **f64) if return_grad else None,
                        torch.empty((B, K, d), **f64) if return_grad else None),
               'res_d': res, 'res_h': torch.empty(B * width, dtype=torch.float64).pin_memory(),
               'out': (res[:B], res[B:].view(B, d) if return_grad else None,
                       torch.empty(1, dtype=torch.int32, device=dev))}
        _closure_buffers[key] = buf
    return buf


def _packed(fitinfo, emu, theta, y, x, return_grad, chunk=4096):
    """numpy (B, 1) or (B, 1+d) from one fused device evaluation: theta goes up through pinned memory, the packed
    [ll | dll] table comes back with one copy, and every intermediate lives in per-shape cached buffers."""
    ops = _ops()
    torch = ops._torch()
    theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    B, d = theta.shape
    if B > chunk:                                     # large one-off populations: the general path
        ll, dll, _ = loglik_device(fitinfo, emu, theta, y, x, return_grad=return_grad, chunk=chunk)
        if return_grad:
            return torch.cat([ll[:, None], dll], dim=1).cpu().numpy()
        return ll.cpu().numpy().reshape(-1, 1)
    consts = woodbury_constants(fitinfo, emu, x, y)
    st = _device_state(emu)['_b200']
    if st.get('shard') is not None:
        # factors live on their owner ranks (PCGPwM_B200 fit option factors='owner'): collective PC-sharded predict,
        # then every rank evaluates the cheap k-space Woodbury for the whole population ("by PC")
        ll, dll, _ = loglik_device(fitinfo, emu, theta, y, x, return_grad=return_grad, chunk=chunk)
        if return_grad:
            return ops.to_host([torch.cat([ll[:, None], dll], dim=1)])[0]
        return ops.to_host([ll])[0].reshape(-1, 1)
    K = st['pw'].shape[0]
    buf = _buffers(torch, consts['phiT'].device, B, d, K, return_grad)
    buf['theta_h'].numpy()[...] = theta
    buf['theta_d'].copy_(buf['theta_h'], non_blocking=True)
    pv, pvar, dpv, dpvar = ops.gp_predict(st, buf['theta_d'], return_grad, out=buf['pred'])
    ops.woodbury_loglik(consts, pv, pvar, dpv, dpvar, out=buf['out'])
    buf['res_h'].copy_(buf['res_d'], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    host = buf['res_h'].numpy()
    if return_grad:
        return np.concatenate([host[:B, None], host[B:].reshape(B, d)], axis=1)
    return host[:B].reshape(-1, 1).copy()


def _evaluate(fitinfo, emu, theta, y, x, return_grad):
    """Optionally shard the theta population over the torch.distributed ranks (fitinfo['shard_theta']).

    Every rank must then call with the same theta (samplers run replicated with identical seeds); each rank
    evaluates its slice, the batch-wide unmodelled-variance flag (dbw.py:277-284) is OR-reduced so that it is
    still decided over the whole population, and the (B, 1+d) table is all-gathered (NCCL).
    """
    if not _pc_family(emu):
        if return_grad:
            raise ValueError("directbayeswoodbury_B200: the log-likelihood gradient needs an emulator fitted with "
                             "method='PCGPwM_B200' (or the reference's 'PCGPwM'); other emulators are evaluated without.")
        return _loglik_any_emulator(fitinfo, emu, np.atleast_2d(theta), y, x)
    if fitinfo.get('shard_theta') and _device_state(emu)['_b200'].get('shard') is None:
        from .. import _dist
        comm = _dist.Comm()
        if comm.world > 1:
            return _evaluate_sharded(fitinfo, emu, np.atleast_2d(theta), y, x, return_grad, comm)
    return _packed(fitinfo, emu, theta, y, x, return_grad)


def _evaluate_sharded(fitinfo, emu, theta, y, x, return_grad, comm):
    from .. import _dist
    ops = _ops()
    torch = ops._torch()
    consts = woodbury_constants(fitinfo, emu, x, y)
    info = _device_state(emu)
    sizes, starts = _dist.block_partition(theta.shape[0], comm.world)
    lo, hi = int(starts[comm.rank]), int(starts[comm.rank + 1])
    dev = consts['phiT'].device
    fired = torch.zeros(1, dtype=torch.int32, device=dev)
    pred = None
    if hi > lo:
        pred = _emu_method.predict_pcspace(info, theta[lo:hi], return_grad, population=theta.shape[0])
        fired = ops.woodbury_trigger(consts, pred[1])
    comm.dist.all_reduce(fired, op=comm.dist.ReduceOp.MAX)
    width = 1 + (theta.shape[1] if return_grad else 0)
    local = np.zeros((0, width))
    if pred is not None:
        ll, dll, _ = ops.woodbury_loglik(consts, *pred, variant=int(fired.item()))
        local = (torch.cat([ll[:, None], dll], dim=1) if return_grad else ll[:, None]).cpu().numpy()
    return comm.allgather_rows(local, sizes)


class DeviceLogpost:
    """What a device-resident sampler (utilitiesmethods/PTLMC_B200) needs from the log-posterior closure: the prior's
    box, if the prior is uniform on one, and a fixed-shape, allocation-free, synchronisation-free evaluation of the
    log-likelihood and its gradient that can be enqueued thousands of times (and captured in a CUDA graph).

    A prior object declares its box with an attribute `box = (lo, hi)` (arrays of length d): lpdf is constant inside
    and -inf outside.  Without it the sampler keeps calling the host closure (any Python prior works there)."""

    def __init__(self, fitinfo, emu, x, y, thetaprior, extra_logp=None):
        self.fitinfo, self.emu, self.x, self.y, self.prior = fitinfo, emu, x, y, thetaprior
        box = getattr(thetaprior, 'box', None)
        self.box = None
        if box is not None and extra_logp is None:
            lo, hi = (np.asarray(b, dtype=np.float64).reshape(-1) for b in box)
            if lo.shape == hi.shape and np.all(hi > lo):
                self.box = (lo, hi)

    def evaluator(self, B, d):
        return _DeviceEvaluator(self.fitinfo, self.emu, self.x, self.y, B, d)

    def record(self, **stats):
        self.fitinfo['_b200sampler'] = stats


_PEER_READY = {'done': False}


class _DeviceEvaluator:
    """ll (B), dll (B,d) of the proposals in `thetap` (B,d), rows masked by `valid` (int32) -- all device buffers of
    fixed address.  launch() enqueues predict + Woodbury on the current stream.  With factors sharded over ranks
    (PCGPwM_B200 fit option factors='owner') every rank evaluates the components whose factor it owns straight into
    its block of one buffer, the blocks are exchanged -- through NVLink peer memory by this library's own kernels
    (symmetric buffers, sb200_peer_allgather_post / _wait; csrc/peer.cu), or with an NCCL all-gather where symmetric
    memory is not available (SB200_PEER_EXCHANGE=0 forces that) -- and the Woodbury kernel reads the gathered blocks as
    they lie (segmented layout): no repacking."""

    def __init__(self, fitinfo, emu, x, y, B, d):
        ops = _ops()
        torch = ops._torch()
        self.ops, self.torch = ops, torch
        consts = woodbury_constants(fitinfo, emu, x, y)
        st = _device_state(emu)['_b200']
        dev = consts['phiT'].device
        f64 = dict(dtype=torch.float64, device=dev)
        self.B, self.d, self.st = B, d, st
        self.thetap = torch.zeros((B, d), **f64)
        self.valid = torch.ones(B, dtype=torch.int32, device=dev)
        self.ll, self.dll = torch.zeros(B, **f64), torch.zeros((B, d), **f64)
        self.fired = torch.zeros(1, dtype=torch.int32, device=dev)
        sh = st.get('shard')
        if sh is None:
            K = st['pw'].shape[0]
            self.consts, self.segments, self.comm, self.peer = consts, None, None, None
            self.pred = (torch.empty((B, K), **f64), torch.empty((B, K), **f64), torch.empty((B, K, d), **f64),
                         torch.empty((B, K, d), **f64))
            return
        from .. import _dist
        self.comm = _dist.Comm()
        if self.comm.world != sh['world'] or self.comm.rank != sh['rank']:
            raise RuntimeError('the emulator\'s factors are sharded over %d ranks; calibrate in the same process group'
                               % sh['world'])
        world, rank = sh['world'], sh['rank']
        kmax = max(max(sh['counts']), 1)
        chunk = B * kmax * (2 + 2 * d)
        chunk += chunk & 1                                     # even: blocks stay 16-byte aligned
        self.world, self.rank, self.chunk = world, rank, chunk
        self.kl, self.kmax = int(st['local']['pw'].shape[0]), kmax
        self.segments = (kmax, chunk)
        self.peer = None
        # Exchange of the blocks: NCCL all-gather by default.  The peer-memory exchange (csrc/peer.cu) is 2-3 % faster per
        # step (4 x B200: 0.797 vs 0.817 ms) but torch's first symmetric-memory rendezvous costs ~1.5 s per process, more
        # than a short chain gains: it is used when asked for (SB200_PEER_EXCHANGE=1) or once a process has paid that
        # price already (a second calibration in the same process); SB200_PEER_EXCHANGE=0 never uses it.
        mode = os.environ.get('SB200_PEER_EXCHANGE', 'auto')
        if mode == '1' or (mode != '0' and _PEER_READY['done']):
            try:
                self._init_peer(torch, dev)
                _PEER_READY['done'] = True
            except Exception:
                if mode == '1':
                    raise
                self.peer = None
        if self.peer is None:
            self.gathered = torch.zeros(world * chunk, **f64)
            self._blocks = [self._init_blocks(self.gathered.view(world, chunk))]
        # where every real component sits among the padded, rank-ordered slots: the Woodbury kernel visits only those, so
        # its k x k work does not grow with the number of ranks (padding to world x kmax slots doubled it on 4 ranks)
        K = st['pw'].shape[0]
        pos = np.zeros(K, dtype=np.int32)
        offs = np.concatenate([[0], np.cumsum(sh['counts'])])
        order = np.argsort(sh['owner'], kind='stable')
        for r in range(world):
            for c, j in enumerate(order[offs[r]:offs[r + 1]]):
                pos[j] = r * kmax + c
        self.seg_pos = torch.from_numpy(pos).to(dev)
        self.consts = consts

    def _views(self, block):
        B, kmax, d = self.B, self.kmax, self.d
        o = B * kmax
        return (block[:o].view(B, kmax), block[o:2 * o].view(B, kmax), block[2 * o:2 * o + o * d].view(B, kmax, d),
                block[2 * o + o * d:2 * o + 2 * o * d].view(B, kmax, d))

    def _init_blocks(self, blocks):
        """blocks: (world, chunk) view of one gathered buffer.  Padded components: predvar 1, everything else 0."""
        o = self.B * self.kmax
        blocks.zero_()
        blocks[:, o:2 * o] = 1.0
        return {'all': blocks, 'local': blocks[self.rank], 'pred_local': self._views(blocks[self.rank]),
                'pred': self._views(blocks[0])}

    def _init_peer(self, torch, dev):
        """Blocks in symmetric memory (every rank's buffer mapped into every peer), exchanged by sb200_peer_allgather_*
        over NVLink instead of an NCCL all-gather; two buffers, used by alternate steps (see csrc/peer.cu)."""
        import torch.distributed._symmetric_memory as symm
        from .._lib import lib
        if not hasattr(lib, 'sb200_peer_allgather_post'):
            raise RuntimeError('library without the peer exchange')
        world, chunk = self.world, self.chunk
        dist = self.comm.dist
        total = 2 * world * chunk + world                       # [parity 0 | parity 1 | world flag words]
        buf = symm.empty(total, dtype=torch.float64, device=dev)
        group = dist.group.WORLD
        try:
            hdl = symm.rendezvous(buf, group)
        except Exception:
            hdl = symm.rendezvous(buf, group.group_name)
        buf.zero_()
        halves = [self._init_blocks(buf[p * world * chunk:(p + 1) * world * chunk].view(world, chunk)) for p in (0, 1)]
        torch.cuda.synchronize()
        dist.barrier()                                          # nobody posts into a buffer that is still being initialised
        ptrs = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=dev)
        self._blocks = halves
        self.peer = {'buf': buf, 'hdl': hdl, 'ptrs': ptrs, 'flag_off': 2 * world * chunk, 'par': 0,
                     'epoch': torch.zeros(1, dtype=torch.int64, device=dev),
                     'arrive': torch.zeros(world, dtype=torch.int32, device=dev),
                     'error': torch.zeros(1, dtype=torch.int32, device=dev)}

    def check(self):
        """After the chain: raise if a peer's block did not arrive in time (the kernels never hang on it)."""
        if self.segments is not None and self.peer is not None and int(self.peer['error'].item()) != 0:
            raise RuntimeError('directbayeswoodbury_B200: a rank did not deliver its prediction block in time')

    def launch(self):
        ops = self.ops
        if self.segments is None:
            ops.gp_predict(self.st, self.thetap, True, out=self.pred, population=self.B)
            ops.woodbury_loglik(self.consts, *self.pred, out=(self.ll, self.dll, self.fired), row_valid=self.valid)
            return
        if self.peer is None:
            blk = self._blocks[0]
            if self.kl:
                ops.gp_predict(self.st['local'], self.thetap, True, out=blk['pred_local'], population=self.B,
                               out_ld=self.kmax)
            self.comm.dist.all_gather_into_tensor(self.gathered, blk['local'])
        else:
            from .._lib import lib, check
            pr = self.peer
            par = pr['par']
            pr['par'] ^= 1
            blk = self._blocks[par]
            if self.kl:
                ops.gp_predict(self.st['local'], self.thetap, True, out=blk['pred_local'], population=self.B,
                               out_ld=self.kmax)
            stream = ops._stream(self.torch)
            check(lib.sb200_peer_allgather_post(blk['local'].data_ptr(), self.chunk, pr['ptrs'].data_ptr(),
                                                (par * self.world + self.rank) * self.chunk, pr['flag_off'], self.rank,
                                                self.world, pr['epoch'].data_ptr(), pr['arrive'].data_ptr(), stream),
                  'sb200_peer_allgather_post')
            check(lib.sb200_peer_allgather_wait(pr['buf'].data_ptr() + 8 * pr['flag_off'], self.rank, self.world,
                                                pr['epoch'].data_ptr(), pr['error'].data_ptr(), stream),
                  'sb200_peer_allgather_wait')
        ops.woodbury_loglik(self.consts, *blk['pred'], out=(self.ll, self.dll, self.fired), row_valid=self.valid,
                            segments=self.segments, seg_pos=self.seg_pos)


def loglik(fitinfo, emu, theta, y, x):
    """(B,1) log-likelihoods; same contract as directbayeswoodbury.loglik (dbw.py:261-306)."""
    return _evaluate(fitinfo, emu, theta, y, x, False)


def loglik_grad(fitinfo, emu, theta, y, x):
    """(B,1), (B,d); same contract as directbayeswoodbury.loglik_grad (dbw.py:309-383)."""
    packed = _evaluate(fitinfo, emu, theta, y, x, True)
    return packed[:, :1].copy(), packed[:, 1:].copy()


def make_logpost(fitinfo, emu, x, y, fd_step=10 ** (-6), extra_logp=None, with_grad=True):
    """The closure handed to the sampler (dbw.py:100-129) and the initial-draw function (132-146).

    fd_step: step of the finite-difference prior gradient used when the prior has no lpdf_grad (dbw.py:82-98).
    extra_logp: optional (value(theta) -> (B,1), grad(theta) -> (B,d)) pair added on the rows with finite prior
    (mlbayeswoodbury's classifier term).  with_grad=False: the closure never returns gradients."""
    thetaprior = fitinfo['thetaprior']
    if with_grad and 'lpdf_grad' not in dir(thetaprior):
        def lpdf_grad(theta):
            base = thetaprior.lpdf(theta)
            inds = np.where(np.isfinite(base))[0]
            grad = np.zeros((theta.shape[0], theta.shape[1]))
            for k in range(theta.shape[1]):
                prop = copy.copy(theta)
                prop[:, k] += fd_step
                grad[inds, k] = (1 / fd_step) * (thetaprior.lpdf(prop[inds, :]) - base[inds]).reshape(len(inds),)
            return grad
        thetaprior.lpdf_grad = lpdf_grad

    def logpostfull_wgrad(theta, return_grad=True):
        logpost = thetaprior.lpdf(theta)
        inds = np.where(np.isfinite(logpost))[0]
        if with_grad and return_grad:
            dlogpost = thetaprior.lpdf_grad(theta)
            if len(inds):
                ll, dll = loglik_grad(fitinfo, emu, theta[inds, :], y, x)
                logpost[inds] += ll
                dlogpost[inds] += dll
                if extra_logp is not None:
                    logpost[inds] += extra_logp[0](theta[inds, :])
                    dlogpost[inds] += extra_logp[1](theta[inds, :])
            return logpost, dlogpost
        if len(inds) > 0:
            logpost[inds] += loglik(fitinfo, emu, theta[inds, :], y, x)
            if extra_logp is not None:
                logpost[inds] += extra_logp[0](theta[inds, :])
        return logpost

    def draw_func(n):
        p = thetaprior.rnd(1).shape[1]
        theta0 = np.array([]).reshape(0, p)
        if 'thetarnd' in fitinfo:
            theta0 = fitinfo['thetarnd']
        if '_emulator__theta' in dir(emu):
            theta0 = np.vstack((theta0, copy.copy(emu._emulator__theta)))
        n0 = len(theta0)
        if n0 < n:
            theta0 = np.vstack((thetaprior.rnd(n - n0), theta0))
        else:
            theta0 = theta0[np.random.randint(theta0.shape[0], size=n), :]
        return theta0

    # lets a device-resident sampler plugin (PTLMC_B200) keep the whole chain on the GPU when the prior allows it
    logpostfull_wgrad.b200 = (DeviceLogpost(fitinfo, emu, x, y, thetaprior, extra_logp)
                              if with_grad and _pc_family(emu) else None)
    return logpostfull_wgrad, draw_func


def fit(fitinfo, emu, x, y, **bayeswoodbury_args):
    """Posterior sampling with surmise's own sampler plugins (dbw.py:7-163)."""
    try:
        from surmise.utilities import sampler        # the reference's dispatcher, used unmodified
        import sys
        from .. import UTILITIES_METHODS, register
        name = bayeswoodbury_args.get('sampler')
        if name in UTILITIES_METHODS and 'surmise.utilitiesmethods.' + name not in sys.modules:
            register()                               # our sampler plugins, asked for before register() was called
    except ImportError:
        from ..utilities import sampler              # same contract, for installations without surmise
    if 'shard_theta' in bayeswoodbury_args:          # our options, not sampler options
        fitinfo['shard_theta'] = bool(bayeswoodbury_args.pop('shard_theta'))
    if bayeswoodbury_args.pop('shard_factors', False):
        # every rank holds the same emulator: keep only this rank's block of the full-data factors and evaluate
        # PC-sharded from here on (collective; all ranks run the same sampler with the same seed)
        _emu_method.shard_state(_device_state(emu))
    if _pc_family(emu):
        woodbury_constants(fitinfo, emu, x, y)
    logpost, draw_func = make_logpost(fitinfo, emu, x, y)
    sampler_obj = sampler(logpost_func=logpost, draw_func=draw_func, **bayeswoodbury_args)
    theta = sampler_obj.sampler_info['theta']
    ladj = logpost(theta, return_grad=False)
    mladj = np.max(ladj)
    fitinfo['lpdfapproxnorm'] = np.log(np.mean(np.exp(ladj - mladj))) + mladj
    fitinfo['thetarnd'] = theta
    fitinfo['y'] = y
    fitinfo['x'] = x
    fitinfo['emu'] = emu
    return


def predict(predinfo, fitinfo, emu, x, args=None):
    """Posterior predictive summaries at x (dbw.py:166-216)."""
    theta = fitinfo['thetarnd']
    if theta.ndim == 1:
        theta = theta.reshape((1, theta.shape[0]))
    pred = emu.predict(x, theta)
    mean = pred.mean()
    cxh = pred.covxhalf()
    rnd = copy.deepcopy(mean).T
    z = np.random.standard_normal(size=(theta.shape[0], cxh.shape[2]))
    rnd += np.einsum('xbk,bk->bx', cxh, z)
    predinfo['rnd'] = rnd
    predinfo['mean'] = np.mean(mean, 1)
    predinfo['var'] = np.mean(np.sum(np.square(cxh), axis=2), 1) + np.var(mean, 1)
    return


def thetarnd(fitinfo, s=100, args=None):
    """s draws from the stored posterior sample (dbw.py:219-238)."""
    return fitinfo['thetarnd'][np.random.choice(fitinfo['thetarnd'].shape[0], size=s), :]


def thetalpdf(fitinfo, theta, args=None):
    """Normalised log-posterior at theta (dbw.py:241-258)."""
    emu, y, x = fitinfo['emu'], fitinfo['y'], fitinfo['x']
    logpost = fitinfo['thetaprior'].lpdf(theta)
    if logpost.ndim > 0.5 and logpost.shape[0] > 1.5:
        inds = np.where(np.isfinite(logpost))[0]
        logpost[inds] += loglik(fitinfo, emu, theta[inds], y, x)
    elif np.isfinite(logpost):
        logpost += loglik(fitinfo, emu, theta, y, x)
    return logpost - fitinfo['lpdfapproxnorm']
