"""Multi-GPU plumbing: the path shards only by independent units.

  * fit         -- principal components are independent given the previous sweep's leaders: each rank fits
                   its own PCs, then every rank all-gathers the (k, p+1) table [hyper-parameters | hypind];
  * calibration -- theta proposals are independent: each rank evaluates a contiguous slice of the (B, d)
                   population and the (B, 1+d) results are all-gathered.
One process per GPU, torch.distributed with NCCL (tensors stay on the device); the same code runs over gloo
with CPU tensors, which is how tests/test_dist_gloo.py exercises it without a GPU.  No data-path collective
exists beyond these small all-gathers (<= a few KB per call: latency-bound, NVLink bandwidth is irrelevant).
"""
import numpy as np


class Comm:
    """Thin view of the default torch.distributed group (or a single-process stand-in)."""

    def __init__(self, enabled=True):
        self.rank, self.world, self.dist = 0, 1, None
        if not enabled:
            return
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.dist, self.rank, self.world = dist, dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass

    def _device(self):
        import torch
        if self.dist.get_backend() == 'nccl':
            return torch.device('cuda', torch.cuda.current_device())
        return torch.device('cpu')

    def allgather_rows(self, local, counts):
        """Concatenate per-rank float64 row blocks; counts[r] = rows owned by rank r (known to every rank)."""
        local = np.ascontiguousarray(local, dtype=np.float64)
        if self.world == 1:
            return local
        import torch
        width = local.shape[1]
        dev = self._device()
        pad = max(counts)
        buf = torch.zeros((pad, width), dtype=torch.float64, device=dev)
        if local.shape[0]:
            buf[:local.shape[0]] = torch.from_numpy(local).to(dev)
        out = [torch.empty_like(buf) for _ in range(self.world)]
        self.dist.all_gather(out, buf)
        return np.concatenate([out[r][:counts[r]].cpu().numpy() for r in range(self.world)], axis=0)

    def allgather_tensor(self, local, counts):
        """allgather_rows for tensors that stay where they are (device tensors over NCCL, CPU tensors over gloo):
        rank r contributes local (counts[r], w); returns the (sum(counts), w) concatenation in rank order."""
        if self.world == 1:
            return local
        import torch
        pad = max(max(counts), 1)
        buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        if local.shape[0]:
            buf[:local.shape[0]] = local
        out = torch.empty((self.world * pad,) + tuple(buf.shape[1:]), dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(out, buf)              # concatenated along dim 0 (NCCL and gloo agree on this)
        out = out.view((self.world, pad) + tuple(buf.shape[1:]))
        if all(c == pad for c in counts):
            return out.reshape((-1,) + tuple(local.shape[1:]))
        return torch.cat([out[r, :counts[r]] for r in range(self.world)], 0)

    def broadcast_array(self, arr, src=0):
        """Rank src's int64 array to everyone (used to agree on the random sub-samples)."""
        if self.world == 1:
            return np.asarray(arr)
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64)).to(self._device())
        self.dist.broadcast(t, src=src)
        return t.cpu().numpy()


def block_partition(count, world):
    """Contiguous, near-equal blocks: rank r owns [starts[r], starts[r+1])."""
    base, extra = divmod(count, world)
    sizes = [base + (1 if r < extra else 0) for r in range(world)]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    return sizes, starts


def owner_plan(fidx, starts, rank):
    """PC sharding that follows factor ownership ("factors stay on their owner").

    fidx (k,): distinct-factor index of every component; starts: block_partition boundaries of the factors over the
    ranks.  Component j lives on the rank that owns factor fidx[j].  Returns
      owner (k,)      rank of every component,
      counts          components per rank,
      order (k,)      component ids in gathered order (rank-major, component order within a rank),
      inverse (k,)    position of component j in the gathered order (gathered[inverse] is in component order),
      pcs_local       this rank's components, ascending."""
    fidx = np.asarray(fidx)
    world = len(starts) - 1
    owner = np.searchsorted(np.asarray(starts), fidx, side='right') - 1
    owner = np.minimum(owner, world - 1)
    counts = np.bincount(owner, minlength=world).astype(int).tolist()
    order = np.argsort(owner, kind='stable')
    inverse = np.empty_like(order)
    inverse[order] = np.arange(len(order))
    return {'owner': owner, 'counts': counts, 'order': order, 'inverse': inverse,
            'pcs_local': np.where(owner == rank)[0]}


def gather_blocks(comm, local, sizes, out):
    """Assemble a batch whose contiguous blocks were computed by their owner ranks.

    out: preallocated (sum(sizes), ...) tensor, same shape on every rank; `local` holds this rank's rows
    [starts[rank], starts[rank+1]) (None if it owns nothing).  Every owner broadcasts its slice (NCCL over NVLink for
    the full-data factors: C4 = 40 x 512 MB, against 0.5 s to factor them all on one GPU)."""
    starts = np.concatenate([[0], np.cumsum(sizes)])
    lo, hi = int(starts[comm.rank]), int(starts[comm.rank + 1])
    if hi > lo:
        out[lo:hi] = local
    if comm.world > 1:
        for r in range(comm.world):
            if sizes[r]:
                comm.dist.broadcast(out[int(starts[r]):int(starts[r + 1])], src=r)
    return out


def jacobi_sweeps(fit_one, k, p, comm, draw_rows, sweeps=3, start_hyps=None, start_inds=None, run_block=None):
    """Sharded variant of PCGPwM.__fitGPs (PCGPwM.py:678-722).

    The reference sweeps Gauss-Seidel (PC j sees the PCs fitted earlier in the same sweep); sharding PCs over
    ranks needs the Jacobi form: in sweep s every PC sees the leaders as they stood after sweep s-1.
    fit_one(j, cand (c,p) or None, cand_ids (c,) or None, rows) -> (hyp (p,), hypind or -1)
    draw_rows(j) -> the random sub-sample for PC j (drawn on rank 0 in PC order, broadcast, so the result
    does not depend on the number of ranks).
    run_block(list of zero-argument callables) -> list of results: how this rank's PCs are executed; default is
    one after the other, the GPU plugin passes a lock-step runner that batches their likelihood launches.
    Returns (hyps (k,p), hypinds (k,) int, rows list) identical on every rank.
    """
    sizes, starts = block_partition(k, comm.world)
    mine = range(int(starts[comm.rank]), int(starts[comm.rank + 1]))
    hyps = np.zeros((k, p)) if start_hyps is None else np.array(start_hyps, dtype=np.float64)
    hypinds = -1 * np.ones(k) if start_inds is None else np.array(start_inds, dtype=np.float64)
    rows_used = [None] * k
    for _sweep in range(sweeps):
        drawn = [np.asarray(draw_rows(j)) for j in range(k)] if comm.rank == 0 else None
        if comm.world > 1:
            flat = np.concatenate(drawn) if comm.rank == 0 else np.zeros(0, dtype=np.int64)
            lens = comm.broadcast_array(np.array([len(r) for r in drawn]) if comm.rank == 0 else np.zeros(k, dtype=np.int64))
            if comm.rank != 0:
                flat = np.zeros(int(lens.sum()), dtype=np.int64)
            flat = comm.broadcast_array(flat)
            cuts = np.concatenate([[0], np.cumsum(lens)])
            drawn = [flat[cuts[j]:cuts[j + 1]] for j in range(k)]
        own = np.arange(k)
        leaders = np.where(hypinds == own)[0]
        cand = hyps[leaders] if leaders.size else None
        local = np.zeros((len(mine), p + 1))
        lead_ids = leaders if leaders.size else None
        jobs = [(lambda j=j: fit_one(j, cand, lead_ids, drawn[j])) for j in mine]
        outs = run_block(jobs) if run_block is not None else [job() for job in jobs]
        for t, j in enumerate(mine):
            hyp, hi = outs[t]
            hi = min(j, hi)
            local[t, :p] = hyp
            local[t, p] = j if hi < -0.5 else hi
        table = comm.allgather_rows(local, sizes)
        hyps, hypinds = table[:, :p].copy(), table[:, p].copy()
        rows_used = drawn
    return hyps, canonical_hypinds(hyps), rows_used


def canonical_hypinds(hyps):
    """hypind[j] = the smallest index whose hyper-parameter row is byte-identical to row j.

    In a Jacobi sweep component j can adopt leader L's hyper-parameters of the PREVIOUS sweep while L itself moves on
    in the same sweep, which would leave hypind[j] == L with hyp[j] != hyp[L].  The reference's consumers rely on
    emulist[j]['hyp'] == emulist[hypind[j]]['hyp'] (PCGPwM.predict indexes rsave by hypind, PCGPwM.py:219-230), so the
    indices are recomputed from the final table -- identically on every rank."""
    first, out = {}, np.zeros(len(hyps), dtype=int)
    for j, row in enumerate(np.ascontiguousarray(hyps, dtype=np.float64)):
        out[j] = first.setdefault(row.tobytes(), j)
    return out


def sharded_rows(fn, theta, comm):
    """Evaluate fn on this rank's slice of theta (B, d) and all-gather the (B, w) results in order."""
    theta = np.atleast_2d(theta)
    if comm.world == 1:
        return fn(theta)
    sizes, starts = block_partition(theta.shape[0], comm.world)
    lo, hi = int(starts[comm.rank]), int(starts[comm.rank + 1])
    if hi > lo:
        local = np.asarray(fn(theta[lo:hi]), dtype=np.float64)
        width = np.array([local.shape[1]])
    else:
        local, width = None, np.array([0])
    # ranks with an empty slice learn the width from rank 0 (which always owns rows when B >= 1)
    width = int(comm.broadcast_array(width)[0])
    if local is None:
        local = np.zeros((0, width))
    return comm.allgather_rows(local, sizes)
