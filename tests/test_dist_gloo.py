"""CPU, world_size 2 over gloo: the multi-GPU sharding logic (PCs during fit, theta during calibration)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.fixture(scope='module')
def two_rank_results(tmp_path_factory):
    import torch.multiprocessing as mp
    import dist_worker
    out = str(tmp_path_factory.mktemp('dist'))
    mp.spawn(dist_worker.worker, args=(2, _free_port(), out), nprocs=2, join=True)
    return [dict(np.load(os.path.join(out, 'rank%d.npz' % r))) for r in range(2)]


def test_block_partition():
    from surmise_b200._dist import block_partition
    sizes, starts = block_partition(7, 3)
    assert sizes == [3, 2, 2] and list(starts) == [0, 3, 5, 7]
    sizes, starts = block_partition(2, 4)
    assert sizes == [1, 1, 0, 0]


def test_sharded_fit_equals_single_process(two_rank_results):
    """PCs sharded over 2 ranks: identical table on both ranks and identical to the 1-rank Jacobi run."""
    import dist_worker
    r0, r1 = two_rank_results
    assert np.array_equal(r0['hyps'], r1['hyps']) and np.array_equal(r0['inds'], r1['inds'])
    assert np.array_equal(r0['rows'], r1['rows'])                   # same random sub-samples everywhere
    hyps, inds, rows = dist_worker.run_sweeps()                     # single process, Comm() falls back to world=1
    assert np.allclose(r0['hyps'], hyps, rtol=0, atol=1e-12)
    assert np.array_equal(r0['inds'], inds) and np.array_equal(r0['rows'], rows)
    assert set(inds.tolist()) <= set(range(len(inds)))
    assert all(inds[j] <= j for j in range(len(inds)))             # hypind = min(j, leader)  (PCGPwM.py:717)
    # the invariant PCGPwM.predict relies on (rsave[hypind], PCGPwM.py:219-230): a component's hyper-parameters ARE
    # its hypind's hyper-parameters, also when its leader moved on in the same Jacobi sweep
    assert all(np.array_equal(hyps[j], hyps[inds[j]]) for j in range(len(inds)))


def test_canonical_hypinds():
    from surmise_b200._dist import canonical_hypinds
    h = np.array([[1.0, 2.0], [3.0, 4.0], [1.0, 2.0], [3.0, 4.0 + 1e-16], [5.0, 6.0], [3.0, 4.0]])
    assert canonical_hypinds(h).tolist() == [0, 1, 0, 1, 4, 1]
    h[3, 1] = np.nextafter(4.0, 5.0)                                # one ulp away is a different leader
    assert canonical_hypinds(h).tolist() == [0, 1, 0, 3, 4, 1]


@pytest.mark.parametrize('B', [1, 2, 7, 64])
def test_sharded_theta_evaluation(two_rank_results, B):
    for r in two_rank_results:
        assert r['got%d' % B].shape == r['want%d' % B].shape
        assert np.array_equal(r['got%d' % B], r['want%d' % B])


def test_ragged_allgather(two_rank_results):
    want = np.vstack([np.zeros((1, 3)), np.ones((2, 3))])
    for r in two_rank_results:
        assert np.array_equal(r['ragged'], want)


def test_gather_blocks_assembles_owner_slices(two_rank_results):
    """_dist.gather_blocks (the sharded final factorisation's exchange): every rank ends with the full batch."""
    want = np.arange(5 * 12, dtype=np.float64).reshape(5, 3, 4)
    for r in two_rank_results:
        assert np.array_equal(r['blocks'], want)


def test_owner_layout_gathers_in_component_order(two_rank_results):
    """_dist.owner_plan + Comm.allgather_tensor (factors stay with their owner; results travel): rank-major gathered rows,
    re-ordered by plan['inverse'], come out in component order on every rank."""
    want = np.arange(11, dtype=np.float64)[:, None].repeat(3, 1)
    for r in two_rank_results:
        assert np.array_equal(r['owner_table'], want)


def test_owner_plan_partitions_components():
    from surmise_b200._dist import owner_plan, block_partition
    fidx = np.array([0, 1, 0, 2, 3, 1, 4, 4, 2])
    sizes, starts = block_partition(5, 3)
    plans = [owner_plan(fidx, starts, r) for r in range(3)]
    assert sorted(np.concatenate([p['pcs_local'] for p in plans]).tolist()) == list(range(9))
    assert plans[0]['counts'] == [4, 3, 2]
    for r, p in enumerate(plans):                      # a component lives where its factor lives
        assert all(starts[r] <= fidx[j] < starts[r + 1] for j in p['pcs_local'])
    sizes, starts = block_partition(2, 4)              # more ranks than factors: the late ranks own nothing
    p = owner_plan(np.array([0, 1, 1, 0]), starts, 3)
    assert p['counts'] == [2, 2, 0, 0] and p['pcs_local'].size == 0
