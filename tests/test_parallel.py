"""Sharding logic of the multi-GPU path, on CPU with the gloo backend (world_size 2)."""

import os
import socket

import numpy as np
import pytest

from sparselm_b200.parallel import GridShard, assign_columns


@pytest.mark.parametrize("F,K,W", [(5, 100, 8), (5, 100, 4), (5, 100, 2), (5, 10, 8), (3, 7, 2), (2, 3, 8), (5, 100, 1)])
def test_assignment_is_a_balanced_partition(F, K, W):
    owner = assign_columns(F, K, W)
    assert owner.shape == (F, K) and owner.min() >= 0 and owner.max() < W
    counts = np.bincount(owner.ravel(), minlength=W)
    assert counts.sum() == F * K and counts.max() - counts.min() <= 1 + (F * K < W)
    # every rank touches the minimum number of folds its capacity range spans
    total = F * K
    for r in range(W):
        lo, hi = (r * total) // W, ((r + 1) * total) // W
        if hi > lo:
            spanned = set(range(lo // K, (hi - 1) // K + 1))
            assert set(np.unique(np.nonzero(owner == r)[0])) <= spanned
    # alphas of a fold are interleaved among its ranks: each rank's columns span the grid
    if W > 1 and K >= 20:
        for f in range(F):
            for r in np.unique(owner[f]):
                cols = np.flatnonzero(owner[f] == r)
                if len(cols) >= 4:
                    assert cols.min() < K / 2 < cols.max()


def test_row_ranges_partition_rows():
    n = 20003
    ranges = [GridShard(r, 8).row_range(n) for r in range(8)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(7))
    # what engine.prepare builds on every rank: its 1/world slice of EVERY test fold's rows
    row_ptr = np.array([0, 4001, 8001, 12002, 16002, 20003])
    covered = np.zeros(n, dtype=int)
    for r in range(8):
        fr = GridShard(r, 8).fold_row_ranges(row_ptr)
        assert len(fr) == 5
        for f, (lo, hi) in enumerate(fr):
            assert row_ptr[f] <= lo <= hi <= row_ptr[f + 1]
            assert abs((hi - lo) - (row_ptr[f + 1] - row_ptr[f]) / 8) < 1  # balanced per fold
            covered[lo:hi] += 1
    assert np.all(covered == 1)
    # more ranks than rows in a fold: empty slices are fine
    fr = [GridShard(r, 8).fold_row_ranges(np.array([0, 3, 5])) for r in range(8)]
    assert sum(hi - lo for ranges_ in fr for lo, hi in ranges_) == 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shard = GridShard(rank, world)
        F, K, p = 5, 24, 6
        rng = np.random.default_rng(0)
        full = rng.standard_normal((K, F))  # the table a single process would produce
        mine = shard.my_columns(F, K)
        table = np.full((K, F), np.nan)
        for f in range(F):
            table[mine[f], f] = full[mine[f], f]
        summed = shard.allreduce_sum_numpy(np.nan_to_num(table, nan=0.0))
        # row-sharded Gram: partial X^T X of the rank's rows, summed
        X = rng.standard_normal((101, p))
        r0, r1 = shard.row_range(101)
        part = torch.from_numpy(X[r0:r1].T @ X[r0:r1])
        shard.allreduce_sum_(part)
        ok = np.allclose(summed, full, rtol=0, atol=0) and np.allclose(part.numpy(), X.T @ X, rtol=1e-13)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_score_gather_and_gram_allreduce():
    import torch.multiprocessing as mp

    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
