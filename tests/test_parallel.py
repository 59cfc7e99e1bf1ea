"""Sharding logic of the multi-GPU path, on CPU with the gloo backend (world_size 2)."""

import os
import socket

import numpy as np
import pytest

from sparselm_b200.parallel import GridShard, assign_columns, gather_layout, scoring_plan


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


@pytest.mark.parametrize("F,K,W", [(5, 100, 8), (5, 100, 2), (3, 7, 4), (2, 3, 8)])
def test_gather_layout_is_aligned_and_disjoint(F, K, W):
    owner = assign_columns(F, K, W)
    kfr, pad, off, ldg = gather_layout(owner, W)
    assert kfr.sum() == F * K and ldg % 8 == 0 and ldg >= 8
    assert np.all(pad % 8 == 0) and np.all(pad >= kfr) and np.all(pad - kfr < 8)
    assert np.all(off % 8 == 0) and np.all(off[:, -1] <= ldg)
    for r in range(W):
        for f in range(F):
            assert off[r, f] + pad[r, f] == off[r, f + 1]  # groups are laid out back to back


def _exchange_worker(rank, world, port, out):
    """The data movement of model_selection._sharded_residual_sums with numpy standing in for the GPU scorer:
    compact all-gather of the solved columns, this rank's row slices of every fold against every rank's groups,
    sum of the partial residual tables == the single-process table."""
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shard = GridShard(rank, world)
        F, K, p, n = 3, 13, 5, 47
        rng = np.random.default_rng(3)
        X, y = rng.standard_normal((n, p)), rng.standard_normal(n)
        coef = rng.standard_normal((F, p, K))      # what the solves of (fold, column) produced, all ranks agree
        icpt = rng.standard_normal((F, K))
        row_ptr = np.array([0, 16, 31, 47])
        owner = assign_columns(F, K, world)
        mine = shard.my_columns(F, K)
        kfr, pad, off, ldg = gather_layout(owner, world)
        send = torch.zeros((p + 1, ldg), dtype=torch.float64)
        for f in range(F):
            kf = len(mine[f])
            o = int(off[rank, f])
            send[:p, o:o + kf] = torch.from_numpy(coef[f][:, mine[f]])
            send[p, o:o + kf] = torch.from_numpy(icpt[f][mine[f]])
        recv = torch.empty((world, p + 1, ldg), dtype=torch.float64)
        shard.all_gather_(send, recv)
        recv = recv.numpy()
        sse = np.zeros((F, K))
        tsse = np.zeros((F, K))
        for lo, hi, r, f, kind, cols in scoring_plan(owner, shard.fold_row_ranges(row_ptr), world, want_train=True):
            o, kf = int(off[r, f]), len(cols)
            B, b0 = recv[r, :p, o:o + kf], recv[r, p, o:o + kf]
            res = y[lo:hi, None] - X[lo:hi] @ B - b0[None, :]
            (sse if kind == 0 else tsse)[f][cols] += (res ** 2).sum(axis=0)
        both = shard.allreduce_sum_numpy(np.stack([sse, tsse]))
        ref = np.zeros((2, F, K))
        for f in range(F):
            te = np.arange(row_ptr[f], row_ptr[f + 1])
            tr = np.setdiff1d(np.arange(n), te)
            for kind, idx in ((0, te), (1, tr)):
                res = y[idx, None] - X[idx] @ coef[f] - icpt[f][None, :]
                ref[kind, f] = (res ** 2).sum(axis=0)
        out[rank] = bool(np.allclose(both, ref, rtol=1e-12, atol=1e-12))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_row_sharded_scoring_exchange():
    import torch.multiprocessing as mp

    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_exchange_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
