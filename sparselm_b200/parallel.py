"""Multi-GPU sharding of one CV search: one process per GPU, torch.distributed.

Two independent shardings compose:

* rows    : every rank builds, for every test fold, the Gram block of its own 1/world
            slice of that fold's rows; the upper triangles of all blocks are packed into one
            buffer and summed with ONE all-reduce (NCCL over NVLink / NVSwitch).
* grid    : the (fold, alpha) problems are independent (the reference treats them as
            independent joblib tasks, model_selection.py:304-323).  Each rank solves a
            subset chosen so that it touches as few Grams as possible (fold-major
            capacity ranges) while the alphas of a fold are interleaved among the ranks
            sharing it (small alphas need more iterations).  The only data-path traffic
            is one all-reduce of the tiny zero-padded score / info arrays.
"""

from __future__ import annotations

import functools

import numpy as np

__all__ = ["GridShard", "assign_columns", "gather_layout", "scoring_plan"]


@functools.lru_cache(maxsize=64)
def _assign_columns_cached(n_folds: int, n_cols: int, world: int):
    owner = _assign_columns(n_folds, n_cols, world)
    owner.setflags(write=False)
    return owner


def assign_columns(n_folds: int, n_cols: int, world: int):
    """owner[f][k] in [0, world): rank that solves column k of fold f (read-only, cached)."""
    return _assign_columns_cached(int(n_folds), int(n_cols), int(world))


def _assign_columns(n_folds: int, n_cols: int, world: int):
    """owner[f][k] in [0, world): rank that solves column k of fold f.

    Fold-major capacity ranges give the quota q[f][r] of fold f's columns owned by rank
    r (each rank touches the minimum number of folds); within a fold the columns are
    dealt to its ranks proportionally and evenly spaced (Bresenham), so every rank gets
    a representative sample of the alpha grid."""
    total = n_folds * n_cols
    bounds = [(r * total) // world for r in range(world + 1)]
    owner = np.zeros((n_folds, n_cols), dtype=np.int64)
    for f in range(n_folds):
        lo, hi = f * n_cols, (f + 1) * n_cols
        quota = np.array([max(0, min(hi, bounds[r + 1]) - max(lo, bounds[r])) for r in range(world)])
        ranks = np.flatnonzero(quota)
        given = np.zeros(world)
        for k in range(n_cols):
            # rank whose share is furthest behind its proportional target after k+1 columns
            deficit = quota[ranks] * (k + 1) / n_cols - given[ranks]
            r = ranks[int(np.argmax(deficit))]
            owner[f, k] = r
            given[r] += 1
    return owner


def gather_layout(owner, world: int):
    """Layout of the coefficient exchange of a row-sharded scoring (model_selection._sharded_residual_sums).

    Every rank sends ONE [p + 1][ldg] block: the columns it solved, fold after fold, every fold's group padded to
    a multiple of 8 columns (the scoring GEMM reads a group through a 16-byte aligned pointer with row stride ldg).
    Returns (kfr, pad, off, ldg): kfr[r][f] = columns rank r solved on fold f, pad = kfr rounded up to 8,
    off[r][f] = first column of that group in rank r's block, ldg = common block width (>= 8, multiple of 8)."""
    owner = np.asarray(owner)
    n_splits = owner.shape[0]
    kfr = np.array([[int((owner[f] == r).sum()) for f in range(n_splits)] for r in range(world)], dtype=np.int64)
    pad = (kfr + 7) // 8 * 8
    off = np.concatenate([np.zeros((world, 1), dtype=np.int64), np.cumsum(pad, axis=1)], axis=1)
    ldg = max(8, int(off[:, -1].max()))
    return kfr, pad, off, ldg


def scoring_plan(owner, rows, world: int, want_train: bool = False):
    """The scoring problems of ONE rank of a row-sharded scoring: rows[f] = (lo, hi) is the rank's slice of test fold
    f's rows.  Every (fold f, solving rank r) group of columns is scored on the rank's slice of fold f (kind 0: test
    residual sums) and, for train scores, on its slices of the other folds (kind 1).  Returns a list of
    (lo, hi, r, f, kind, cols): cols = the grid columns of the group (owner[f] == r), in the order they were sent.
    Summed over the ranks, kind 0 covers every (column, test row) pair exactly once."""
    owner = np.asarray(owner)
    n_splits = owner.shape[0]
    plan = []
    for f in range(n_splits):
        for r in range(world):
            cols = np.flatnonzero(owner[f] == r)
            if len(cols) == 0:
                continue
            lo, hi = rows[f]
            if hi > lo:
                plan.append((int(lo), int(hi), r, f, 0, cols))
            if want_train:
                for f2 in range(n_splits):
                    lo2, hi2 = rows[f2]
                    if f2 != f and hi2 > lo2:
                        plan.append((int(lo2), int(hi2), r, f, 1, cols))
    return plan


class GridShard:
    """Rank-local view of a sharded CV search."""

    def __init__(self, rank: int, world: int, group=None):
        self.rank, self.world, self.group = int(rank), int(world), group

    def row_range(self, n: int, start: int = 0):
        """This rank's contiguous slice of the rows [start, start + n)."""
        return start + (self.rank * n) // self.world, start + ((self.rank + 1) * n) // self.world

    def fold_row_ranges(self, row_ptr):
        """Per test fold f (rows row_ptr[f]..row_ptr[f+1]) the slice this rank builds."""
        return [self.row_range(int(row_ptr[f + 1] - row_ptr[f]), int(row_ptr[f])) for f in range(len(row_ptr) - 1)]

    def my_columns(self, n_folds: int, n_cols: int):
        """list over folds of the column indices this rank solves."""
        owner = assign_columns(n_folds, n_cols, self.world)
        return [np.flatnonzero(owner[f] == self.rank) for f in range(n_folds)]

    def comm_ptr(self, device=None):
        """The ncclComm_t of the process group (as an integer) for the engine's own collectives
        (slm_gram_allreduce / slm_allreduce_sum / slm_gather_results), or None when the group is not an
        initialised NCCL group (gloo on CPU, lazily created communicator): callers then go through
        torch.distributed."""
        if self.world == 1:
            return None
        if "_comm" not in self.__dict__:
            ptr = None
            try:
                import torch
                import torch.distributed as dist

                pg = self.group if self.group is not None else dist.distributed_c10d._get_default_group()
                dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
                ptr = int(pg._get_backend(dev)._comm_ptr())
            except Exception:
                ptr = None
            self.__dict__["_comm"] = ptr or None
        return self.__dict__["_comm"]

    def all_gather_(self, send, recv):
        """recv[r] = send of rank r (torch tensors; recv has a leading world dimension)."""
        import torch.distributed as dist

        # concatenated form [world * d0, ...]: the one every backend accepts (gloo rejects the stacked shape)
        flat = recv.view((recv.shape[0] * send.shape[0],) + tuple(send.shape[1:])) if send.dim() >= 1 else recv
        dist.all_gather_into_tensor(flat, send.contiguous(), group=self.group)
        return recv

    def allreduce_sum_(self, tensor):
        """In-place sum over ranks of a torch tensor (cuda -> NCCL, cpu -> gloo)."""
        if self.world == 1:
            return tensor
        import torch.distributed as dist

        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)
        return tensor

    def allreduce_sum_numpy(self, arr, device=None):
        """Sum over ranks of a (small) numpy array, returned as numpy."""
        if self.world == 1:
            return arr
        import torch

        t = torch.from_numpy(np.ascontiguousarray(arr))
        if device is not None:
            t = t.to(device)
        self.allreduce_sum_(t)
        return t.cpu().numpy()
