"""How much contraction work would a per-slab column-block mask save in the row-sparse Gram apply?

Re-runs the C3 search with max_iter = 1, 2, ... and reads the iterates back (the engine is deterministic), so the
support pattern of every iteration is known.  For every iteration the contraction work (rows x columns, in units
of 8-column blocks x 16-row slabs) is counted for
  narrow : chunks of 32 columns, one support list per chunk (what the engine runs mid-solve)
  wide   : one chunk per fold contracting over the union support
  masked : one chunk per fold, support rows sorted by their number of active 8-column blocks, every 16-row slab
           skips the column blocks none of its rows touches (floor: WMIN blocks per slab = the L2->shared-memory
           time of the slab's band of G)
and weighted with the measured tile efficiencies (0.72 for 32-column tiles, 0.88 for 104-column tiles)."""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sklearn.base import clone  # noqa: E402
from sklearn.model_selection import KFold  # noqa: E402

from sparselm_b200 import engine as E  # noqa: E402
from sparselm_b200.model_selection import batched_cv  # noqa: E402

warnings.simplefilter("ignore")
wl = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
X, y, est, alphas, F = wl["X"], wl["y"], wl["est"], wl["alphas"], wl["F"]
n, p = X.shape
engine = E.get_engine(0)
Xd = torch.from_numpy(X).to(engine.device)
folds = [te for _, te in KFold(F).split(X)]
ests = [clone(est).set_params(alpha=a) for a in alphas]
specs = [e._problem_spec(p) for e in ests]
opts = est._engine_options()
WMIN = 5


def run(max_iter):
    o = dict(opts)
    o["max_iter"] = max_iter
    r = batched_cv(engine, Xd, y, folds, ests, specs, o, "neg_root_mean_squared_error")
    B = r["warm"]._batches[0][0]
    return (B != 0).cpu().numpy(), r


full_nz, rfull = run(20000)
n_iter = rfull["info"]["n_iter"]  # [n_cand, n_splits], candidate order
order = np.argsort([s.strength for s in specs], kind="stable")  # batch column order
T = int(n_iter.max())
K = len(specs)
prev = np.zeros_like(full_nz)
tot = {"narrow": 0.0, "wide": 0.0, "masked": 0.0, "masked_nofloor": 0.0, "exact": 0.0, "engine_like": 0.0}
print(f"iterations {T}, columns per fold {K}")
for t in range(1, T + 1):
    nz, _ = run(t)
    S = nz | prev  # support of the extrapolated point
    prev = nz
    row = {k: 0.0 for k in tot}
    for f in range(F):
        # columns still iterating (checked every 10 iterations; compaction keeps the order)
        t_check = (t - 1) // 10 * 10
        act = np.flatnonzero(n_iter[order, f] > t_check)
        if len(act) == 0:
            continue
        Sf = S[f][:, act]  # [p, Ka]
        Ka = Sf.shape[1]
        nb = (Ka + 7) // 8
        blk = np.zeros((p, nb), dtype=bool)
        for b in range(nb):
            blk[:, b] = Sf[:, 8 * b:8 * b + 8].any(axis=1)
        row["exact"] += Sf.sum()
        # narrow: chunks of 4 blocks
        nar = 0.0
        for c in range(0, nb, 4):
            rows = blk[:, c:c + 4].any(axis=1).sum()
            nar += rows * 8 * min(4, nb - c)
        wid = blk.any(axis=1).sum() * 8 * nb
        row["narrow"] += nar / 0.72
        row["wide"] += wid / 0.88
        row["engine_like"] += min(nar / 0.72, wid / 0.88)
        pc = blk.sum(axis=1)
        rows_sorted = np.argsort(-pc, kind="stable")
        rows_sorted = rows_sorted[pc[rows_sorted] > 0]
        m = 0.0
        m0 = 0.0
        for s0 in range(0, len(rows_sorted), 16):
            msk = blk[rows_sorted[s0:s0 + 16]].any(axis=0).sum()
            m += 16 * 8 * max(msk, min(WMIN, nb))
            m0 += 16 * 8 * msk
        row["masked"] += m / 0.88
        row["masked_nofloor"] += m0 / 0.88
    for k in tot:
        tot[k] += row[k]
    if t <= 12 or t % 5 == 0:
        print(f"it {t:3d}: " + "  ".join(f"{k} {v / 1e6:8.2f}" for k, v in row.items()), flush=True)
print("totals (M row-columns, efficiency-weighted except exact): " + "  ".join(f"{k} {v / 1e6:9.1f}" for k, v in tot.items()))
print(f"masked / engine_like = {tot['masked'] / tot['engine_like']:.3f}   masked_nofloor / engine_like = "
      f"{tot['masked_nofloor'] / tot['engine_like']:.3f}")
