"""Phase breakdown of one device-resident CV step per rank (run under torchrun for N > 1).

    python tools/scale_probe.py [c3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/scale_probe.py c3

Every phase method is wrapped with a synchronize on both sides (this removes overlap, so the
sum of the phases is an upper bound of the step; the un-instrumented step time is printed
next to it)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sklearn.base import clone  # noqa: E402
from sklearn.model_selection import KFold  # noqa: E402

from sparselm_b200 import engine as E  # noqa: E402
from sparselm_b200.model_selection import batched_cv  # noqa: E402

rank, world, local = bench.dist_init(0)
wl = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
X, y, est, alphas, F = wl["X"], wl["y"], wl["est"], wl["alphas"], wl["F"]
n, p = X.shape
engine = E.get_engine(local)
shard = None
if world > 1:
    from sparselm_b200.parallel import GridShard

    shard = GridShard(rank, world)
Xd = torch.from_numpy(X).to(engine.device)
folds = [te for _, te in KFold(F).split(X)]
ests = [clone(est).set_params(alpha=a) for a in alphas]
specs = [e._problem_spec(p) for e in ests]
opts = est._engine_options()


def step():
    return batched_cv(engine, Xd, y, folds, ests, specs, dict(opts), "neg_root_mean_squared_error", shard=shard)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


for _ in range(3):
    res = step()
times = []
for _ in range(5):
    barrier()
    t0 = time.perf_counter()
    res = step()
    torch.cuda.synchronize()
    times.append((time.perf_counter() - t0) * 1e3)
plain_ms = float(np.mean(times))

acc = {}


def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    label = label or name

    def inner(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize()
        acc[label] = acc.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
        return out

    setattr(obj, name, inner)


for nm in ("pack", "gram_blocks", "gram_complement", "gram_center", "lipschitz", "solve", "cv_score"):
    wrap(engine, nm)
if shard is not None:
    wrap(shard, "allreduce_sum_")
    wrap(shard, "allreduce_sum_numpy")
reps = 3
tot = []
for _ in range(reps):
    barrier()
    t0 = time.perf_counter()
    res = step()
    torch.cuda.synchronize()
    tot.append((time.perf_counter() - t0) * 1e3)
out = {"rank": rank, "world": world, "plain_step_ms": plain_ms, "instrumented_step_ms": float(np.mean(tot)),
       "phases_ms": {k: v / reps for k, v in acc.items()}, "iters": int(res["iters_run"])}
out["phases_ms"]["other(host+unwrapped)"] = out["instrumented_step_ms"] - sum(out["phases_ms"].values())
if world > 1:
    # the sharded search must reproduce the single-process score table
    for nm in ("pack", "gram_blocks", "gram_complement", "gram_center", "lipschitz", "solve", "cv_score"):
        pass
    ref = batched_cv(engine, Xd, y, folds, ests, specs, dict(opts), "neg_root_mean_squared_error")
    err = float(np.abs(ref["test_scores"] - res["test_scores"]).max() / np.abs(ref["test_scores"]).max())
    out["sharded_vs_single_rel_err"] = err
    assert err < 1e-8, err
    # the same through the public API on a host-resident X (rows travel only as far as needed)
    from sparselm_b200.model_selection import GridSearchCV

    gs = GridSearchCV(clone(est), {"alpha": list(alphas)}, cv=F)
    gs._shard = shard
    gs.fit(X, y)
    ms = gs.cv_results_["mean_test_score"]
    err2 = float(np.abs(ms - ref["test_scores"].mean(1)).max() / np.abs(ms).max())
    out["sharded_api_vs_single_rel_err"] = err2
    assert err2 < 1e-8, err2
for r in range(world):
    barrier()
    if r == rank:
        print(json.dumps(out), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
