"""Host-side cost of one device-resident CV step: cProfile over 20 steps of batched_cv (C3), sorted by
own time, with the waits inside the library / .cpu() calls listed separately.  Shows what the Python side
does while the GPU is idle (tools/timeline.py shows the gaps from the device side)."""
import cProfile
import io
import os
import pstats
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


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for _ in range(N):
    step()
torch.cuda.synchronize()
plain = (time.perf_counter() - t0) / N * 1e3
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
torch.cuda.synchronize()
pr.disable()
if rank == 0:
    print(f"step wall {plain:.3f} ms (unprofiled), world {world}")
    s = io.StringIO()
    ps = pstats.Stats(pr, stream=s).sort_stats("tottime")
    ps.print_stats(45)
    txt = s.getvalue()
    # per-step milliseconds
    print(txt[:9000])
if world > 1:
    torch.distributed.destroy_process_group()
