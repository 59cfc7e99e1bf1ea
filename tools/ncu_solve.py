"""ncu target: the bench workload's device-resident CV step, run twice (the second one is the
one to profile: pick launches of it with --launch-skip).  Usage: python tools/ncu_solve.py [c3]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sklearn.base import clone  # noqa: E402
from sklearn.model_selection import KFold  # noqa: E402

from sparselm_b200.engine import get_engine  # noqa: E402
from sparselm_b200.model_selection import batched_cv  # noqa: E402

wl = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
X, y, est, alphas, F = wl["X"], wl["y"], wl["est"], wl["alphas"], wl["F"]
n, p = X.shape
engine = get_engine(0)
Xd = torch.from_numpy(X).to(engine.device)
folds = [te for _, te in KFold(F).split(X)]
ests = [clone(est).set_params(alpha=a) for a in alphas]
specs = [e._problem_spec(p) for e in ests]
opts = est._engine_options()
for i in range(int(os.environ.get("NCU_STEPS", 2))):
    res = batched_cv(engine, Xd, y, folds, ests, specs, dict(opts), "neg_root_mean_squared_error")
    torch.cuda.synchronize()
    print("step", i, "iters", res["iters_run"], "launches", engine.launch_count())
