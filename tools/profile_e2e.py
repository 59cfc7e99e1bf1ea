"""cProfile of one end-to-end GridSearchCV.fit on the C3 workload (host-side overheads)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sklearn.base import clone  # noqa: E402

from sparselm_b200.model_selection import GridSearchCV  # noqa: E402

wl = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
X, y, est, alphas, F = wl["X"], wl["y"], wl["est"], wl["alphas"], wl["F"]
Xh = torch.from_numpy(X).pin_memory().numpy()
grid = {"alpha": list(alphas)}
for _ in range(2):
    GridSearchCV(clone(est), grid, cv=F).fit(Xh, y)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
gs = GridSearchCV(clone(est), grid, cv=F).fit(Xh, y)
torch.cuda.synchronize()
pr.disable()
print("e2e ms", (time.perf_counter() - t0) * 1e3)
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
