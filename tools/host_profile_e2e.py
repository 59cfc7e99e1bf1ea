"""Host-side cost of the public call: cProfile over GridSearchCV.fit on a pinned host design (C3), sorted by
own time.  What Python does around the GPU work of bench.py's e2e number."""
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

from sparselm_b200.model_selection import GridSearchCV  # noqa: E402

wl = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
X, y, est, alphas, F = wl["X"], wl["y"], wl["est"], wl["alphas"], wl["F"]
Xh = torch.from_numpy(X).pin_memory().numpy()
grid = {"alpha": list(alphas)}


def step():
    return GridSearchCV(clone(est), grid, cv=F).fit(Xh, y)


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for _ in range(N):
    step()
torch.cuda.synchronize()
print(f"fit wall {(time.perf_counter() - t0) / N * 1e3:.3f} ms (unprofiled)")
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40)
print(s.getvalue()[:8000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(45)
print(s.getvalue()[:9000])
