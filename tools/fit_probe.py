"""Wall time of a plain estimator.fit from pinned host memory (C3 design, one alpha)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sklearn.base import clone  # noqa: E402

wl = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
X, y, est, alphas = wl["X"], wl["y"], wl["est"], wl["alphas"]
Xh = torch.from_numpy(X).pin_memory().numpy()
for ai in (30, 60):
    e = clone(est).set_params(alpha=alphas[ai])
    ts = []
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e.fit(Xh, y)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"alpha[{ai}]: fit from pinned host {min(ts[1:]):.2f} ms (H2D of X alone is {X.nbytes / 55e9 * 1e3:.1f} ms at 55 GB/s), "
          f"iterations {e.solver_info_['iterations']}, nnz {int((e.coef_ != 0).sum())}")
