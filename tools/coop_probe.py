"""Single-column solves on the C3 design: cooperative iterations against the regular kernels.

    python tools/coop_probe.py [c3|c2]

Per alpha (three points of the grid): wall time of the solve (Gram resident), iterations,
launches, support size, time inside the iteration kernels."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sklearn.base import clone  # noqa: E402

from sparselm_b200 import engine as E  # noqa: E402
from sparselm_b200.model._base import solve_specs  # noqa: E402

wl = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
X, y, est, alphas = wl["X"], wl["y"], wl["est"], wl["alphas"]
n, p = X.shape
eng = E.get_engine(0)
spec0 = clone(est).set_params(alpha=alphas[0])._problem_spec(p)
fd = eng.prepare(torch.from_numpy(X).to(eng.device), y, None, False, None, col_perm=spec0.col_perm)
torch.cuda.synchronize()
for ai in (20, 50, 80):
    spec = clone(est).set_params(alpha=alphas[ai])._problem_spec(p)
    for on in (1, 0):
        eng.set_option("coop", on)
        for rep in range(3):
            eng.timing_enable(True)
            eng.timing_reset()
            l0 = eng.launch_count()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = solve_specs(eng, fd, [spec], use_full=True, tol=1e-9)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            tim = eng.timing_read()
            eng.timing_enable(False)
        nnz = int((out["coef"][0, :, 0] != 0).sum().item())
        print(f"alpha[{ai}] coop={on}: {dt:7.3f} ms  iters_run={out['iters_run']} n_iter={int(out['n_iter'][0, 0])} "
              f"launches={eng.launch_count() - l0} nnz={nnz} prox_ms={tim['prox']['ms']:.3f} "
              f"apply_ms={tim['gram_apply']['ms']:.3f} gap_ms={tim['gap']['ms']:.3f}", flush=True)
