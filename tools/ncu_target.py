"""Short workload for ncu: a few launches of each hot kernel at the C3 shape."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparselm_b200.engine import PenaltyGrid, get_engine  # noqa: E402

eng = get_engine()
dev = eng.device
n, p, F, K = int(os.environ.get("NCU_N", 20000)), int(os.environ.get("NCU_P", 4096)), 5, 100
X = torch.randn(n, p, dtype=torch.float64, device=dev)
w = torch.zeros(p, dtype=torch.float64, device=dev)
w[: p // 10] = 100 * torch.rand(p // 10, dtype=torch.float64, device=dev)
y = X @ w + 10 * torch.randn(n, dtype=torch.float64, device=dev)
folds = [np.arange(f * n // F, (f + 1) * n // F) for f in range(F)]
fd = eng.prepare(X, y, test_folds=folds)
c = fd.G_full[p, :p].cpu().numpy()
alphas = np.abs(c).max() / n * np.logspace(0, -3, K)
Gn = 200
gptr = np.linspace(0, p, Gn + 1).astype(np.int32)
grid = PenaltyGrid(p=p, lam1=0.5 * alphas, gptr=gptr, W2=np.tile((0.5 * alphas)[None, :], (Gn, 1)))
res = eng.solve(fd.G_train, p, fd.n_train, fd.lipschitz(eng, list(range(F))), [grid] * F, tol=1e-9, max_iter=int(os.environ.get("NCU_ITERS", 12)))
torch.cuda.synchronize()
print("iters", res["iters_run"])
