"""Cost of the dense factorisations an active-manifold Newton phase would need (see
tools/newton_polish_proto.py): FP64 Cholesky of n x n, single and batched, on the device."""
import json
import sys
import time

import torch

dev = torch.device("cuda", 0)
out = {}
for n in (500, 1000, 1961):
    A = torch.randn(n, n, dtype=torch.float64, device=dev)
    H = A @ A.T + n * torch.eye(n, dtype=torch.float64, device=dev)
    g = torch.randn(n, 1, dtype=torch.float64, device=dev)
    for batch in (1, 8, 32):
        Hb = H.expand(batch, n, n).contiguous() if batch > 1 else H
        gb = g.expand(batch, n, 1).contiguous() if batch > 1 else g
        for _ in range(2):
            L, info = torch.linalg.cholesky_ex(Hb)
            x = torch.cholesky_solve(gb, L)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        reps = 5
        e0.record()
        for _ in range(reps):
            L, info = torch.linalg.cholesky_ex(Hb)
        e1.record()
        for _ in range(reps):
            x = torch.cholesky_solve(gb, L)
        e2.record()
        torch.cuda.synchronize()
        out[f"n{n}_b{batch}"] = {"cholesky_ms": e0.elapsed_time(e1) / reps, "solve_ms": e1.elapsed_time(e2) / reps,
                                 "tflops": batch * n**3 / 3 / (e0.elapsed_time(e1) / reps * 1e-3) / 1e12}
print(json.dumps(out, indent=1))
