"""Cost of the dense factorisations of the Newton phase (sparselm_b200/newton.py): FP64 Cholesky and
Cholesky solve of n x n, single and batched, on the device.  Measured on a B200 (round 1):
n = 1961: 0.90 ms single / 0.39 ms per matrix in a batch of 32 (6.5 TFLOP/s); cholesky_solve of one
right-hand side 0.30 / 0.36 ms; n = 1000: 0.44 / 0.095 ms; n = 500: 0.22 / 0.031 ms."""
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
