"""Gram-build (SYRK) tile/pipeline variants at the C3 shape: ms and TFLOP/s (SYRK count)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparselm_b200.engine import Engine  # noqa: E402

n, p, F = 20000, 4000, 5
dev = torch.device("cuda", 0)
X = torch.randn(n, p, dtype=torch.float64, device=dev)
y = np.random.default_rng(0).standard_normal(n)
row_ptr = np.linspace(0, n, F + 1).astype(np.int64)
ref = None
for sid in [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3]:
    os.environ["SLM_FORCE_SYRK_SHAPE"] = str(sid)
    eng = Engine(0)
    Xa = eng.pack(X, y)
    pa = Xa.shape[1]
    ts = []
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        G = eng.gram_blocks(Xa, row_ptr)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    if ref is None:
        ref = G.clone()
    err = float((G - ref).abs().max() / ref.abs().max())
    print(f"syrk shape {sid}: {ms:.3f} ms  {n * pa * (pa + 1) / ms / 1e9:.2f} TFLOP/s (SYRK count)  max rel diff vs shape 0 {err:.1e}",
          flush=True)
    del eng, Xa, G
