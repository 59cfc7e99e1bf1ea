"""Row-sparse Gram apply at a mid-solve C3 scenario (5 folds x 4 chunks of 32 columns, chunk
supports 1950/420/388/341 rows of 4000): ms per launch for forced tile variants."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparselm_b200.engine import Engine  # noqa: E402

p, F, K, ldz = 4000, 5, 100, 104
dev = torch.device("cuda", 0)
pa = 4008
G = torch.randn(F, pa, pa, dtype=torch.float64, device=dev)
G = G + G.transpose(1, 2)
rng = np.random.default_rng(0)
Zh = np.zeros((F, p, ldz))
supp = [1950, 420, 388, 341]
for f in range(F):
    for cc, s in enumerate(supp):
        rows = rng.choice(p, s, replace=False)
        c0, c1 = cc * 32, min(cc * 32 + 32, K)
        Zh[f, rows[:, None], np.arange(c0, c1)[None, :]] = rng.standard_normal((s, c1 - c0))
Z = torch.from_numpy(Zh).to(dev)
exec_flops = sum(2.0 * p * s * (min(cc * 32 + 32, K) - cc * 32) for cc, s in enumerate(supp)) * F
ref = None
for sid in [int(a) for a in sys.argv[1:]] or [3, 12, 13]:
    os.environ["SLM_FORCE_SPARSE_SHAPE"] = str(sid)
    eng = Engine(0)
    ts = []
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.gram_apply_rowsparse(G, p, [K] * F, Z, chunk_w=32)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    if ref is None:
        ref = out.clone()
    err = float((out - ref).abs().max() / ref.abs().max())
    print(f"sparse shape {sid}: {min(ts):.4f} ms (incl. list build + zero fill)  {exec_flops / min(ts) / 1e9:.2f} TFLOP/s executed  "
          f"rel diff {err:.1e}", flush=True)
    del eng
