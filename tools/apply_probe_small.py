"""Row-sparse Gram apply in the latency regime of an 8-GPU rank of C3: 62 columns of one fold (or 38 + 25 of
two folds), nested supports from 40 to 2400 rows of 4000, alphas interleaved.  ms per launch for forced tile
variants and chunk widths (20 back-to-back launches, CUDA events around the train)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparselm_b200.engine import Engine  # noqa: E402

p, pa = 4000, 4008
dev = torch.device("cuda", 0)
two = len(sys.argv) > 1 and sys.argv[1] == "two"
Ks = [38, 25] if two else [62]
F = len(Ks)
ldz = 64
G = torch.randn(F, pa, pa, dtype=torch.float64, device=dev)
G = G + G.transpose(1, 2)
rng = np.random.default_rng(0)
Zh = np.zeros((F, p, ldz))
tot = 0.0
for f, K in enumerate(Ks):
    perm = rng.permutation(p)
    sizes = np.round(np.geomspace(2400, 40, K)).astype(int)  # column 0 = smallest alpha
    for k, s in enumerate(sizes):
        Zh[f, perm[:s], k] = rng.standard_normal(s)
        tot += 2.0 * p * s
Z = torch.from_numpy(Zh).to(dev)
ref = None
REP = 20
for cw in (16, 32, 64):
    for sid in range(14):
        os.environ["SLM_FORCE_SPARSE_SHAPE"] = str(sid)
        eng = Engine(0)
        try:
            for _ in range(3):
                out = eng.gram_apply_rowsparse(G, p, Ks, Z, chunk_w=cw)
            torch.cuda.synchronize()
            eng.timing_enable(True)
            eng.timing_reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(REP):
                out = eng.gram_apply_rowsparse(G, p, Ks, Z, chunk_w=cw)
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / REP
        except Exception as ex:  # a shape narrower than the chunk is refused
            print(f"cw {cw} shape {sid}: {str(ex)[:80]}", flush=True)
            del eng
            continue
        if ref is None:
            ref = out.clone()
        err = float((out - ref).abs().max() / ref.abs().max())
        tm = eng.timing_read()["gram_apply"]
        tk = tm["ms"] / max(1, tm["launches"])
        print(f"cw {cw:3d} shape {sid:2d}: call {t * 1e3:7.1f} us  GEMM kernel {tk * 1e3:7.1f} us x{tm['launches'] // REP}  "
              f"column-exact {tot / tk / 1e9:6.2f} TF/s  rel diff {err:.1e}", flush=True)
        del eng
