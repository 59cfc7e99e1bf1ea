"""Gram-apply tile / feeding variants at the C3 shapes, one process, pure kernel time (the engine's
own CUDA-event family timers): dense phase (every row in the support), mid-solve (per-chunk supports),
tail (few columns).  TMA-fed kernels against the cp.async ones.

    python tools/gemm_probe.py > gpurun_out/gemm_probe.txt
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparselm_b200.engine import Engine  # noqa: E402

p, F, K, ldz, pa = 4000, 5, 100, 104, 4008
dev = torch.device("cuda", 0)
eng = Engine(0)
G = torch.randn(F, pa, pa, dtype=torch.float64, device=dev)
G = G + G.transpose(1, 2)
rng = np.random.default_rng(0)


def scenario(supp, width, Kf=K):
    Zh = np.zeros((F, p, ldz))
    flops = 0.0
    for f in range(F):
        for cc, s in enumerate(supp):
            c0, c1 = cc * width, min(cc * width + width, Kf)
            if c1 <= c0:
                continue
            rows = rng.choice(p, s, replace=False)
            Zh[f, rows[:, None], np.arange(c0, c1)[None, :]] = rng.standard_normal((s, c1 - c0))
            flops += 2.0 * p * s * (c1 - c0)
    return torch.from_numpy(Zh).to(dev), flops


def timed(fn, reps=6, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    eng.timing_enable(True)
    eng.timing_reset()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    t = eng.timing_read()
    eng.timing_enable(False)
    return t["gram_apply"]["ms"] / reps


def line(tag, ms, flops, ref, out):
    err = float((out - ref).abs().max() / ref.abs().max()) if ref is not None else 0.0
    print(f"{tag:62s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:6.2f} TFLOP/s executed   rel diff {err:.1e}", flush=True)


# ---- A: dense phase -----------------------------------------------------------------------------
Zd, _ = scenario([p], 128)
fl = 2.0 * p * p * K * F
ref = None
for tma in (31, 7, 0):
    eng.set_option("tma", tma)
    for sid, nm in ((-1, "auto"), (0, "128x128 (2,4,8,4)"), (1, "128x104 (8,1,2,13)"), (2, "128x104 (16,1,1,13)")):
        eng.set_option("force_apply_shape", sid)
        ms = timed(lambda: eng.gram_apply(G, p, [K] * F, Zd))
        out = eng.gram_apply(G, p, [K] * F, Zd)
        ref = out.clone() if ref is None else ref
        line(f"A dense apply, tiled, tma={tma}, shape {nm}", ms, fl, ref, out)
    eng.set_option("force_apply_shape", -1)
    for sid, nm in ((-1, "auto"), (10, "(16,1,1,13)"), (12, "(8,1,2,13)"), (11, "128x128 (2,4,8,4)")):
        eng.set_option("force_sparse_shape", sid)
        ms = timed(lambda: eng.gram_apply_rowsparse(G, p, [K] * F, Zd, chunk_w=104))
        out = eng.gram_apply_rowsparse(G, p, [K] * F, Zd, chunk_w=104)
        line(f"A dense apply through the row-sparse kernel cw=104, tma={tma}, shape {nm}", ms, fl, ref, out)
    eng.set_option("force_sparse_shape", -1)

# ---- B: mid-solve --------------------------------------------------------------------------------
for supp32 in ([1950, 420, 388, 341], [3200, 1500, 900, 600], [900, 300, 250, 200]):
    Zm, flm = scenario(supp32, 32)
    ref = None
    for tma in (15, 7, 0):
        eng.set_option("tma", tma)
        for cw, sids in ((32, (-1, 3, 7)), (64, (-1, 13, 6)), (104, (-1, 10))):
            for sid in sids:
                eng.set_option("force_sparse_shape", sid)
                ms = timed(lambda: eng.gram_apply_rowsparse(G, p, [K] * F, Zm, chunk_w=cw))
                out = eng.gram_apply_rowsparse(G, p, [K] * F, Zm, chunk_w=cw)
                ref = out.clone() if ref is None else ref
                line(f"B mid-solve supports {supp32}, cw={cw}, tma={tma}, sparse shape {sid}", ms, flm, ref, out)
        eng.set_option("force_sparse_shape", -1)

# ---- C: tail ------------------------------------------------------------------------------------------
for Kt, s in ((16, 700), (8, 500)):
    Zt, flt = scenario([s], 32, Kf=Kt)
    ref = None
    for tma in (15, 0):
        eng.set_option("tma", tma)
        for sid in (-1, 0, 1, 2):
            eng.set_option("force_sparse_shape", sid)
            ms = timed(lambda: eng.gram_apply_rowsparse(G, p, [Kt] * F, Zt, chunk_w=32))
            out = eng.gram_apply_rowsparse(G, p, [Kt] * F, Zt, chunk_w=32)
            ref = out.clone() if ref is None else ref
            line(f"C tail K={Kt} support {s}, tma={tma}, sparse shape {sid}", ms, flt, ref, out)
    eng.set_option("force_sparse_shape", -1)
