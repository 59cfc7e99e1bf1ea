"""cuBLAS DSYRK / DGEMM on the Gram-build blocks next to the engine's own SYRK kernel.

SURVEY section 7 step 4 names cuBLAS DSYRK as the kernel to beat.  Shapes: C3 (5 fold blocks
of 4000 x 4008) and one C5 block (80000 x 8008).  All times are CUDA events on the launching
stream, best of 4 after 2 warm-ups; TFLOP/s by the SYRK count rows*pa*(pa+1) for every arm
(a kernel that does the full 2*rows*pa^2 gets no extra credit).

    python tools/dsyrk_probe.py [c3] [c5]  ->  one JSON line per shape
"""
import ctypes
import glob
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparselm_b200.engine import Engine  # noqa: E402


def load_cublas():
    import nvidia.cublas  # noqa: F401  (torch's bundled wheel)

    base = list(nvidia.cublas.__path__)[0]
    path = sorted(glob.glob(os.path.join(base, "lib", "libcublas.so*")))[0]
    lib = ctypes.CDLL(path)
    h = ctypes.c_void_p()
    assert lib.cublasCreate_v2(ctypes.byref(h)) == 0
    return lib, h


def timed(fn, reps=4, warm=2):
    ts = []
    for i in range(warm + reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    return min(ts)


def run(shape):
    if shape == "c3":
        rows, p, F = 4000, 4000, 5
    else:
        rows, p, F = 80000, 8000, 1
    dev = torch.device("cuda", 0)
    eng = Engine(0)
    n = rows * F
    X = torch.randn(n, p, dtype=torch.float64, device=dev)
    y = np.random.default_rng(0).standard_normal(n)
    Xa = eng.pack(X, y)
    del X
    pa = Xa.shape[1]
    row_ptr = np.linspace(0, n, F + 1).astype(np.int64)
    syrk_flops = float(n) * pa * (pa + 1)
    out = {"shape": shape, "rows_per_block": rows, "blocks": F, "pa": pa}

    ms = timed(lambda: eng.gram_blocks(Xa, row_ptr))
    out["engine_syrk_ms"] = ms
    out["engine_syrk_tflops"] = syrk_flops / ms / 1e9
    G = eng.gram_blocks(Xa, row_ptr)

    lib, h = load_cublas()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.cublasSetStream_v2(h, stream) == 0
    C = torch.zeros((F, pa, pa), dtype=torch.float64, device=dev)
    one, zero = ctypes.c_double(1.0), ctypes.c_double(0.0)
    # row-major Xa[rows, pa] is the column-major pa x rows matrix A: G = A A^T, trans = N
    CUBLAS_FILL_LOWER, CUBLAS_OP_N, CUBLAS_OP_T = 0, 0, 1

    def dsyrk():
        for f in range(F):
            a = ctypes.c_void_p(Xa.data_ptr() + 8 * int(row_ptr[f]) * pa)
            c = ctypes.c_void_p(C[f].data_ptr())
            rc = lib.cublasDsyrk_v2(h, CUBLAS_FILL_LOWER, CUBLAS_OP_N, pa, rows, ctypes.byref(one), a, pa,
                                    ctypes.byref(zero), c, pa)
            assert rc == 0, rc

    ms = timed(dsyrk)
    out["cublas_dsyrk_ms"] = ms
    out["cublas_dsyrk_tflops"] = syrk_flops / ms / 1e9
    # column-major lower triangle == row-major upper triangle
    iu = torch.triu_indices(pa, pa, device=dev)
    err = float((C[0][iu[0], iu[1]] - G[0][iu[0], iu[1]]).abs().max() / G[0].abs().max())
    out["dsyrk_vs_engine_max_rel_diff"] = err

    def dgemm():
        for f in range(F):
            a = ctypes.c_void_p(Xa.data_ptr() + 8 * int(row_ptr[f]) * pa)
            c = ctypes.c_void_p(C[f].data_ptr())
            rc = lib.cublasDgemm_v2(h, CUBLAS_OP_N, CUBLAS_OP_T, pa, pa, rows, ctypes.byref(one), a, pa, a, pa,
                                    ctypes.byref(zero), c, pa)
            assert rc == 0, rc

    ms = timed(dgemm)
    out["cublas_dgemm_full_ms"] = ms
    out["cublas_dgemm_full_tflops_syrk_count"] = syrk_flops / ms / 1e9
    out["cublas_dgemm_full_tflops_executed"] = 2.0 * n * pa * pa / ms / 1e9
    lib.cublasDestroy_v2(h)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    for s in sys.argv[1:] or ["c3", "c5"]:
        run(s)
