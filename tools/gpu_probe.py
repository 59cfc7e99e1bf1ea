"""One-shot GPU probe: FP64 peaks (cuBLAS DGEMM as the roofline denominator) and the
engine's dominant kernels at the C3 shape.  Writes gpurun_out/probe.json."""

import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparselm_b200.engine import PenaltyGrid, get_engine  # noqa: E402


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def main():
    out = {}
    eng = get_engine()
    dev = eng.device
    out["gpu"] = torch.cuda.get_device_name(0)
    out["sm_count"] = eng.sm_count
    # cuBLAS DGEMM peak
    N = 8192
    a = torch.randn(N, N, dtype=torch.float64, device=dev)
    b = torch.randn(N, N, dtype=torch.float64, device=dev)
    best, med = ev_time(lambda: torch.matmul(a, b), reps=10, warm=3)
    out["cublas_dgemm_8192_tflops_best"] = 2 * N ** 3 / best / 1e9
    out["cublas_dgemm_8192_tflops_median"] = 2 * N ** 3 / med / 1e9
    t0 = time.time()
    cnt = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(5):
            torch.matmul(a, b)
        cnt += 5
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    out["cublas_dgemm_8192_tflops_sustained"] = cnt * 2 * N ** 3 / e0.elapsed_time(e1) / 1e9
    del a, b
    # skinny cuBLAS: (4096x4096) @ (4096x104), batched x5 -- what the library does on the apply shape
    p, K, F = 4096, 100, 5
    Gc = torch.randn(F, p, p, dtype=torch.float64, device=dev)
    Zc = torch.randn(F, p, 104, dtype=torch.float64, device=dev)
    best, med = ev_time(lambda: torch.bmm(Gc, Zc), reps=10, warm=3)
    out["cublas_bmm_apply_shape_tflops_best"] = 2 * F * p * p * 104 / best / 1e9
    out["cublas_bmm_apply_ms"] = best
    del Gc, Zc

    # engine: C3 shape
    n = 20000
    X = torch.randn(n, p, dtype=torch.float64, device=dev)
    w = torch.zeros(p, dtype=torch.float64, device=dev)
    w[torch.randperm(p, device=dev)[: p // 10]] = 100 * torch.rand(p // 10, dtype=torch.float64, device=dev)
    y = X @ w + 10 * torch.randn(n, dtype=torch.float64, device=dev)
    Xa = eng.pack(X, y)
    pa = Xa.shape[1]
    row_ptr = np.linspace(0, n, F + 1).astype(np.int64)
    G = eng.gram_blocks(Xa, row_ptr, extra=1)
    best, med = ev_time(lambda: eng.lib.slm_gram_blocks(eng.h, eng._ptr(Xa), pa, row_ptr.ctypes.data_as(
        __import__("ctypes").POINTER(__import__("ctypes").c_int64)), F, eng._ptr(G), eng.stream), reps=5, warm=1)
    syrk_flops = n * pa * (pa + 1)
    out["gram_build_ms_best"] = best
    out["gram_build_tflops_syrk_count"] = syrk_flops / best / 1e9
    out["gram_build_tflops_executed"] = (2.0 * n * pa * pa * (33 * 34 / 2) / (33 * 33)) / best / 1e9
    # cuBLAS reference for the same job (full GEMM X^T X per block)
    Xs = X[: n // F]
    best_c, _ = ev_time(lambda: torch.matmul(Xs.T, Xs), reps=5, warm=2)
    out["cublas_xtx_block_ms"] = best_c
    out["cublas_xtx_block_tflops_full_count"] = 2 * (n // F) * p * p / best_c / 1e9
    eng.gram_complement(G, F, out=G[F])
    # apply at full width for different K
    for Kc in (100, 104, 64, 32, 8):
        ldz = max(8, (Kc + 7) // 8 * 8)
        Z = torch.randn(F, p, ldz, dtype=torch.float64, device=dev)
        GZ = torch.empty_like(Z)
        import ctypes
        Karr = (ctypes.c_int32 * F)(*([Kc] * F))
        fn = lambda: eng.lib.slm_gram_apply(eng.h, eng._ptr(G), pa * pa, pa, p, F, Karr, eng._ptr(Z), ldz,
                                            eng._ptr(GZ), eng.stream)
        best, med = ev_time(fn, reps=10, warm=3)
        out[f"apply_K{Kc}_ms_best"] = best
        out[f"apply_K{Kc}_ms_median"] = med
        out[f"apply_K{Kc}_tflops"] = 2.0 * F * p * p * Kc / best / 1e9
        out[f"apply_K{Kc}_GBps"] = 8.0 * F * p * p / best / 1e6
    # per-shape sweep at K=104 (forced tile shape): per-tile efficiency of each warp layout
    import ctypes as _ct
    names = ["128x128/8w", "128x104/8w", "128x104/16w", "128x56/8w x2", "128x128/16w", "128x64/8w x2",
             "64x104/8w x2", "128x32/8w x2"]
    from sparselm_b200.engine import Engine
    sweep = {}
    for Kc in (104, 64, 32):
        ldz = Kc
        Z = torch.randn(F, p, ldz, dtype=torch.float64, device=dev)
        GZ = torch.empty_like(Z)
        Karr = (_ct.c_int32 * F)(*([Kc] * F))
        for sid in range(len(names)):
            os.environ["SLM_FORCE_APPLY_SHAPE"] = str(sid)
            e2 = Engine(0)
            fn = lambda: e2.lib.slm_gram_apply(e2.h, e2._ptr(G), pa * pa, pa, p, F, Karr, e2._ptr(Z), ldz,
                                               e2._ptr(GZ), e2.stream)
            best, med = ev_time(fn, reps=10, warm=3)
            sweep[f"K{Kc} {names[sid]}"] = {"ms": best, "useful_tflops": 2.0 * F * p * p * Kc / best / 1e9}
            del e2
    os.environ.pop("SLM_FORCE_APPLY_SHAPE", None)
    for sid in range(2):
        os.environ["SLM_FORCE_SYRK_SHAPE"] = str(sid)
        e2 = Engine(0)
        fn = lambda: e2.lib.slm_gram_blocks(e2.h, e2._ptr(Xa), pa, row_ptr.ctypes.data_as(_ct.POINTER(_ct.c_int64)), F,
                                            e2._ptr(G), e2.stream)
        Gsave = G.clone()
        best, med = ev_time(fn, reps=3, warm=1)
        sweep[f"syrk shape {sid}"] = {"ms": best, "tflops_syrk_count": syrk_flops / best / 1e9}
        G.copy_(Gsave)
        del e2, Gsave
    os.environ.pop("SLM_FORCE_SYRK_SHAPE", None)
    os.environ.pop("SLM_FORCE_APPLY_SHAPE", None)
    out["shape_sweep"] = sweep
    # full C3-like solve (SGL, 200 groups x ~20)
    torch.cuda.synchronize()
    t0 = time.time()
    fd = eng.prepare(X, y, test_folds=[np.arange(row_ptr[f], row_ptr[f + 1]) for f in range(F)])
    torch.cuda.synchronize()
    out["prepare_s"] = time.time() - t0
    c = fd.G_full[p, :p].cpu().numpy()
    amax = np.abs(c).max() / n
    alphas = amax * np.logspace(0, -3, K)
    Gn = 200
    gptr = np.linspace(0, p, Gn + 1).astype(np.int32)
    grid = PenaltyGrid(p=p, lam1=0.5 * alphas, gptr=gptr, W2=np.tile((0.5 * alphas)[None, :], (Gn, 1)))
    eng.timing_enable(True)
    for tol in (1e-9, 1e-11):
        eng.timing_reset()
        torch.cuda.synchronize()
        t0 = time.time()
        res = eng.solve(fd.G_train, p, fd.n_train, fd.lipschitz(eng, list(range(F))), [grid] * F, tol=tol)
        torch.cuda.synchronize()
        dt = time.time() - t0
        tim = eng.timing_read()
        out[f"solve_tol{tol:g}"] = {
            "wall_s": dt, "iters_run": res["iters_run"], "n_unconverged": res["n_unconverged"],
            "n_iter_mean": float(res["n_iter"][:, :K].mean()), "n_iter_max": int(res["n_iter"][:, :K].max()),
            "fits_per_s": F * K / dt, "timing": tim,
            "max_rel_gap": float((res["gap"][:, :K] / np.abs(res["primal"][:, :K])).max()),
        }
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.json", "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
