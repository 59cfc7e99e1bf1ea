"""Kernel timeline of one device-resident CV step (CUPTI through torch.profiler; no nsys here).

    python tools/timeline.py [c3]                # or under torchrun for N > 1

Prints, per rank: step wall time, summed kernel time, idle time of the GPU inside the step,
time per kernel name, and the largest idle gaps with the kernels on either side.  The chrome
trace goes to gpurun_out/timeline_w{world}_r{rank}.json."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sklearn.base import clone  # noqa: E402
from sklearn.model_selection import KFold  # noqa: E402

from sparselm_b200 import engine as E  # noqa: E402
from sparselm_b200.model_selection import batched_cv  # noqa: E402

rank, world, local = bench.dist_init(0)
wl = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
X, y, est, alphas, F = wl["X"], wl["y"], wl["est"], wl["alphas"], wl["F"]
n, p = X.shape
engine = E.get_engine(local)
shard = None
if world > 1:
    from sparselm_b200.parallel import GridShard

    shard = GridShard(rank, world)
Xd = torch.from_numpy(X).to(engine.device)
folds = [te for _, te in KFold(F).split(X)]
ests = [clone(est).set_params(alpha=a) for a in alphas]
specs = [e._problem_spec(p) for e in ests]
opts = est._engine_options()


E2E = len(sys.argv) > 2 and sys.argv[2] == "e2e"  # public API on pinned host arrays (refit included)
if E2E:
    from sparselm_b200.model_selection import GridSearchCV  # noqa: E402

    Xh = torch.from_numpy(X).pin_memory().numpy()
    grid = {"alpha": list(alphas)}


def step():
    if E2E:
        gs = GridSearchCV(clone(est), grid, cv=F)
        if shard is not None:
            gs._shard = shard
        return gs.fit(Xh, y)
    return batched_cv(engine, Xd, y, folds, ests, specs, dict(opts), "neg_root_mean_squared_error", shard=shard)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


for _ in range(3):
    step()
barrier()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    barrier()
    t0 = time.perf_counter()
    step()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
os.makedirs("gpurun_out", exist_ok=True)
# the chrome trace of rank 0 is kept (gpurun_out/ travels back, 64 MiB at most); other ranks parse theirs from /tmp
path = f"{'gpurun_out' if rank == 0 else '/tmp'}/timeline{'_e2e' if E2E else ''}_w{world}_r{rank}.json"
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
ks = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e],
            key=lambda e: e["ts"])
# keep the kernels of the profiled step: after the last barrier-ish gap; simply take all (the
# profile region holds one barrier + one step)
busy = 0.0
end = None
gaps = []
by_name = {}
for i, e in enumerate(ks):
    s, d = e["ts"], e["dur"]
    nm = e["name"][:70]
    by_name.setdefault(nm, [0, 0.0])
    by_name[nm][0] += 1
    by_name[nm][1] += d
    if end is not None and s > end:
        gaps.append((s - end, ks[i - 1]["name"][:50], nm[:50]))
    if end is None or s + d > end:
        busy += (s + d - max(s, end)) if end is not None and end > s else d
        end = s + d
span = (ks[-1]["ts"] + ks[-1]["dur"] - ks[0]["ts"]) if ks else 0.0
out = {"rank": rank, "world": world, "step_wall_ms": wall, "kernel_span_ms": span / 1e3, "gpu_busy_ms": busy / 1e3,
       "gpu_idle_ms": (span - busy) / 1e3, "n_kernels": len(ks)}
lines = [json.dumps(out)]
for nm, (c, d) in sorted(by_name.items(), key=lambda kv: -kv[1][1])[:25]:
    lines.append(f"  {d / 1e3:9.3f} ms  x{c:<5d} {nm}")
lines.append("  largest idle gaps (us): prev kernel -> next kernel")
for g, a, b in sorted(gaps, reverse=True)[:25]:
    lines.append(f"  {g:9.1f}  {a} -> {b}")
lines.append(f"  gaps > 20us: {sum(1 for g in gaps if g[0] > 20)} totalling {sum(g[0] for g in gaps if g[0] > 20) / 1e3:.3f} ms;"
             f" gaps <= 20us: {sum(1 for g in gaps if g[0] <= 20)} totalling {sum(g[0] for g in gaps if g[0] <= 20) / 1e3:.3f} ms")
# what the host was doing during the large gaps: CPU-side events overlapping each gap window
cpu = [e for e in ev if e.get("cat") in ("cpu_op", "cuda_runtime", "user_annotation", "python_function") and "dur" in e]
lines.append("  host activity inside the gaps > 20us (name: us overlapped)")
gi = []
end = None
for i, e in enumerate(ks):
    s_, d_ = e["ts"], e["dur"]
    if end is not None and s_ - end > 20:
        gi.append((end, s_, ks[i - 1]["name"][:40], e["name"][:40]))
    end = s_ + d_ if end is None else max(end, s_ + d_)
for a, b, pn, nn in sorted(gi, key=lambda g: g[0] - g[1])[:14]:
    acts = {}
    for c in cpu:
        o = min(b, c["ts"] + c["dur"]) - max(a, c["ts"])
        if o > 2:
            acts[c["name"][:48]] = acts.get(c["name"][:48], 0.0) + o
    top = sorted(acts.items(), key=lambda kv: -kv[1])[:7]
    lines.append(f"  gap {b - a:8.1f} us  [{pn} -> {nn}]: " + "; ".join(f"{k}: {v:.0f}" for k, v in top))
for r in range(world):
    barrier()
    if r == rank:
        print("\n".join(lines), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
