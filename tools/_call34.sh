set -x
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -q -m gpu --timeout=150 > gpurun_out/r02s_gpu_tests.log 2>&1
tail -30 gpurun_out/r02s_gpu_tests.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-tall > gpurun_out/r02s_bench_1gpu.json 2> gpurun_out/r02s_bench_1gpu.err
SLM_FUSED_PROX=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-tall > gpurun_out/r02s_bench_1gpu_classic.json 2> gpurun_out/r02s_bench_1gpu_classic.err
python - <<PY
import json
for f in ("gpurun_out/r02s_bench_1gpu.json","gpurun_out/r02s_bench_1gpu_classic.json"):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["ms_per_step"],2), "unconv", d["config"]["unconverged"], "iters", d["config"]["iterations_per_step"], d["config"]["mean_iterations_per_fit"], "frac", round(d["roofline"]["frac"],3), "sf", round(d["roofline"]["support_fraction"],3), d["roofline"]["step_ms_by_kernel_family"])
PY
tail -n 3 gpurun_out/r02s_bench_1gpu.err
