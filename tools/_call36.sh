set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_estimators.py tests/test_gpu_edge.py tests/test_gpu_coop.py -q --timeout=150 > gpurun_out/r02t_tests.log 2>&1
tail -5 gpurun_out/r02t_tests.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-tall > gpurun_out/r02t_bench_1gpu.json 2> gpurun_out/r02t_bench_1gpu.err
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02t_bench_c2.json 2> gpurun_out/r02t_bench_c2.err
SLM_FUSED_PROX=0 timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02t_bench_c2_classic.json 2> gpurun_out/r02t_bench_c2_classic.err
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu > gpurun_out/r02t_bench_c4.json 2> gpurun_out/r02t_bench_c4.err
python - <<PY
import json
for f in ("gpurun_out/r02t_bench_1gpu.json","gpurun_out/r02t_bench_c2.json","gpurun_out/r02t_bench_c2_classic.json","gpurun_out/r02t_bench_c4.json"):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["ms_per_step"],2), "unconv", d["config"]["unconverged"], "iters", d["config"]["iterations_per_step"], d["config"]["mean_iterations_per_fit"], "frac", round(d["roofline"]["frac"],3), d["roofline"]["kernel"][:30], d["roofline"]["step_ms_by_kernel_family"])
PY
tail -n 3 gpurun_out/r02t_bench_1gpu.err
