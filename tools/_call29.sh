set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_newton.py tests/test_gpu_fullsize.py tests/test_gpu_tma.py tests/test_gpu_engine.py -x -q --timeout=200 > gpurun_out/r02o_tests.log 2>&1
tail -5 gpurun_out/r02o_tests.log | cut -c1-250
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu > gpurun_out/r02o_bench_c4.json 2> gpurun_out/r02o_bench_c4.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-tall > gpurun_out/r02o_bench_c3.json 2> gpurun_out/r02o_bench_c3.err
python - <<PY
import json
for f in ("gpurun_out/r02o_bench_c4.json","gpurun_out/r02o_bench_c3.json"):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["ms_per_step"],2), "unconv", d["config"]["unconverged"], "frac", round(d["roofline"]["frac"],3), d["roofline"]["step_ms_by_kernel_family"], d["roofline"].get("newton_phase",{}).get("ms_per_step"))
PY
tail -3 gpurun_out/r02o_bench_c4.err gpurun_out/r02o_bench_c3.err
timeout 300 python tools/timeline.py c4 > gpurun_out/r02o_timeline_c4.txt 2> gpurun_out/r02o_timeline_c4.err
cut -c1-200 gpurun_out/r02o_timeline_c4.txt | head -12
rm -f gpurun_out/timeline_w1_r0.json
