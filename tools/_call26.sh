set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_newton.py tests/test_gpu_fullsize.py -x -q --timeout=200 > gpurun_out/r02n_newton_tests.log 2>&1
tail -8 gpurun_out/r02n_newton_tests.log | cut -c1-250
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu > gpurun_out/r02n_bench_c4.json 2> gpurun_out/r02n_bench_c4.err
python - <<PY
import json
for l in open("gpurun_out/r02n_bench_c4.json"):
    if l.startswith("{"):
        d=json.loads(l); print("C4:", round(d["ms_per_step"],1), "ms e2e", round(d["e2e"]["ms_per_step"],1), "unconv", d["config"]["unconverged"], d["roofline"].get("newton_phase"))
PY
tail -3 gpurun_out/r02n_bench_c4.err
timeout 300 python tools/timeline.py c4 > gpurun_out/r02n_timeline_c4.txt 2> gpurun_out/r02n_timeline_c4.err
cut -c1-200 gpurun_out/r02n_timeline_c4.txt | head -16
rm -f gpurun_out/timeline_w1_r0.json
