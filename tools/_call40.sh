set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-tall > gpurun_out/r02zz_bench_1gpu.json 2> gpurun_out/r02zz_bench_1gpu.err
timeout 300 python -m pytest tests/test_gpu_estimators.py -q --timeout=150 -x > gpurun_out/r02zz_tests.log 2>&1
tail -3 gpurun_out/r02zz_tests.log
python - <<PY
import json
for l in open("gpurun_out/r02zz_bench_1gpu.json"):
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["ms_per_step"],2), "unconv", d["config"]["unconverged"])
PY
tail -n 3 gpurun_out/r02zz_bench_1gpu.err
