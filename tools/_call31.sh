set -x
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -q -m gpu --timeout=150 > gpurun_out/r02p_gpu_tests.log 2>&1
tail -8 gpurun_out/r02p_gpu_tests.log | cut -c1-250
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu > gpurun_out/r02p_bench_c4.json 2> gpurun_out/r02p_bench_c4.err
python - <<PY
import json
for f in ("gpurun_out/r02p_bench_c4.json",):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["ms_per_step"],2), "unconv", d["config"]["unconverged"], d["roofline"].get("newton_phase",{}).get("ms_per_step"))
PY
