set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02k_gpu_tests.log 2>&1
tail -15 gpurun_out/r02k_gpu_tests.log
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02k_bench_c4.json 2> gpurun_out/r02k_bench_c4.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02k_bench_c4.json") if l.startswith('{')][-1])
print("c4", round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['config']['unconverged'], d['roofline'].get('newton_phase',{}).get('ms_per_step'), {k:round(x,2) for k,x in d['roofline']['step_ms_by_kernel_family'].items()})
PY
tail -3 gpurun_out/r02k_bench_c4.err
timeout 900 bash tools/profile_round2.sh r02 2>&1 | tail -8
