set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02h_gpu_tests.log 2>&1
tail -8 gpurun_out/r02h_gpu_tests.log
for v in 1 0; do
SLM_PROX2=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-tall > gpurun_out/r02h_bench_prox$v.json 2> gpurun_out/r02h_bench_prox$v.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02h_bench_prox$v.json") if l.startswith('{')][-1])
print("prox2=$v", round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), round(d['roofline']['frac'],3), d['config']['unconverged'], d['config']['iterations_per_step'], {k:round(x,2) for k,x in d['roofline']['step_ms_by_kernel_family'].items()})
PY
done
SLM_PROX2=1 timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02h_bench_c2.json 2> gpurun_out/r02h_bench_c2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02h_bench_c2.json") if l.startswith('{')][-1])
print("c2", round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['config']['unconverged'], {k:round(x,2) for k,x in d['roofline']['step_ms_by_kernel_family'].items()})
PY
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02h_bench_c4.json 2> gpurun_out/r02h_bench_c4.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02h_bench_c4.json") if l.startswith('{')][-1])
print("c4", round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['config']['unconverged'], d['roofline'].get('newton_phase',{}).get('ms_per_step'), {k:round(x,2) for k,x in d['roofline']['step_ms_by_kernel_family'].items()})
PY
