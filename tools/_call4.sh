set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tma.py -x -q > gpurun_out/r02c_tma_tests.log 2>&1
TMA_RC=$?
tail -12 gpurun_out/r02c_tma_tests.log
timeout 600 python tools/gemm_probe.py > gpurun_out/r02c_gemm_probe.txt 2> gpurun_out/r02c_gemm_probe.err
cat gpurun_out/r02c_gemm_probe.txt; tail -3 gpurun_out/r02c_gemm_probe.err
for m in 15 31 7; do
  SLM_TMA=$m timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-tall > gpurun_out/r02c_bench_tma$m.json 2> gpurun_out/r02c_bench_tma$m.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/r02c_bench_tma$m.json'))
print('mask $m', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['step_ms_by_kernel_family'], d['config']['unconverged'])"
  tail -2 gpurun_out/r02c_bench_tma$m.err
done
python tools/dsyrk_probe.py c3 > gpurun_out/r02c_dsyrk.jsonl 2>&1; SLM_TMA=31 python tools/dsyrk_probe.py c3 >> gpurun_out/r02c_dsyrk.jsonl 2>&1; cat gpurun_out/r02c_dsyrk.jsonl | cut -c1-250
