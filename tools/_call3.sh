set -x
mkdir -p gpurun_out
timeout 600 python tools/gemm_probe.py > gpurun_out/r02b_gemm_probe.txt 2> gpurun_out/r02b_gemm_probe.err
cat gpurun_out/r02b_gemm_probe.txt; tail -3 gpurun_out/r02b_gemm_probe.err
timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -x -k "certified or tall" > gpurun_out/r02b_fullsize.log 2>&1
tail -15 gpurun_out/r02b_fullsize.log
SLM_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r02b_trace_bench.json 2> gpurun_out/r02b_trace.err
grep -c "slm" gpurun_out/r02b_trace.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02b_launches.log 2>&1
tail -2 gpurun_out/r02b_launches.log | cut -c1-300
