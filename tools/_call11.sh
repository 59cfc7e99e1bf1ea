set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02g_gpu_tests.log 2>&1
tail -12 gpurun_out/r02g_gpu_tests.log
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02g_bench_c4.json 2> gpurun_out/r02g_bench_c4.err
tail -c 1500 gpurun_out/r02g_bench_c4.json; tail -3 gpurun_out/r02g_bench_c4.err
