set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_newton.py tests/test_gpu_edge.py tests/test_gpu_tma.py -q -m gpu > gpurun_out/r02e_tests.log 2>&1
tail -40 gpurun_out/r02e_tests.log
timeout 600 python tools/check_estimators.py > gpurun_out/r02e_check_estimators.txt 2> gpurun_out/r02e_check_estimators.err
cat gpurun_out/r02e_check_estimators.txt | cut -c1-400; tail -3 gpurun_out/r02e_check_estimators.err
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02e_bench_c4.json 2> gpurun_out/r02e_bench_c4.err
tail -c 1800 gpurun_out/r02e_bench_c4.json; tail -3 gpurun_out/r02e_bench_c4.err
SLM_NEWTON_TORCH=1 timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02e_bench_c4_torch.json 2> gpurun_out/r02e_bench_c4_torch.err
tail -c 1000 gpurun_out/r02e_bench_c4_torch.json
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_bench_1gpu.json 2> gpurun_out/r02e_bench_1gpu.err
tail -c 3500 gpurun_out/r02e_bench_1gpu.json; tail -3 gpurun_out/r02e_bench_1gpu.err
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/r02e_bench_c2.json 2> gpurun_out/r02e_bench_c2.err
tail -c 1500 gpurun_out/r02e_bench_c2.json; tail -3 gpurun_out/r02e_bench_c2.err
