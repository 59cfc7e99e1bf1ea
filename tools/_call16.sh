set -x
mkdir -p gpurun_out
timeout 800 python -m pytest tests -q -m gpu --timeout=150 > gpurun_out/r02l_gpu_tests.log 2>&1
tail -25 gpurun_out/r02l_gpu_tests.log | cut -c1-250
