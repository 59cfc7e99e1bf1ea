set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout=150 > gpurun_out/r02zz_gpu_tests_final.log 2>&1
tail -4 gpurun_out/r02zz_gpu_tests_final.log | cut -c1-300
