set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02j_gpu_tests.log 2>&1
tail -15 gpurun_out/r02j_gpu_tests.log
timeout 900 bash tools/profile_round2.sh r02 2>&1 | tail -8
