set -x
mkdir -p gpurun_out
timeout 120 python tools/newton_debug.py 203 > gpurun_out/r02f_newton_debug.txt 2>&1
cat gpurun_out/r02f_newton_debug.txt | tail -30
timeout 120 python tools/newton_debug.py 64 2>&1 | tail -12
timeout 300 python -m pytest tests/test_gpu_newton.py -q -m gpu -x 2>&1 | tail -15
