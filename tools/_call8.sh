set -x
timeout 300 python -m pytest tests/test_gpu_newton.py -q -m gpu -k "kernels_match" 2>&1 | tail -30
