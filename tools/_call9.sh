SLM_TRACE=1 timeout 300 python -m pytest tests/test_gpu_newton.py -q -m gpu -k "kernels_match" -s 2>&1 | grep -E "slm newton|passed|failed" | head -40
