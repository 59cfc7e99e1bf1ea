set -x
mkdir -p gpurun_out
export SAN_TIMEOUT=700
for tool in memcheck racecheck; do
  T0=$(date +%s)
  if [ $tool = memcheck ]; then TESTS="tests/test_gpu_engine.py tests/test_gpu_coop.py tests/test_gpu_newton.py tests/test_gpu_tma.py"; else TESTS="tests/test_gpu_engine.py tests/test_gpu_newton.py tests/test_gpu_tma.py"; fi
  timeout 800 compute-sanitizer --tool $tool --print-limit 20 \
      python -m pytest $TESTS -q -m gpu -p no:cacheprovider > gpurun_out/r02z_san_${tool}.log 2>&1
  echo "rc=$? elapsed $(( $(date +%s) - T0 )) s" >> gpurun_out/r02z_san_${tool}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=|elapsed" gpurun_out/r02z_san_${tool}.log | tail -6
done
