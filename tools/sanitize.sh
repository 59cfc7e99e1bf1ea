#!/bin/bash
# compute-sanitizer over the solver-kernel tests (run on the GPU box through gpurun).
# memcheck: out-of-bounds / misaligned global+shared accesses of every kernel of the engine;
# racecheck: shared-memory hazards (the CTA-level reductions, the cp.async/TMA tile rings, the
# cooperative kernels' shared buffers); synccheck: divergent barriers.
# Global-memory protocols (stream-K tile flags, grid barrier) are outside racecheck's model;
# they are covered by the bit-reproducibility tests (tests/test_gpu_engine.py).
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
  # memcheck sees every kernel family; the shared-memory tools skip the cooperative-launch tests
  # (grid barriers under racecheck instrumentation run for minutes per launch)
  # SAN_MORE: further test files for both tools (end of round 2: "tests/test_gpu_newton.py tests/test_gpu_tma.py";
  # racecheck on the shared-memory Cholesky kernels of test_gpu_newton.py needs more than 10 minutes)
  if [ $tool = memcheck ]; then TESTS="tests/test_gpu_engine.py tests/test_gpu_coop.py $SAN_MORE"; else TESTS="tests/test_gpu_engine.py $SAN_MORE"; fi
  timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool --print-limit 20 \
      python -m pytest $TESTS -q -m gpu -p no:cacheprovider ${SAN_PYTEST_ARGS} > $OUT/${TAG}_san_${tool}.log 2>&1
  echo "rc=$?" >> $OUT/${TAG}_san_${tool}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" $OUT/${TAG}_san_${tool}.log | tail -5
done
