set -x
mkdir -p gpurun_out
SLM_TRACE=1 timeout 300 python bench.py --workload c4 --steps 1 --warmup 0 --no-cpu > gpurun_out/r02m_c4_trace.json 2> gpurun_out/r02m_c4_trace.err
grep "newton step" gpurun_out/r02m_c4_trace.err | head -60 | cut -c1-260
grep -c "newton step" gpurun_out/r02m_c4_trace.err
