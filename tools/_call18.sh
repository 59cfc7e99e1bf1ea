set -x
mkdir -p gpurun_out
timeout 300 python tools/apply_probe_small.py > gpurun_out/r02m_apply_small_one.txt 2>&1
cat gpurun_out/r02m_apply_small_one.txt | cut -c1-200
timeout 300 python tools/apply_probe_small.py two > gpurun_out/r02m_apply_small_two.txt 2>&1
cat gpurun_out/r02m_apply_small_two.txt | cut -c1-200
