set -x
mkdir -p gpurun_out
timeout 600 python tools/staircase_estimate.py c3 > gpurun_out/r02m_staircase_estimate.txt 2> gpurun_out/r02m_staircase_estimate.err
cut -c1-220 gpurun_out/r02m_staircase_estimate.txt | head -60; tail -5 gpurun_out/r02m_staircase_estimate.err
