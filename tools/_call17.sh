set -x
mkdir -p gpurun_out
timeout 300 python tools/timeline.py c3 > gpurun_out/r02m_timeline_1gpu.txt 2> gpurun_out/r02m_timeline_1gpu.err
cut -c1-400 gpurun_out/r02m_timeline_1gpu.txt | head -80; tail -3 gpurun_out/r02m_timeline_1gpu.err
timeout 300 python tools/timeline.py c3 e2e > gpurun_out/r02m_timeline_e2e_1gpu.txt 2> gpurun_out/r02m_timeline_e2e_1gpu.err
cut -c1-400 gpurun_out/r02m_timeline_e2e_1gpu.txt | head -80
rm -f gpurun_out/timeline_e2e_w1_r0.json
