set -x
mkdir -p gpurun_out
timeout 300 python tools/timeline.py c4 > gpurun_out/r02m_timeline_c4.txt 2> gpurun_out/r02m_timeline_c4.err
cut -c1-330 gpurun_out/r02m_timeline_c4.txt | head -75; tail -3 gpurun_out/r02m_timeline_c4.err
rm -f gpurun_out/timeline_w1_r0.json
