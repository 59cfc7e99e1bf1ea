set -x
mkdir -p gpurun_out
timeout 300 python tools/host_profile.py c3 > gpurun_out/r02m_host_profile_1gpu.txt 2> gpurun_out/r02m_host_profile_1gpu.err
cut -c1-200 gpurun_out/r02m_host_profile_1gpu.txt | head -90; tail -3 gpurun_out/r02m_host_profile_1gpu.err
