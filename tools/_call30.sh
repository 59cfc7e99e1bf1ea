set -x
mkdir -p gpurun_out
timeout 300 python tools/host_profile_e2e.py c3 > gpurun_out/r02o_host_profile_e2e.txt 2> gpurun_out/r02o_host_profile_e2e.err
cut -c1-180 gpurun_out/r02o_host_profile_e2e.txt | head -120; tail -3 gpurun_out/r02o_host_profile_e2e.err
