set -x
mkdir -p gpurun_out
timeout 300 python tools/timeline.py c3 > gpurun_out/r02s_timeline_1gpu.txt 2> gpurun_out/r02s_timeline_1gpu.err
cut -c1-200 gpurun_out/r02s_timeline_1gpu.txt | head -24
rm -f gpurun_out/timeline_w1_r0.json
NCU_STEPS=2 timeout 500 ncu --set full --clock-control none --import-source on -k regex:prox_fused -s 75 -c 1 -f \
    -o gpurun_out/r02s_prox_fused python tools/ncu_solve.py c3 > gpurun_out/r02s_prox_fused.log 2>&1
ls -la gpurun_out/r02s_prox_fused.ncu-rep
