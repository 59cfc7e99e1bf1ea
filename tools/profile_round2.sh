#!/bin/bash
# ncu --set full captures of round 2's dominant kernels out of the second device-resident C3 CV step
# (tools/ncu_solve.py): the TMA-fed Gram build, a dense-phase and a mid-solve launch of the TMA gather4-fed
# row-sparse apply, a prox_fused launch.  Outputs under gpurun_out/ (read here with ncu -i ... --page raw --csv).
set -x
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
NB="--kernel-name-base demangled"
NCU_STEPS=2 ncu --set full --clock-control none --import-source on $NB -k 'regex:gemm_f64_tma_kernel<\(int\)2, \(int\)4, \(int\)8, \(int\)4, \(int\)32' -s 1 -c 1 -f \
    -o $OUT/${TAG}_syrk python tools/ncu_solve.py c3 > $OUT/${TAG}_syrk.log 2>&1
NCU_STEPS=2 ncu --set full --clock-control none --import-source on $NB -k 'regex:gemm_f64_tma_kernel<\(int\)8, \(int\)1, \(int\)2, \(int\)4, \(int\)16, \(int\)3' -s 48 -c 1 -f \
    -o $OUT/${TAG}_apply_dense python tools/ncu_solve.py c3 > $OUT/${TAG}_apply_dense.log 2>&1
NCU_STEPS=2 ncu --set full --clock-control none --import-source on $NB -k 'regex:gemm_f64_tma_kernel<\(int\)8, \(int\)1, \(int\)2, \(int\)4, \(int\)16, \(int\)3' -s 68 -c 1 -f \
    -o $OUT/${TAG}_apply_mid python tools/ncu_solve.py c3 > $OUT/${TAG}_apply_mid.log 2>&1
NCU_STEPS=2 ncu --set full --clock-control none --import-source on -k regex:prox_fused -s 80 -c 1 -f \
    -o $OUT/${TAG}_prox python tools/ncu_solve.py c3 > $OUT/${TAG}_prox.log 2>&1
ls -la $OUT/${TAG}_*.ncu-rep
