#!/bin/bash
# ncu evidence of one round (run on the GPU box through gpurun; outputs under gpurun_out/).
#   launch list of one device-resident C3 CV step, --set full captures of the Gram build, a
#   mid-solve row-sparse apply and the fused small-design kernel.
set -x
TAG=${1:-r01b}
OUT=gpurun_out
NCU_STEPS=2 ncu --metrics gpu__time_duration.sum --clock-control none -s 360 -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/ncu_solve.py c3 > $OUT/${TAG}_launches.log 2>&1
NCU_STEPS=2 ncu --set full --clock-control none --import-source on -k regex:gemm_f64 -s 80 -c 1 -f \
    -o $OUT/${TAG}_syrk python tools/ncu_solve.py c3 > $OUT/${TAG}_syrk.log 2>&1
NCU_STEPS=2 ncu --set full --clock-control none --import-source on -k regex:gemm_f64 -s 118 -c 1 -f \
    -o $OUT/${TAG}_apply python tools/ncu_solve.py c3 > $OUT/${TAG}_apply.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fista_small -s 12 -c 1 -f \
    -o $OUT/${TAG}_small python bench.py --workload c1 --steps 1 --warmup 3 --no-cpu > $OUT/${TAG}_small.log 2>&1
ls -la $OUT/${TAG}_*
