#!/bin/bash
# ncu evidence of one round (run on the GPU box through gpurun; outputs under gpurun_out/).
#   * launch list of the bench command itself (every kernel of bench.py --steps 2);
#   * --set full captures out of the second device-resident C3 CV step (tools/ncu_solve.py):
#     the Gram build, a mid-solve row-sparse apply, a mid-solve prox_main / prox_momentum pair,
#     and one cooperative few-column launch of a single fit (tools/coop_probe.py).
set -x
TAG=${1:-r01c}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/${TAG}_launches.log 2>&1
NCU_STEPS=2 ncu --set full --clock-control none --import-source on -k regex:gemm_f64 -s 80 -c 1 -f \
    -o $OUT/${TAG}_syrk python tools/ncu_solve.py c3 > $OUT/${TAG}_syrk.log 2>&1
NCU_STEPS=2 ncu --set full --clock-control none --import-source on -k regex:gemm_f64 -s 118 -c 1 -f \
    -o $OUT/${TAG}_apply python tools/ncu_solve.py c3 > $OUT/${TAG}_apply.log 2>&1
NCU_STEPS=2 ncu --set full --clock-control none --import-source on -k regex:prox_main -s 85 -c 1 -f \
    -o $OUT/${TAG}_prox python tools/ncu_solve.py c3 > $OUT/${TAG}_prox.log 2>&1
NCU_STEPS=2 ncu --set full --clock-control none --import-source on -k regex:prox_momentum -s 85 -c 1 -f \
    -o $OUT/${TAG}_mom python tools/ncu_solve.py c3 > $OUT/${TAG}_mom.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fista_coop -s 4 -c 1 -f \
    -o $OUT/${TAG}_coop python tools/coop_probe.py c3 > $OUT/${TAG}_coop.log 2>&1
ls -la $OUT/${TAG}_*
