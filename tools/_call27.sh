set -x
mkdir -p gpurun_out
NCU_STEPS=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:chol_diag -s 10 -c 2 -f \
    -o gpurun_out/r02n_chol_diag python tools/ncu_solve.py c4 > gpurun_out/r02n_chol_diag.log 2>&1
tail -5 gpurun_out/r02n_chol_diag.log
NCU_STEPS=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:chol_panel -s 10 -c 1 -f \
    -o gpurun_out/r02n_chol_panel python tools/ncu_solve.py c4 > gpurun_out/r02n_chol_panel.log 2>&1
ls -la gpurun_out/*.ncu-rep
