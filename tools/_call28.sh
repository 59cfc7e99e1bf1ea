set -x
mkdir -p gpurun_out
NB="--kernel-name-base demangled"
NCU_STEPS=1 timeout 500 ncu --set full --clock-control none --import-source on $NB -k 'regex:gemm_f64_tma_kernel<\(int\)2, \(int\)4, \(int\)8, \(int\)4, \(int\)32' -s 3 -c 1 -f \
    -o gpurun_out/r02n_trail_early python tools/ncu_solve.py c4 > gpurun_out/r02n_trail_early.log 2>&1
NCU_STEPS=1 timeout 500 ncu --set full --clock-control none --import-source on $NB -k 'regex:gemm_f64_tma_kernel<\(int\)2, \(int\)4, \(int\)8, \(int\)4, \(int\)32' -s 18 -c 1 -f \
    -o gpurun_out/r02n_trail_mid python tools/ncu_solve.py c4 > gpurun_out/r02n_trail_mid.log 2>&1
ls -la gpurun_out/r02n_trail*.ncu-rep
# launch list of the whole step (durations per launch)
NCU_STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none $NB -k 'regex:gemm_f64_tma_kernel<\(int\)2, \(int\)4, \(int\)8, \(int\)4, \(int\)32|chol_' -c 3000 --csv --log-file gpurun_out/r02n_newton_launches.csv python tools/ncu_solve.py c4 > gpurun_out/r02n_newton_launches.log 2>&1
wc -l gpurun_out/r02n_newton_launches.csv
