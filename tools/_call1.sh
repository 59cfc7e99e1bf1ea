set -x
mkdir -p gpurun_out
nvidia-smi -L
python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_c5_1gpu.json 2> gpurun_out/r02_bench_c5_1gpu.err
tail -c 600 gpurun_out/r02_bench_c5_1gpu.json
python tools/dsyrk_probe.py c3 c5 > gpurun_out/r02_dsyrk.jsonl 2> gpurun_out/r02_dsyrk.err
cat gpurun_out/r02_dsyrk.jsonl; tail -3 gpurun_out/r02_dsyrk.err
SAN_TOOLS="memcheck" SAN_TIMEOUT=420 bash tools/sanitize.sh r02
SAN_TOOLS="racecheck" SAN_TIMEOUT=300 bash tools/sanitize.sh r02
