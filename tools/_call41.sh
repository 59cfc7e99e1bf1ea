set -x
mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 500 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02z_bench_${N}gpu.json 2> gpurun_out/r02z_bench_${N}gpu.err
tail -c 600 gpurun_out/r02z_bench_${N}gpu.json; tail -n 3 gpurun_out/r02z_bench_${N}gpu.err
