set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-tall --no-weak > gpurun_out/r02zz_bench_8gpu.json 2> gpurun_out/r02zz_bench_8gpu.err
tail -c 900 gpurun_out/r02zz_bench_8gpu.json; tail -n 3 gpurun_out/r02zz_bench_8gpu.err
