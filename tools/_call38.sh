set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 500 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02z_bench_8gpu.json 2> gpurun_out/r02z_bench_8gpu.err
tail -c 1200 gpurun_out/r02z_bench_8gpu.json; tail -n 3 gpurun_out/r02z_bench_8gpu.err
timeout 200 $TR tools/timeline.py c3 > gpurun_out/r02z_timeline_8gpu.txt 2> gpurun_out/r02z_timeline_8gpu.err
head -30 gpurun_out/r02z_timeline_8gpu.txt | cut -c1-200
rm -f gpurun_out/timeline_w8_r0.json
