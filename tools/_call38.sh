set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02z_bench_8gpu.json 2> gpurun_out/r02z_bench_8gpu.err
tail -c 1500 gpurun_out/r02z_bench_8gpu.json; tail -n 3 gpurun_out/r02z_bench_8gpu.err
timeout 300 $TR tools/timeline.py c3 > gpurun_out/r02z_timeline_8gpu.txt 2> gpurun_out/r02z_timeline_8gpu.err
head -40 gpurun_out/r02z_timeline_8gpu.txt | cut -c1-200
timeout 300 $TR tools/timeline.py c3 e2e > gpurun_out/r02z_timeline_e2e_8gpu.txt 2> gpurun_out/r02z_timeline_e2e_8gpu.err
rm -f gpurun_out/timeline_e2e_w8_r0.json gpurun_out/timeline_w8_r0.json
timeout 400 $TR bench.py --gpus 8 --workload c5 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02z_bench_c5_8gpu.json 2> gpurun_out/r02z_bench_c5_8gpu.err
tail -c 600 gpurun_out/r02z_bench_c5_8gpu.json
