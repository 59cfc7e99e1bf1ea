set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02q_bench_8gpu.json 2> gpurun_out/r02q_bench_8gpu.err
tail -c 3500 gpurun_out/r02q_bench_8gpu.json; tail -3 gpurun_out/r02q_bench_8gpu.err
timeout 300 $TR tools/timeline.py c3 > gpurun_out/r02q_timeline_8gpu.txt 2> gpurun_out/r02q_timeline_8gpu.err
head -64 gpurun_out/r02q_timeline_8gpu.txt | cut -c1-330; tail -3 gpurun_out/r02q_timeline_8gpu.err
timeout 300 $TR tools/timeline.py c3 e2e > gpurun_out/r02q_timeline_e2e_8gpu.txt 2> gpurun_out/r02q_timeline_e2e_8gpu.err
head -30 gpurun_out/r02q_timeline_e2e_8gpu.txt | cut -c1-200
rm -f gpurun_out/timeline_e2e_w8_r0.json gpurun_out/timeline_w8_r0.json
