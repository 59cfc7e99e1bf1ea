set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 --no-tall --no-weak > gpurun_out/r02d_bench_8gpu.json 2> gpurun_out/r02d_bench_8gpu.err
tail -c 2500 gpurun_out/r02d_bench_8gpu.json; tail -3 gpurun_out/r02d_bench_8gpu.err
timeout 300 $TR tools/timeline.py c3 > gpurun_out/r02d_timeline_8gpu.txt 2> gpurun_out/r02d_timeline_8gpu.err
head -75 gpurun_out/r02d_timeline_8gpu.txt | cut -c1-160; tail -3 gpurun_out/r02d_timeline_8gpu.err
timeout 300 $TR tools/timeline.py c3 e2e > gpurun_out/r02d_timeline_e2e_8gpu.txt 2> gpurun_out/r02d_timeline_e2e_8gpu.err
head -45 gpurun_out/r02d_timeline_e2e_8gpu.txt | cut -c1-160
rm -f gpurun_out/timeline_e2e_w8_r0.json
