set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02i_bench_2gpu.json 2> gpurun_out/r02i_bench_2gpu.err
grep "^{" gpurun_out/r02i_bench_2gpu.json | tail -c 3000; tail -5 gpurun_out/r02i_bench_2gpu.err
timeout 300 $TR tools/scale_probe.py c3 > gpurun_out/r02i_scale_probe_2gpu.txt 2>&1
tail -25 gpurun_out/r02i_scale_probe_2gpu.txt | cut -c1-300
