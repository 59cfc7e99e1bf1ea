set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/timeline.py c3 > gpurun_out/r02m_timeline_2gpu.txt 2> gpurun_out/r02m_timeline_2gpu.err
cut -c1-420 gpurun_out/r02m_timeline_2gpu.txt | head -75; tail -3 gpurun_out/r02m_timeline_2gpu.err
