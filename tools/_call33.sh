set -x
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -q -m gpu --timeout=150 -x > gpurun_out/r02r_gpu_tests.log 2>&1
tail -5 gpurun_out/r02r_gpu_tests.log | cut -c1-250
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-tall > gpurun_out/r02r_bench_2gpu.json 2> gpurun_out/r02r_bench_2gpu.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-tall > gpurun_out/r02r_bench_1gpu.json 2> gpurun_out/r02r_bench_1gpu.err
python - <<PY
import json
for f in ("gpurun_out/r02r_bench_2gpu.json","gpurun_out/r02r_bench_1gpu.json"):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, round(d["ms_per_step"],2), "ms e2e", round(d["e2e"]["ms_per_step"],2), "unconv", d["config"]["unconverged"], d["roofline"]["step_ms_by_kernel_family"], d.get("sharded_parity"), d.get("weak_scaling",{}).get("ms_per_step"))
PY
tail -n 3 gpurun_out/r02r_bench_2gpu.err
