set -x
mkdir -p gpurun_out
T=r02z
timeout 1000 python -m pytest tests -q -m gpu --timeout=150 > gpurun_out/${T}_gpu_tests.log 2>&1
tail -4 gpurun_out/${T}_gpu_tests.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
for w in c1 c2 c4; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
done
timeout 400 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_bench_c5_1gpu.json 2> gpurun_out/${T}_bench_c5_1gpu.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l)
            if "ms_per_step" not in d: print(f, d); continue
            print(f, round(d["ms_per_step"],2), "ms e2e", (d.get("e2e") or {}).get("ms_per_step"), "unconv", (d.get("config") or {}).get("unconverged"), "frac", (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-tall > gpurun_out/${T}_launches.log 2>&1
wc -l gpurun_out/${T}_launches.csv
bash tools/profile_round2.sh ${T} > gpurun_out/${T}_profile.log 2>&1
for k in syrk apply_dense apply_mid prox; do
  ncu -i gpurun_out/${T}_$k.ncu-rep --page raw --csv > gpurun_out/${T}_${k}_raw.csv 2>/dev/null
done
rm -f gpurun_out/${T}_syrk.ncu-rep gpurun_out/${T}_apply_dense.ncu-rep gpurun_out/${T}_prox.ncu-rep
ls -la gpurun_out | head -40; du -sh gpurun_out
