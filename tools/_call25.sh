set -x
mkdir -p gpurun_out
for cfg in "400 200" "300 200" "300 100" "250 100" "200 100" "200 200"; do
  set -- $cfg
  SLM_NEWTON_FIRST=$1 SLM_NEWTON_LATER=$2 timeout 200 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu > gpurun_out/tune_c4_$1_$2.json 2> gpurun_out/tune_c4_$1_$2.err
  python - <<PY
import json
for l in open("gpurun_out/tune_c4_$1_$2.json"):
    if l.startswith("{"):
        d=json.loads(l); print("first $1 later $2:", round(d["ms_per_step"],1), "ms  iters", d["config"]["iterations_per_step"], "unconv", d["config"]["unconverged"], d["roofline"].get("newton_phase",{}).get("ms_per_step"), d["roofline"].get("newton_phase",{}).get("factorizations_per_step"))
PY
done
