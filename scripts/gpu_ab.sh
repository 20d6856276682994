#!/bin/bash
# A/B of an environment switch on one box: alternating bench runs (40 steps each), ms/step and SM clock per run.
# usage: scripts/gpu_ab.sh TAG VAR=VALUE [rounds]
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-ab}; SW=${2:-GRAFP_NO_MR_FUSED=1}; R=${3:-3}
for i in $(seq 1 $R); do
  timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-train --no-bf16 --no-db > $OUT/${TAG}_on_$i.json 2>/dev/null
  env $SW timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-train --no-bf16 --no-db > $OUT/${TAG}_off_$i.json 2>/dev/null
  python - <<PY
import json
a=json.load(open("$OUT/${TAG}_on_$i.json")); b=json.load(open("$OUT/${TAG}_off_$i.json"))
print("round $i: default %.3f ms (%d MHz)   $SW %.3f ms (%d MHz)" % (a["ms_per_step"], a["clocks"]["sm_mhz"], b["ms_per_step"], b["clocks"]["sm_mhz"]))
PY
done
