#!/bin/bash
# Same-box A/B of an environment switch:  tools/ab_env.sh "<bench args>" VAR=VAL_A VAR=VAL_B [VAR=VAL_A ...]
args=$1; shift
i=0
for kv in "$@"; do
  i=$((i+1))
  env $kv timeout 600 python bench.py $args --no-cpu --no-secondary > gpurun_out/ab_env_$i.json 2>/dev/null
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/ab_env_$i.json") if l.startswith("{")][-1])
    print("$args | $kv: value %.2f e2e %.2f ms/step %.2f clk %s host_gap %s"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["per_rank"][0]["host_gap_ms"]), flush=True)
except Exception as e:
    print("$args | $kv: FAILED", e)
PY
done
