#!/bin/bash
# Same-box A/B of whole source trees (git archives built under _ab/<ref>/, see DESIGN.md "Measurement"):
#   tools/ab_trees.sh <workload> <tag:dir[:ENV=VAL]>...   -> gpurun_out/ab_<workload>_<tag>.json + one summary line each
WL=$1; shift
root=$(pwd)
for spec in "$@"; do
  IFS=: read tag dir envs <<< "$spec"
  (cd "$dir" && env $envs timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary --workload $WL > "$root/gpurun_out/ab_${WL}_$tag.json" 2>/dev/null)
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$root/gpurun_out/ab_${WL}_$tag.json") if l.startswith("{")][-1])
    k=d["roofline"]["kernel_ms"]
    print("$WL $tag: value %.3f e2e %.3f ms/step %.1f clk %s | "%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"]) + " ".join(f"{a}={b:.1f}" for a,b in k.items() if b>1), flush=True)
except Exception as e:
    print("$WL $tag: FAILED", e)
PY
done
