run() { # tag env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary --workload $WL > gpurun_out/ab_${WL}_$tag.json 2>/dev/null
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/ab_${WL}_$tag.json") if l.startswith("{")][-1])
k=d["roofline"]["kernel_ms"]
print("$WL $tag: value %.3f ms/step %.1f clk %s | "%(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"]) + " ".join(f"{a}={b:.1f}" for a,b in k.items() if b>1))
PY
}
for WL in sample_d30 score_d16; do
  run cur_lnf A=1
  run cur_nolnf VAR_B200_LNF=0
  run v1_lnf VAR_B200_LIB=var_b200/_variants/lib_lnf_v1.so
  run v1_nolnf VAR_B200_LIB=var_b200/_variants/lib_lnf_v1.so VAR_B200_LNF=0
  run cur_lnf2 A=1
done
