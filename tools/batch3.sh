set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_c.log 2>&1; tail -3 gpurun_out/r02_pytest_c.log
python bench.py --no-cpu --no-secondary --steps 3 --warmup 3 > gpurun_out/r02_ab_attn_fast.json 2>/dev/null
VAR_B200_ATTN_FAST=0 python bench.py --no-cpu --no-secondary --steps 3 --warmup 3 > gpurun_out/r02_ab_attn_general.json 2>/dev/null
python - <<'PY'
import json
for t in ("fast","general"):
    d=json.loads([l for l in open(f"gpurun_out/r02_ab_attn_{t}.json") if l.startswith("{")][-1])
    print("attn", t, round(d["value"],2), round(d["ms_per_step"],1), d["clocks"]["sm_mhz"], {k:round(v,1) for k,v in d["roofline"]["kernel_ms"].items() if v>5})
PY
python tools/kernels_one.py embed 10
for wl in score_d16 sample_d16 sample_d36_512; do python bench.py --workload $wl --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_$wl.json 2>gpurun_out/r02_bench_$wl.err; tail -c 300 gpurun_out/r02_bench_$wl.err; done
O=gpurun_out/ev2; mkdir -p $O
NCU="ncu --set full --clock-control none"
$NCU -k regex:quant_search -s 30 -c 10 -o $O/ncu_quant_search10 -f python tools/kernels_one.py quant 1 > $O/ncu_quant_search10.log 2>&1
ncu -i $O/ncu_quant_search10.ncu-rep --page details --csv > $O/ncu_quant_search10.details.csv; ncu -i $O/ncu_quant_search10.ncu-rep --page raw --csv > $O/ncu_quant_search10.raw.csv; rm -f $O/ncu_quant_search10.ncu-rep
$NCU --kernel-name-base demangled -k 'regex:gemm_bf16_kernel<[0-9]+, *4,' -s 1 -c 1 -o $O/ncu_qkv -f python tools/kernels_one.py qkv 1 > $O/ncu_qkv.log 2>&1
ncu -i $O/ncu_qkv.ncu-rep --page details --csv > $O/ncu_qkv.details.csv; ncu -i $O/ncu_qkv.ncu-rep --page raw --csv > $O/ncu_qkv.raw.csv; rm -f $O/ncu_qkv.ncu-rep
du -sh gpurun_out
