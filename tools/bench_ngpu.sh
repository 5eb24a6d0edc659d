#!/bin/bash
# The driver's torchrun command on N GPUs of one box:  gpurun --gpus N -- bash tools/bench_ngpu.sh N
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo rc=$?
tail -c 400 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_${N}gpu.json") if l.startswith("{")][-1])
print("weak", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],1))
print("per_rank ms", [r["ms_per_step"] for r in d["per_rank"]], "clk", [r["sm_mhz"] for r in d["per_rank"]])
s=d.get("strong_config5"); print("strong", round(s["value"],1), "per gpu batch", s["per_gpu_batch"], "ms", round(s["ms_per_step"],1), "e2e", round(s["e2e"],1)) if s else None
s=d.get("secondary"); print("scoring", round(s["value"],2), "e2e", round(s["e2e"],2), "ms", round(s["ms_per_step"],1), [r["ms_per_step"] for r in s["per_rank"]]) if s else None
PY
