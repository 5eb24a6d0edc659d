#!/bin/bash
# The driver's torchrun command on 2 GPUs (gpurun --gpus 2 -- bash tools/bench_2gpu.sh): the default bench line (weak d30
# sampling + strong_config5 + class-sharded scoring with the all-gather) and the reference arm (rank 0 only).
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_2gpu_ref.json 2> gpurun_out/r02_bench_2gpu_ref.err; echo rc=$?; cut -c1-400 gpurun_out/r02_bench_2gpu_ref.json
