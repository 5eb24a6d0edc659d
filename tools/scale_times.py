"""Measurement aid: time of the transformer blocks (var_b200_blocks, KV-cached) per pyramid scale for a CFG batch,
with the achieved fraction of the tensor peak at the clocks the run saw.  usage: scale_times.py [depth=30] [B=32] [reps=5]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from var_b200 import build_vae_var  # noqa: E402
from var_b200.init_utils import dense_init_  # noqa: E402

depth, B, reps = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (30, 32, 5)
dev = "cuda"
vae, var = build_vae_var(dev, depth=depth)
dense_init_(var, seed=2)
var.eval()
pm = var._model()
n_seq = 2 * B
C = 64 * depth
labels = torch.randint(0, 1000, (n_seq,), device=dev, dtype=torch.int32)
ada = pm.ada_params(labels)
kv = pm.kv_cache(n_seq)
tot = 0.0
cur = 0
print(f"d{depth} B={B} (n_seq={n_seq}): per-scale time of the {depth} blocks, GEMM FLOPs only in the TFLOP/s column")
for si, pn in enumerate(var.patch_nums):
    l = pn * pn
    x = torch.randn(n_seq * l, C, device=dev) * 0.5
    for _ in range(2):
        pm.blocks_cached(x, ada, n_seq, l, cur, kv, labels)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        pm.blocks_cached(x, ada, n_seq, l, cur, kv, labels)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 24.0 * C * C * depth * n_seq * l
    print(f"scale {si} l={l:4d} M={n_seq * l:6d}: {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s  ({1e3 * ms / (5 * depth):6.1f} us per launch)")
    tot += ms
    cur += l
print(f"total {tot:.2f} ms")
