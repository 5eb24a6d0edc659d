"""Measurement aid: time the attention kernel alone on the d16 scoring shape (125 seqs x 16 heads x 680 tokens)."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from var_b200 import lib as L  # noqa: E402

PN = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)
ends = list(np.cumsum([p * p for p in PN]))
n_seq, H, Lq = 125, 16, 680
q = torch.nn.functional.normalize(torch.randn(n_seq, H, Lq, 64, device="cuda"), dim=-1).mul(6).bfloat16()
k = torch.nn.functional.normalize(torch.randn(n_seq, H, Lq, 64, device="cuda"), dim=-1).bfloat16()
v = torch.randn(n_seq, H, Lq, 64, device="cuda").bfloat16()
out = torch.empty(n_seq, Lq, H * 64, device="cuda", dtype=torch.bfloat16)
BOUND = float(os.environ.get("ATTN_BOUND", "6"))  # 0 = general kernel
QLOG2 = int(os.environ.get("ATTN_QLOG2", "1" if BOUND > 0 else "0"))  # q pre-multiplied by log2(e)
if QLOG2:
    q = (q.float() * 1.4426950408889634).bfloat16()
arr = (C.c_int * 10)(*[int(e) for e in ends])
lib = L.load()


def run():
    L.check(lib.var_b200_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), n_seq, H, Lq, Lq, 0, 10, arr,
                                   BOUND, QLOG2, L.current_stream()))


for _ in range(5):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
print(os.environ.get("VAR_B200_LIB", "default"), f"bound={BOUND} q_log2={QLOG2} attention {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call")
