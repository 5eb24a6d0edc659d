"""Measurement aid: the attention kernel alone at every KV-cached scale of a d30 CFG batch (n_seq x 30 heads).
usage: attn_scale_times.py [n_seq=512] [H=30]"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from var_b200 import lib as L  # noqa: E402

n_seq, H = (int(a) for a in sys.argv[1:3]) if len(sys.argv) > 2 else (512, 30)
PN = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)
ends = list(np.cumsum([p * p for p in PN]))
Lmax = ends[-1]
k = torch.nn.functional.normalize(torch.randn(n_seq, H, Lmax, 64, device="cuda"), dim=-1).bfloat16()
v = torch.randn(n_seq, H, Lmax, 64, device="cuda").bfloat16()
arr = (C.c_int * 10)(*[int(e) for e in ends])
lib = L.load()
tot = 0.0
pos0 = 0
for si, pn in enumerate(PN):
    l = pn * pn
    q = (torch.nn.functional.normalize(torch.randn(n_seq, H, l, 64, device="cuda"), dim=-1) * 6 * 1.4426950408889634).bfloat16()
    out = torch.empty(n_seq, l, H * 64, device="cuda", dtype=torch.bfloat16)

    def run():
        L.check(lib.var_b200_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), n_seq, H, l, Lmax, pos0, 10, arr,
                                       6.0, 1, L.current_stream()))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    n_qt = (l + 127) // 128
    n_kt = (ends[si] + 63) // 64
    tiles = n_seq * H * n_qt * n_kt
    pairs = n_seq * H * l * ends[si]
    print(f"scale {si} l={l:3d} keys={ends[si]:3d}: {us:8.1f} us  items={n_seq * H * n_qt} tiles/item={n_kt:2d}  "
          f"{us * 1e3 / tiles * 148:7.1f} ns*SM per tile  {pairs * 256 / us / 1e6:7.1f} TFLOP/s on visible pairs")
    tot += us
    pos0 += l
print(f"total {tot / 1e3:.2f} ms per layer-step; x30 layers = {tot * 30 / 1e3:.1f} ms")
