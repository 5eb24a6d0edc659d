"""Measurement aid for ncu: a few launches of one GEMM flavour at a d16 shape (M=85000).
usage: gemm_prof_one.py gelu|bias|proj|fc2 [depth=16]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from var_b200 import lib as L  # noqa: E402

lib = L.load()
kind = sys.argv[1] if len(sys.argv) > 1 else "gelu"
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 16
Cd = 64 * depth
n_seq, l = (125, 680) if depth == 16 else (512, 256)
M = n_seq * l
N, K, epi = {"gelu": (4 * Cd, Cd, L.EPI_GELU_BF16), "bias": (4 * Cd, Cd, L.EPI_BIAS_BF16),
             "proj": (Cd, Cd, L.EPI_GATE_RESID), "fc2": (Cd, 4 * Cd, L.EPI_GATE_RESID)}[kind]
A = (torch.randn(M, K, device="cuda") * 0.05).bfloat16()
W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
bias = torch.zeros(N, device="cuda")
a = L.GemmArgs()
a.A, a.W, a.M, a.N, a.K, a.epilogue = A.data_ptr(), W.data_ptr(), M, N, K, epi
a.bias = bias.data_ptr()
if epi == L.EPI_GATE_RESID:
    x = torch.zeros(M, N, device="cuda")
    gate = torch.ones(n_seq, N, device="cuda")
    a.out, a.resid, a.gate, a.gate_ld, a.rows_per_seq = x.data_ptr(), x.data_ptr(), gate.data_ptr(), N, l
else:
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    a.out = out.data_ptr()
for _ in range(8):
    L.check(lib.var_b200_gemm_bf16(C.byref(a), L.current_stream()))
torch.cuda.synchronize()
print("ok")
