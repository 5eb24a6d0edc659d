"""Measurement aid for ncu: a few launches of one GEMM flavour at the d16 fc1 shape (M=85000, N=4096, K=1024)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from var_b200 import lib as L  # noqa: E402

lib = L.load()
M, N, K = 85000, 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
epi = {"gelu": L.EPI_GELU_BF16, "bias": L.EPI_BIAS_BF16}[sys.argv[1] if len(sys.argv) > 1 else "gelu"]
A = (torch.randn(M, K, device="cuda") * 0.05).bfloat16()
W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
bias = torch.zeros(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
a = L.GemmArgs()
a.A, a.W, a.M, a.N, a.K, a.epilogue = A.data_ptr(), W.data_ptr(), M, N, K, epi
a.bias, a.out = bias.data_ptr(), out.data_ptr()
for _ in range(8):
    L.check(lib.var_b200_gemm_bf16(C.byref(a), L.current_stream()))
torch.cuda.synchronize()
print("ok")
