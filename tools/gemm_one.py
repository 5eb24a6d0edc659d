"""Measurement aid: run the fc1 GEMM (GELU epilogue) alone at a given shape (default: d30, B=256, last scale)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from var_b200 import lib as L  # noqa: E402

M, N, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (131072, 7680, 1920)
lib = L.load()
A = (torch.randn(M, K, device="cuda") * 0.05).bfloat16()
W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
bias = torch.zeros(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
a = L.GemmArgs()
a.A, a.W, a.M, a.N, a.K, a.epilogue = A.data_ptr(), W.data_ptr(), M, N, K, L.EPI_GELU_BF16
a.bias, a.out = bias.data_ptr(), out.data_ptr()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    L.check(lib.var_b200_gemm_bf16(C.byref(a), L.current_stream()))
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    L.check(lib.var_b200_gemm_bf16(C.byref(a), L.current_stream()))
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10 * 1e-3
print(f"M={M} N={N} K={K}: {t * 1e6:.1f} us, {2.0 * M * N * K / t / 1e12:.1f} TFLOP/s, algorithmic bytes {(M * K + N * K + M * N) * 2 / 1e9:.3f} GB")
