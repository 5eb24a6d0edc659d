"""Measurement aid: is the short-K GEMM (d16, K=1024) bound by its epilogue or by per-tile overheads?
Sweeps K and the epilogue flavour at M=85000, N=4096 (d16 fc1 shape) with enough repetitions to reach steady clocks."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from var_b200 import lib as L  # noqa: E402

lib = L.load()
M, N = 85000, 4096


def run(K, epi, reps=60):
    A = (torch.randn(M, K, device="cuda") * 0.05).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.zeros(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if epi == L.EPI_BIAS_F32 else torch.bfloat16)
    a = L.GemmArgs()
    a.A, a.W, a.M, a.N, a.K, a.epilogue = A.data_ptr(), W.data_ptr(), M, N, K, epi
    a.bias, a.out = bias.data_ptr(), out.data_ptr()
    for _ in range(20):
        L.check(lib.var_b200_gemm_bf16(C.byref(a), L.current_stream()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        L.check(lib.var_b200_gemm_bf16(C.byref(a), L.current_stream()))
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps * 1e-3
    return t, 2.0 * M * N * K / t / 1e12


for K in (1024, 2048, 4096):
    for name, epi in (("bias->bf16", L.EPI_BIAS_BF16), ("gelu->bf16", L.EPI_GELU_BF16), ("bias->f32", L.EPI_BIAS_F32)):
        t, tf = run(K, epi)
        print(f"K={K:5d} {name:11s}: {t * 1e6:8.1f} us  {tf:7.1f} TFLOP/s")
