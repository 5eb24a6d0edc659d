"""Measurement aid: run each GEMM flavour of a transformer block alone at the d30 (or a given) shape.

usage: python tools/gemm_one.py [depth=30] [n_seq=512] [l=256]      (last AR scale of a B=256 CFG batch by default)
"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from var_b200 import lib as L  # noqa: E402

depth, n_seq, l = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (30, 512, 256)
FORCE_BN = int(sys.argv[4]) if len(sys.argv) > 4 else 0  # 0 = the launcher's choice
Cd, H, Lmax = 64 * depth, depth, 680
M = n_seq * l
lib = L.load()


def timed(a, flops, label):
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
    print(f"{label:28s} M={a.M} N={a.N} K={a.K}: {t * 1e6:8.1f} us  {flops / t / 1e12:7.1f} TFLOP/s")


def base(N, K, epi):
    A = (torch.randn(M, K, device="cuda") * 0.05).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.zeros(N, device="cuda")
    a = L.GemmArgs()
    a.A, a.W, a.M, a.N, a.K, a.epilogue = A.data_ptr(), W.data_ptr(), M, N, K, epi
    a.bias = bias.data_ptr()
    a.force_bn = FORCE_BN
    return a, (A, W, bias)


# deferred-LayerNorm inputs (gemm_sm100.cuh LNF): partial row sums as a producer launch writes them, class tables
N_CLS = 1001
parts = lib.var_b200_gemm_ln_parts(M, Cd)
ln_part = torch.rand(M, parts, 2, device="cuda")
ln_part[..., 1] += 2.0
labels = torch.randint(0, N_CLS, (n_seq,), device="cuda", dtype=torch.int32)


def consumer(a, N):
    u = torch.randn(N_CLS, N, device="cuda") * 0.1
    v = torch.randn(N_CLS, N, device="cuda") * 0.1
    a.ln_part_in, a.ln_parts, a.ln_C, a.ln_eps = ln_part.data_ptr(), parts, Cd, 1e-6
    a.ln_u, a.ln_v, a.ln_labels, a.rows_per_seq = u.data_ptr(), v.data_ptr(), labels.data_ptr(), l
    return u, v


# fc1 (GELU)
a, keep = base(4 * Cd, Cd, L.EPI_GELU_BF16)
out = torch.empty(M, 4 * Cd, device="cuda", dtype=torch.bfloat16)
a.out = out.data_ptr()
timed(a, 2.0 * M * 4 * Cd * Cd, "fc1  (GELU -> bf16)")
keep2 = consumer(a, 4 * Cd)
timed(a, 2.0 * M * 4 * Cd * Cd, "fc1  + deferred LN")
del out, keep2
# qkv
a, keep = base(3 * Cd, Cd, L.EPI_QKV)
q = torch.empty(n_seq, H, l, 64, device="cuda", dtype=torch.bfloat16)
kc = torch.empty(n_seq, H, Lmax, 64, device="cuda", dtype=torch.bfloat16)
vc = torch.empty_like(kc)
scale = torch.full((H,), 4.0, device="cuda")
a.q_out, a.k_cache, a.v_cache, a.q_scale = q.data_ptr(), kc.data_ptr(), vc.data_ptr(), scale.data_ptr()
a.C, a.H, a.pos0, a.Lmax, a.rows_per_seq = Cd, H, Lmax - l, Lmax, l
timed(a, 2.0 * M * 3 * Cd * Cd, "qkv  (norm/scale/scatter)")
keep2 = consumer(a, 3 * Cd)
timed(a, 2.0 * M * 3 * Cd * Cd, "qkv  + deferred LN")
del q, kc, vc, keep2
# proj / fc2 (gate + residual, fp32 in place)
for K, name in ((Cd, "proj (resid + g*acc, fp32)"), (4 * Cd, "fc2  (resid + g*acc, fp32)")):
    a, keep = base(Cd, K, L.EPI_GATE_RESID)
    x = torch.zeros(M, Cd, device="cuda")
    gate = torch.ones(n_seq, Cd, device="cuda")
    a.out, a.resid, a.gate, a.gate_ld, a.rows_per_seq = x.data_ptr(), x.data_ptr(), gate.data_ptr(), Cd, l
    timed(a, 2.0 * M * Cd * K, name)
    if not FORCE_BN:
        a_out = torch.empty(M, Cd, device="cuda", dtype=torch.bfloat16)
        a.ln_a_out, a.ln_scale, a.ln_part_out = a_out.data_ptr(), gate.data_ptr(), ln_part.data_ptr()
        timed(a, 2.0 * M * Cd * K, name[:4] + " + deferred LN outputs")
        del a_out
    del x
