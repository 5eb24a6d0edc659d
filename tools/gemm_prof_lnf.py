"""Measurement aid for ncu: the three deferred-LayerNorm GEMM flavours of a d30 block at the last AR scale of a B=256 CFG
batch (M = 131072), two launches each in the order fc1+LN (consumer), proj+LN outputs, fc2+LN outputs (producers).
usage: gemm_prof_lnf.py [depth=30] [n_seq=512] [l=256]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from var_b200 import lib as L  # noqa: E402

depth, n_seq, l = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (30, 512, 256)
Cd = 64 * depth
M = n_seq * l
lib = L.load()
parts = lib.var_b200_gemm_ln_parts(M, Cd)
ln_part = torch.rand(M, parts, 2, device="cuda")
ln_part[..., 1] += 2.0
labels = torch.randint(0, 1001, (n_seq,), device="cuda", dtype=torch.int32)


def base(N, K, epi):
    A = (torch.randn(M, K, device="cuda") * 0.05).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.zeros(N, device="cuda")
    a = L.GemmArgs()
    a.A, a.W, a.M, a.N, a.K, a.epilogue = A.data_ptr(), W.data_ptr(), M, N, K, epi
    a.bias = bias.data_ptr()
    return a, (A, W, bias)


def run(a, n=2):
    for _ in range(n):
        L.check(lib.var_b200_gemm_bf16(C.byref(a), L.current_stream()))
    torch.cuda.synchronize()


# fc1 + deferred LN (consumer)
a, keep = base(4 * Cd, Cd, L.EPI_GELU_BF16)
out = torch.empty(M, 4 * Cd, device="cuda", dtype=torch.bfloat16)
u = torch.randn(1001, 4 * Cd, device="cuda") * 0.1
v = torch.randn(1001, 4 * Cd, device="cuda") * 0.1
a.out = out.data_ptr()
a.ln_part_in, a.ln_parts, a.ln_C, a.ln_eps = ln_part.data_ptr(), parts, Cd, 1e-6
a.ln_u, a.ln_v, a.ln_labels, a.rows_per_seq = u.data_ptr(), v.data_ptr(), labels.data_ptr(), l
run(a)
del out, u, v, keep
# proj / fc2 + deferred LN outputs (producers)
for K in (Cd, 4 * Cd):
    a, keep = base(Cd, K, L.EPI_GATE_RESID)
    x = torch.zeros(M, Cd, device="cuda")
    gate = torch.ones(n_seq, Cd, device="cuda")
    a_out = torch.empty(M, Cd, device="cuda", dtype=torch.bfloat16)
    a.out, a.resid, a.gate, a.gate_ld, a.rows_per_seq = x.data_ptr(), x.data_ptr(), gate.data_ptr(), Cd, l
    a.ln_a_out, a.ln_scale, a.ln_part_out = a_out.data_ptr(), gate.data_ptr(), ln_part.data_ptr()
    run(a)
    del x, a_out, keep
print("ok")
