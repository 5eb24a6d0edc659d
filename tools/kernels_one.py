"""Measurement aid (CUDA events here, ncu around it): the non-GEMM kernels of the path at the headline shapes, one
family per invocation, each timed alone with its algorithmic bytes -> GB/s against MEASURED_PEAKS.json hbm_gbs.

usage: kernels_one.py ln|sample|embed|quant|qkv|next  [reps]
  ln      ln_modulate_kernel, d30 last AR scale of a B=256 CFG batch (M = 131 072 rows x 1920)
  sample  sample_kernel<false>, B=256 images x 256 positions, V=4096, top_k=900 (CFG mix + top-k + multinomial)
  embed   embed_kernel, d30 last scale (word_embed K=32 GEMV + level/position embeddings)
  quant   var_b200_quant_encode B=64 (quant_kernel + quant_search_kernel, 10 scales; config 2)
  qkv     gemm_bf16_kernel<.,QKV,2>: d30 last scale, the K/V-cache append lives in this epilogue
  next    quant_kernel single-scale step (bicubic up-sample + Phi conv + area down-sample), B=256, every scale
"""
import ctypes as C
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from var_b200 import build_vae_var, lib as L  # noqa: E402
from var_b200.init_utils import dense_init_  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "ln"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
lib = L.load()
dev = "cuda"
pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2


def timed(fn, nbytes, what):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    t = sorted(ts)[len(ts) // 2]
    print(json.dumps(dict(kernel=what, us=round(t * 1e6, 1), algorithmic_MB=round(nbytes / 1e6, 2),
                          GBps=round(nbytes / t / 1e9, 1), frac_of_measured_hbm=round(nbytes / t / 1e9 / pk, 3), l2="flushed")))


if kind == "ln":
    n_seq, l, Cd = 512, 256, 1920
    M = n_seq * l
    x = torch.randn(M, Cd, device=dev)
    ada = torch.randn(n_seq, 6 * Cd, device=dev)
    out = torch.empty(M, Cd, device=dev, dtype=torch.bfloat16)
    timed(lambda: L.check(lib.var_b200_ln_modulate(x.data_ptr(), ada.data_ptr(), ada[:, Cd:].data_ptr(), 6 * Cd, l, out.data_ptr(),
                                                  M, Cd, 1e-6, L.current_stream())), M * Cd * 6, f"ln_modulate M={M} C={Cd}")
elif kind == "sample":
    B, l, V = 256, 256, 4096
    logits = torch.randn(2 * B, l, V, device=dev)
    q = torch.empty(B * l, V, device=dev).exponential_(1.0)
    idx = torch.empty(B, l, dtype=torch.int64, device=dev)
    timed(lambda: L.check(lib.var_b200_cfg_topk_sample(logits.data_ptr(), B, l, V, 1, 1.5, q.data_ptr(), 900, 0.0, idx.data_ptr(),
                                                      None, L.current_stream())), B * l * V * 12, f"sample B={B} l={l} V={V} top_k=900")
elif kind in ("embed", "qkv"):
    _, var = build_vae_var(dev, depth=30)
    dense_init_(var, seed=2)
    var.eval()
    pm = var._model()
    n_seq, l, Cd = 512, 256, 1920
    labels = torch.randint(0, 1001, (n_seq,), device=dev, dtype=torch.int32)
    nxt = torch.randn(n_seq // 2, l, 32, device=dev)
    if kind == "embed":
        timed(lambda: pm.embed(nxt, n_seq // 2, labels, n_seq, l, 0, 424), n_seq * l * Cd * 4 + nxt.numel() * 4 + l * Cd * 4,
              f"embed n_seq={n_seq} l={l} C={Cd}")
    else:
        M = n_seq * l
        A = (torch.randn(M, Cd, device=dev) * 0.05).bfloat16()
        bw = pm.blocks_arr[0]
        qb = torch.empty(M * Cd, device=dev, dtype=torch.bfloat16)
        kv = torch.zeros(2 * n_seq * Cd * 680, device=dev, dtype=torch.bfloat16)
        a = L.GemmArgs()
        a.A, a.W, a.M, a.N, a.K, a.epilogue = A.data_ptr(), bw.w_qkv, M, 3 * Cd, Cd, L.EPI_QKV
        a.bias, a.q_out, a.k_cache, a.v_cache, a.q_scale = bw.b_qkv, qb.data_ptr(), kv.data_ptr(), kv[n_seq * Cd * 680:].data_ptr(), bw.q_scale
        a.C, a.H, a.pos0, a.Lmax, a.rows_per_seq = Cd, 30, 424, 680, l
        t_bytes = M * Cd * 2 + 3 * Cd * Cd * 2 + 3 * M * Cd * 2
        timed(lambda: L.check(lib.var_b200_gemm_bf16(C.byref(a), L.current_stream())), t_bytes,
              f"gemm QKV M={M} N={3 * Cd} K={Cd} (KV append = {2 * M * Cd * 2 / 1e6:.0f} MB of the {t_bytes / 1e6:.0f} MB; "
              f"{2.0 * M * 3 * Cd * Cd / 1e12:.2f} TFLOP: tensor-bound, see roofline)")
elif kind in ("quant", "next"):
    vae, _ = build_vae_var(dev, depth=2)
    dense_init_(vae, seed=1)
    qz = vae.quantize
    if kind == "quant":
        B = 64
        f = torch.randn(B, 32, 16, 16, device=dev) * 1.5
        for mode in (0, 1):
            qz.search_mode = mode
            timed(lambda: qz.f_to_idxBl_or_fhat(f, to_fhat=False), B * (32 * 256 * 4 + 680 * 8),
                  f"quant_encode B={B} search_mode={mode} ({'tensor-core filter + exact re-rank' if mode == 0 else 'fused fp32'}; "
                  "dependency chain over 10 scales, not an HBM kernel)")
        qz.search_mode = 0
    else:
        B = 256
        f_hat = torch.zeros(B, 32, 16, 16, device=dev)
        idxs = [torch.randint(0, 4096, (B, p * p), device=dev) for p in qz.v_patch_nums]

        def run():
            for si in range(10):
                qz.get_next_autoregressive_input(si, 10, f_hat, idx_Bl=idxs[si], token_major=True)
        timed(run, 10 * B * 32 * 256 * 4 * 2, f"quant_next_input x10 scales B={B} (f_hat read + write per scale)")
print("ok")
