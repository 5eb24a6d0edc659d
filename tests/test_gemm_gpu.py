"""GPU parity of the tcgen05 GEMM family (through the C-ABI) against plain torch fp32 math on the same bf16 inputs."""
import ctypes as C
import math

import pytest
import torch

from var_b200 import lib as L

pytestmark = pytest.mark.gpu


def _gemm(A, W, epi, force_bn=0, **kw):
    lib = L.load()
    a = L.GemmArgs()
    a.A, a.W = A.data_ptr(), W.data_ptr()
    a.M, a.K = A.shape
    a.N = W.shape[0]
    a.epilogue, a.force_bn = epi, force_bn
    for k, v in kw.items():
        setattr(a, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    L.check(lib.var_b200_gemm_bf16(C.byref(a), L.current_stream()), "gemm")
    torch.cuda.synchronize()


def _ref(A, W):
    return A.float() @ W.float().t()


@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("N", [64, 128, 256])
def test_umma_probe(N, b_mn):
    torch.manual_seed(N + b_mn)
    lib = L.load()
    A = torch.randn(128, 64, device="cuda").bfloat16()
    B = (torch.randn(64, N, device="cuda") if b_mn else torch.randn(N, 64, device="cuda")).bfloat16()
    D = torch.full((128, N), float("nan"), device="cuda")
    L.check(lib.var_b200_umma_probe(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, b_mn, L.current_stream()), "probe")
    torch.cuda.synchronize()
    ref = A.float() @ (B.float() if b_mn else B.float().t())
    err = (D - ref).abs().max().item()
    assert err < 1e-3, f"N={N} b_mn={b_mn} max err {err}"


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 256, 64, 0), (128, 256, 256, 0), (300, 1024, 1024, 0), (5440, 3072, 1024, 0),
    (680, 1920, 1920, 0), (680, 1920, 1920, 128), (1000, 4096, 1024, 256), (257, 5760, 1920, 0),
    (4096, 7680, 1920, 0), (128, 1024, 4096, 128),
    (700, 1920, 1920, 256), (300, 2080, 512, 256), (300, 352, 256, 0), (129, 160, 128, 128),  # narrow last N tiles: 128 / 32 / 160 / 32 wide
    # the same shapes through the 1-CTA kernel (bit 16 of force_bn)
    (300, 1024, 1024, 0x10000 | 256), (680, 1920, 1920, 0x10000 | 192), (1000, 4096, 1024, 0x10000 | 128),
    (129, 256, 64, 0), (257, 256, 128, 0), (7000, 3072, 1024, 0),
])
def test_gemm_bias_f32(M, N, K, bn):
    torch.manual_seed(0)
    A = (torch.randn(M, K, device="cuda") / math.sqrt(K)).bfloat16()
    W = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda")
    _gemm(A, W, L.EPI_BIAS_F32, force_bn=bn, bias=bias, out=out)
    ref = _ref(A, W) + bias
    err = (out - ref).abs().max().item()
    assert torch.isfinite(out).all()
    assert err < 2e-3, f"max err {err}"


def test_gemm_gelu_bf16():
    torch.manual_seed(1)
    M, N, K = 1360, 4096, 1024
    A = (torch.randn(M, K, device="cuda") / math.sqrt(K)).bfloat16()
    W = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    _gemm(A, W, L.EPI_GELU_BF16, bias=bias, out=out)
    ref = torch.nn.functional.gelu(_ref(A, W) + bias, approximate="tanh")
    err = (out.float() - ref).abs().max().item()
    assert err < 3e-2, f"max err {err}"


def test_gemm_gate_resid_inplace():
    torch.manual_seed(2)
    n_seq, l, N, K = 6, 100, 1024, 4096
    M = n_seq * l
    A = (torch.randn(M, K, device="cuda") / math.sqrt(K)).bfloat16()
    W = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    x = torch.randn(M, N, device="cuda")
    ada = torch.randn(n_seq, 6 * N, device="cuda")
    ref = x + (_ref(A, W) + bias) * ada[:, N:2 * N].repeat_interleave(l, dim=0)
    _gemm(A, W, L.EPI_GATE_RESID, bias=bias, out=x, resid=x, gate=ada[:, N:], rows_per_seq=l, gate_ld=6 * N)
    err = (x - ref).abs().max().item()
    assert err < 5e-3, f"max err {err}"


def test_gemm_qkv_scatter():
    torch.manual_seed(3)
    n_seq, l, H, Lmax, pos0 = 4, 36, 16, 100, 55
    Cdim = H * 64
    M = n_seq * l
    A = (torch.randn(M, Cdim, device="cuda") / math.sqrt(Cdim)).bfloat16()
    W = torch.randn(3 * Cdim, Cdim, device="cuda").bfloat16()
    bias = torch.randn(3 * Cdim, device="cuda")
    bias[Cdim:2 * Cdim] = 0
    scale = torch.rand(H, device="cuda") * 5 + 1
    q = torch.zeros(n_seq, H, l, 64, device="cuda", dtype=torch.bfloat16)
    kc = torch.zeros(n_seq, H, Lmax, 64, device="cuda", dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    _gemm(A, W, L.EPI_QKV, bias=bias, q_out=q, k_cache=kc, v_cache=vc, q_scale=scale, C=Cdim, H=H, pos0=pos0, Lmax=Lmax,
          rows_per_seq=l)
    qkv = (_ref(A, W) + bias).view(n_seq, l, 3, H, 64).permute(2, 0, 3, 1, 4)
    qr = torch.nn.functional.normalize(qkv[0], dim=-1) * scale.view(1, H, 1, 1)
    kr = torch.nn.functional.normalize(qkv[1], dim=-1)
    vr = qkv[2]
    assert (q.float() - qr).abs().max().item() < 4e-2
    assert (kc[:, :, pos0:pos0 + l].float() - kr).abs().max().item() < 1e-2
    assert (vc[:, :, pos0:pos0 + l].float() - vr).abs().max().item() < 4e-2
    assert kc[:, :, :pos0].abs().max().item() == 0 and vc[:, :, pos0 + l:].abs().max().item() == 0


@pytest.mark.parametrize("n_seq,l,Cdim,K", [(6, 100, 1024, 1024), (3, 256, 1920, 7680), (40, 4, 1024, 4096), (130, 1, 640, 640),
                                            (5, 169, 2304, 2304)])
def test_gemm_deferred_layernorm_chain(n_seq, l, Cdim, K):
    """Deferred LayerNorm (gemm_sm100.cuh, LNF): the GATE_RESID producer's extra outputs (x*(1+scale) in bf16, per-row
    partial sums) and the GELU / QKV consumers that finish LN(x)(1+s)+sh = rstd*(acc - mean*U) + V, against plain fp32
    torch math: LayerNorm -> modulate -> Linear (basic_var.py:157-158)."""
    torch.manual_seed(n_seq * 1000 + l)
    lib = L.load()
    M, n_cls, H = n_seq * l, 9, Cdim // 64
    A = (torch.randn(M, K, device="cuda") / math.sqrt(K)).bfloat16()
    W = torch.randn(Cdim, K, device="cuda").bfloat16()
    bias = torch.randn(Cdim, device="cuda")
    x = torch.randn(M, Cdim, device="cuda") * 3 + 0.7          # residual stream with a DC offset
    labels = torch.randint(0, n_cls, (n_seq,), device="cuda", dtype=torch.int32)
    ada_cls = torch.randn(n_cls, 6 * Cdim, device="cuda") * 0.3   # per-class gamma1,gamma2,scale1,scale2,shift1,shift2
    ada = ada_cls[labels.long()].contiguous()
    x_ref = x + (_ref(A, W) + bias) * ada[:, Cdim:2 * Cdim].repeat_interleave(l, dim=0)
    parts = lib.var_b200_gemm_ln_parts(M, Cdim)
    a_out = torch.full((M, Cdim), float("nan"), device="cuda", dtype=torch.bfloat16)
    part = torch.full((M, parts, 2), float("nan"), device="cuda")
    _gemm(A, W, L.EPI_GATE_RESID, bias=bias, out=x, resid=x, gate=ada[:, Cdim:], rows_per_seq=l, gate_ld=6 * Cdim,
          ln_a_out=a_out, ln_scale=ada[:, 3 * Cdim:], ln_part_out=part)
    assert (x - x_ref).abs().max().item() < 5e-3
    scale2 = ada[:, 3 * Cdim:4 * Cdim].repeat_interleave(l, dim=0)
    shift2 = ada[:, 5 * Cdim:6 * Cdim].repeat_interleave(l, dim=0)
    assert (a_out.float() - x_ref * (1 + scale2)).abs().max().item() < 2 ** -8 * (x_ref * (1 + scale2)).abs().max().item() + 1e-2
    assert (part[..., 0].sum(1) - x_ref.sum(1)).abs().max().item() < 2e-2 * math.sqrt(Cdim)
    assert (part[..., 1].sum(1) - (x_ref * x_ref).sum(1)).abs().max().item() < 1e-4 * (x_ref * x_ref).sum(1).max().item()
    # reference of the consumer: Linear(LN(x)(1+scale)+shift), everything fp32
    a_ref = torch.nn.functional.layer_norm(x_ref, (Cdim,), eps=1e-6) * (1 + scale2) + shift2
    # --- GELU consumer (fc1)
    N1 = 4 * Cdim
    W1 = (torch.randn(N1, Cdim, device="cuda") / math.sqrt(Cdim)).bfloat16()
    b1 = torch.randn(N1, device="cuda") * 0.1
    u = ((1 + ada_cls[:, 3 * Cdim:4 * Cdim]).bfloat16().float() @ W1.float().t()).contiguous()
    v = (ada_cls[:, 5 * Cdim:6 * Cdim].bfloat16().float() @ W1.float().t() + b1).contiguous()
    h = torch.full((M, N1), float("nan"), device="cuda", dtype=torch.bfloat16)
    _gemm(a_out, W1, L.EPI_GELU_BF16, bias=b1, out=h, rows_per_seq=l, ln_part_in=part, ln_parts=parts, ln_C=Cdim, ln_eps=1e-6,
          ln_u=u, ln_v=v, ln_labels=labels)
    pre = a_ref @ W1.float().t() + b1
    ref_h = torch.nn.functional.gelu(pre, approximate="tanh")
    # the unfused path rounds the normalised operand to bf16 as well: compare the two error levels
    unf = torch.nn.functional.gelu(a_ref.bfloat16().float() @ W1.float().t() + b1, approximate="tanh")
    e_f, e_u = (h.float() - ref_h).abs().max().item(), (unf.bfloat16().float() - ref_h).abs().max().item()
    print(f"deferred-LN GELU consumer n_seq={n_seq} l={l} C={Cdim}: max err {e_f:.4f} (unfused bf16 operand: {e_u:.4f})")
    assert torch.isfinite(h.float()).all() and e_f < max(2.5 * e_u, 4e-2)
    # --- QKV consumer
    Wq = (torch.randn(3 * Cdim, Cdim, device="cuda") / math.sqrt(Cdim)).bfloat16()
    bq = torch.randn(3 * Cdim, device="cuda") * 0.1
    bq[Cdim:2 * Cdim] = 0
    uq = ((1 + ada_cls[:, 3 * Cdim:4 * Cdim]).bfloat16().float() @ Wq.float().t()).contiguous()
    vq = (ada_cls[:, 5 * Cdim:6 * Cdim].bfloat16().float() @ Wq.float().t() + bq).contiguous()
    qs = torch.rand(H, device="cuda") * 5 + 1
    Lmax, pos0 = l + 7, 3
    q = torch.zeros(n_seq, H, l, 64, device="cuda", dtype=torch.bfloat16)
    kc = torch.zeros(n_seq, H, Lmax, 64, device="cuda", dtype=torch.bfloat16)
    vc = torch.zeros_like(kc)
    _gemm(a_out, Wq, L.EPI_QKV, bias=bq, q_out=q, k_cache=kc, v_cache=vc, q_scale=qs, C=Cdim, H=H, pos0=pos0, Lmax=Lmax,
          rows_per_seq=l, ln_part_in=part, ln_parts=parts, ln_C=Cdim, ln_eps=1e-6, ln_u=uq, ln_v=vq, ln_labels=labels)
    qkv = (a_ref @ Wq.float().t() + bq).view(n_seq, l, 3, H, 64).permute(2, 0, 3, 1, 4)
    qr = torch.nn.functional.normalize(qkv[0], dim=-1) * qs.view(1, H, 1, 1)
    kr = torch.nn.functional.normalize(qkv[1], dim=-1)
    assert (q.float() - qr).abs().max().item() < 6e-2
    assert (kc[:, :, pos0:pos0 + l].float() - kr).abs().max().item() < 1.5e-2
    assert (vc[:, :, pos0:pos0 + l].float() - qkv[2]).abs().max().item() < 6e-2


def test_gemm_score_partials():
    torch.manual_seed(4)
    M, N, K = 700, 4096, 1024
    A = (torch.randn(M, K, device="cuda") / math.sqrt(K)).bfloat16()
    W = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    gt = torch.randint(0, N, (M,), device="cuda", dtype=torch.int32)
    bn = L.load().var_b200_gemm_tile_n(N)
    nt = (N + bn - 1) // bn
    part = torch.zeros(M, L.GEMM_EPI_PARTS * nt, 2, device="cuda")
    gl = torch.zeros(M, device="cuda")
    _gemm(A, W, L.EPI_SCORE, bias=bias, gt=gt, part=part, gt_logit=gl)
    logits = _ref(A, W) + bias
    m = part[..., 0].max(dim=1).values
    lse = m + torch.log((part[..., 1] * torch.exp(part[..., 0] - m[:, None])).sum(1))
    ref_lse = torch.logsumexp(logits, dim=1)
    ref_gl = logits.gather(1, gt.long()[:, None])[:, 0]
    assert (lse - ref_lse).abs().max().item() < 2e-3
    assert (gl - ref_gl).abs().max().item() < 2e-3


def pack_conv3x3(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> bf16 [Cout, 9 * kpt * 64], tap-major, channels zero-padded (include/var_b200.h)."""
    Cout, Cin = w.shape[:2]
    kp = (Cin + 63) // 64 * 64
    out = torch.zeros(Cout, 9, kp, device=w.device, dtype=torch.bfloat16)
    out[:, :, :Cin] = w.permute(0, 2, 3, 1).reshape(Cout, 9, Cin).to(torch.bfloat16)
    return out.reshape(Cout, 9 * kp).contiguous()


@pytest.mark.parametrize("B,H,W,Cin,Cout,resid", [
    (2, 16, 16, 32, 640, False), (2, 16, 16, 640, 640, True), (3, 32, 32, 320, 320, True), (1, 64, 64, 320, 160, False),
    (2, 128, 128, 160, 160, True), (1, 256, 256, 160, 192, False), (1, 8, 16, 64, 64, True), (5, 16, 16, 160, 128, False),
])
def test_conv3x3_nhwc_vs_torch(B, H, W, Cin, Cout, resid):
    """Implicit-GEMM 3x3 convolution (nine shifted TMA boxes, zero-filled borders) vs torch conv2d on the same bf16 data."""
    torch.manual_seed(B * 1000 + H + Cin)
    x = torch.randn(B, H, W, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(9 * Cin)).bfloat16()
    bias = torch.randn(Cout, device="cuda")
    r = torch.randn(B, H, W, Cout, device="cuda").bfloat16() if resid else None
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.load().var_b200_conv3x3_nhwc(x.data_ptr(), pack_conv3x3(w).data_ptr(), bias.data_ptr(),
                                           r.data_ptr() if resid else None, out.data_ptr(), B, H, W, Cin, Cout,
                                           L.current_stream()), "conv3x3")
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1)
    if resid:
        ref = ref.bfloat16().float() + r.float()
    err = (out.float() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all() and err <= 2 ** -7 * ref.abs().max().item() + 1e-2, f"max err {err}"  # bf16 ulp


@pytest.mark.parametrize("B,HW,Cdim", [(3, 256, 640), (2, 256, 128), (1, 1024, 64), (5, 256, 320)])
def test_vae_attn_block_vs_torch(B, HW, Cdim):
    """AttnBlock.forward (models/basic_vae.py:63-92) on var_b200's kernels (GroupNorm, five GEMM launches with the
    per-image products as block-diagonal batches, row softmax) vs the fp32 PyTorch module on the same bf16 input."""
    from var_b200.basic_vae import AttnBlock
    torch.manual_seed(B + HW + Cdim)
    side = int(math.isqrt(HW))
    blk = AttnBlock(Cdim).cuda()
    for prm in blk.parameters():
        torch.nn.init.normal_(prm, std=0.5 if prm.dim() == 1 else 1.5 / math.sqrt(Cdim))
    with torch.no_grad():
        blk.norm.weight.add_(1.0)
    x = (torch.randn(B, Cdim, side, side, device="cuda") * 1.3 + 0.2).bfloat16()
    with torch.no_grad():
        ref = blk(x.float())                                                       # [B, C, H, W] fp32
    xn = x.permute(0, 2, 3, 1).contiguous()                                        # NHWC
    out = torch.full_like(xn, float("nan"))
    lib = L.load()
    ws = torch.empty(lib.var_b200_vae_attn_workspace(B, HW, Cdim, 32), dtype=torch.uint8, device="cuda")
    wq = blk.qkv.weight.detach().reshape(3 * Cdim, Cdim).bfloat16().contiguous()
    wp = blk.proj_out.weight.detach().reshape(Cdim, Cdim).bfloat16().contiguous()
    f32 = lambda t: t.detach().float().contiguous()
    args = (f32(blk.norm.weight), f32(blk.norm.bias), f32(blk.qkv.bias), f32(blk.proj_out.bias))
    L.check(lib.var_b200_vae_attn_block(xn.data_ptr(), args[0].data_ptr(), args[1].data_ptr(), 32, blk.norm.eps, wq.data_ptr(),
                                        args[2].data_ptr(), wp.data_ptr(), args[3].data_ptr(), out.data_ptr(), B, HW, Cdim,
                                        ws.data_ptr(), ws.numel(), L.current_stream()), "vae_attn_block")
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    # five chained bf16 roundings (g, q/k, P, O, out) on values of size ~|ref|: a few bf16 ulps
    tol = 4 * 2 ** -8 * ref.abs().max().item() + 2e-2
    print(f"vae attn B={B} HW={HW} C={Cdim}: max err {err.max():.4f} mean {err.mean():.5f} (|ref| max {ref.abs().max():.2f})")
    assert torch.isfinite(got).all() and err.max().item() <= tol, f"max err {err.max().item()} > {tol}"
    # the attention term itself must be there (not just the shortcut)
    assert (ref - x.float()).abs().mean().item() > 10 * err.mean().item()


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 256, 256, 160, 160), (1, 128, 128, 160, 160), (3, 64, 64, 320, 320),
                                            (2, 32, 32, 320, 320), (1, 32, 32, 64, 32)])
def test_conv3x3_stride2_nhwc_vs_torch(B, H, W, Cin, Cout):
    """Downsample2x (models/basic_vae.py:31-37: zero pad (0,1,0,1), 3x3, stride 2) through stride-2 TMA boxes vs torch."""
    torch.manual_seed(B + H + Cin)
    x = torch.randn(B, H, W, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(9 * Cin)).bfloat16()
    bias = torch.randn(Cout, device="cuda")
    out = torch.full((B, H // 2, W // 2, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.load().var_b200_conv3x3_s2_nhwc(x.data_ptr(), pack_conv3x3(w).data_ptr(), bias.data_ptr(), out.data_ptr(), B, H, W,
                                              Cin, Cout, L.current_stream()), "conv3x3_s2")
    torch.cuda.synchronize()
    xp = torch.nn.functional.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1))
    ref = torch.nn.functional.conv2d(xp, w.float(), bias, stride=2).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all() and err <= 2 ** -7 * ref.abs().max().item() + 1e-2, f"max err {err}"


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 256, 256, 160, 32), (1, 16, 16, 640, 32), (2, 64, 64, 8, 160)])
def test_conv3x3_narrow_channels(B, H, W, Cin, Cout):
    """32-wide output tiles (decoder conv_out padded 3 -> 32, encoder conv_out 640 -> 32) and an 8-channel input (the
    image padded 3 -> 8): channel tails are TMA zero fill on the activation side and zero rows / columns in the weights."""
    torch.manual_seed(H + Cin + Cout)
    x = torch.randn(B, H, W, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(9 * Cin)).bfloat16()
    if Cout == 32:
        w[3:] = 0  # as the padded conv_out weights
    bias = torch.randn(Cout, device="cuda")
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.load().var_b200_conv3x3_nhwc(x.data_ptr(), pack_conv3x3(w).data_ptr(), bias.data_ptr(), None, out.data_ptr(), B, H,
                                           W, Cin, Cout, L.current_stream()), "conv3x3")
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all() and err <= 2 ** -7 * ref.abs().max().item() + 1e-2, f"max err {err}"


@pytest.mark.parametrize("n_pix,Cin,Cout,resid", [(512, 32, 32, False), (3 * 1024, 640, 320, False), (300, 320, 160, True),
                                                  (128 * 128, 160, 32, False)])
def test_conv1x1_nhwc_vs_torch(n_pix, Cin, Cout, resid):
    """1x1 convolution = GEMM over the pixels; Cin below / not a multiple of the 64-wide K block is zero-filled by TMA."""
    torch.manual_seed(n_pix + Cin)
    x = torch.randn(n_pix, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, Cin, device="cuda") / math.sqrt(Cin)).bfloat16()
    kp = (Cin + 63) // 64 * 64
    wp = torch.zeros(Cout, kp, device="cuda", dtype=torch.bfloat16)
    wp[:, :Cin] = w
    bias = torch.randn(Cout, device="cuda")
    r = torch.randn(n_pix, Cout, device="cuda").bfloat16() if resid else None
    out = torch.full((n_pix, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.load().var_b200_conv1x1_nhwc(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), r.data_ptr() if resid else None,
                                           out.data_ptr(), n_pix, Cin, Cout, L.current_stream()), "conv1x1")
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + bias
    if resid:
        ref = ref.bfloat16().float() + r.float()
    err = (out.float() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all() and err <= 2 ** -7 * ref.abs().max().item() + 1e-2, f"max err {err}"


@pytest.mark.parametrize("B,H,W,Cin,Cout,resid", [(2, 16, 16, 640, 640, True), (3, 32, 32, 320, 320, False),
                                                  (1, 256, 256, 160, 160, True), (2, 64, 64, 640, 320, False)])
def test_conv3x3_groupnorm_statistics_from_the_epilogue(B, H, W, Cin, Cout, resid):
    """var_b200_conv3x3_gn_nhwc: same output as the plain convolution (bit for bit) plus the GroupNorm(32) (sum, sum of
    squares) of exactly the stored bf16 values; var_b200_gn_apply_nhwc on them equals the two-pass GroupNorm kernel."""
    torch.manual_seed(H + Cin + Cout)
    lib = L.load()
    x = torch.randn(B, H, W, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(9 * Cin)).bfloat16()
    bias = torch.randn(Cout, device="cuda")
    r = torch.randn(B, H, W, Cout, device="cuda").bfloat16() if resid else None
    wp = pack_conv3x3(w)
    ref_out = torch.empty((B, H, W, Cout), device="cuda", dtype=torch.bfloat16)
    L.check(lib.var_b200_conv3x3_nhwc(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), r.data_ptr() if resid else None,
                                      ref_out.data_ptr(), B, H, W, Cin, Cout, L.current_stream()), "conv3x3")
    out = torch.full_like(ref_out, float("nan"))
    sums = torch.full((B, 32, 2), float("nan"), device="cuda")
    ws = torch.empty(lib.var_b200_conv3x3_gn_workspace(B, H, W, Cout), dtype=torch.uint8, device="cuda")
    L.check(lib.var_b200_conv3x3_gn_nhwc(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), r.data_ptr() if resid else None,
                                         out.data_ptr(), B, H, W, Cin, Cout, 32, sums.data_ptr(), ws.data_ptr(), ws.numel(),
                                         L.current_stream()), "conv3x3_gn")
    torch.cuda.synchronize()
    assert torch.equal(out, ref_out)
    y = out.double().reshape(B, H * W, 32, Cout // 32)
    ref_s = torch.stack((y.sum(dim=(1, 3)), (y * y).sum(dim=(1, 3))), dim=-1)
    rel = ((sums.double() - ref_s).abs() / ref_s.abs().clamp_min(1.0)).max().item()
    assert rel < 2e-5, f"statistics rel err {rel}"   # fp32 accumulation of up to 65536 x 5 values per group
    # apply from the sums vs the two-pass kernel
    gamma, beta = torch.randn(Cout, device="cuda"), torch.randn(Cout, device="cuda")
    y1, y2 = torch.empty_like(out), torch.empty_like(out)
    L.check(lib.var_b200_gn_apply_nhwc(out.data_ptr(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y1.data_ptr(), B, H * W,
                                       Cout, 32, 1e-6, 1, L.current_stream()), "gn_apply")
    ws2 = torch.empty(lib.var_b200_gn_workspace(B, H * W, Cout, 32), dtype=torch.uint8, device="cuda")
    L.check(lib.var_b200_gn_silu_nhwc(out.data_ptr(), None, gamma.data_ptr(), beta.data_ptr(), y2.data_ptr(), B, H * W, Cout, 32,
                                      1e-6, 1, ws2.data_ptr(), ws2.numel(), L.current_stream()), "gn_silu")
    torch.cuda.synchronize()
    d = (y1.float() - y2.float()).abs().max().item()
    assert d <= 2 ** -6 * y2.float().abs().max().item(), f"GroupNorm from epilogue statistics differs from the two-pass kernel by {d}"
    # deterministic
    sums2 = torch.empty_like(sums)
    L.check(lib.var_b200_conv3x3_gn_nhwc(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), r.data_ptr() if resid else None,
                                         out.data_ptr(), B, H, W, Cin, Cout, 32, sums2.data_ptr(), ws.data_ptr(), ws.numel(),
                                         L.current_stream()), "conv3x3_gn")
    torch.cuda.synchronize()
    assert torch.equal(sums, sums2)
