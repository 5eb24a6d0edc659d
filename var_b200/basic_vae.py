"""CNN encoder / decoder of the VQVAE (models/basic_vae.py:99-226).

Two forms live here. The `nn.Module` tree (`Encoder`, `Decoder`, `ResnetBlock`, `AttnBlock`) reproduces the reference's
state_dict keys so `vae_ch160v4096z32.pth` loads with strict=True, and is the fp32 PyTorch parity reference (its encoder
produces the bit-exact token indices). The execution plans `NHWCDecoder` / `NHWCEncoder` run the same functions on bf16
channels-last tensors entirely on var_b200's own kernels (implicit-GEMM convolutions on the tcgen05 GEMM, stride-2 TMA
boxes, block-diagonally batched AttnBlock, GroupNorm with statistics from the producing convolution's epilogue); cuDNN is
only the fallback for shapes the kernels do not tile and the `own_conv=False` A/B switch.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F


def _gn(ch: int) -> nn.GroupNorm:
    return nn.GroupNorm(32, ch, eps=1e-6, affine=True)


class ResnetBlock(nn.Module):
    """GN-SiLU-conv3x3 twice plus a (1x1-projected) skip (models/basic_vae.py:40-60)."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.norm1, self.conv1 = _gn(cin), nn.Conv2d(cin, cout, 3, 1, 1)
        self.norm2, self.conv2 = _gn(cout), nn.Conv2d(cout, cout, 3, 1, 1)
        self.nin_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else nn.Identity()

    def forward(self, x):
        y = self.conv1(F.silu(self.norm1(x)))
        y = self.conv2(F.silu(self.norm2(y)))
        return self.nin_shortcut(x) + y


class AttnBlock(nn.Module):
    """Single-head spatial self-attention with scale C^-0.5 (models/basic_vae.py:63-92)."""

    def __init__(self, ch: int):
        super().__init__()
        self.C = ch
        self.norm = _gn(ch)
        self.qkv = nn.Conv2d(ch, 3 * ch, 1)
        self.proj_out = nn.Conv2d(ch, ch, 1)

    def forward(self, x):
        B, C, H, W = x.shape
        q, k, v = self.qkv(self.norm(x)).view(B, 3, C, H * W).transpose(2, 3).unbind(1)  # each [B, HW, C]
        o = F.scaled_dot_product_attention(q, k, v, scale=float(C) ** -0.5)
        return x + self.proj_out(o.transpose(1, 2).reshape(B, C, H, W))


class _Resample(nn.Module):
    def __init__(self, ch: int, down: bool):
        super().__init__()
        self.down = down
        self.conv = nn.Conv2d(ch, ch, 3, 2 if down else 1, 0 if down else 1)

    def forward(self, x):
        if self.down:  # asymmetric pad then stride-2 conv (models/basic_vae.py:31-37)
            return self.conv(F.pad(x, (0, 1, 0, 1)))
        return self.conv(F.interpolate(x, scale_factor=2, mode="nearest"))  # models/basic_vae.py:22-28


class _Level(nn.Module):
    def __init__(self, cin: int, cout: int, n_blocks: int, with_attn: bool, resample: str | None):
        super().__init__()
        self.block = nn.ModuleList()
        self.attn = nn.ModuleList()
        for _ in range(n_blocks):
            self.block.append(ResnetBlock(cin, cout))
            cin = cout
            if with_attn:
                self.attn.append(AttnBlock(cout))
        if resample == "down":
            self.downsample = _Resample(cout, True)
        elif resample == "up":
            self.upsample = _Resample(cout, False)

    def run(self, h):
        for i, blk in enumerate(self.block):
            h = blk(h)
            if len(self.attn):
                h = self.attn[i](h)
        return h


class _Mid(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.block_1, self.attn_1, self.block_2 = ResnetBlock(ch, ch), AttnBlock(ch), ResnetBlock(ch, ch)

    def forward(self, h):
        return self.block_2(self.attn_1(self.block_1(h)))


class Encoder(nn.Module):
    def __init__(self, ch=160, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, in_channels=3, z_channels=32):
        super().__init__()
        n = len(ch_mult)
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
        mults = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList(
            _Level(ch * mults[i], ch * ch_mult[i], num_res_blocks, with_attn=(i == n - 1),
                   resample=("down" if i != n - 1 else None)) for i in range(n))
        top = ch * ch_mult[-1]
        self.mid = _Mid(top)
        self.norm_out = _gn(top)
        self.conv_out = nn.Conv2d(top, z_channels, 3, 1, 1)

    def forward(self, x):
        h = self.conv_in(x)
        for lvl in self.down:
            h = lvl.run(h)
            if hasattr(lvl, "downsample"):
                h = lvl.downsample(h)
        return self.conv_out(F.silu(self.norm_out(self.mid(h))))


class Decoder(nn.Module):
    def __init__(self, ch=160, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, in_channels=3, z_channels=32):
        super().__init__()
        n = len(ch_mult)
        top = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, top, 3, 1, 1)
        self.mid = _Mid(top)
        levels, cin = [], top
        for i in reversed(range(n)):
            cout = ch * ch_mult[i]
            levels.insert(0, _Level(cin, cout, num_res_blocks + 1, with_attn=(i == n - 1),
                                    resample=("up" if i != 0 else None)))
            cin = cout
        self.up = nn.ModuleList(levels)
        self.norm_out = _gn(cin)
        self.conv_out = nn.Conv2d(cin, in_channels, 3, 1, 1)

    def forward(self, z):
        h = self.mid(self.conv_in(z))
        for lvl in reversed(self.up):
            h = lvl.run(h)
            if hasattr(lvl, "upsample"):
                h = lvl.upsample(h)
        return self.conv_out(F.silu(self.norm_out(h)))


class NHWCDecoder:
    """16-bit channels-last execution plan of a `Decoder` (+ post_quant_conv). The 3x3 convolutions run on var_b200's own
    implicit-GEMM tcgen05 kernel (`var_b200_conv3x3_nhwc`: nine shifted TMA boxes per K sweep, bias and the ResnetBlock
    shortcut fused into the epilogue; 3-channel ends padded to 8 inputs / 32 outputs) and the 1x1 convolutions on the plain
    GEMM; shapes the kernel cannot tile (odd widths) and `own_conv=False` fall back to bias-free cuDNN NHWC convolutions.
    The AttnBlock is `var_b200_vae_attn_block`. Everything between
    the convolutions is var_b200's NHWC glue (csrc/groupnorm.cu): GroupNorm+SiLU in two passes, residual adds with the
    convolution biases folded in, nearest-2x up-sampling. Same function as
    Decoder.forward(post_quant_conv(f_hat)) up to 16-bit rounding
    (tests/test_parity_gpu.py::test_nhwc_decoder_matches_pytorch_decoder)."""

    def __init__(self, decoder: Decoder, post_quant_conv: nn.Conv2d, dtype=torch.bfloat16):
        if dtype != torch.bfloat16:
            raise NotImplementedError("the NHWC glue kernels are bf16")
        self.dtype = dtype
        self.dec = decoder
        self.post = post_quant_conv
        self._w = {}
        self.own_conv = True   # False: every convolution through cuDNN (A/B measurements)
        self.gn_from_conv = True  # GroupNorm statistics from the producing convolution's epilogue (False: statistics pass)
        from . import lib as L
        self.L, self.lib = L, L.load()

    # ---- cached 16-bit / fp32 parameter copies
    def _cw(self, m: nn.Conv2d):
        w = self._w.get(id(m))
        if w is None or w[2] != (m.weight._version, m.bias._version):
            w = (m.weight.detach().to(self.dtype).contiguous(memory_format=torch.channels_last),
                 m.bias.detach().float().contiguous(), (m.weight._version, m.bias._version))
            self._w[id(m)] = w
        return w

    def _conv(self, x, m: nn.Conv2d):
        """bias-free convolution; the caller folds `self._cw(m)[1]` into the consumer kernel"""
        return F.conv2d(x, self._cw(m)[0], None, stride=m.stride, padding=m.padding)

    def _bias(self, m):
        return self._cw(m)[1]

    # ---- own tcgen05 convolutions
    def _own_ok(self, x, m: nn.Conv2d) -> bool:
        """Can this convolution run on var_b200's implicit-GEMM kernels? (3x3 stride 1 pad 1, the 3x3 stride-2
        Downsample2x on a (0,1,0,1)-padded input, 1x1.) Channel counts are padded where the packing allows it: 3 input
        channels to 8 (conv_in of the encoder), 3 output channels to 32 (conv_out of the decoder)."""
        B, Cin, H, W = x.shape if isinstance(x, torch.Tensor) else x
        if not self.own_conv or (m.out_channels % 32 and m.out_channels != 3) or (Cin % 8 and Cin != 3):
            return False
        if m.kernel_size == (1, 1):
            return m.padding == (0, 0) and m.stride == (1, 1)
        if m.kernel_size != (3, 3):
            return False
        if m.stride == (2, 2):  # Downsample2x: H, W are the unpadded input extent
            Ho, Wo = H // 2, W // 2
            return m.padding == (0, 0) and H % 2 == 0 and W % 2 == 0 and Wo <= 128 and 128 % Wo == 0 and (Ho * Wo) % 128 == 0
        return (m.stride == (1, 1) and m.padding == (1, 1) and (H * W) % 128 == 0
                and (128 % W == 0 if W <= 128 else W % 128 == 0))

    def _packed(self, m: nn.Conv2d):
        """bf16 weights [Cout', taps, ceil64(Cin')] (tap-major, zero padded; Cin' = Cin rounded up to 8, Cout' = Cout
        rounded up to 32) + fp32 bias [Cout']"""
        key = ("own", id(m))
        w = self._w.get(key)
        if w is None or w[2] != (m.weight._version, m.bias._version):
            Cout, Cin, kh, kw = m.weight.shape
            kp = (Cin + 63) // 64 * 64
            cop = (Cout + 31) // 32 * 32
            wp = torch.zeros((cop, kh * kw, kp), dtype=torch.bfloat16, device=m.weight.device)
            wp[:Cout, :, :Cin] = m.weight.detach().permute(0, 2, 3, 1).reshape(Cout, kh * kw, Cin).to(torch.bfloat16)
            bias = torch.zeros(cop, dtype=torch.float32, device=m.weight.device)
            bias[:Cout] = m.bias.detach().float()
            w = (wp.reshape(cop, kh * kw * kp).contiguous(), bias, (m.weight._version, m.bias._version))
            self._w[key] = w
        return w

    def _own_conv(self, x, m: nn.Conv2d, resid=None, stats=False):
        """conv + bias (+ resid) on the tcgen05 kernels; x / resid / result: bf16 channels_last. A 3-channel output comes
        back as the 32-channel padded tensor's first three channels (a view). stats=True returns (y, sums): the
        convolution's epilogue also leaves the GroupNorm(32) statistics of y behind (sums is None when the shape does not
        allow it), so the GroupNorm that follows needs no statistics pass."""
        if stats:
            B, Cin, H, W = x.shape
            if (self.gn_from_conv and m.kernel_size == (3, 3) and m.stride == (1, 1) and (H * W) % 256 == 0 and Cin % 8 == 0
                    and m.out_channels % 32 == 0):
                wp, bias, _ = self._packed(m)
                Cout = m.out_channels
                y = torch.empty((B, Cout, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
                sums = torch.empty((B, 32, 2), dtype=torch.float32, device=x.device)
                ws = torch.empty(self.lib.var_b200_conv3x3_gn_workspace(B, H, W, Cout), dtype=torch.uint8, device=x.device)
                self.L.check(self.lib.var_b200_conv3x3_gn_nhwc(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), self.L.ptr(resid),
                                                               y.data_ptr(), B, H, W, Cin, Cout, 32, sums.data_ptr(), ws.data_ptr(),
                                                               ws.numel(), self.L.current_stream()), "conv3x3_gn_nhwc")
                return y, sums
            return self._own_conv(x, m, resid), None
        B, Cin, H, W = x.shape
        assert x.is_contiguous(memory_format=torch.channels_last) and x.dtype == torch.bfloat16
        if Cin % 8:  # 3 image channels -> 8 (zeros): TMA rows are at least 16 bytes
            x8 = torch.zeros((B, 8, H, W), dtype=x.dtype, device=x.device).contiguous(memory_format=torch.channels_last)
            x8[:, :Cin] = x
            x, Cin = x8, 8
        wp, bias, _ = self._packed(m)
        Cout = wp.shape[0]  # padded to a multiple of 32
        st = self.L.current_stream()
        if m.kernel_size == (3, 3) and m.stride == (2, 2):
            assert resid is None
            y = torch.empty((B, Cout, H // 2, W // 2), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
            self.L.check(self.lib.var_b200_conv3x3_s2_nhwc(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W,
                                                           Cin, Cout, st), "conv3x3_s2_nhwc")
        elif m.kernel_size == (3, 3):
            y = torch.empty((B, Cout, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
            self.L.check(self.lib.var_b200_conv3x3_nhwc(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), self.L.ptr(resid),
                                                        y.data_ptr(), B, H, W, Cin, Cout, st), "conv3x3_nhwc")
        else:  # 1x1: a plain GEMM over the pixels
            y = torch.empty((B, Cout, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
            self.L.check(self.lib.var_b200_conv1x1_nhwc(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), self.L.ptr(resid),
                                                        y.data_ptr(), B * H * W, Cin, Cout, st), "conv1x1_nhwc")
        return y if Cout == m.out_channels else y[:, :m.out_channels]

    def _conv_any(self, x, m: nn.Conv2d):
        """conv + bias: own kernel when the shape allows it, else cuDNN + the bias-add glue kernel"""
        if self._own_ok(x, m):
            return self._own_conv(x, m)
        y = self._conv(x, m)
        return self._add(y, self._bias(m), None, None, out=y)

    def _gn(self, x, m: nn.GroupNorm, silu: bool, pre_bias=None, sums=None):
        B, Cc, H, W = x.shape
        assert x.is_contiguous(memory_format=torch.channels_last) and x.dtype == torch.bfloat16
        p = self._w.get(id(m))
        if p is None or p[2] != (m.weight._version, m.bias._version):
            p = (m.weight.detach().float().contiguous(), m.bias.detach().float().contiguous(),
                 (m.weight._version, m.bias._version))
            self._w[id(m)] = p
        y = torch.empty_like(x)  # preserves channels_last
        if sums is not None and pre_bias is None and m.num_groups == 32:  # statistics came with the producing convolution
            self.L.check(self.lib.var_b200_gn_apply_nhwc(x.data_ptr(), sums.data_ptr(), p[0].data_ptr(), p[1].data_ptr(),
                                                         y.data_ptr(), B, H * W, Cc, m.num_groups, m.eps, int(silu),
                                                         self.L.current_stream()), "gn_apply_nhwc")
            return y
        ws = torch.empty(self.lib.var_b200_gn_workspace(B, H * W, Cc, m.num_groups), dtype=torch.uint8, device=x.device)
        self.L.check(self.lib.var_b200_gn_silu_nhwc(x.data_ptr(), self.L.ptr(pre_bias), p[0].data_ptr(), p[1].data_ptr(),
                                                    y.data_ptr(), B, H * W, Cc, m.num_groups, m.eps, int(silu),
                                                    ws.data_ptr(), ws.numel(), self.L.current_stream()), "gn_silu_nhwc")
        return y

    def _add(self, a, bias_a, b, bias_b, out=None):
        B, Cc, H, W = a.shape
        out = torch.empty_like(a) if out is None else out
        self.L.check(self.lib.var_b200_add_bias_nhwc(a.data_ptr(), self.L.ptr(bias_a), self.L.ptr(b), self.L.ptr(bias_b),
                                                     out.data_ptr(), B * H * W, Cc, self.L.current_stream()), "add_bias_nhwc")
        return out

    def _res(self, x, blk: ResnetBlock, x_sums=None):
        """-> (block output, GroupNorm statistics of it or None). x_sums: statistics of x if its producer left them."""
        g1 = self._gn(x, blk.norm1, True, sums=x_sums)
        B_, _, H_, W_ = g1.shape
        if self._own_ok(g1, blk.conv1) and self._own_ok((B_, blk.conv1.out_channels, H_, W_), blk.conv2):
            h, h_sums = self._own_conv(g1, blk.conv1, stats=True)
            g2 = self._gn(h, blk.norm2, True, sums=h_sums)
            if isinstance(blk.nin_shortcut, nn.Identity):
                sc = x
            elif self._own_ok(x, blk.nin_shortcut):
                sc = self._own_conv(x, blk.nin_shortcut)
            else:
                sc = self._conv(x, blk.nin_shortcut)
                sc = self._add(sc, self._bias(blk.nin_shortcut), None, None, out=sc)
            return self._own_conv(g2, blk.conv2, resid=sc, stats=True)
        h = self._conv(g1, blk.conv1)
        h = self._conv(self._gn(h, blk.norm2, True, pre_bias=self._bias(blk.conv1)), blk.conv2)
        if isinstance(blk.nin_shortcut, nn.Identity):
            return self._add(h, self._bias(blk.conv2), x, None, out=h), None
        sc = self._conv(x, blk.nin_shortcut)
        return self._add(h, self._bias(blk.conv2), sc, self._bias(blk.nin_shortcut), out=h), None

    def _attn(self, x, blk: AttnBlock):
        B, Cc, H, W = x.shape
        if self.own_conv and (H * W) % 256 == 0 and H * W <= 4096 and Cc % 64 == 0:
            # the whole block on var_b200's kernels (csrc/vae_attn.cu): GroupNorm, the 1x1 convolutions and both
            # per-image products on the tcgen05 GEMM (block-diagonal batching), row softmax, shortcut in the epilogue
            key = ("attn", id(blk))
            w = self._w.get(key)
            ver = tuple(p_._version for p_ in blk.parameters())
            if w is None or w[-1] != ver:
                f32 = lambda t: t.detach().float().contiguous()
                w = (blk.qkv.weight.detach().reshape(3 * Cc, Cc).to(torch.bfloat16).contiguous(), f32(blk.qkv.bias),
                     blk.proj_out.weight.detach().reshape(Cc, Cc).to(torch.bfloat16).contiguous(), f32(blk.proj_out.bias),
                     f32(blk.norm.weight), f32(blk.norm.bias), ver)
                self._w[key] = w
            y = torch.empty_like(x)
            ws = torch.empty(self.lib.var_b200_vae_attn_workspace(B, H * W, Cc, blk.norm.num_groups), dtype=torch.uint8,
                             device=x.device)
            self.L.check(self.lib.var_b200_vae_attn_block(x.data_ptr(), w[4].data_ptr(), w[5].data_ptr(), blk.norm.num_groups,
                                                          blk.norm.eps, w[0].data_ptr(), w[1].data_ptr(), w[2].data_ptr(),
                                                          w[3].data_ptr(), y.data_ptr(), B, H * W, Cc, ws.data_ptr(), ws.numel(),
                                                          self.L.current_stream()), "vae_attn_block")
            return y
        qkv = self._conv(self._gn(x, blk.norm, False), blk.qkv)               # [B,3C,H,W] channels_last, no bias yet
        qkv = self._add(qkv, self._bias(blk.qkv), None, None, out=qkv)
        q, k, v = qkv.permute(0, 2, 3, 1).reshape(B, H * W, 3 * Cc).chunk(3, dim=-1)
        o = F.scaled_dot_product_attention(q, k, v, scale=float(Cc) ** -0.5)   # [B,HW,C]
        o = o.reshape(B, H, W, Cc).permute(0, 3, 1, 2)                         # channels_last view
        p = self._conv(o, blk.proj_out)
        return self._add(p, self._bias(blk.proj_out), x, None, out=p)

    def _level(self, h, lvl, sums=None):
        for i, blk in enumerate(lvl.block):
            h, sums = self._res(h, blk, sums)
            if len(lvl.attn):
                h, sums = self._attn(h, lvl.attn[i]), None
        return h, sums

    def _upsample(self, h, conv: nn.Conv2d):
        B, Cc, H, W = h.shape
        up = torch.empty((B, Cc, 2 * H, 2 * W), dtype=h.dtype, device=h.device, memory_format=torch.channels_last)
        self.L.check(self.lib.var_b200_upsample2x_nhwc(h.data_ptr(), None, up.data_ptr(), B, H, W, Cc,
                                                       self.L.current_stream()), "upsample2x_nhwc")
        if self._own_ok(up, conv):
            return self._own_conv(up, conv, stats=True)
        return self._conv_any(up, conv), None

    @torch.no_grad()
    def __call__(self, f_hat: torch.Tensor) -> torch.Tensor:
        d = self.dec
        x = f_hat.to(self.dtype).contiguous(memory_format=torch.channels_last)
        h = self._conv_any(x, self.post)
        h, sums = self._own_conv(h, d.conv_in, stats=True) if self._own_ok(h, d.conv_in) else (self._conv_any(h, d.conv_in), None)
        h, _ = self._res(h, d.mid.block_1, sums)
        h = self._attn(h, d.mid.attn_1)
        h, sums = self._res(h, d.mid.block_2)
        for lvl in reversed(d.up):
            h, sums = self._level(h, lvl, sums)
            if hasattr(lvl, "upsample"):
                h, sums = self._upsample(h, lvl.upsample.conv)
        g = self._gn(h, d.norm_out, True, sums=sums)
        if self._own_ok(g, d.conv_out):  # 3 output channels as a 32-wide tile of the implicit GEMM
            h = self._own_conv(g, d.conv_out)
        else:
            w, b, _ = self._cw(d.conv_out)
            h = F.conv2d(g, w, b.to(self.dtype), padding=d.conv_out.padding)
        return h.float().contiguous()


class NHWCEncoder(NHWCDecoder):
    """Channels-last 16-bit plan of `Encoder` (+ quant_conv): same glue kernels as NHWCDecoder; Downsample2x is the
    reference's asymmetric (0,1,0,1) zero pad followed by a stride-2 convolution (models/basic_vae.py:31-37)."""

    def __init__(self, encoder: Encoder, quant_conv: nn.Conv2d, dtype=torch.bfloat16):
        super().__init__(encoder, quant_conv, dtype)
        self.enc = encoder
        self.qconv = quant_conv

    @torch.no_grad()
    def __call__(self, img: torch.Tensor) -> torch.Tensor:
        e = self.enc
        x = img.to(self.dtype).contiguous(memory_format=torch.channels_last)
        if self._own_ok(x, e.conv_in):  # 3 image channels zero-padded to 8
            h = self._own_conv(x, e.conv_in)
        else:
            w, b, _ = self._cw(e.conv_in)
            h = F.conv2d(x, w, b.to(self.dtype), padding=e.conv_in.padding)
        sums = None
        for lvl in e.down:
            h, sums = self._level(h, lvl, sums)
            if hasattr(lvl, "downsample"):
                dc = lvl.downsample.conv
                sums = None
                if self._own_ok(h, dc):  # stride-2 TMA boxes; the (0,1,0,1) zero pad is TMA's out-of-bounds fill
                    h = self._own_conv(h, dc)
                else:
                    y = self._conv(F.pad(h, (0, 1, 0, 1)), dc)
                    h = self._add(y, self._bias(dc), None, None, out=y)
        h, _ = self._res(h, e.mid.block_1, sums)
        h = self._attn(h, e.mid.attn_1)
        h, sums = self._res(h, e.mid.block_2)
        h = self._conv_any(self._conv_any(self._gn(h, e.norm_out, True, sums=sums), e.conv_out), self.qconv)
        return h.float().contiguous()  # fp32 NCHW features for the quantizer
