"""Host-side mirror of VQVAE (/root/reference/models/vqvae.py): same constructor, methods and state_dict keys.
The quantizer runs on the sm_100a kernels; the CNN encoder/decoder stay PyTorch (boundary helpers)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from .basic_vae import Decoder, Encoder
from .quant import VectorQuantizer2


class VQVAE(nn.Module):
    def __init__(self, vocab_size=4096, z_channels=32, ch=128, dropout=0.0, beta=0.25, using_znorm=False,
                 quant_conv_ks=3, quant_resi=0.5, share_quant_resi=4, default_qresi_counts=0,
                 v_patch_nums=(1, 2, 3, 4, 5, 6, 8, 10, 13, 16), test_mode=True):
        super().__init__()
        self.test_mode = test_mode
        self.decoder_dtype = None  # set to torch.bfloat16 / float16 to run fhat_to_img through a 16-bit decoder copy
        self.decoder_nhwc = True   # with decoder_dtype=bfloat16: channels-last plan with the fused GroupNorm+SiLU kernel
        self.decoder_own_conv = True  # ... and the 3x3 convolutions on var_b200's implicit-GEMM tcgen05 kernel (else cuDNN)
        self.encoder_dtype = None  # torch.bfloat16: channels-last 16-bit encoder plan (token indices then depend on bf16
        #                            rounding of the features; the default fp32 encoder keeps them reference-exact)
        self.V, self.Cvae = vocab_size, z_channels
        cfg = dict(ch=ch, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, in_channels=3, z_channels=z_channels)
        self.encoder = Encoder(**cfg)
        self.decoder = Decoder(**cfg)
        self.vocab_size = vocab_size
        self.downsample = 2 ** (len(cfg["ch_mult"]) - 1)
        self.quantize = VectorQuantizer2(vocab_size=vocab_size, Cvae=z_channels, using_znorm=using_znorm, beta=beta,
                                         default_qresi_counts=default_qresi_counts, v_patch_nums=v_patch_nums,
                                         quant_resi=quant_resi, share_quant_resi=share_quant_resi)
        self.quant_conv = nn.Conv2d(z_channels, z_channels, quant_conv_ks, 1, quant_conv_ks // 2)
        self.post_quant_conv = nn.Conv2d(z_channels, z_channels, quant_conv_ks, 1, quant_conv_ks // 2)
        if test_mode:
            self.eval()
            for p in self.parameters():
                p.requires_grad_(False)

    def forward(self, inp, ret_usages=False):
        raise NotImplementedError("VQVAE.forward is VAE training (out of scope, SURVEY.md §2 row 3)")

    # ---- decode side (vqvae.py:62-63, 77-90)
    def fhat_to_img(self, f_hat: torch.Tensor):
        if self.decoder_dtype is torch.bfloat16 and f_hat.is_cuda and self.decoder_nhwc:
            if getattr(self, "_nhwc_dec", None) is None:
                from .basic_vae import NHWCDecoder
                self._nhwc_dec = NHWCDecoder(self.decoder, self.post_quant_conv)
            self._nhwc_dec.own_conv = bool(getattr(self, "decoder_own_conv", True))
            return self._nhwc_dec(f_hat).clamp_(-1, 1)
        if self.decoder_dtype is not None and f_hat.is_cuda:
            post, dec = self._low_precision_decoder()
            return dec(post(f_hat.to(self.decoder_dtype))).float().clamp_(-1, 1)
        return self.decoder(self.post_quant_conv(f_hat)).clamp_(-1, 1)

    def _low_precision_decoder(self):
        """bf16/fp16 copy of the CNN decoder (boundary helper, cuDNN): unlike autocast, GroupNorm/SiLU also read and
        write 16-bit tensors, which halves the elementwise HBM traffic of the 256-px stages."""
        import copy
        key = (self.decoder_dtype, tuple(p._version for p in self.decoder.parameters()),
               tuple(p._version for p in self.post_quant_conv.parameters()), self.decoder.conv_in.weight.data_ptr())
        if getattr(self, "_lp_key", None) != key:
            self._lp = (copy.deepcopy(self.post_quant_conv).to(self.decoder_dtype),
                        copy.deepcopy(self.decoder).to(self.decoder_dtype))
            self._lp_key = key
        return self._lp

    def idxBl_to_img(self, ms_idx_Bl: List[torch.Tensor], same_shape: bool, last_one=False):
        if not same_shape:
            raise NotImplementedError("same_shape=False is the reference's experimental path (quant.py:122-131)")
        if last_one:
            return self.fhat_to_img(self.quantize.idxBl_to_fhat(ms_idx_Bl, last_one=True))
        return [self.fhat_to_img(f) for f in self.quantize.idxBl_to_fhat(ms_idx_Bl, last_one=False)]

    def embed_to_img(self, ms_h_BChw: List[torch.Tensor], all_to_max_scale: bool, last_one=False):
        """vqvae.py:86-90 on arbitrary per-scale embedding maps."""
        fs = self.quantize.embed_to_fhat(ms_h_BChw, all_to_max_scale=all_to_max_scale, last_one=last_one)
        return self.fhat_to_img(fs) if last_one else [self.fhat_to_img(f) for f in fs]

    # ---- encode side (vqvae.py:65-75, 92-98)
    def img_to_post(self, inp_img_no_grad: torch.Tensor, v_patch_nums=None):
        if self.encoder_dtype is torch.bfloat16 and inp_img_no_grad.is_cuda:
            if getattr(self, "_nhwc_enc", None) is None:
                from .basic_vae import NHWCEncoder
                self._nhwc_enc = NHWCEncoder(self.encoder, self.quant_conv)
            self._nhwc_enc.own_conv = bool(getattr(self, "decoder_own_conv", True))
            return self._nhwc_enc(inp_img_no_grad)
        return self.quant_conv(self.encoder(inp_img_no_grad))

    def img_to_idxBl(self, inp_img_no_grad: torch.Tensor,
                     v_patch_nums: Optional[Sequence[Union[int, Tuple[int, int]]]] = None) -> List[torch.LongTensor]:
        return self.quantize.f_to_idxBl_or_fhat(self.img_to_post(inp_img_no_grad), to_fhat=False,
                                                v_patch_nums=v_patch_nums)

    def img_to_fhat(self, inp_img_no_grad: torch.Tensor, v_patch_nums=None) -> List[torch.Tensor]:
        return self.quantize.f_to_idxBl_or_fhat(self.img_to_post(inp_img_no_grad), to_fhat=True,
                                                v_patch_nums=v_patch_nums)

    def img_to_reconstructed_img(self, x, v_patch_nums=None, last_one=False):
        fs = self.img_to_fhat(x, v_patch_nums)
        if last_one:
            return self.fhat_to_img(fs[-1])
        return [self.fhat_to_img(f) for f in fs]

    def load_state_dict(self, state_dict: Dict[str, Any], strict=True, assign=False):
        k = "quantize.ema_vocab_hit_SV"  # vqvae.py:100-103
        if k in state_dict and state_dict[k].shape[0] != self.quantize.ema_vocab_hit_SV.shape[0]:
            state_dict[k] = self.quantize.ema_vocab_hit_SV
        return super().load_state_dict(state_dict=state_dict, strict=strict, assign=assign)
