"""var_b200 — B200-native (sm_100a) hot path of VAR next-scale prediction behind the reference's Python API.

    from var_b200 import build_vae_var
    vae, var = build_vae_var(device="cuda", depth=16)          # same signature as models/__init__.py:9-39
    idx = vae.img_to_idxBl(img); logits = var(label_B, vae.quantize.idxBl_to_var_input(idx))
    imgs = var.autoregressive_infer_cfg(B=8, label_B=labels, cfg=1.5, top_k=900, g_seed=0)

All arithmetic of the hot path runs in var_b200/libvar_b200.so (build: `python -m var_b200.build`).
"""
from __future__ import annotations

from typing import Tuple

from . import ops  # noqa: F401  (registers the var_b200:: torch.library ops)
from .quant import VectorQuantizer2
from .var import VAR
from .vqvae import VQVAE

__all__ = ["build_vae_var", "VAR", "VQVAE", "VectorQuantizer2"]


def build_vae_var(
    device, patch_nums=(1, 2, 3, 4, 5, 6, 8, 10, 13, 16),
    V=4096, Cvae=32, ch=160, share_quant_resi=4,
    num_classes=1000, depth=16, shared_aln=False, attn_l2_norm=True,
    flash_if_available=True, fused_if_available=True,
    init_adaln=0.5, init_adaln_gamma=1e-5, init_head=0.02, init_std=-1,
) -> Tuple[VQVAE, VAR]:
    """Drop-in for models/__init__.py:9-39 (heads = depth, width = 64*depth, dpr = 0.1*depth/24).
    Unlike the reference this does not monkey-patch reset_parameters process-wide."""
    import torch
    heads, width, dpr = depth, depth * 64, 0.1 * depth / 24
    with torch.device(device):  # parameters are created (and initialised) directly on the target device
        return _build(device, patch_nums, V, Cvae, ch, share_quant_resi, num_classes, depth, shared_aln, attn_l2_norm,
                      flash_if_available, fused_if_available, init_adaln, init_adaln_gamma, init_head, init_std, heads,
                      width, dpr)


def _build(device, patch_nums, V, Cvae, ch, share_quant_resi, num_classes, depth, shared_aln, attn_l2_norm,
           flash_if_available, fused_if_available, init_adaln, init_adaln_gamma, init_head, init_std, heads, width, dpr):
    vae_local = VQVAE(vocab_size=V, z_channels=Cvae, ch=ch, test_mode=True, share_quant_resi=share_quant_resi,
                      v_patch_nums=patch_nums).to(device)
    var_wo_ddp = VAR(vae_local=vae_local, num_classes=num_classes, depth=depth, embed_dim=width, num_heads=heads,
                     drop_rate=0., attn_drop_rate=0., drop_path_rate=dpr, norm_eps=1e-6, shared_aln=shared_aln,
                     cond_drop_rate=0.1, attn_l2_norm=attn_l2_norm, patch_nums=patch_nums,
                     flash_if_available=flash_if_available, fused_if_available=fused_if_available).to(device)
    var_wo_ddp.init_weights(init_adaln=init_adaln, init_adaln_gamma=init_adaln_gamma, init_head=init_head,
                            init_std=init_std)
    return vae_local, var_wo_ddp
