"""Shared fixtures for the parity tests: seeded models, oracle adapters, golden loaders."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from oracle.quant_oracle import QuantOracle
from oracle.var_oracle import VarCfg
from var_b200 import build_vae_var
from var_b200.init_utils import dense_init_

GOLDEN = Path(__file__).resolve().parent / "golden"
PATCH_NUMS = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)


def golden(name: str):
    return np.load(GOLDEN / name)


def replay_noise_more_smooth(seed: int, B: int, V: int = 4096, patch_nums=PATCH_NUMS, device="cpu"):
    """Generator use of autoregressive_infer_cfg(more_smooth=True): per scale the sampler's Exp(1) draw, then the
    Gumbel draw of gumbel_softmax_with_rng (helpers.py:26). Returns (q list, g list)."""
    g = torch.Generator(device=device).manual_seed(seed)
    qs, gs = [], []
    for pn in patch_nums:
        qs.append(torch.empty(B * pn * pn, V, device=device).exponential_(1, generator=g))
        gs.append(torch.empty(B * pn * pn, V, device=device).exponential_(1, generator=g))
    return qs, gs


def split_scales(flat_BL: np.ndarray, patch_hws=None):
    hws = patch_hws or [(p, p) for p in PATCH_NUMS]
    out, off = [], 0
    for h, w in hws:
        out.append(np.ascontiguousarray(flat_BL[:, off:off + h * w]).astype(np.int64))
        off += h * w
    return out


_cache = {}


def seeded_models(depth: int = 2, shared_aln: bool = False, device: str = "cpu", vae_seed: int = 1, var_seed: int = 2):
    """Our modules with the same dense init the golden generator applied to the reference (oracle/gen_golden.py)."""
    key = (depth, shared_aln, device, vae_seed, var_seed)
    if key not in _cache:
        vae, var = build_vae_var(device="cpu", depth=depth, shared_aln=shared_aln)
        dense_init_(vae, seed=vae_seed)
        dense_init_(var, seed=var_seed)
        var.eval(); vae.eval(); var.cond_drop_rate = 0
        if device != "cpu":
            vae, var = vae.to(device), var.to(device)
        _cache[key] = (vae, var)
    return _cache[key]


def quant_oracle_of(vae) -> QuantOracle:
    q = vae.quantize
    phis = q.quant_resi.phis()
    return QuantOracle(q.embedding.weight.detach().cpu().numpy(),
                       np.stack([p.weight.detach().cpu().numpy() for p in phis]),
                       np.stack([p.bias.detach().cpu().numpy() for p in phis]), q.v_patch_nums, resi=abs(q.quant_resi_ratio))


def var_cfg_of(var) -> VarCfg:
    return VarCfg(depth=var.depth, patch_nums=tuple(var.patch_nums), num_classes=var.num_classes, V=var.V, Cvae=var.Cvae,
                  shared_aln=var.shared_aln, attn_l2_norm=bool(var.blocks[0].attn.attn_l2_norm))


def sd_cpu(var):
    return {k: v.detach().cpu().float() if v.is_floating_point() else v.detach().cpu() for k, v in var.state_dict().items()}


def replay_noise(seed: int, B: int, V: int = 4096, patch_nums=PATCH_NUMS, device="cpu", skip_scales=()):
    """The Exp(1) draws torch.multinomial makes inside sample_with_top_k_top_p_. Scales listed in
    skip_scales draw nothing (VAR.inpainting skips the sampler where every token is kept, var.py:313-314)."""
    g = torch.Generator(device=device).manual_seed(seed)
    return [None if si in skip_scales else torch.empty(B * pn * pn, V, device=device).exponential_(1, generator=g)
            for si, pn in enumerate(patch_nums)]


def embed_inputs(B: int = 2, seed: int = 5, patch_nums=PATCH_NUMS):
    """Arbitrary per-scale maps for embed_to_fhat (shared by oracle/gen_golden_embed.py and the tests)."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, 32, p, p, generator=g) for p in patch_nums]


def logits_inputs(C: int, B: int = 2, l: int = 9, seed: int = 6):
    """(h [B,l,C], labels) for get_logits (shared by oracle/gen_golden_embed.py and the tests)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, l, C, generator=g), torch.tensor([7, 1000])
