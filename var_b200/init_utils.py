"""Seeded dense initialisation shared by the parity harness and the benchmark.

The reference's default init makes the transformer body numerically invisible (adaLN gamma ~ 1e-7) and leaves the
VQVAE convolutions uninitialised, so parity on it would test nothing. `dense_init_` draws every
parameter from a per-name seeded normal so that any two modules with the same state_dict keys (ours, the
reference's) receive identical values, independent of parameter registration order.
"""
from __future__ import annotations

import math
import zlib

import torch
import torch.nn as nn


def _gen(name: str, seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _normal_(p: torch.Tensor, std: float, g: torch.Generator, mean: float = 0.0) -> None:
    with torch.no_grad():
        tmp = torch.empty(p.shape, dtype=torch.float32, device=p.device)
        tmp.normal_(mean, std, generator=g)
        p.copy_(tmp)


@torch.no_grad()
def dense_init_(module: nn.Module, seed: int = 0) -> nn.Module:
    """In-place dense init of a VAR or VQVAE module (ours or the reference's). Recipe: SURVEY.md Appendix B."""
    for name, p in sorted(module.named_parameters(), key=lambda kv: kv[0]):
        g = _gen(name, seed, p.device)
        leaf = name.rsplit(".", 1)[-1]
        if p.ndim == 4:  # conv weight [out, in, kh, kw]
            _normal_(p, 1.0 / math.sqrt(p.shape[1] * p.shape[2] * p.shape[3]), g)
        elif "norm" in name and p.ndim == 1:  # GroupNorm affine
            p.fill_(1.0 if leaf == "weight" else 0.0)
        elif name.endswith("embedding.weight") or name.endswith("class_emb.weight"):
            _normal_(p, 1.0, g)
        elif "ada_lin" in name and leaf == "weight":
            _normal_(p, 0.5 / math.sqrt(p.shape[1]), g)
        elif "ada_lin" in name and leaf == "bias":
            _normal_(p, 0.3, g)
        elif leaf == "ada_gss":
            _normal_(p, 0.3, g)
        elif leaf in ("q_bias", "v_bias"):
            _normal_(p, 0.1, g)
        elif leaf == "scale_mul_1H11":
            _normal_(p, 0.3, g, mean=math.log(4.0))
        elif leaf in ("pos_start", "pos_1LC") or name.endswith("lvl_embed.weight"):
            _normal_(p, 0.5, g)
        elif p.ndim == 2:  # linear weight [out, in]
            _normal_(p, 1.0 / math.sqrt(p.shape[1]), g)
        elif p.ndim == 1:  # remaining biases
            _normal_(p, 0.02, g)
        else:
            _normal_(p, 0.02, g)
    return module
