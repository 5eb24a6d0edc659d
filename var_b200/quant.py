"""Host-side mirror of VectorQuantizer2 (/root/reference/models/quant.py) over the sm_100a quantizer kernels.

Same method names, argument meaning and return types as the reference inference methods; the training `forward`
(EMA / all_reduce, quant.py:52-104) is out of scope. Parameters keep the reference's state_dict keys
(`embedding.weight`, `quant_resi.qresi_ls.{k}.weight|bias`, buffer `ema_vocab_hit_SV`).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import lib as L


class Phi(nn.Conv2d):
    """3x3 residual mixing conv (quant.py:199-206); applied inside the CUDA kernel, never called as a module."""

    def __init__(self, embed_dim: int, quant_resi: float):
        super().__init__(embed_dim, embed_dim, 3, 1, 1)
        self.resi_ratio = abs(quant_resi)


class _PhiBank(nn.Module):
    """PhiShared / PhiPartiallyShared / PhiNonShared selector (quant.py:209-243) with identical parameter names."""

    def __init__(self, phis: Sequence[Phi], mode: str):
        super().__init__()
        self.mode = mode
        if mode == "shared":
            self.qresi = phis[0]
        elif mode == "partial":
            self.qresi_ls = nn.ModuleList(phis)
        else:  # non-shared: the reference subclasses ModuleList, so keys are "0.weight", ...
            for i, p in enumerate(phis):
                self.add_module(str(i), p)
        self.n = len(phis)
        K = self.n
        self.ticks = (np.linspace(1 / 3 / K, 1 - 1 / 3 / K, K) if K == 4 else np.linspace(1 / 2 / K, 1 - 1 / 2 / K, K))

    def phis(self) -> List[Phi]:
        if self.mode == "shared":
            return [self.qresi]
        if self.mode == "partial":
            return list(self.qresi_ls)
        return [getattr(self, str(i)) for i in range(self.n)]

    def index(self, at_from_0_to_1: float) -> int:
        if self.mode == "shared":
            return 0
        return int(np.argmin(np.abs(self.ticks - at_from_0_to_1)).item())

    def __getitem__(self, at_from_0_to_1: float) -> Phi:
        return self.phis()[self.index(at_from_0_to_1)]


def _hw(pn) -> Tuple[int, int]:
    return (pn, pn) if isinstance(pn, int) else (int(pn[0]), int(pn[1]))


class VectorQuantizer2(nn.Module):
    def __init__(self, vocab_size, Cvae, using_znorm, beta: float = 0.25, default_qresi_counts=0, v_patch_nums=None,
                 quant_resi=0.5, share_quant_resi=4):
        super().__init__()
        if using_znorm:
            raise NotImplementedError("using_znorm=True (cosine search, quant.py:151-153) is not on the hot path")
        if abs(quant_resi) <= 1e-6:
            raise NotImplementedError("quant_resi=0 (Phi = Identity) is not supported")
        self.vocab_size, self.Cvae, self.using_znorm = vocab_size, Cvae, using_znorm
        self.v_patch_nums = tuple(v_patch_nums)
        self.quant_resi_ratio = quant_resi
        if share_quant_resi == 0:
            n, mode = default_qresi_counts or len(self.v_patch_nums), "nonshared"
        elif share_quant_resi == 1:
            n, mode = 1, "shared"
        else:
            n, mode = share_quant_resi, "partial"
        self.quant_resi = _PhiBank([Phi(Cvae, quant_resi) for _ in range(n)], mode)
        self.register_buffer("ema_vocab_hit_SV", torch.full((len(self.v_patch_nums), vocab_size), fill_value=0.0))
        self.beta = beta
        self.embedding = nn.Embedding(vocab_size, Cvae)
        self.prog_si = -1
        # 0: tensor-core distance filter + exact fp32 re-rank (default); 1: fused fp32 CUDA-core search. Same indices.
        self.search_mode = 0
        self._pack_key = None
        self._packed = None

    def forward(self, *a, **k):
        raise NotImplementedError("VectorQuantizer2.forward is VAE training (out of scope, SURVEY.md §2 row 3)")

    def _device(self) -> torch.device:
        return self.embedding.weight.device

    def codebook_distances(self):
        """(dists [V,V] fp32, order [V,V] int32): pairwise L2 distances of the code vectors and, per code, all codes
        by increasing distance (models/var.py:459-462, var_analysis.py:256). Cached per codebook version."""
        w = self.embedding.weight
        key = (w._version, w.data_ptr())
        if getattr(self, "_dist_key", None) != key:
            E = w.detach().float()
            d = torch.cdist(E, E, p=2).contiguous()
            self._dist = (d, torch.argsort(d, dim=1).to(torch.int32).contiguous())
            self._dist_key = key
        return self._dist

    def _check_tokens(self, idx: torch.Tensor, V: int):
        if idx.numel():  # the reference's embedding lookup raises / device-asserts on a bad token
            lo, hi = (int(v) for v in torch.aminmax(idx))
            if lo < 0 or hi >= V:
                raise IndexError(f"token index out of range [0, {V}) (embedding lookup, models/quant.py:180)")

    # ------------------------------------------------------------------ packing
    def _weights(self):
        ps = self.quant_resi.phis()
        key = (self.embedding.weight._version, self.embedding.weight.data_ptr(),
               tuple((p.weight._version, p.weight.data_ptr(), p.bias._version) for p in ps))
        if key != self._pack_key:
            w = torch.stack([p.weight.detach().float() for p in ps]).contiguous()
            b = torch.stack([p.bias.detach().float() for p in ps]).contiguous()
            self._packed = (self.embedding.weight.detach().float().contiguous(), w, b)
            self._pack_key = key
        return self._packed

    def _qargs(self, patch_hws: Sequence[Tuple[int, int]]):
        """(codebook, phi_w, phi_b, ph, pw, phi_of_scale, resi): the quantizer arguments of the var_b200:: custom ops."""
        cb, w, b = self._weights()
        if not cb.is_cuda:
            raise L.VarB200Error("var_b200 quantizer needs its parameters on a CUDA device (no CPU path)")
        S = len(patch_hws)
        if S > L.MAX_SCALES:
            raise ValueError(f"at most {L.MAX_SCALES} scales are supported, got {S}")
        phi_of = [self.quant_resi.index(i / (S - 1) if S > 1 else 0.0) for i in range(S)]
        return cb, w, b, [h for h, _ in patch_hws], [wd for _, wd in patch_hws], phi_of, abs(self.quant_resi_ratio)

    # ------------------------------------------------------------------ reference API
    @L.device_guard
    def f_to_idxBl_or_fhat(self, f_BChw: torch.Tensor, to_fhat: bool,
                           v_patch_nums: Optional[Sequence[Union[int, Tuple[int, int]]]] = None):
        """quant.py:135-166."""
        B, Cc, H, W = f_BChw.shape
        patch_hws = [_hw(pn) for pn in (v_patch_nums or self.v_patch_nums)]
        assert patch_hws[-1][0] == H and patch_hws[-1][1] == W, f'{patch_hws[-1]=} != ({H=}, {W=})'
        f = f_BChw.detach().float().contiguous()
        idx, fh = torch.ops.var_b200.quant_encode(f, *self._qargs(patch_hws), bool(to_fhat), int(self.search_mode))
        if to_fhat:
            return [fh[i] for i in range(len(patch_hws))]
        out, off = [], 0
        for h, w in patch_hws:
            out.append(idx[off:off + B * h * w].view(B, h * w))
            off += B * h * w
        return out

    def _flat_idx(self, ms_idx_Bl: List[torch.Tensor]) -> torch.Tensor:
        return torch.cat([t.reshape(-1).to(torch.int64) for t in ms_idx_Bl]).contiguous()

    @L.device_guard
    def _decode(self, ms_idx_Bl, want_input: bool, want_list: bool):
        B = ms_idx_Bl[0].shape[0]
        hws = [_hw(pn) for pn in self.v_patch_nums]
        assert len(ms_idx_Bl) == len(hws)
        idx = self._flat_idx(ms_idx_Bl)
        self._check_tokens(idx, self.vocab_size)
        vin, fl, last = torch.ops.var_b200.quant_decode(idx, B, *self._qargs(hws), bool(want_input), bool(want_list))
        return (vin if want_input else None), (fl if want_list else None), last

    def idxBl_to_var_input(self, gt_ms_idx_Bl: List[torch.Tensor]) -> torch.Tensor:
        """quant.py:169-184 -> [B, L - first_l, Cvae] fp32."""
        if len(self.v_patch_nums) == 1:
            return None
        return self._decode(gt_ms_idx_Bl, True, False)[0]

    def idxBl_to_fhat(self, ms_idx_Bl: List[torch.Tensor], last_one: bool = False):
        """embed_to_fhat(all_to_max_scale=True) on codebook lookups (vqvae.py:77-84 + quant.py:107-121)."""
        _, fl, last = self._decode(ms_idx_Bl, False, not last_one)
        return last if last_one else [fl[i] for i in range(fl.shape[0])]

    @L.device_guard
    def embed_to_fhat(self, ms_h_BChw: List[torch.Tensor], all_to_max_scale=True, last_one=False):
        """quant.py:107-133 on arbitrary per-scale maps h_si [B, Cvae, ph, pw]. The decode kernel reads them through a
        "virtual codebook" (row = one (scale, image, position) vector, identity indices), so the arithmetic is the
        one `idxBl_to_fhat` runs on codebook lookups."""
        if not all_to_max_scale:
            raise NotImplementedError("all_to_max_scale=False is the reference's experimental path (quant.py:122-131)")
        hws = [_hw(pn) for pn in self.v_patch_nums]
        assert len(ms_h_BChw) == len(hws)
        B, dev = ms_h_BChw[0].shape[0], ms_h_BChw[0].device
        for h, (ph, pw) in zip(ms_h_BChw, hws):
            assert tuple(h.shape) == (B, self.Cvae, ph, pw), f"{tuple(h.shape)=} != {(B, self.Cvae, ph, pw)}"
        rows = torch.cat([h.detach().float().reshape(B, self.Cvae, -1).transpose(1, 2).reshape(-1, self.Cvae)
                          for h in ms_h_BChw]).contiguous()
        _, w, b, ph, pw, phi_of, resi = self._qargs(hws)
        idx = torch.arange(rows.shape[0], device=dev, dtype=torch.int64)
        _, fl, last = torch.ops.var_b200.quant_decode(idx, B, rows, w, b, ph, pw, phi_of, resi, False, not last_one)
        return last if last_one else [fl[i] for i in range(fl.shape[0])]

    @L.device_guard
    def get_next_autoregressive_input(self, si: int, SN: int, f_hat: torch.Tensor, h_BChw: torch.Tensor = None, *,
                                      idx_Bl: torch.Tensor = None, token_major: bool = False):
        """quant.py:187-196. The kernel path takes the sampled indices (h = embedding[idx], models/var.py:177,182);
        f_hat is updated in place. Returns (f_hat, next) with next = area(f_hat) as NCHW (reference layout) or
        [B, l_next, Cvae] when token_major."""
        hws = [_hw(pn) for pn in self.v_patch_nums]
        assert SN == len(hws)
        cb, w, b, ph, pw, phi_of, resi = self._qargs(hws)
        B = f_hat.shape[0]
        assert f_hat.dtype == torch.float32 and f_hat.is_contiguous()
        if idx_Bl is None:
            # arbitrary h map (the more_smooth soft embeddings, var.py:178-182): the same kernel reads h through a
            # "virtual codebook" whose row b*l + t is h[b, :, t] and an identity index
            if h_BChw is None:
                raise ValueError("get_next_autoregressive_input needs h_BChw or idx_Bl")
            cb = h_BChw.detach().float().reshape(B, self.Cvae, ph[si] * pw[si]).transpose(1, 2).reshape(-1, self.Cvae).contiguous()
            idx_Bl = torch.arange(B * ph[si] * pw[si], device=f_hat.device, dtype=torch.int64).view(B, ph[si] * pw[si])
        idx = idx_Bl.to(torch.int64).contiguous()
        nxt = torch.ops.var_b200.quant_next_input(si, f_hat, idx, cb, w, b, ph, pw, phi_of, resi, bool(token_major))
        return f_hat, (nxt if si != SN - 1 else f_hat)
