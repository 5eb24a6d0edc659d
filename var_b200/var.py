"""Host-side mirror of VAR (/root/reference/models/var.py, models/basic_var.py) over the sm_100a kernels.

Same constructor, public methods (`forward`, `autoregressive_infer_cfg`, `get_logits`) and state_dict keys/shapes as
the reference, so `var_d{16,20,24,30,36}.pth` load with strict=True. The nn.Module tree only *holds* parameters;
all arithmetic runs in libvar_b200.so through the C-ABI (include/var_b200.h). There is no PyTorch fallback: calling
the model with CPU parameters, or without the built library, raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from . import lib as L
from .vqvae import VQVAE


# --------------------------------------------------------------------------------------------------
# parameter containers (names follow models/basic_var.py so the state_dict keys match)
# --------------------------------------------------------------------------------------------------
class _SelfAttention(nn.Module):
    def __init__(self, C: int, H: int, attn_l2_norm: bool):
        super().__init__()
        self.num_heads, self.head_dim, self.attn_l2_norm = H, C // H, attn_l2_norm
        if attn_l2_norm:
            self.scale_mul_1H11 = nn.Parameter(torch.full((1, H, 1, 1), 4.0).log())
            self.max_scale_mul = math.log(100)
        self.mat_qkv = nn.Linear(C, 3 * C, bias=False)
        self.q_bias, self.v_bias = nn.Parameter(torch.zeros(C)), nn.Parameter(torch.zeros(C))
        self.register_buffer("zero_k_bias", torch.zeros(C))
        self.proj = nn.Linear(C, C)

    def kv_caching(self, enable: bool):  # the cache lives in VAR._kv (preallocated), kept for API compatibility
        pass


class _FFN(nn.Module):
    def __init__(self, C: int, hidden: int):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(C, hidden), nn.Linear(hidden, C)


class _AdaLNBlock(nn.Module):
    def __init__(self, C: int, D: int, H: int, mlp_ratio: float, shared_aln: bool, attn_l2_norm: bool):
        super().__init__()
        self.attn = _SelfAttention(C, H, attn_l2_norm)
        self.ffn = _FFN(C, round(C * mlp_ratio))
        self.shared_aln = shared_aln
        if shared_aln:
            self.ada_gss = nn.Parameter(torch.randn(1, 1, 6, C) / C ** 0.5)
        else:
            self.ada_lin = nn.Sequential(nn.SiLU(inplace=False), nn.Linear(D, 6 * C))


class _AdaLNBeforeHead(nn.Module):
    def __init__(self, C: int, D: int):
        super().__init__()
        self.ada_lin = nn.Sequential(nn.SiLU(inplace=False), nn.Linear(D, 2 * C))


# Measurement hook: bench.py arms it for the one step an `ncu --profile-from-start off` launch list should cover; the AR
# loop then calls cudaProfilerStart when it reaches scale VAR_B200_PROFILE_FROM_SCALE (default 0 = the whole step).
_PROFILE_FROM_SCALE = int(os.environ.get("VAR_B200_PROFILE_FROM_SCALE", "0"))
_PROFILE_ARMED = [False]


class VAR(nn.Module):
    def __init__(self, vae_local: VQVAE, num_classes=1000, depth=16, embed_dim=1024, num_heads=16, mlp_ratio=4.,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., norm_eps=1e-6, shared_aln=False, cond_drop_rate=0.1,
                 attn_l2_norm=False, patch_nums=(1, 2, 3, 4, 5, 6, 8, 10, 13, 16), flash_if_available=True,
                 fused_if_available=True):
        super().__init__()
        assert embed_dim % num_heads == 0
        if embed_dim // num_heads != 64:
            raise NotImplementedError("the attention kernel is specialised for head_dim 64 (models/__init__.py:19-20)")
        if mlp_ratio != 4.:
            raise NotImplementedError("mlp_ratio must be 4")
        if len(patch_nums) > L.MAX_SCALES:
            raise ValueError(f"at most {L.MAX_SCALES} scales")
        self.Cvae, self.V = vae_local.Cvae, vae_local.vocab_size
        self.depth, self.C, self.D, self.num_heads = depth, embed_dim, embed_dim, num_heads
        self.cond_drop_rate = cond_drop_rate
        self.drop_path_rate = drop_path_rate
        self.prog_si = -1
        self.norm_eps = norm_eps
        self.shared_aln = shared_aln
        self.patch_nums: Tuple[int, ...] = tuple(patch_nums)
        self.L = sum(pn ** 2 for pn in self.patch_nums)
        self.first_l = self.patch_nums[0] ** 2
        self.begin_ends, cur = [], 0
        for pn in self.patch_nums:
            self.begin_ends.append((cur, cur + pn * pn))
            cur += pn * pn
        self.num_stages_minus_1 = len(self.patch_nums) - 1
        self._rng: Optional[torch.Generator] = None

        self.vae_proxy: Tuple[VQVAE] = (vae_local,)
        self.vae_quant_proxy = (vae_local.quantize,)
        self.word_embed = nn.Linear(self.Cvae, self.C)

        init_std = math.sqrt(1 / self.C / 3)
        self.num_classes = num_classes
        self.class_emb = nn.Embedding(num_classes + 1, self.C)
        nn.init.trunc_normal_(self.class_emb.weight.data, mean=0, std=init_std)
        self.pos_start = nn.Parameter(torch.empty(1, self.first_l, self.C))
        nn.init.trunc_normal_(self.pos_start.data, mean=0, std=init_std)
        self.pos_1LC = nn.Parameter(torch.empty(1, self.L, self.C))
        nn.init.trunc_normal_(self.pos_1LC.data, mean=0, std=init_std)
        self.lvl_embed = nn.Embedding(len(self.patch_nums), self.C)
        nn.init.trunc_normal_(self.lvl_embed.weight.data, mean=0, std=init_std)

        self.shared_ada_lin = (nn.Sequential(nn.SiLU(inplace=False), nn.Linear(self.D, 6 * self.C)) if shared_aln
                               else nn.Identity())
        self.blocks = nn.ModuleList(_AdaLNBlock(self.C, self.D, num_heads, mlp_ratio, shared_aln, attn_l2_norm)
                                    for _ in range(depth))

        d = torch.cat([torch.full((pn * pn,), i) for i, pn in enumerate(self.patch_nums)]).view(1, self.L, 1)
        dT = d.transpose(1, 2)
        self.register_buffer("lvl_1L", dT[:, 0].contiguous())
        # kept only for state_dict compatibility (var.py:111-112): the kernels compute the mask from level ends
        self.register_buffer("attn_bias_for_masking",
                             torch.where(d >= dT, 0., -torch.inf).reshape(1, 1, self.L, self.L).contiguous())

        self.head_nm = _AdaLNBeforeHead(self.C, self.D)
        self.head = nn.Linear(self.C, self.V)
        self._pack_key = None
        self._packed = None

    # ------------------------------------------------------------------ reference-compatible helpers
    @property
    def rng(self) -> torch.Generator:
        dev = self.lvl_1L.device
        if self._rng is None or self._rng.device != dev:
            self._rng = torch.Generator(device=dev)
        return self._rng

    def init_weights(self, init_adaln=0.5, init_adaln_gamma=1e-5, init_head=0.02, init_std=0.02, conv_std_or_gain=0.02):
        """models/var.py:577-627 (same distributions; used by build_vae_var)."""
        if init_std < 0:
            init_std = (1 / self.C / 3) ** 0.5
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight.data, std=init_std)
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.Embedding):
                nn.init.trunc_normal_(m.weight.data, std=init_std)
        if init_head >= 0:
            self.head.weight.data.mul_(init_head)
            self.head.bias.data.zero_()
        self.head_nm.ada_lin[-1].weight.data.mul_(init_adaln)
        self.head_nm.ada_lin[-1].bias.data.zero_()
        depth = len(self.blocks)
        for b in self.blocks:
            b.attn.proj.weight.data.div_(math.sqrt(2 * depth))
            b.ffn.fc2.weight.data.div_(math.sqrt(2 * depth))
            if hasattr(b, "ada_lin"):
                b.ada_lin[-1].weight.data[2 * self.C:].mul_(init_adaln)
                b.ada_lin[-1].weight.data[:2 * self.C].mul_(init_adaln_gamma)
                b.ada_lin[-1].bias.data.zero_()
            else:
                b.ada_gss.data[:, :, 2:].mul_(init_adaln)
                b.ada_gss.data[:, :, :2].mul_(init_adaln_gamma)

    # ------------------------------------------------------------------ weight packing for the C-ABI
    def _version_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def repack(self):
        self._pack_key = None

    def _model(self) -> "PackedModel":
        key = self._version_key()
        if key != self._pack_key:
            self._packed = PackedModel(self)
            self._pack_key = key
        return self._packed

    def _device(self) -> torch.device:
        return self.lvl_1L.device

    # ------------------------------------------------------------------ forward passes
    def _labels_i32(self, label_B: torch.Tensor, n: int) -> torch.Tensor:
        if label_B.numel():  # the reference's embedding lookup raises (CPU) / device-asserts (CUDA) on a bad label
            lo, hi = (int(v) for v in torch.aminmax(label_B.detach().reshape(-1)))
            if lo < 0 or hi > self.num_classes:
                raise IndexError(f"class label out of range [0, {self.num_classes}] (class_emb, models/var.py:61)")
        lab = label_B.reshape(-1).to(device=self.lvl_1L.device, dtype=torch.int32)
        if lab.numel() == 1 and n > 1:
            lab = lab.expand(n)
        if lab.numel() != n:
            raise RuntimeError(f"label_B has {lab.numel()} entries but the batch is {n} (models/var.py:199-203)")
        return lab.contiguous()

    @torch.no_grad()
    @L.device_guard
    def get_logits(self, h_or_h_and_residual, cond_BD=None, *, labels: Optional[torch.Tensor] = None):
        """models/var.py:118-124: head(head_nm(h.float(), cond_BD)) -> fp32 logits [B, l, V]. The head's (scale, shift)
        come from `cond_BD` (SiLU -> Linear on the tcgen05 GEMM) or, keyword-only, from `labels`
        (cond_BD = class_emb(labels), the table `forward` uses)."""
        if not isinstance(h_or_h_and_residual, torch.Tensor):  # (h, residual) pair of the fused add-norm path
            h, resi = h_or_h_and_residual
            h = resi + h  # drop_path is the identity at inference
        else:
            h = h_or_h_and_residual
        pm = self._model()
        B, l, _ = h.shape
        x = h.detach().to(self.lvl_1L.device).float().contiguous()
        if labels is not None:
            ada = pm.ada_params(self._labels_i32(labels, B))
        elif cond_BD is not None:
            ada = pm.head_ada_from_cond(cond_BD.reshape(B, self.D))
        else:
            raise ValueError("get_logits needs cond_BD or labels")
        return pm.head_logits(x, ada, B, l)

    @torch.no_grad()
    @L.device_guard
    def forward(self, label_B: torch.LongTensor, x_BLCv_wo_first_l: torch.Tensor, *, return_blocks: bool = False):
        """models/var.py:192-234: teacher-forced logits [B, L, V] fp32."""
        if self.prog_si >= 0:
            raise NotImplementedError("progressive training (prog_si >= 0) is not supported (inference hot path)")
        if self.training and self.drop_path_rate > 0:
            raise RuntimeError("var_b200 implements the inference path: call .eval() first (drop_path is training only)")
        pm = self._model()
        B = x_BLCv_wo_first_l.shape[0]
        dev = self.lvl_1L.device
        if tuple(x_BLCv_wo_first_l.shape) != (B, self.L - self.first_l, self.Cvae):
            raise RuntimeError(f"x_BLCv_wo_first_l must be [B, {self.L - self.first_l}, {self.Cvae}], got "
                               f"{tuple(x_BLCv_wo_first_l.shape)} (models/var.py:204-207)")
        # same RNG side effect and label dropout as the reference (var.py:201)
        drop = torch.rand(B, device=dev) < self.cond_drop_rate
        labels = self._labels_i32(label_B, B)
        labels = torch.where(drop, self.num_classes, labels).to(torch.int32).contiguous()
        x_in = x_BLCv_wo_first_l.to(dev).float().contiguous()
        x = pm.embed(x_in, B, labels, B, self.L, self.first_l, 0)
        ada = pm.ada_params(labels)
        dump = torch.empty((self.depth, B * self.L, self.C), device=dev) if return_blocks else None
        pm.blocks_teacher(x, ada, B, dump, labels)
        logits = pm.head_logits(x, ada, B, self.L)
        if return_blocks:
            return logits, [dump[i].view(B, self.L, self.C) for i in range(self.depth)]
        return logits

    @torch.no_grad()
    @L.device_guard
    def autoregressive_infer_cfg(self, B: int, label_B: Optional[Union[int, torch.LongTensor]], g_seed: Optional[int] = None,
                                 cfg=1.5, top_k=0, top_p=0.0, more_smooth=False, *, forced_idx=None, return_trace=False,
                                 decode=True, cuda_graph=False):
        """models/var.py:126-190: KV-cached CFG sampling; returns images [B,3,H,W] in [0,1].
        Extras (keyword-only): forced_idx = tokens to feed instead of the sampled ones and return_trace = also return
        dict(idx, logits, f_hat) (parity harness); decode=False skips the CNN decoder (returns f_hat);
        cuda_graph=True replays the whole 10-scale loop (~130 launches per scale and block) as one captured CUDA graph
        per (B, cfg, top_k, top_p) — same tokens as the eager path for the same seed."""
        dev = self.lvl_1L.device
        if more_smooth and (cuda_graph or forced_idx is not None):
            raise ValueError("more_smooth does not combine with cuda_graph / forced_idx")
        if g_seed is None:
            rng = None
        else:
            self.rng.manual_seed(g_seed)
            rng = self.rng
        if label_B is None:
            uniform = torch.full((1, self.num_classes), 1.0 / self.num_classes, dtype=torch.float32, device=dev)
            label_B = torch.multinomial(uniform, num_samples=B, replacement=True, generator=rng).reshape(B)
        elif isinstance(label_B, int):
            label_B = torch.full((B,), fill_value=self.num_classes if label_B < 0 else label_B, device=dev)
        label_B = label_B.to(dev)
        labels = self._labels_i32(torch.cat((label_B, torch.full_like(label_B, self.num_classes))), 2 * B)
        if cuda_graph:
            if forced_idx is not None or return_trace:
                raise ValueError("cuda_graph=True does not support forced_idx / return_trace")
            f_hat = self._ar_graph_replay(B, labels, rng, float(cfg), int(top_k), float(top_p), g_seed)
            trace = None
        else:
            trace = dict(idx=[], logits=[]) if return_trace else None
            f_hat = self._ar_loop(B, labels, rng, cfg, top_k, top_p, forced_idx, trace, more_smooth=more_smooth)
        if return_trace:
            trace["f_hat"] = f_hat
        img = self.vae_proxy[0].fhat_to_img(f_hat).add_(1).mul_(0.5) if decode else f_hat
        return (img, trace) if return_trace else img

    @torch.no_grad()
    @L.device_guard
    def inpainting(self, gt_tokens: torch.Tensor, mask: torch.Tensor, label: Optional[Union[int, torch.LongTensor]] = None,
                   g_seed: Optional[int] = None, cfg: float = 1.5, top_k: int = 0, top_p: float = 0.0,
                   more_smooth: bool = False, *, forced_idx=None, return_trace=False, decode=True):
        """models/var.py:236-364: KV-cached CFG sampling that keeps the tokens where `mask` [B, L] is True (they are
        taken from `gt_tokens` [B, L], the output of vae.img_to_idxBl concatenated over scales) and samples the rest;
        a scale whose tokens are all kept runs the blocks (its K/V must enter the cache) but neither the head nor the
        sampler, and draws no noise (var.py:313-314). Returns images [B,3,H,W] in [0,1]. Keyword extras as in
        autoregressive_infer_cfg."""
        if mask.shape != gt_tokens.shape:
            raise ValueError("Mask shape must match the latent token shape obtained from vae.img_to_idxBl")
        if more_smooth:
            raise NotImplementedError("more_smooth inside inpainting (var.py:333-341) reads the logits of a possibly skipped "
                                      "scale in the reference; use autoregressive_infer_cfg(more_smooth=True)")
        dev = self.lvl_1L.device
        B = gt_tokens.shape[0]
        if gt_tokens.shape[1] != self.L:
            raise ValueError(f"gt_tokens must hold {self.L} tokens per image, got {gt_tokens.shape[1]}")
        if label is None:  # var.py:274: the reference draws from the global generator here
            uniform = torch.full((1, self.num_classes), 1.0 / self.num_classes, dtype=torch.float32, device=dev)
            label = torch.multinomial(uniform, num_samples=B, replacement=True).reshape(B)
        elif isinstance(label, int):
            label = torch.full((B,), fill_value=label, device=dev)
        label = label.to(dev)
        if g_seed is None:
            rng = None
        else:
            self.rng.manual_seed(g_seed)
            rng = self.rng
        labels = self._labels_i32(torch.cat((label, torch.full_like(label, self.num_classes))), 2 * B)
        gt = gt_tokens.to(dev).to(torch.int64).contiguous()
        keep = mask.to(dev).bool().contiguous()
        if int(gt.min()) < 0 or int(gt.max()) >= self.V:
            raise ValueError(f"gt_tokens out of range [0, {self.V})")
        trace = dict(idx=[], logits=[]) if return_trace else None
        f_hat = self._ar_loop(B, labels, rng, cfg, top_k, top_p, forced_idx, trace, gt, keep)
        if return_trace:
            trace["f_hat"] = f_hat
        img = self.vae_proxy[0].fhat_to_img(f_hat).add_(1).mul_(0.5) if decode else f_hat
        return (img, trace) if return_trace else img

    @torch.no_grad()
    @L.device_guard
    def smooth_sampling(self, gt_tokens: torch.Tensor, n: int, label: Optional[Union[int, torch.LongTensor]] = None,
                        g_seed: Optional[int] = None, cfg: float = 1.5, more_smooth: bool = False,
                        neighbor_threshold: Optional[float] = None, *, forced_idx=None, return_trace=False, decode=True):
        """models/var.py:366-575: KV-cached CFG loop in which every token is the arg-max of the mixed log-probabilities
        among the codebook neighbours of the ground-truth token (`1 + int((n-1)*ratio)` nearest, or those within
        d_min + (neighbor_threshold - d_min)*ratio). Returns (images [B,3,H,W] in [0,1], sum_log_likelihood,
        sum_distance_log_likelihood); as in the reference the first sum adds the log-probabilities truncated to
        integers (int64 `new_tensor`, var.py:536). No randomness is consumed (g_seed only reseeds self.rng)."""
        if more_smooth:
            raise NotImplementedError("more_smooth (Gumbel-softmax visualisation path, var.py:543-551) is out of scope")
        import ctypes as C
        from . import lib as L
        lib = L.load()
        dev = self.lvl_1L.device
        B = gt_tokens.shape[0]
        if gt_tokens.shape[1] != self.L:
            raise ValueError(f"gt_tokens must hold {self.L} tokens per image, got {gt_tokens.shape[1]}")
        if not 1 <= int(n) <= self.V:
            raise ValueError(f"n must be in [1, {self.V}]")
        if label is None:
            uniform = torch.full((1, self.num_classes), 1.0 / self.num_classes, dtype=torch.float32, device=dev)
            label = torch.multinomial(uniform, num_samples=B, replacement=True).reshape(B)
        elif isinstance(label, int):
            label = torch.full((B,), fill_value=label, device=dev)
        label = label.to(dev)
        if g_seed is not None:
            self.rng.manual_seed(g_seed)
        labels = self._labels_i32(torch.cat((label, torch.full_like(label, self.num_classes))), 2 * B)
        gt = gt_tokens.to(dev).to(torch.int32).contiguous()
        if int(gt.min()) < 0 or int(gt.max()) >= self.V:
            raise ValueError(f"gt_tokens out of range [0, {self.V})")
        quant = self.vae_quant_proxy[0]
        dists, order = quant.codebook_distances()                                      # var.py:459-462 (cached per codebook)
        neighbors = order[:, :n].contiguous()
        pm = self._model()
        S = len(self.patch_nums)
        ada = pm.ada_params(labels)
        kv = pm.kv_cache(2 * B)
        H = W = self.patch_nums[-1]
        f_hat = torch.zeros((B, self.Cvae, H, W), dtype=torch.float32, device=dev)
        sum_ll = torch.zeros((), dtype=torch.int64, device=dev)
        sum_dll = torch.zeros((), dtype=torch.float32, device=dev)
        trace = dict(idx=[], sel=[], logp=[], dlogp=[]) if return_trace else None
        cur, nxt = 0, None
        for si, pn in enumerate(self.patch_nums):
            l = pn * pn
            ratio = si / self.num_stages_minus_1 if self.num_stages_minus_1 > 0 else 0.0
            x = pm.embed(None, 0, labels, 2 * B, l, self.first_l, 0) if si == 0 else pm.embed(nxt, B, labels, 2 * B, l, 0, cur)
            pm.blocks_cached(x, ada, 2 * B, l, cur, kv, labels)
            logits = pm.head_logits(x, ada, 2 * B, l)
            idx = torch.empty((B, l), dtype=torch.int64, device=dev)
            lp = torch.empty((B, l), dtype=torch.float32, device=dev)
            dlp = torch.empty((B, l), dtype=torch.float32, device=dev)
            gt_seg = gt[:, cur:cur + l].contiguous()
            L.check(lib.var_b200_neighbor_select(
                logits.data_ptr(), B, l, self.V, float(cfg * ratio), gt_seg.data_ptr(), neighbors.data_ptr(), dists.data_ptr(),
                int(n), 1 + int((n - 1) * ratio), int(neighbor_threshold is not None),
                float(neighbor_threshold if neighbor_threshold is not None else 0.0), float(ratio), idx.data_ptr(),
                lp.data_ptr(), dlp.data_ptr(), L.current_stream()), "neighbor_select")
            sum_ll = sum_ll + lp.to(torch.int64).sum()
            sum_dll = sum_dll + dlp.sum()
            sel = idx
            if forced_idx is not None:
                idx = forced_idx[si].to(dev).to(torch.int64).contiguous()
            if trace is not None:
                trace["idx"].append(idx); trace["sel"].append(sel); trace["logp"].append(lp); trace["dlogp"].append(dlp)
            _, nxt = quant.get_next_autoregressive_input(si, S, f_hat, idx_Bl=idx, token_major=True)
            cur += l
        if trace is not None:
            trace["f_hat"] = f_hat
        img = self.vae_proxy[0].fhat_to_img(f_hat).add_(1).mul_(0.5) if decode else f_hat
        return (img, sum_ll, sum_dll, trace) if return_trace else (img, sum_ll, sum_dll)

    def _ar_loop(self, B, labels, rng, cfg, top_k, top_p, forced_idx=None, trace=None, gt_tokens=None, keep_mask=None,
                 more_smooth=False):
        """The 10 strictly sequential scale steps (var.py:160-187) on labels [2B] int32 (cond rows, then uncond).
        gt_tokens / keep_mask [B, L]: inpainting (var.py:303-328)."""
        pm = self._model()
        dev = self.lvl_1L.device
        quant = self.vae_quant_proxy[0]
        S = len(self.patch_nums)
        ada = pm.ada_params(labels)
        kv = pm.kv_cache(2 * B)
        H = W = self.patch_nums[-1]
        f_hat = torch.zeros((B, self.Cvae, H, W), dtype=torch.float32, device=dev)
        cur, nxt = 0, None
        all_kept = None
        if keep_mask is not None:  # one host read for the whole loop: which scales keep every token
            offs = [0]
            for p_ in self.patch_nums:
                offs.append(offs[-1] + p_ * p_)
            all_kept = torch.stack([keep_mask[:, a:b].all() for a, b in zip(offs[:-1], offs[1:])]).tolist()
        for si, pn in enumerate(self.patch_nums):
            l = pn * pn
            if _PROFILE_ARMED[0] and si == _PROFILE_FROM_SCALE:  # measurement hook (bench.py VAR_B200_PROFILE_STEP)
                _PROFILE_ARMED[0] = False
                torch.cuda.profiler.start()
            if si == 0:
                x = pm.embed(None, 0, labels, 2 * B, l, self.first_l, 0)
            else:
                x = pm.embed(nxt, B, labels, 2 * B, l, 0, cur)
            pm.blocks_cached(x, ada, 2 * B, l, cur, kv, labels)
            if all_kept is not None and all_kept[si]:
                idx, mixed = gt_tokens[:, cur:cur + l].contiguous(), None
            else:
                logits = pm.head_logits(x, ada, 2 * B, l)
                q = torch.empty((B * l, self.V), dtype=torch.float32, device=dev).exponential_(1.0, generator=rng)
                t = cfg * (si / self.num_stages_minus_1) if self.num_stages_minus_1 > 0 else 0.0
                mixed = torch.empty((B, l, self.V), dtype=torch.float32, device=dev) if trace is not None else None
                h_soft = None
                if more_smooth:  # var.py:178-180: soft embedding from Gumbel noise drawn after the sampler's noise
                    q2 = torch.empty((B * l, self.V), dtype=torch.float32, device=dev).exponential_(1.0, generator=rng)
                    ratio = si / self.num_stages_minus_1 if self.num_stages_minus_1 > 0 else 0.0
                    idx, h_soft = pm.sample_smooth(logits, B, l, t, q, top_k, top_p, q2, max(0.27 * (1 - ratio * 0.95), 0.005),
                                                   1 + ratio, quant.embedding.weight.detach().float().contiguous(), mixed)
                    if trace is not None:
                        trace.setdefault("h", []).append(h_soft)
                else:
                    idx = pm.sample(logits, B, l, t, q, top_k, top_p, mixed)
                if keep_mask is not None:
                    idx = torch.where(keep_mask[:, cur:cur + l], gt_tokens[:, cur:cur + l], idx).contiguous()
            if forced_idx is not None:
                idx = forced_idx[si].to(dev).to(torch.int64).contiguous()
            if trace is not None:
                trace["idx"].append(idx)
                trace["logits"].append(mixed)
            if more_smooth:
                _, nxt = quant.get_next_autoregressive_input(si, S, f_hat, h_soft.transpose(1, 2).reshape(B, self.Cvae, pn, pn),
                                                             token_major=True)
            else:
                _, nxt = quant.get_next_autoregressive_input(si, S, f_hat, idx_Bl=idx, token_major=True)
            cur += l
        return f_hat

    def _ar_graph_replay(self, B, labels, rng, cfg, top_k, top_p, g_seed=None):
        """Capture-once / replay of _ar_loop. Static inputs: the label buffer; the generator state is registered with
        the graph so a re-seeded generator drives the replayed Exp(1) draws.
        The captured kernels hold raw pointers into the packed model's shared workspaces (KV cache, block scratch) and
        its bf16 weight copies: the graphs therefore live ON the PackedModel (they die with a repack) and carry the
        workspace generation they were captured under; any later reallocation of a workspace (a call with another
        batch size) bumps `ws_gen`, which drops every captured graph before it could replay into freed memory."""
        pm = self._model()
        key = (B, cfg, top_k, top_p, rng is not None)
        if pm.graph_gen != pm.ws_gen:
            pm.graphs.clear()
        ent = pm.graphs.get(key)
        if ent is None:
            static_labels = labels.clone()
            cur_stream = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur_stream)
            with torch.cuda.stream(side):  # warm-up outside capture: workspaces, KV cache, kernel attributes
                self._ar_loop(B, static_labels, rng, cfg, top_k, top_p)
            cur_stream.wait_stream(side)
            torch.cuda.synchronize()
            if pm.graph_gen != pm.ws_gen:  # the warm-up (re)allocated a workspace: older graphs point into freed memory
                pm.graphs.clear()
            gen = pm.ws_gen
            graph = torch.cuda.CUDAGraph()
            if rng is not None:
                graph.register_generator_state(rng)
            with torch.cuda.graph(graph):
                f_hat = self._ar_loop(B, static_labels, rng, cfg, top_k, top_p)
            if pm.ws_gen != gen:
                raise L.VarB200Error("a workspace was reallocated during CUDA-graph capture")
            pm.graph_gen = gen
            ent = (graph, static_labels, f_hat)
            pm.graphs[key] = ent
        graph, static_labels, f_hat = ent
        static_labels.copy_(labels)
        if rng is not None:
            rng.manual_seed(g_seed)  # warm-up and capture advanced the generator: replay from the caller's seed
        graph.replay()
        return f_hat.clone()

    def extra_repr(self):
        return f"depth={self.depth}, C={self.C}, shared_aln={self.shared_aln}, drop_path_rate={self.drop_path_rate:g}"


# --------------------------------------------------------------------------------------------------
# packed weights + thin typed wrappers over the C-ABI
# --------------------------------------------------------------------------------------------------
class PackedModel:
    """bf16/fp32 copies of the parameters in the layout var_b200_model_t expects, plus call wrappers."""

    def __init__(self, var: VAR):
        p0 = var.head.weight
        if not p0.is_cuda:
            raise L.VarB200Error("var_b200.VAR needs its parameters on a CUDA device: there is no CPU fallback path")
        self.lib = L.load()
        self.dev = p0.device
        self.var_cfg = (var.depth, var.C, var.num_heads, var.V, var.Cvae, var.L, var.first_l)
        C_, depth = var.C, var.depth
        bf = lambda t: t.detach().to(torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().float().contiguous()
        keep: List[torch.Tensor] = []
        self.blocks_arr = (L.BlockWeights * depth)()
        # |q.k| <= per-head scale exp(min(scale_mul, ln 100)) (basic_var.py:101-105), a model constant: with the largest
        # one <= 43 the attention kernel runs its bounded-score variant, and log2(e) is folded into q_scale so that the
        # scores leave the tensor core as base-2 exponents (one host read at pack time)
        l2 = bool(var.blocks[0].attn.attn_l2_norm)
        if l2:
            max_score = float(torch.stack([b.attn.scale_mul_1H11.detach().clamp_max(b.attn.max_scale_mul).exp().max()
                                           for b in var.blocks]).max().item())
        else:  # attn_l2_norm=False (basic_var.py:72): unnormalised q, k -> unbounded scores, the general attention kernel
            max_score = 0.0
        q_log2 = 0.0 < max_score <= 43.0
        q_mul = math.log2(math.e) if q_log2 else 1.0
        for i, b in enumerate(var.blocks):
            a = b.attn
            ts = dict(
                w_qkv=bf(a.mat_qkv.weight),
                b_qkv=f32(torch.cat((a.q_bias, torch.zeros_like(a.q_bias), a.v_bias))),
                q_scale=(f32(a.scale_mul_1H11.clamp_max(a.max_scale_mul).exp().reshape(-1) * q_mul) if l2 else
                         torch.full((var.num_heads,), 0.25 / math.sqrt(a.head_dim), dtype=torch.float32, device=p0.device)),
                w_proj=bf(a.proj.weight), b_proj=f32(a.proj.bias),
                w_fc1=bf(b.ffn.fc1.weight), b_fc1=f32(b.ffn.fc1.bias),
                w_fc2=bf(b.ffn.fc2.weight), b_fc2=f32(b.ffn.fc2.bias))
            for k, t in ts.items():
                setattr(self.blocks_arr[i], k, t.data_ptr())
                keep.append(t)
        if var.shared_aln:
            w_ada = torch.cat((var.shared_ada_lin[1].weight, var.head_nm.ada_lin[1].weight))
            b_ada = torch.cat((var.shared_ada_lin[1].bias, var.head_nm.ada_lin[1].bias))
            gss = f32(torch.cat([b.ada_gss.reshape(1, 6 * C_) for b in var.blocks]))
        else:
            w_ada = torch.cat([b.ada_lin[1].weight for b in var.blocks] + [var.head_nm.ada_lin[1].weight])
            b_ada = torch.cat([b.ada_lin[1].bias for b in var.blocks] + [var.head_nm.ada_lin[1].bias])
            gss = None
        t = dict(w_ada=bf(w_ada), b_ada=f32(b_ada), w_head=bf(var.head.weight), b_head=f32(var.head.bias),
                 w_word=f32(var.word_embed.weight.t()), b_word=f32(var.word_embed.bias), class_emb=f32(var.class_emb.weight),
                 pos_start=f32(var.pos_start.reshape(var.first_l, C_)),
                 lvl_pos=f32(var.lvl_embed.weight[var.lvl_1L.reshape(-1)] + var.pos_1LC.reshape(var.L, C_)))
        m = L.ModelDesc()
        m.depth, m.C, m.H, m.V, m.Cvae = depth, C_, var.num_heads, var.V, var.Cvae
        m.n_scales, m.num_classes, m.shared_aln, m.norm_eps = len(var.patch_nums), var.num_classes, int(var.shared_aln), var.norm_eps
        for i, pn in enumerate(var.patch_nums):
            m.patch_nums[i] = pn
        m.blocks = C.cast(self.blocks_arr, C.POINTER(L.BlockWeights))
        for k, v in t.items():
            setattr(m, k, v.data_ptr())
            keep.append(v)
        m.ada_rows = w_ada.shape[0]
        m.attn_max_score, m.attn_q_log2, m.attn_no_l2norm = max_score, int(q_log2), int(not l2)
        m.ada_gss = gss.data_ptr() if gss is not None else None
        if gss is not None:
            keep.append(gss)
        self.m, self._keep = m, keep
        self.w_ada, self.b_ada = t["w_ada"], t["b_ada"]
        self.depth, self.C, self.H, self.V, self.Cvae, self.L, self.first_l = self.var_cfg
        self.ada_ld = self.lib.var_b200_ada_ld(C.byref(m))
        from . import ops
        self.handle = ops.register_model(self)  # what the var_b200:: custom ops take as `model`
        self._ws = {}
        self.ws_gen = 0       # bumped whenever a workspace buffer is (re)allocated
        self.graphs = {}      # captured AR loops (VAR._ar_graph_replay), valid while graph_gen == ws_gen
        self.graph_gen = 0
        self.ln_fused = False
        self._build_ln_tables(var)

    def _build_ln_tables(self, var: "VAR"):
        """Deferred LayerNorm (include/var_b200.h: var_b200_ln_tables): per block and class the vectors
        U = W (1 + scale_c), V = W shift_c + bias for the QKV and fc1 GEMMs, so that no LayerNorm pass runs in front of
        them (basic_var.py:157-158). 14 C floats per block and class (d30: 3.2 GB for the 1001 classes); skipped when
        VAR_B200_LNF=0 or the tables would exceed VAR_B200_LNF_MAX_GB (default 12): the blocks then run the LayerNorm
        pass as before."""
        import os
        n_cls = var.num_classes + 1
        per_block = n_cls * 14 * self.C
        if os.environ.get("VAR_B200_LNF", "1") == "0" or \
                per_block * self.depth * 4 > float(os.environ.get("VAR_B200_LNF_MAX_GB", "12")) * 2 ** 30:
            return
        with torch.cuda.device(self.dev):
            labels = torch.arange(n_cls, dtype=torch.int32, device=self.dev)
            ada_all = self.ada_params(labels)
            ws = torch.empty(self.lib.var_b200_ln_tables_workspace(C.byref(self.m), n_cls), dtype=torch.uint8, device=self.dev)
            for i in range(self.depth):
                t = torch.empty(per_block, dtype=torch.float32, device=self.dev)
                parts = torch.split(t, [n_cls * 3 * self.C, n_cls * 3 * self.C, n_cls * 4 * self.C, n_cls * 4 * self.C])
                L.check(self.lib.var_b200_ln_tables(C.byref(self.m), i, ada_all.data_ptr(), n_cls, *[p.data_ptr() for p in parts],
                                                    ws.data_ptr(), ws.numel(), L.current_stream()), "ln_tables")
                for name, p in zip(("u_qkv", "v_qkv", "u_fc1", "v_fc1"), parts):
                    setattr(self.blocks_arr[i], name, p.data_ptr())
                self._keep.append(t)
            torch.cuda.current_stream().synchronize()  # ws / ada_all are released on return
        self.ln_fused = True

    # ---- scratch management: buffers are cached per (tag) and grown on demand
    def _buf(self, tag: str, nbytes: int) -> torch.Tensor:
        b = self._ws.get(tag)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.dev)
            self._ws[tag] = b
            self.ws_gen += 1
        return b

    # ---- every call below goes through the torch.library custom-op layer (var_b200/ops.py -> C-ABI)
    def ada_params(self, labels_i32: torch.Tensor) -> torch.Tensor:
        return torch.ops.var_b200.ada_params(self.handle, labels_i32)

    def head_ada_from_cond(self, cond: torch.Tensor) -> torch.Tensor:
        """adaLN table with only the head_nm columns filled: Linear(SiLU(cond)) (basic_var.py:173) on the GEMM kernel."""
        n = cond.shape[0]
        a_in = torch.nn.functional.silu(cond.detach().to(self.dev).float()).to(torch.bfloat16).contiguous()
        rows = 2 * self.C
        w, b = self.w_ada[-rows:], self.b_ada[-rows:]  # head_nm.ada_lin is the last 2C rows of the packed adaLN weights
        out = torch.empty((n, rows), dtype=torch.float32, device=self.dev)
        g = L.GemmArgs()
        g.A, g.W, g.M, g.N, g.K, g.epilogue = a_in.data_ptr(), w.data_ptr(), n, rows, a_in.shape[1], L.EPI_BIAS_F32
        g.bias, g.out = b.data_ptr(), out.data_ptr()
        L.check(self.lib.var_b200_gemm_bf16(C.byref(g), L.current_stream()), "gemm(head ada)")
        ada = torch.zeros((n, self.ada_ld), dtype=torch.float32, device=self.dev)
        ada[:, self.ada_ld - rows:] = out
        return ada

    def embed(self, x_in, n_x, labels_i32, n_seq, l, first_rows, pos0) -> torch.Tensor:
        return torch.ops.var_b200.embed(self.handle, x_in, labels_i32, n_seq, l, first_rows, pos0)

    def _blocks_ws(self, n_seq, l, score=False):
        fn = self.lib.var_b200_score_workspace if score else self.lib.var_b200_blocks_workspace
        return self._buf("blocks", fn(C.byref(self.m), n_seq, l))

    def blocks_teacher(self, x, ada, n_seq, dump=None, labels=None):
        kv = self._buf("kv_scratch", 2 * n_seq * self.C * self.L * 2)
        torch.ops.var_b200.blocks(self.handle, x, ada, labels if self.ln_fused else None, n_seq, self.L, 0, kv, 0, self.L, dump)

    def kv_cache(self, n_seq: int) -> torch.Tensor:
        """Preallocated zero-initialised cache [depth][2][n_seq,H,L,64] bf16 (replaces torch.cat, basic_var.py:107-109)."""
        n = self.depth * 2 * n_seq * self.C * self.L
        kv = self._ws.get("kv_cache")
        if kv is None or kv.numel() != n:
            self._ws["kv_cache"] = None
            self.graphs.clear()  # frees the graphs' private pools before the (large) new cache is allocated
            kv = torch.zeros(n, dtype=torch.bfloat16, device=self.dev)
            self._ws["kv_cache"] = kv
            self.ws_gen += 1
        return kv

    def blocks_cached(self, x, ada, n_seq, l, pos0, kv, labels=None):
        torch.ops.var_b200.blocks(self.handle, x, ada, labels if self.ln_fused else None, n_seq, l, pos0, kv,
                                  2 * n_seq * self.C * self.L, self.L, None)

    def head_logits(self, x, ada, n_seq, l) -> torch.Tensor:
        return torch.ops.var_b200.head_logits(self.handle, x, ada, n_seq, l)

    def head_score(self, x, ada, n_seq, gt_i32, first_pos=0, per_scale=False, tok_logp=False):
        scores, ps, tl = torch.ops.var_b200.head_score(self.handle, x, ada, n_seq, gt_i32, first_pos, per_scale, tok_logp)
        return scores, (ps if per_scale else None), (tl if tok_logp else None)

    def sample(self, logits, B, l, t, q, top_k, top_p, mixed=None, use_cfg=True) -> torch.Tensor:
        idx, mx = torch.ops.var_b200.cfg_topk_sample(logits, B, l, bool(use_cfg), float(t), q, int(top_k), float(top_p),
                                                     mixed is not None)
        if mixed is not None:
            mixed.copy_(mx)
        return idx

    def sample_smooth(self, logits, B, l, t, q, top_k, top_p, q_gumbel, tau, logit_mul, codebook, mixed=None):
        """Sampler + more_smooth soft embedding (var.py:178-180): returns (idx [B,l], h [B,l,Cvae])."""
        idx, h, mx = torch.ops.var_b200.cfg_topk_sample_smooth(logits, B, l, float(t), q, int(top_k), float(top_p), q_gumbel,
                                                               float(tau), float(logit_mul), codebook, mixed is not None)
        if mixed is not None:
            mixed.copy_(mx)
        return idx, h
