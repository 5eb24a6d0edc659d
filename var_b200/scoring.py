"""Per-class likelihood scoring (the fork's addition: /root/reference/eval_prob.py:417-465,600-601 `bayesian` mode;
per-scale sums as in var_analysis.py:435-466), class-batched and optionally class-sharded over the GPUs of one box.

The reference runs one class per forward and recomputes idxBl_to_var_input per class (eval_prob.py:436, loop
invariant). Here the token pyramid of the image is embedded once and broadcast across a batch of candidate classes;
the head GEMM fuses log-softmax + gather + sum so logits [K,680,4096] are never written.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from .var import VAR
from .vqvae import VQVAE


def _on_var_device(fn):
    """Run with the model's device current (the kernels launch on that device's current stream)."""
    import functools

    @functools.wraps(fn)
    def wrapper(var, *args, **kwargs):
        dev = var._device()
        if dev.type != "cuda":
            return fn(var, *args, **kwargs)
        with torch.cuda.device(dev):
            return fn(var, *args, **kwargs)
    return wrapper


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@torch.no_grad()
@_on_var_device
def class_log_likelihoods(var: VAR, gt_idx_list: Sequence[torch.Tensor], labels: torch.Tensor, *, class_batch: int = 125,
                          first_pos: int = 0, per_scale: bool = False):
    """sum_t log p(gt_t | class) for every label in `labels` (one image: gt_idx_list[si] is [1, pn^2]).
    Returns scores [K] fp32 (and [K, S] per-scale sums when per_scale). eval_prob.py:436-463."""
    assert gt_idx_list[0].shape[0] == 1, "one image at a time (eval_prob.py:268: batch_size=1)"
    pm = var._model()
    dev = pm.dev
    quant = var.vae_quant_proxy[0]
    x_in = quant.idxBl_to_var_input(list(gt_idx_list))  # [1, L-first_l, Cvae], computed once per image
    gt = torch.cat([g.reshape(-1) for g in gt_idx_list]).to(device=dev, dtype=torch.int32).contiguous()
    labels = var._labels_i32(labels.reshape(-1), labels.numel())  # range-checked like the reference's embedding lookup
    K = labels.numel()
    scores = torch.empty(K, dtype=torch.float32, device=dev)
    ps_all = torch.empty((K, len(var.patch_nums)), dtype=torch.float32, device=dev) if per_scale else None
    for lo in range(0, K, class_batch):
        lab = labels[lo:lo + class_batch].contiguous()
        n = lab.numel()
        x = pm.embed(x_in, 1, lab, n, var.L, var.first_l, 0)
        ada = pm.ada_params(lab)
        pm.blocks_teacher(x, ada, n, labels=lab)
        s, ps, _ = pm.head_score(x, ada, n, gt, first_pos=first_pos, per_scale=per_scale)
        scores[lo:lo + n] = s
        if per_scale:
            ps_all[lo:lo + n] = ps
    return (scores, ps_all) if per_scale else scores


@torch.no_grad()
@_on_var_device
def class_log_likelihoods_cfg(var: VAR, gt_idx_list: Sequence[torch.Tensor], labels: torch.Tensor, cfg: float, *,
                              class_batch: int = 64, first_pos: int = 0):
    """CFG-mixed likelihood scores (var_analysis.py:320-346,437-466): the teacher-forced logits of every candidate class
    are mixed with the unconditional (label 1000) logits, t = cfg * si/(S-1) per position, before log-softmax.
    Returns (total [K], per_scale [K, S], tok_logp [K, L]); accumulated / conditional sums of var_analysis.py:447-462
    are prefix / suffix sums of per_scale."""
    import ctypes as C
    from . import lib as L
    lib = L.load()
    assert gt_idx_list[0].shape[0] == 1
    pm = var._model()
    dev = pm.dev
    quant = var.vae_quant_proxy[0]
    x_in = quant.idxBl_to_var_input(list(gt_idx_list))
    gt = torch.cat([g.reshape(-1) for g in gt_idx_list]).to(device=dev, dtype=torch.int32).contiguous()
    S = len(var.patch_nums)
    t_row = torch.cat([torch.full((pn * pn,), cfg * (si / (S - 1))) for si, pn in enumerate(var.patch_nums)]).to(dev).float()
    ends = (C.c_int * S)(*[e for _, e in var.begin_ends])

    def logits_of(lab):
        n = lab.numel()
        x = pm.embed(x_in, 1, lab, n, var.L, var.first_l, 0)
        ada = pm.ada_params(lab)
        pm.blocks_teacher(x, ada, n, labels=lab)
        return pm.head_logits(x, ada, n, var.L)

    lu = logits_of(torch.tensor([var.num_classes], device=dev, dtype=torch.int32))
    labels = var._labels_i32(labels.reshape(-1), labels.numel())  # range-checked like the reference's embedding lookup
    K = labels.numel()
    tok = torch.empty((K, var.L), dtype=torch.float32, device=dev)
    for lo in range(0, K, class_batch):
        lab = labels[lo:lo + class_batch].contiguous()
        lc = logits_of(lab)
        L.check(lib.var_b200_cfg_token_logprob(lc.data_ptr(), lu.data_ptr(), gt.data_ptr(), t_row.data_ptr(), lab.numel(),
                                               var.L, var.V, tok[lo:].data_ptr(), L.current_stream()), "cfg_token_logprob")
    per_scale = torch.empty((K, S), dtype=torch.float32, device=dev)
    total = torch.empty(K, dtype=torch.float32, device=dev)
    L.check(lib.var_b200_scale_sums(tok.data_ptr(), K, var.L, S, ends, first_pos, per_scale.data_ptr(), total.data_ptr(),
                                    L.current_stream()), "scale_sums")
    return total, per_scale, tok


@torch.no_grad()
@_on_var_device
def class_expected_distances(var: VAR, gt_idx_list: Sequence[torch.Tensor], labels: torch.Tensor, cfg: float = 0.0, *,
                             top_k: Optional[int] = None, class_batch: int = 64, first_pos: int = 0):
    """--mode l2_dist of var_analysis.py (:252-256,468-524): per candidate class the negated expected codebook distance
    -sum_v p_v * ||E_gt - E_v||_2 of every position, p = softmax of the teacher-forced logits (CFG-mixed with the
    label-1000 logits when cfg > 0, :336-342; restricted to the top_k most probable tokens and renormalised when top_k
    is given, :476-486). Returns (total [K], per_scale [K, S], tok [K, L]); larger is better, as in the reference."""
    import ctypes as C
    from . import lib as L
    lib = L.load()
    assert gt_idx_list[0].shape[0] == 1
    pm = var._model()
    dev = pm.dev
    quant = var.vae_quant_proxy[0]
    x_in = quant.idxBl_to_var_input(list(gt_idx_list))
    gt = torch.cat([g.reshape(-1) for g in gt_idx_list]).to(device=dev, dtype=torch.int32).contiguous()
    S = len(var.patch_nums)
    dists, _ = quant.codebook_distances()  # var_analysis.py:256, cached per codebook (V x V fp32 = 64 MB, L2 resident)
    t_row = torch.cat([torch.full((pn * pn,), cfg * (si / (S - 1))) for si, pn in enumerate(var.patch_nums)]).to(dev).float()
    ends = (C.c_int * S)(*[e for _, e in var.begin_ends])

    def logits_of(lab):
        n = lab.numel()
        x = pm.embed(x_in, 1, lab, n, var.L, var.first_l, 0)
        ada = pm.ada_params(lab)
        pm.blocks_teacher(x, ada, n, labels=lab)
        return pm.head_logits(x, ada, n, var.L)

    lu = logits_of(torch.tensor([var.num_classes], device=dev, dtype=torch.int32)) if cfg > 0 else None
    labels = var._labels_i32(labels.reshape(-1), labels.numel())  # range-checked like the reference's embedding lookup
    K = labels.numel()
    tok = torch.empty((K, var.L), dtype=torch.float32, device=dev)
    for lo in range(0, K, class_batch):
        lab = labels[lo:lo + class_batch].contiguous()
        lc = logits_of(lab)
        L.check(lib.var_b200_cfg_token_expected_dist(lc.data_ptr(), lu.data_ptr() if lu is not None else None, gt.data_ptr(),
                                                     t_row.data_ptr(), dists.data_ptr(), lab.numel(), var.L, var.V,
                                                     int(top_k or 0), tok[lo:].data_ptr(), L.current_stream()),
                "cfg_token_expected_dist")
    tok.neg_()
    per_scale = torch.empty((K, S), dtype=torch.float32, device=dev)
    total = torch.empty(K, dtype=torch.float32, device=dev)
    L.check(lib.var_b200_scale_sums(tok.data_ptr(), K, var.L, S, ends, first_pos, per_scale.data_ptr(), total.data_ptr(),
                                    L.current_stream()), "scale_sums")
    return total, per_scale, tok


@torch.no_grad()
def classify_image(var: VAR, vae: VQVAE, img: torch.Tensor, num_classes: Optional[int] = None, *, class_batch: int = 125,
                   first_pos: int = 0, group: Optional[dist.ProcessGroup] = None):
    """eval_prob.py:417-465,600-601 for one image [1,3,H,W] in [-1,1]: tokenise, score every class, argmax.
    With torch.distributed initialised the classes are sharded over the ranks and the per-class scores are combined
    with one all-gather. Returns (pred, scores[num_classes])."""
    K = num_classes or var.num_classes
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    gt_idx_list = vae.img_to_idxBl(img)
    lo, hi = shard_range(K, rank, world)
    labels = torch.arange(lo, hi, device=img.device)
    local = class_log_likelihoods(var, gt_idx_list, labels, class_batch=class_batch, first_pos=first_pos)
    scores = gather_class_scores(local, K, rank, world, group)
    return int(torch.argmax(scores).item()), scores


def gather_class_scores(local: torch.Tensor, K: int, rank: int, world: int, group=None) -> torch.Tensor:
    """One all-gather of the per-rank score slices (padded to equal length), reassembled in class order."""
    if world == 1:
        return local
    per = (K + world - 1) // world
    pad = torch.full((per,), float("-inf"), dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    out = torch.empty(world * per, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(K, r, world)
        parts.append(out[r * per:r * per + (hi - lo)])
    return torch.cat(parts)
