"""`torch.library` custom-op layer over the C-ABI (include/var_b200.h): the thin layer BASELINE.json's north_star
names ("PyTorch is the host calling a thin C-ABI torch custom-op layer").

Every op is registered in the `var_b200::` namespace for the CUDA dispatch key only. There is no CPU kernel and no
fallback: calling an op with CPU tensors fails in the dispatcher ("no kernel for the CPU backend"), a missing
libvar_b200.so fails in `lib.load()`. The implementations allocate the outputs with torch (device memory is
PyTorch's), pass raw pointers + sizes + the current stream to the C entry point, and check its status code.

Model-level ops take the packed model as an integer handle (`PackedModel.handle`): the C structure
`var_b200_model_t` holds device pointers of the packed weights, which are owned by the PackedModel instance.
The host mirrors (`var.VAR`, `quant.VectorQuantizer2`, `scoring`) call these ops; nothing else calls the ctypes
binding for these entry points.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import List, Optional, Tuple

import torch
from torch.library import Library

from . import lib as L

_LIB = Library("var_b200", "DEF")
_MODELS: "weakref.WeakValueDictionary[int, object]" = weakref.WeakValueDictionary()
_next_handle = [1]


def register_model(pm) -> int:
    h = _next_handle[0]
    _next_handle[0] += 1
    _MODELS[h] = pm
    return h


def _pm(handle: int):
    pm = _MODELS.get(handle)
    if pm is None:
        raise L.VarB200Error(f"var_b200 op: stale model handle {handle} (the packed model was freed / repacked)")
    return pm


def _op(schema: str):
    name = schema.split("(", 1)[0]

    def deco(fn):
        _LIB.define(schema)
        _LIB.impl(name, fn, "CUDA")
        return fn
    return deco


def _quant_desc(codebook, phi_w, phi_b, ph, pw, phi_of_scale, resi) -> L.QuantDesc:
    d = L.QuantDesc()
    d.Cvae, d.V, d.n_scales = codebook.shape[1], codebook.shape[0], len(ph)
    for i in range(len(ph)):
        d.ph[i], d.pw[i], d.phi_of_scale[i] = ph[i], pw[i], phi_of_scale[i]
    d.n_phi, d.resi = phi_w.shape[0], resi
    d.codebook, d.phi_w, d.phi_b = codebook.data_ptr(), phi_w.data_ptr(), phi_b.data_ptr()
    return d


# ------------------------------------------------------------------------------------------------ transformer
@_op("ada_params(int model, Tensor labels) -> Tensor")
def ada_params(model: int, labels: torch.Tensor) -> torch.Tensor:
    """var_b200_ada_params: SiLU -> Linear of every block + head for n sequences (basic_var.py:156,173), one GEMM."""
    pm = _pm(model)
    n = labels.numel()
    out = torch.empty((n, pm.ada_ld), dtype=torch.float32, device=labels.device)
    ws = pm._buf("ada", pm.lib.var_b200_ada_workspace(C.byref(pm.m), n))
    L.check(pm.lib.var_b200_ada_params(C.byref(pm.m), labels.data_ptr(), n, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                       L.current_stream()), "ada_params")
    return out


@_op("embed(int model, Tensor? x_in, Tensor labels, int n_seq, int l, int first_rows, int pos0) -> Tensor")
def embed(model: int, x_in: Optional[torch.Tensor], labels: torch.Tensor, n_seq: int, l: int, first_rows: int,
          pos0: int) -> torch.Tensor:
    """var_b200_embed: class / position / level embeddings + word_embed (var.py:200-207,151-154,185-187)."""
    pm = _pm(model)
    out = torch.empty((n_seq, l, pm.C), dtype=torch.float32, device=labels.device)
    n_x, l_in = (x_in.shape[0], x_in.shape[1]) if x_in is not None else (0, 0)
    L.check(pm.lib.var_b200_embed(C.byref(pm.m), L.ptr(x_in), n_x, l_in, labels.data_ptr(), n_seq, l, first_rows, pos0,
                                  out.data_ptr(), L.current_stream()), "embed")
    return out


@_op("blocks(int model, Tensor(a!) x, Tensor ada, Tensor? labels, int n_seq, int l, int pos0, Tensor(b!) kv, "
     "int kv_layer_stride, int Lmax, Tensor(c!)? dump) -> ()")
def blocks(model: int, x: torch.Tensor, ada: torch.Tensor, labels: Optional[torch.Tensor], n_seq: int, l: int, pos0: int,
           kv: torch.Tensor, kv_layer_stride: int, Lmax: int, dump: Optional[torch.Tensor]) -> None:
    """var_b200_blocks: all AdaLNSelfAttn blocks in place on x (basic_var.py:152-159), K/V appended at pos0. labels
    (int32, the classes `ada` was computed from) enable the deferred-LayerNorm path (no LayerNorm pass per block)."""
    pm = _pm(model)
    ws = pm._blocks_ws(n_seq, l)
    L.check(pm.lib.var_b200_blocks(C.byref(pm.m), x.data_ptr(), ada.data_ptr(), L.ptr(labels), n_seq, l, pos0, kv.data_ptr(),
                                   kv_layer_stride, Lmax, L.ptr(dump), ws.data_ptr(), ws.numel(), L.current_stream()),
            "blocks")


@_op("head_logits(int model, Tensor x, Tensor ada, int n_seq, int l) -> Tensor")
def head_logits(model: int, x: torch.Tensor, ada: torch.Tensor, n_seq: int, l: int) -> torch.Tensor:
    """var_b200_head_logits: head(head_nm(x, cond)) -> fp32 logits (var.py:118-124)."""
    pm = _pm(model)
    ws = pm._blocks_ws(n_seq, l)
    out = torch.empty((n_seq, l, pm.V), dtype=torch.float32, device=x.device)
    L.check(pm.lib.var_b200_head_logits(C.byref(pm.m), x.data_ptr(), ada.data_ptr(), n_seq, l, out.data_ptr(),
                                        ws.data_ptr(), ws.numel(), L.current_stream()), "head_logits")
    return out


@_op("head_score(int model, Tensor x, Tensor ada, int n_seq, Tensor gt, int first_pos, bool per_scale, bool tok_logp) "
     "-> (Tensor, Tensor, Tensor)")
def head_score(model: int, x: torch.Tensor, ada: torch.Tensor, n_seq: int, gt: torch.Tensor, first_pos: int,
               per_scale: bool, tok_logp: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """var_b200_head_score: head GEMM with the log-softmax / gather / sum of eval_prob.py:446-463 fused (logits are
    never written). Returns (scores [n_seq], per-scale sums [n_seq,S] or empty, token log-probs [n_seq,L] or empty)."""
    pm = _pm(model)
    ws = pm._blocks_ws(n_seq, pm.L, score=True)
    scores = torch.empty(n_seq, dtype=torch.float32, device=x.device)
    ps = torch.empty((n_seq, pm.m.n_scales) if per_scale else (0,), dtype=torch.float32, device=x.device)
    tl = torch.empty((n_seq, pm.L) if tok_logp else (0,), dtype=torch.float32, device=x.device)
    L.check(pm.lib.var_b200_head_score(C.byref(pm.m), x.data_ptr(), ada.data_ptr(), n_seq, pm.L, gt.data_ptr(), gt.numel(),
                                       first_pos, scores.data_ptr(), ps.data_ptr() if per_scale else None,
                                       tl.data_ptr() if tok_logp else None, ws.data_ptr(), ws.numel(), L.current_stream()),
            "head_score")
    return scores, ps, tl


@_op("attention(Tensor q, Tensor k, Tensor v, int n_seq, int H, int Lq, int Lmax, int q_pos0, int[] level_end, "
     "float max_score, bool q_log2) -> Tensor")
def attention(q, k, v, n_seq: int, H: int, Lq: int, Lmax: int, q_pos0: int, level_end: List[int], max_score: float,
              q_log2: bool) -> torch.Tensor:
    """var_b200_attention: block-causal softmax(q k^T) v over the token pyramid (basic_var.py:98-117, var.py:107-112)."""
    out = torch.empty((n_seq, Lq, H * 64), dtype=torch.bfloat16, device=q.device)
    ends = (C.c_int * len(level_end))(*level_end)
    L.check(L.load().var_b200_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), n_seq, H, Lq, Lmax, q_pos0,
                                        len(level_end), ends, float(max_score), int(q_log2), L.current_stream()), "attention")
    return out


@_op("ln_modulate(Tensor x, Tensor scale, Tensor shift, int ada_ld, int rows_per_seq, float eps) -> Tensor")
def ln_modulate(x, scale, shift, ada_ld: int, rows_per_seq: int, eps: float) -> torch.Tensor:
    """var_b200_ln_modulate: LN(x) * (1 + scale[seq]) + shift[seq] -> bf16 (basic_var.py:157-158,174)."""
    M, Cc = x.shape
    out = torch.empty((M, Cc), dtype=torch.bfloat16, device=x.device)
    L.check(L.load().var_b200_ln_modulate(x.data_ptr(), scale.data_ptr(), shift.data_ptr(), ada_ld, rows_per_seq,
                                          out.data_ptr(), M, Cc, float(eps), L.current_stream()), "ln_modulate")
    return out


# ------------------------------------------------------------------------------------------------ sampler
@_op("cfg_topk_sample(Tensor logits, int B, int l, bool use_cfg, float t, Tensor q, int top_k, float top_p, "
     "bool want_mixed) -> (Tensor, Tensor)")
def cfg_topk_sample(logits, B: int, l: int, use_cfg: bool, t: float, q, top_k: int, top_p: float,
                    want_mixed: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """var_b200_cfg_topk_sample: CFG mix + top-k / top-p filter + multinomial(1) replayed from the caller's Exp(1)
    noise q (var.py:172-175, helpers.py:6-19). Returns (idx int64 [B,l], mixed logits [B,l,V] or empty)."""
    V = logits.shape[-1]
    idx = torch.empty((B, l), dtype=torch.int64, device=logits.device)
    mixed = torch.empty((B, l, V) if want_mixed else (0,), dtype=torch.float32, device=logits.device)
    L.check(L.load().var_b200_cfg_topk_sample(logits.data_ptr(), B, l, V, int(use_cfg), float(t), q.data_ptr(), int(top_k),
                                              float(top_p), idx.data_ptr(), mixed.data_ptr() if want_mixed else None,
                                              L.current_stream()), "cfg_topk_sample")
    return idx, mixed


@_op("cfg_topk_sample_smooth(Tensor logits, int B, int l, float t, Tensor q, int top_k, float top_p, Tensor q_gumbel, "
     "float tau, float logit_mul, Tensor codebook, bool want_mixed) -> (Tensor, Tensor, Tensor)")
def cfg_topk_sample_smooth(logits, B: int, l: int, t: float, q, top_k: int, top_p: float, q_gumbel, tau: float,
                           logit_mul: float, codebook, want_mixed: bool):
    """Sampler + the more_smooth Gumbel soft embedding (var.py:178-180). Returns (idx, h [B,l,Cvae], mixed or empty)."""
    V = logits.shape[-1]
    idx = torch.empty((B, l), dtype=torch.int64, device=logits.device)
    h = torch.empty((B, l, codebook.shape[1]), dtype=torch.float32, device=logits.device)
    mixed = torch.empty((B, l, V) if want_mixed else (0,), dtype=torch.float32, device=logits.device)
    L.check(L.load().var_b200_cfg_topk_sample_smooth(
        logits.data_ptr(), B, l, V, 1, float(t), q.data_ptr(), int(top_k), float(top_p), idx.data_ptr(),
        mixed.data_ptr() if want_mixed else None, q_gumbel.data_ptr(), float(tau), float(logit_mul), codebook.data_ptr(),
        int(codebook.shape[1]), h.data_ptr(), L.current_stream()), "cfg_topk_sample_smooth")
    return idx, h, mixed


# ------------------------------------------------------------------------------------------------ quantizer
@_op("quant_encode(Tensor f, Tensor codebook, Tensor phi_w, Tensor phi_b, int[] ph, int[] pw, int[] phi_of_scale, "
     "float resi, bool to_fhat, int search_mode) -> (Tensor, Tensor)")
def quant_encode(f, codebook, phi_w, phi_b, ph: List[int], pw: List[int], phi_of_scale: List[int], resi: float,
                 to_fhat: bool, search_mode: int):
    """var_b200_quant_encode: multi-scale residual VQ (quant.py:135-166). Returns (idx int64 [sum_s B*ph*pw] scale-major,
    f_hat list [S,B,C,H,W] or empty)."""
    B, Cc, H, W = f.shape
    d = _quant_desc(codebook, phi_w, phi_b, ph, pw, phi_of_scale, resi)
    Ltot = sum(a * b for a, b in zip(ph, pw))
    idx = torch.empty(B * Ltot, dtype=torch.int64, device=f.device)
    fh = torch.empty((len(ph), B, Cc, H, W) if to_fhat else (0,), dtype=torch.float32, device=f.device)
    lib = L.load()
    work = torch.empty(lib.var_b200_quant_encode_workspace(C.byref(d), B), dtype=torch.uint8, device=f.device)
    L.check(lib.var_b200_quant_encode(C.byref(d), f.data_ptr(), B, idx.data_ptr(), fh.data_ptr() if to_fhat else None,
                                      work.data_ptr(), work.numel(), int(search_mode), L.current_stream()), "quant_encode")
    return idx, fh


@_op("quant_decode(Tensor idx, int B, Tensor codebook, Tensor phi_w, Tensor phi_b, int[] ph, int[] pw, int[] phi_of_scale, "
     "float resi, bool want_input, bool want_list) -> (Tensor, Tensor, Tensor)")
def quant_decode(idx, B: int, codebook, phi_w, phi_b, ph: List[int], pw: List[int], phi_of_scale: List[int], resi: float,
                 want_input: bool, want_list: bool):
    """var_b200_quant_decode: token pyramid -> (var_input [B,L-l0,C] or empty, f_hat list [S,B,C,H,W] or empty, last f_hat)
    (quant.py:107-121,169-184)."""
    d = _quant_desc(codebook, phi_w, phi_b, ph, pw, phi_of_scale, resi)
    Cc, H, W = codebook.shape[1], ph[-1], pw[-1]
    Ltot, l0 = sum(a * b for a, b in zip(ph, pw)), ph[0] * pw[0]
    dev = idx.device
    vin = torch.empty((B, Ltot - l0, Cc) if want_input else (0,), dtype=torch.float32, device=dev)
    fl = torch.empty((len(ph), B, Cc, H, W) if want_list else (0,), dtype=torch.float32, device=dev)
    last = torch.empty((B, Cc, H, W), dtype=torch.float32, device=dev)
    L.check(L.load().var_b200_quant_decode(C.byref(d), idx.data_ptr(), B, vin.data_ptr() if want_input else None,
                                           fl.data_ptr() if want_list else None, last.data_ptr(), L.current_stream()),
            "quant_decode")
    return vin, fl, last


@_op("quant_next_input(int si, Tensor(a!) f_hat, Tensor idx, Tensor codebook, Tensor phi_w, Tensor phi_b, int[] ph, int[] pw, "
     "int[] phi_of_scale, float resi, bool token_major) -> Tensor")
def quant_next_input(si: int, f_hat, idx, codebook, phi_w, phi_b, ph: List[int], pw: List[int], phi_of_scale: List[int],
                     resi: float, token_major: bool):
    """var_b200_quant_next_input: f_hat += Phi(bicubic(E[idx])) in place; returns area(f_hat -> next scale) as
    [B,l,C] (token_major) or [B,C,h,w]; empty after the last scale (quant.py:187-196)."""
    d = _quant_desc(codebook, phi_w, phi_b, ph, pw, phi_of_scale, resi)
    B, Cc = f_hat.shape[0], f_hat.shape[1]
    last = si == len(ph) - 1
    if last:
        nxt = torch.empty((0,), dtype=torch.float32, device=f_hat.device)
    else:
        nh, nw = ph[si + 1], pw[si + 1]
        nxt = torch.empty((B, nh * nw, Cc) if token_major else (B, Cc, nh, nw), dtype=torch.float32, device=f_hat.device)
    L.check(L.load().var_b200_quant_next_input(C.byref(d), si, f_hat.data_ptr(), idx.data_ptr(), B,
                                               nxt.data_ptr() if (token_major and not last) else None,
                                               nxt.data_ptr() if (not token_major and not last) else None,
                                               L.current_stream()), "quant_next_input")
    return nxt


OPS = ["ada_params", "embed", "blocks", "head_logits", "head_score", "attention", "ln_modulate", "cfg_topk_sample",
       "cfg_topk_sample_smooth", "quant_encode", "quant_decode", "quant_next_input"]
