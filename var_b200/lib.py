"""ctypes binding of libvar_b200.so (the C-ABI declared in include/var_b200.h).

The library is the product: there is no Python/CPU fallback. Importing this module never builds anything; if the
shared object is missing, `load()` raises with the build command.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# VAR_B200_LIB: a variant build of the same library (developer A/B measurements); default = the in-tree build
_LIB_PATH = Path(os.environ["VAR_B200_LIB"]).resolve() if os.environ.get("VAR_B200_LIB") else \
    Path(__file__).resolve().parent / "libvar_b200.so"
_lib = None

EPI_BIAS_F32, EPI_BIAS_BF16, EPI_GELU_BF16, EPI_GATE_RESID, EPI_QKV, EPI_SCORE = range(6)
GEMM_EPI_PARTS = 2  # VAR_B200_GEMM_EPI_PARTS: EPI_SCORE partials per row and tile


class VarB200Error(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("W", C.c_void_p),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("epilogue", C.c_int), ("force_bn", C.c_int),
        ("bias", C.c_void_p), ("out", C.c_void_p),
        ("resid", C.c_void_p), ("gate", C.c_void_p), ("rows_per_seq", C.c_int), ("gate_ld", C.c_int),
        ("q_out", C.c_void_p), ("k_cache", C.c_void_p), ("v_cache", C.c_void_p), ("q_scale", C.c_void_p),
        ("C", C.c_int), ("H", C.c_int), ("pos0", C.c_int), ("Lmax", C.c_int), ("no_l2norm", C.c_int),
        ("gt", C.c_void_p), ("gt_mod", C.c_int), ("part", C.c_void_p), ("gt_logit", C.c_void_p),
        ("ln_a_out", C.c_void_p), ("ln_scale", C.c_void_p), ("ln_part_out", C.c_void_p), ("ln_part_in", C.c_void_p),
        ("ln_parts", C.c_int), ("ln_C", C.c_int), ("ln_eps", C.c_float), ("ln_u", C.c_void_p), ("ln_v", C.c_void_p),
        ("ln_labels", C.c_void_p),
    ]


MAX_SCALES = 16


class QuantDesc(C.Structure):
    _fields_ = [
        ("Cvae", C.c_int), ("V", C.c_int), ("n_scales", C.c_int),
        ("ph", C.c_int * MAX_SCALES), ("pw", C.c_int * MAX_SCALES), ("phi_of_scale", C.c_int * MAX_SCALES),
        ("n_phi", C.c_int), ("resi", C.c_float),
        ("codebook", C.c_void_p), ("phi_w", C.c_void_p), ("phi_b", C.c_void_p),
    ]


class BlockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("w_qkv", "b_qkv", "q_scale", "w_proj", "b_proj", "w_fc1", "b_fc1", "w_fc2", "b_fc2",
                 "u_qkv", "v_qkv", "u_fc1", "v_fc1")]


class ModelDesc(C.Structure):
    _fields_ = [
        ("depth", C.c_int), ("C", C.c_int), ("H", C.c_int), ("V", C.c_int), ("Cvae", C.c_int), ("n_scales", C.c_int),
        ("num_classes", C.c_int), ("shared_aln", C.c_int), ("norm_eps", C.c_float),
        ("patch_nums", C.c_int * MAX_SCALES),
        ("blocks", C.POINTER(BlockWeights)),
        ("w_ada", C.c_void_p), ("b_ada", C.c_void_p), ("ada_rows", C.c_int), ("ada_gss", C.c_void_p),
        ("w_head", C.c_void_p), ("b_head", C.c_void_p), ("w_word", C.c_void_p), ("b_word", C.c_void_p),
        ("class_emb", C.c_void_p), ("pos_start", C.c_void_p), ("lvl_pos", C.c_void_p),
        ("attn_max_score", C.c_float), ("attn_q_log2", C.c_int), ("attn_no_l2norm", C.c_int),
    ]


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load libvar_b200.so; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise VarB200Error(
            f"{_LIB_PATH} is missing: build it with `python -m var_b200.build` "
            "(var_b200 has no CPU or PyTorch fallback path)")
    lib = C.CDLL(str(_LIB_PATH))
    lib.var_b200_last_error.restype = C.c_char_p
    lib.var_b200_last_error.argtypes = []
    lib.var_b200_launch_count.restype = C.c_longlong
    lib.var_b200_launch_count.argtypes = []
    lib.var_b200_profile_begin.restype = None
    lib.var_b200_profile_begin.argtypes = []
    lib.var_b200_profile_end.restype = C.c_int
    lib.var_b200_profile_end.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]
    _declare(lib)
    _lib = lib
    return lib


def _declare(lib: C.CDLL) -> None:
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    sigs = {
        "var_b200_gemm_bf16": [C.POINTER(GemmArgs), vp],
        "var_b200_gemm_tile_n": [i32],
        "var_b200_gemm_ln_parts": [i32, i32],
        "var_b200_umma_probe": [vp, vp, vp, i32, i32, vp],
    }
    sigs.update(_EXTRA_SIGS(vp, i32, i64, f32))
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = i32
    for name in ("var_b200_ada_workspace", "var_b200_blocks_workspace", "var_b200_score_workspace"):
        fn = getattr(lib, name)
        fn.argtypes = [C.POINTER(ModelDesc), i32] if name == "var_b200_ada_workspace" else [C.POINTER(ModelDesc), i32, i32]
        fn.restype = C.c_size_t
    lib.var_b200_ln_tables_workspace.argtypes = [C.POINTER(ModelDesc), i32]
    lib.var_b200_ln_tables_workspace.restype = C.c_size_t
    lib.var_b200_gn_workspace.argtypes = [i32, i32, i32, i32]
    lib.var_b200_gn_workspace.restype = C.c_size_t
    lib.var_b200_conv3x3_gn_workspace.argtypes = [i32, i32, i32, i32]
    lib.var_b200_conv3x3_gn_workspace.restype = C.c_size_t
    lib.var_b200_vae_attn_workspace.argtypes = [i32, i32, i32, i32]
    lib.var_b200_vae_attn_workspace.restype = C.c_size_t
    lib.var_b200_quant_encode_workspace.argtypes = [C.POINTER(QuantDesc), i32]
    lib.var_b200_quant_encode_workspace.restype = C.c_size_t


def exported_symbols():
    """Every symbol include/var_b200.h declares (used by the CPU-only ABI test)."""
    import re
    hdr = (Path(__file__).resolve().parent.parent / "include" / "var_b200.h").read_text()
    return sorted(set(re.findall(r"VAR_B200_API[^;(]*?\b(var_b200_\w+)\s*\(", hdr)))


def _EXTRA_SIGS(vp, i32, i64, f32):
    sz, dbl = C.c_size_t, C.c_double
    return {
        "var_b200_attention": [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, C.POINTER(C.c_int), f32, i32, vp],
        "var_b200_ln_modulate": [vp, vp, vp, i32, i32, vp, i32, i32, f32, vp],
        "var_b200_quant_encode": [C.POINTER(QuantDesc), vp, i32, vp, vp, vp, sz, i32, vp],
        "var_b200_quant_decode": [C.POINTER(QuantDesc), vp, i32, vp, vp, vp, vp],
        "var_b200_quant_next_input": [C.POINTER(QuantDesc), i32, vp, vp, i32, vp, vp, vp],
        "var_b200_cfg_topk_sample": [vp, i32, i32, i32, i32, dbl, vp, i32, f32, vp, vp, vp],
        "var_b200_conv3x3_nhwc": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "var_b200_cfg_topk_sample_smooth": [vp, i32, i32, i32, i32, dbl, vp, i32, f32, vp, vp, vp, f32, f32, vp, i32, vp, vp],
        "var_b200_cfg_token_logprob": [vp, vp, vp, vp, i32, i32, i32, vp, vp],
        "var_b200_neighbor_select": [vp, i32, i32, i32, dbl, vp, vp, vp, i32, i32, i32, f32, f32, vp, vp, vp, vp],
        "var_b200_cfg_token_expected_dist": [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp],
        "var_b200_scale_sums": [vp, i32, i32, i32, C.POINTER(C.c_int), i32, vp, vp, vp],
        "var_b200_gn_silu_nhwc": [vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, i32, vp, sz, vp],
        "var_b200_add_bias_nhwc": [vp, vp, vp, vp, vp, C.c_longlong, i32, vp],
        "var_b200_upsample2x_nhwc": [vp, vp, vp, i32, i32, i32, i32, vp],
        "var_b200_conv3x3_s2_nhwc": [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "var_b200_conv1x1_nhwc": [vp, vp, vp, vp, vp, C.c_longlong, i32, i32, vp],
        "var_b200_conv3x3_gn_nhwc": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, sz, vp],
        "var_b200_gn_apply_nhwc": [vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, i32, vp],
        "var_b200_vae_attn_block": [vp, vp, vp, i32, f32, vp, vp, vp, vp, vp, i32, i32, i32, vp, sz, vp],
        "var_b200_ada_ld": [C.POINTER(ModelDesc)],
        "var_b200_ada_params": [C.POINTER(ModelDesc), vp, i32, vp, vp, sz, vp],
        "var_b200_embed": [C.POINTER(ModelDesc), vp, i32, i32, vp, i32, i32, i32, i32, vp, vp],
        "var_b200_blocks": [C.POINTER(ModelDesc), vp, vp, vp, i32, i32, i32, vp, sz, i32, vp, vp, sz, vp],
        "var_b200_ln_tables": [C.POINTER(ModelDesc), i32, vp, i32, vp, vp, vp, vp, vp, sz, vp],
        "var_b200_head_logits": [C.POINTER(ModelDesc), vp, vp, i32, i32, vp, vp, sz, vp],
        "var_b200_head_score": [C.POINTER(ModelDesc), vp, vp, i32, i32, vp, i32, i32, vp, vp, vp, vp, sz, vp],
    }


PROF_KINDS = ["gemm_bias_f32", "gemm_bias_bf16", "gemm_gelu", "gemm_gate_resid", "gemm_qkv", "gemm_score", "attn", "ln_modulate",
              "embed", "cond_silu", "sample", "quant", "score_finalize", "other"]


class kernel_profile:
    """Context manager: CUDA-event time of every kernel class launched inside the block -> .ms / .n dicts."""

    def __enter__(self):
        load().var_b200_profile_begin()
        return self

    def __exit__(self, *exc):
        n = len(PROF_KINDS)
        ms, cnt = (C.c_double * n)(), (C.c_longlong * n)()
        check(load().var_b200_profile_end(ms, cnt, n), "profile_end")
        self.ms = {k: ms[i] for i, k in enumerate(PROF_KINDS) if cnt[i]}
        self.n = {k: cnt[i] for i, k in enumerate(PROF_KINDS) if cnt[i]}
        return False


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().var_b200_last_error().decode(errors="replace")
        raise VarB200Error(f"{what or 'var_b200 call'} failed with code {rc}: {msg}")


def ptr(t) -> int:
    """Device (or host) address of a torch tensor, None -> NULL."""
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def device_guard(fn):
    """Decorator for module methods: run with the device of the module's parameters current, so that the kernels are
    launched on that device's current stream whichever device the caller had selected."""
    import functools

    import torch

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        dev = self._device()
        if dev.type != "cuda":
            return fn(self, *args, **kwargs)  # raises VarB200Error further down: there is no CPU path
        with torch.cuda.device(dev):
            return fn(self, *args, **kwargs)
    return wrapper
