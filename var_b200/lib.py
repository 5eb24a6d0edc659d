"""ctypes binding of libvar_b200.so (the C-ABI declared in include/var_b200.h).

The library is the product: there is no Python/CPU fallback. Importing this module never builds anything; if the
shared object is missing, `load()` raises with the build command.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libvar_b200.so"
_lib = None

EPI_BIAS_F32, EPI_BIAS_BF16, EPI_GELU_BF16, EPI_GATE_RESID, EPI_QKV, EPI_SCORE = range(6)


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
        ("C", C.c_int), ("H", C.c_int), ("pos0", C.c_int), ("Lmax", C.c_int),
        ("gt", C.c_void_p), ("part", C.c_void_p), ("gt_logit", C.c_void_p),
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
    _declare(lib)
    _lib = lib
    return lib


def _declare(lib: C.CDLL) -> None:
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    sigs = {
        "var_b200_gemm_bf16": [C.POINTER(GemmArgs), vp],
        "var_b200_gemm_tile_n": [i32],
        "var_b200_umma_probe": [vp, vp, vp, i32, i32, vp],
    }
    sigs.update(_EXTRA_SIGS(vp, i32, i64, f32))
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = i32


def _EXTRA_SIGS(vp, i32, i64, f32):
    return {}


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
