"""ORACLE (test infrastructure, not product code): ctypes/numpy front-end of oracle/quant_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
Mirrors the call signatures of VectorQuantizer2 (/root/reference/models/quant.py:107-196) on numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libquant_oracle.so"
MAXS = 32


class _Cfg(C.Structure):
    _fields_ = [("B", C.c_int), ("C", C.c_int), ("H", C.c_int), ("W", C.c_int), ("V", C.c_int), ("S", C.c_int),
                ("ph", C.c_int * MAXS), ("pw", C.c_int * MAXS), ("phi_of_scale", C.c_int * MAXS), ("n_phi", C.c_int),
                ("resi", C.c_float), ("codebook", C.c_void_p), ("phi_w", C.c_void_p), ("phi_b", C.c_void_p)]


def build(force: bool = False) -> Path:
    if force or not _SO.exists() or _SO.stat().st_mtime < (_HERE / "quant_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "libquant_oracle.so"], check=True, capture_output=True)
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
        for n in ("vq_oracle_encode", "vq_oracle_decode", "vq_oracle_next_input"):
            getattr(_lib, n).restype = C.c_int
    return _lib


def phi_index_per_scale(n_scales: int, n_phi: int) -> List[int]:
    """PhiPartiallyShared / PhiNonShared selector (quant.py:218-243): nearest tick to si/(S-1).
    n_phi == 1 is PhiShared (always 0)."""
    if n_phi == 1:
        return [0] * n_scales
    K = n_phi
    ticks = np.linspace(1 / 3 / K, 1 - 1 / 3 / K, K) if K == 4 else np.linspace(1 / 2 / K, 1 - 1 / 2 / K, K)
    return [int(np.argmin(np.abs(ticks - si / (n_scales - 1)))) for si in range(n_scales)]


def _hw(patch_nums) -> List[Tuple[int, int]]:
    return [(p, p) if isinstance(p, int) else (int(p[0]), int(p[1])) for p in patch_nums]


class QuantOracle:
    def __init__(self, codebook: np.ndarray, phi_w: np.ndarray, phi_b: np.ndarray, patch_nums: Sequence,
                 resi: float = 0.5):
        self.codebook = np.ascontiguousarray(codebook, dtype=np.float32)
        self.phi_w = np.ascontiguousarray(phi_w, dtype=np.float32)  # [n_phi,C,C,3,3]
        self.phi_b = np.ascontiguousarray(phi_b, dtype=np.float32)  # [n_phi,C]
        self.patch_hws = _hw(patch_nums)
        self.resi = float(resi)
        self.V, self.Cv = self.codebook.shape

    def _cfg(self, B: int, patch_hws=None) -> _Cfg:
        hws = self.patch_hws if patch_hws is None else _hw(patch_hws)
        c = _Cfg()
        c.B, c.C, c.V, c.S = B, self.Cv, self.V, len(hws)
        c.H, c.W = hws[-1]
        for i, (h, w) in enumerate(hws):
            c.ph[i], c.pw[i] = h, w
        for i, k in enumerate(phi_index_per_scale(len(hws), self.phi_w.shape[0])):
            c.phi_of_scale[i] = k
        c.n_phi, c.resi = self.phi_w.shape[0], self.resi
        c.codebook = self.codebook.ctypes.data
        c.phi_w = self.phi_w.ctypes.data
        c.phi_b = self.phi_b.ctypes.data
        return c

    @staticmethod
    def _split(flat: np.ndarray, B: int, hws) -> List[np.ndarray]:
        out, off = [], 0
        for h, w in hws:
            out.append(flat[off:off + B * h * w].reshape(B, h * w).copy())
            off += B * h * w
        return out

    def f_to_idxBl_or_fhat(self, f_BChw: np.ndarray, to_fhat: bool, v_patch_nums=None):
        """quant.py:135-166."""
        f = np.ascontiguousarray(f_BChw, dtype=np.float32)
        B = f.shape[0]
        cfg = self._cfg(B, v_patch_nums)
        hws = self.patch_hws if v_patch_nums is None else _hw(v_patch_nums)
        assert hws[-1] == (f.shape[2], f.shape[3]), f"patch_hws[-1]={hws[-1]} != (H={f.shape[2]}, W={f.shape[3]})"
        L = sum(h * w for h, w in hws)
        idx = np.zeros(B * L, dtype=np.int64)
        fh = np.zeros((len(hws),) + f.shape, dtype=np.float32) if to_fhat else None
        rc = _load().vq_oracle_encode(C.byref(cfg), C.c_void_p(f.ctypes.data), C.c_void_p(idx.ctypes.data),
                                      C.c_void_p(fh.ctypes.data if to_fhat else None))
        assert rc == 0, rc
        return [fh[i] for i in range(len(hws))] if to_fhat else self._split(idx, B, hws)

    def _flat_idx(self, ms_idx_Bl: List[np.ndarray]) -> np.ndarray:
        return np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.int64).reshape(-1) for x in ms_idx_Bl]))

    def idxBl_to_var_input(self, ms_idx_Bl: List[np.ndarray]) -> np.ndarray:
        """quant.py:169-184 -> [B, L - l0, Cvae]."""
        B = ms_idx_Bl[0].shape[0]
        cfg = self._cfg(B)
        L = sum(h * w for h, w in self.patch_hws)
        l0 = self.patch_hws[0][0] * self.patch_hws[0][1]
        out = np.zeros((B, L - l0, self.Cv), dtype=np.float32)
        idx = self._flat_idx(ms_idx_Bl)
        rc = _load().vq_oracle_decode(C.byref(cfg), C.c_void_p(idx.ctypes.data), C.c_void_p(out.ctypes.data), None)
        assert rc == 0, rc
        return out

    def idxBl_to_fhat(self, ms_idx_Bl: List[np.ndarray], last_one: bool = False):
        """embed_to_fhat(all_to_max_scale=True) applied to codebook lookups (quant.py:107-121, vqvae.py:77-84)."""
        B = ms_idx_Bl[0].shape[0]
        cfg = self._cfg(B)
        H, W = self.patch_hws[-1]
        fh = np.zeros((len(self.patch_hws), B, self.Cv, H, W), dtype=np.float32)
        idx = self._flat_idx(ms_idx_Bl)
        rc = _load().vq_oracle_decode(C.byref(cfg), C.c_void_p(idx.ctypes.data), None, C.c_void_p(fh.ctypes.data))
        assert rc == 0, rc
        return fh[-1] if last_one else [fh[i] for i in range(fh.shape[0])]

    def get_next_autoregressive_input(self, si: int, f_hat: np.ndarray, idx_Bl: Optional[np.ndarray],
                                      h: Optional[np.ndarray] = None) -> Optional[np.ndarray]:
        """quant.py:187-196 with h_BChw = codebook[idx] (or an arbitrary h [B,C,ph,pw], the more_smooth soft embeddings of
        var.py:178-182, fed through a virtual codebook with an identity index); f_hat updated in place; returns
        area(f_hat) or None."""
        assert f_hat.dtype == np.float32 and f_hat.flags.c_contiguous
        B = f_hat.shape[0]
        cfg = self._cfg(B)
        if h is not None:
            ph, pw = self.patch_hws[si]
            h_tok = np.ascontiguousarray(h.reshape(B, self.Cv, ph * pw).transpose(0, 2, 1), dtype=np.float32)
            cfg.codebook, cfg.V = h_tok.ctypes.data, B * ph * pw
            idx_Bl = np.arange(B * ph * pw, dtype=np.int64).reshape(B, ph * pw)
        idx = np.ascontiguousarray(idx_Bl, dtype=np.int64)
        nxt = None
        if si != len(self.patch_hws) - 1:
            nh, nw = self.patch_hws[si + 1]
            nxt = np.zeros((B, self.Cv, nh, nw), dtype=np.float32)
        rc = _load().vq_oracle_next_input(C.byref(cfg), si, C.c_void_p(f_hat.ctypes.data), C.c_void_p(idx.ctypes.data),
                                          C.c_void_p(nxt.ctypes.data if nxt is not None else None))
        assert rc == 0, rc
        return nxt
