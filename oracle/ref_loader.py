"""ORACLE tooling (test / measurement infrastructure, not product code): make the UNMODIFIED reference travel.

The reference's hot path is pure Python (`/root/reference/models/*.py` + the top-level `dist.py` they import). It needs
no build; to be callable on the GPU box (where /root/reference does not exist) `fetch()` copies exactly those files,
byte for byte, into the git-ignored `oracle/_ref/` (so they never enter the history, but ship with the gpurun
snapshot like the built .so files). `__graft_entry__.build()` calls `fetch()`; `bench.py --impl reference` and the
`cpu_baseline` leg call `load()`.

Only tests/, __graft_entry__ and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys
import typing
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference")
REF_DST = HERE / "_ref"
FILES = ["dist.py", "models/__init__.py", "models/basic_vae.py", "models/basic_var.py", "models/helpers.py",
         "models/quant.py", "models/var.py", "models/vqvae.py"]


def _sha(p: Path) -> str:
    return hashlib.sha256(p.read_bytes()).hexdigest()


def fetch(verbose: bool = True) -> bool:
    """Copy the reference's model sources into oracle/_ref/ (no-op when /root/reference is absent). Returns True when
    oracle/_ref holds a complete copy afterwards."""
    if REF_SRC.exists():
        for rel in FILES:
            dst = REF_DST / rel
            dst.parent.mkdir(parents=True, exist_ok=True)
            if not dst.exists() or _sha(dst) != _sha(REF_SRC / rel):
                shutil.copyfile(REF_SRC / rel, dst)
        (REF_DST / "MANIFEST.sha256").write_text("".join(f"{_sha(REF_DST / rel)}  {rel}\n" for rel in FILES))
        if verbose:
            print(f"oracle/_ref: {len(FILES)} reference files copied unmodified from {REF_SRC}")
    return available()


def available() -> bool:
    return all((REF_DST / rel).exists() for rel in FILES)


_loaded = None


def load():
    """Import the reference's `models` package from oracle/_ref. Returns (build, ref_var_module, ref_helpers_module);
    build(**kw) -> (vae, var) on CPU in eval mode with nn.*.reset_parameters restored (models/__init__.py:24-25 patches
    them process-wide) — SURVEY.md Appendix B."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where "
                           "/root/reference exists")
    import torch
    import torch.nn as nn
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    if "models" in sys.modules or "dist" in sys.modules:
        raise RuntimeError("a module named `models` or `dist` is already imported")
    sys.path.insert(0, str(REF_DST))
    torch.Optional = typing.Optional  # models/var.py:241-242 needs it (absent in torch 2.11)
    saved = {c: c.reset_parameters for c in (nn.Linear, nn.LayerNorm, nn.BatchNorm2d, nn.SyncBatchNorm, nn.Conv1d,
                                             nn.Conv2d, nn.ConvTranspose1d, nn.ConvTranspose2d)}
    try:
        import models  # noqa: F401
        from models import build_vae_var
        import models.helpers as ref_helpers
        import models.var as ref_var
    finally:
        sys.path.remove(str(REF_DST))

    def build(**kw):
        vae, var = build_vae_var(device="cpu", flash_if_available=False, fused_if_available=False, **kw)
        for c, f in saved.items():
            c.reset_parameters = f
        var.eval(); vae.eval(); var.cond_drop_rate = 0
        return vae, var

    _loaded = (build, ref_var, ref_helpers)
    return _loaded
