"""In-tree build of libvar_b200.so (sm_100a only) with plain nvcc.

`python -m var_b200.build` compiles every .cu under var_b200/csrc into var_b200/_build/*.o (in parallel, skipping
objects newer than their sources) and links var_b200/libvar_b200.so. The .so is git-ignored but travels with the
gpurun snapshot. No torch headers are involved: the library is a plain C-ABI (include/var_b200.h).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
BUILD = ROOT / "_build"
LIB = ROOT / "libvar_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v", "-DVAR_B200_BUILD"]
# files whose arithmetic must match the C oracle bit for bit: no FMA contraction
EXACT_FILES = {"quant.cu"}


def _sources():
    return sorted(p for p in CSRC.glob("*.cu"))


def _headers_mtime() -> float:
    hs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT.parent / "include" / "var_b200.h"]
    return max(h.stat().st_mtime for h in hs if h.exists())


def _compile(src: Path, force: bool) -> str:
    obj = BUILD / (src.stem + ".o")
    if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, _headers_mtime()):
        return f"[up-to-date] {src.name}"
    cmd = [NVCC, *ARCH, *COMMON]
    if src.name in EXACT_FILES:
        cmd += ["-fmad=false"]
    cmd += ["-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = (BUILD / (src.stem + ".log"))
    log.write_text(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    return f"[compiled] {src.name}"


def build(force: bool = False, verbose: bool = True) -> Path:
    BUILD.mkdir(exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for msg in ex.map(lambda s: _compile(s, force), srcs):
            if verbose:
                print(msg, flush=True)
    objs = [str(BUILD / (s.stem + ".o")) for s in srcs]
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[linked] {LIB}", flush=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
