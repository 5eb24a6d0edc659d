#!/usr/bin/env python
"""Opcode census of libvar_b200.so per kernel (cuobjdump -sass): which kernels really carry tcgen05 / TMEM / TMA
instructions. Runs without a GPU:   python tools/sass_summary.py > profiles/sass_summary.txt

Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st,
UTMALDG = cp.async.bulk.tensor (TMA load), UTCBAR = tcgen05.commit -> mbarrier (MULTICAST across the CTA pair),
SYNCS = mbarrier ops, MUFU = special-function unit, F2FP = packed float -> bf16 conversions.
"""
from __future__ import annotations

import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "var_b200" / "libvar_b200.so"
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU", "F2FP", "FFMA2", "HMMA",
       "LDG", "STG", "LDS", "STS", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        per[cur]["_total"] += 1
        base = op.split(".")[0]
        per[cur][base] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            per[cur]["UTCHMMA.2CTA"] += 1
        if op.startswith("UTCBAR") and "MULTICAST" in op:
            per[cur]["UTCBAR.MULTICAST"] += 1
        if op.startswith("UTMALDG") and ".2CTA" in op:
            per[cur]["UTMALDG.2CTA"] += 1
    names = demangle(list(per))
    cols = OPS + ["UTCBAR.MULTICAST", "UTMALDG.2CTA"]
    print(f"# {LIB.name}: SASS opcode counts per kernel (sm_100a), {len(per)} kernels")
    print("# " + " ".join(cols) + " | total instructions | kernel")
    tot = collections.Counter()
    for k, c in per.items():
        tot.update(c)
        short = re.sub(r"\(.*", "", names.get(k, k))
        print(" ".join(f"{c.get(o, 0):5d}" for o in cols) + f" | {c['_total']:6d} | {short}")
    print("# totals: " + ", ".join(f"{o}={tot.get(o, 0)}" for o in cols))


if __name__ == "__main__":
    sys.exit(main())
