"""Measurement aid: time the adaLN LayerNorm-modulate kernel alone (default: d30 AR scales of a B=256 CFG batch)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from var_b200 import lib as L  # noqa: E402

depth, n_seq = (int(a) for a in sys.argv[1:3]) if len(sys.argv) > 2 else (30, 512)
C = 64 * depth
lib = L.load()
for l in (1, 16, 36, 64, 100, 169, 256, 680):
    M = n_seq * l
    x = torch.randn(M, C, device="cuda")
    ada = torch.randn(n_seq, 6 * C, device="cuda")
    out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    def run():
        L.check(lib.var_b200_ln_modulate(x.data_ptr(), ada.data_ptr(), ada[:, C:].data_ptr(), 6 * C, l, out.data_ptr(), M, C,
                                         1e-6, L.current_stream()))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20 * 1e-3
    print(f"l={l:4d} M={M:7d} C={C}: {t * 1e6:8.1f} us  {M * C * 6 / t / 1e12:5.2f} TB/s")
    del x, out
