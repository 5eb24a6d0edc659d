import sys, torch
from pathlib import Path
sys.path.insert(0, "/root/repo")
from var_b200 import lib as L
lib = L.load()
B, l, V = 256, 256, 4096
logits = torch.randn(2 * B, l, V, device="cuda")
q = torch.empty(B * l, V, device="cuda").exponential_(1.0)
idx = torch.empty(B, l, dtype=torch.int64, device="cuda")
for tk, tp in ((900, 0.0), (900, 0.95), (0, 0.95)):
    f = lambda: L.check(lib.var_b200_cfg_topk_sample(logits.data_ptr(), B, l, V, 1, 1.5, q.data_ptr(), tk, tp, idx.data_ptr(), None, L.current_stream()))
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); f(); e1.record(); torch.cuda.synchronize()
    print(f"top_k={tk} top_p={tp}: {e0.elapsed_time(e1)/2:.2f} ms for {B*l} rows")
