"""Measurement aid (not product code): time the CNN decoder variants and list their top CUDA kernels."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from var_b200 import build_vae_var  # noqa: E402
from var_b200.init_utils import dense_init_  # noqa: E402

torch.backends.cudnn.benchmark = True
vae, var = build_vae_var("cuda", depth=2)
dense_init_(vae, 1)
B = 128
f = torch.randn(B, 32, 16, 16, device="cuda")


def timeit(fn, n=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n


vae.decoder_dtype = torch.bfloat16
for nhwc, own in ((False, False), (True, False), (True, True)):
    vae.decoder_nhwc, vae.decoder_own_conv = nhwc, own
    print(f"bf16 decoder nhwc_plan={nhwc} own_conv={own}: {1e3 * timeit(lambda: vae.fhat_to_img(f)) / B:.3f} ms/img")
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    vae.fhat_to_img(f)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=80))
