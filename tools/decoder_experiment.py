import torch, time, copy, sys
sys.path.insert(0, "/root/repo")
from var_b200 import build_vae_var
from var_b200.init_utils import dense_init_
torch.backends.cudnn.benchmark = True
vae, var = build_vae_var("cuda", depth=2); dense_init_(vae, 1)
B = 128
f = torch.randn(B, 32, 16, 16, device="cuda")
def timeit(fn, n=3):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n
dec16 = copy.deepcopy(vae.decoder).bfloat16(); post16 = copy.deepcopy(vae.post_quant_conv).bfloat16()
print("bf16 NCHW      ms/img", 1e3 * timeit(lambda: dec16(post16(f.bfloat16()))) / B)
dcl = copy.deepcopy(dec16).to(memory_format=torch.channels_last); pcl = copy.deepcopy(post16).to(memory_format=torch.channels_last)
print("bf16 NHWC      ms/img", 1e3 * timeit(lambda: dcl(pcl(f.bfloat16().contiguous(memory_format=torch.channels_last)))) / B)
dec_h = copy.deepcopy(vae.decoder).half(); post_h = copy.deepcopy(vae.post_quant_conv).half()
print("fp16 NCHW      ms/img", 1e3 * timeit(lambda: dec_h(post_h(f.half()))) / B)
with torch.autocast("cuda", dtype=torch.bfloat16):
    print("autocast bf16  ms/img", 1e3 * timeit(lambda: vae.decoder(vae.post_quant_conv(f))) / B)
# per-stage profile of the bf16 NCHW decoder
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    dec16(post16(f.bfloat16())); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))
