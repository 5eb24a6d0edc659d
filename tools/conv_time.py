"""Measurement aid: 3x3 convolution layers of the VQVAE decoder (B=64), own implicit GEMM vs cuDNN (channels_last bf16)."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from var_b200 import lib as L  # noqa: E402
from test_gemm_gpu import pack_conv3x3  # noqa: E402

torch.backends.cudnn.benchmark = True
lib = L.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for H, Cin, Cout in ((16, 640, 640), (32, 640, 320), (32, 320, 320), (64, 320, 320), (128, 320, 160), (128, 160, 160), (256, 160, 160)):
    x = torch.randn(B, H, H, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(9 * Cin)).bfloat16()
    wp = pack_conv3x3(w)
    bias = torch.zeros(Cout, device="cuda")
    out = torch.empty(B, H, H, Cout, device="cuda", dtype=torch.bfloat16)
    xc = x.permute(0, 3, 1, 2)  # NCHW view of channels_last memory
    wc = w.contiguous(memory_format=torch.channels_last)

    def ours():
        L.check(lib.var_b200_conv3x3_nhwc(x.data_ptr(), wp.data_ptr(), bias.data_ptr(), None, out.data_ptr(), B, H, H, Cin, Cout,
                                          L.current_stream()))

    def cudnn():
        return torch.nn.functional.conv2d(xc, wc, None, padding=1)

    res = []
    for fn in (ours, cudnn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 10 * 1e-3)
    fl = 2.0 * B * H * H * Cout * Cin * 9
    print(f"B={B} {H}x{H} {Cin}->{Cout}: ours {res[0] * 1e6:8.1f} us {fl / res[0] / 1e12:7.1f} TF/s | cuDNN {res[1] * 1e6:8.1f} us {fl / res[1] / 1e12:7.1f} TF/s")
    del x, out
