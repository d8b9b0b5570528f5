"""GPU microbenchmark of the convolution kernels on the update-block shapes (CUDA-event timed)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from accflow_b200 import _lib as L
from accflow_b200.engine import Kernels, PackedConv, View

torch.set_grad_enabled(False)
B = int(os.environ.get("MB_PAIRS", "6"))
h = w = 64
SHAPES = [  # name, cins, cout, kh, kw
    ("gru_zr 1x5 384->256", [128, 128, 128], 256, 1, 5),
    ("convc2 3x3 256->192", [256], 192, 3, 3),
    ("convc1 1x1 324->256", [324], 256, 1, 1),
    ("flowhead2 3x3 256->2", [256], 2, 3, 3),
    ("gru_q 5x1 384->128", [128, 128, 128], 128, 5, 1),
    ("convm 3x3 256->126", [256], 126, 3, 3),
    ("enc1 3x3 64->64 @256", [64], 64, 3, 3, 256),
    ("enc2 3x3 96->96 @128", [96], 96, 3, 3, 128),
]
# warm the clocks up before the first timing (a cold GPU ramps for ~100 ms)
_w = torch.randn(4096, 4096, device="cuda")
for _ in range(int(os.environ.get("MB_WARM", "60"))):
    _w @ _w
torch.cuda.synchronize()
only = os.environ.get("MB_ONLY")
iters = int(os.environ.get("MB_ITERS", "20"))
for prec in os.environ.get("MB_PREC", "fp32,bf16x3,bf16").split(","):
    K = Kernels(torch.device("cuda:0"), prec)
    for name, cins, cout, kh, kw, *rest in SHAPES:
        if only and only not in name:
            continue
        h = w = rest[0] if rest else 64
        g = torch.Generator().manual_seed(0)
        srcs = [View(torch.randn(B, h, w, c, generator=g).cuda()) for c in cins]
        wt = (torch.randn(cout, sum(cins), kh, kw, generator=g) * 0.05).cuda()
        pc = PackedConv([wt], [torch.zeros(cout).cuda()], 1, (kh // 2, kw // 2))
        out = View(torch.empty(B, h, w, cout, device="cuda"))
        for _ in range(10):
            K.conv(pc, srcs, out, act=L.ACT_RELU)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            K.conv(pc, srcs, out, act=L.ACT_RELU)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        flop = 2.0 * B * h * w * cout * sum(cins) * kh * kw
        print(json.dumps({"precision": prec, "conv": name, "pairs": B, "us": round(us, 1), "tflops": round(flop / us / 1e6, 1)}))
