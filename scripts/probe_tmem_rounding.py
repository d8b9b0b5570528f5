"""Experiment (GPU): how does tcgen05.mma round when it adds into an fp32 TMEM accumulator?

A 1x1 conv with a = 1 everywhere; w[k=0] = 2^24 and w[k = 16, 32, ...] = 1.5 (all bf16-exact),
so each later K=16 MMA adds exactly 1.5 to an accumulator holding 2^24 (ulp = 2):
  round-to-nearest  -> +2 per MMA        truncation -> stuck at 2^24       wide accumulator -> exact sum rounded once
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from accflow_b200 import _lib as L
from accflow_b200.engine import Kernels, PackedConv, View

torch.set_grad_enabled(False)
K = Kernels(torch.device("cuda:0"), "bf16")
for kdim in (64, 256, 1024):
    w = torch.zeros(32, kdim, 1, 1)
    w[:, 0] = 2.0 ** 24
    w[:, 16::16] = 1.5
    w[1, 16::16] = -1.5
    w[2, 16::16] = 0.5            # below half-ulp: lost under RN and RZ alike
    w[3, 16::16] = -0.5           # truncation would drop a full ulp each time
    x = torch.ones(1, 4, 4, kdim).cuda()
    out = torch.empty(1, 4, 4, 32, device="cuda")
    K.conv(PackedConv([w.cuda()], [None]), [View(x)], View(out))
    torch.cuda.synchronize()
    n = kdim // 16 - 1
    r = out[0, 0, 0, :4].double().cpu() - 2.0 ** 24
    print(f"K={kdim}: {n} adds of +1.5/-1.5/+0.5/-0.5 onto 2^24 -> deltas {r.tolist()} "
          f"(RN: {2*n}, {-2*n}, 0, 0   RZ: 0, {-2*n}, 0, {-2*n}   exact: {1.5*n}, {-1.5*n}, {0.5*n}, {-0.5*n})")
