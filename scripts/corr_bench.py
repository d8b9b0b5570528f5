"""GPU: correlation pyramid (GEMM + fused level 1 + pool kernel) timing, TMA-store epilogue on/off."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from accflow_b200.engine import FlowEstimatorEngine, Kernels, View

torch.set_grad_enabled(False)
K = Kernels(torch.device("cuda:0"), "fp16x2")
eng = FlowEstimatorEngine.__new__(FlowEstimatorEngine)
eng.k = K
_w = torch.randn(4096, 4096, device="cuda")
for _ in range(40):
    _w @ _w
for B in (8, 27):
    g = torch.Generator().manual_seed(0)
    f1 = View(torch.randn(B, 64, 64, 256, generator=g).cuda()); f2 = View(torch.randn(B, 64, 64, 256, generator=g).cuda())
    ref = None
    for mode, mm in (("1", "1"), ("0", "1"), ("1", "0"), ("0", "0")):
        os.environ["ACCFLOW_TC_TMA_STORE"] = mode
        os.environ["ACCFLOW_TC_M_MAJOR"] = mm
        for _ in range(3):
            lv = eng.corr_pyramid(f1, f2, f"cb{B}")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lv = eng.corr_pyramid(f1, f2, f"cb{B}")
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        out = [t.clone() for t in lv]
        if ref is None:
            ref = out
        diff = max(float((a - b).abs().max()) for a, b in zip(out, ref))
        mb = B * 4096 * 4096 * 4 * 1.3125 / 1e6
        print(json.dumps({"pairs": B, "tma_store": mode, "m_major": mm, "us_pyramid": round(us, 1), "GBps_written": round(mb / us * 1e3 / 1e3, 1), "max_diff_vs_tma": diff}), flush=True)
