"""GPU: time the correlation lookup at bench shape (planes-only output, coordinates ~ grid + N(0, 4 px));
ACCFLOW_LOOKUP=pairs / fast select the earlier kernels for an A/B run."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from accflow_b200.engine import Kernels, View

torch.set_grad_enabled(False)
if os.environ.get("L2_FETCH"):          # experiment: cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes, default 64)
    from cuda.bindings import runtime as rt
    torch.zeros(1, device="cuda")
    print("setlimit", rt.cudaDeviceSetLimit(rt.cudaLimit.cudaLimitMaxL2FetchGranularity, int(os.environ["L2_FETCH"])),
          rt.cudaDeviceGetLimit(rt.cudaLimit.cudaLimitMaxL2FetchGranularity), file=sys.stderr)
B, h, w = int(os.environ.get("PAIRS", "18")), 64, 64
P = h * w
for prec in ("fp16x2", "fp16"):
    K = Kernels(torch.device("cuda:0"), prec)
    lv = [torch.randn(B * P, (h >> l) * (w >> l), device="cuda") for l in range(4)]
    grid = torch.stack(torch.meshgrid(torch.arange(w), torch.arange(h), indexing="xy"), -1).reshape(1, P, 2).float().cuda()
    coords = grid + torch.randn(B, P, 2, device="cuda") * 4
    corr, flow, mf = View(torch.empty(B, h, w, 324, device="cuda")), torch.empty(B, P, 2, device="cuda"), View(torch.empty(B, h, w, 128, device="cuda"))
    K.planes_ptr(corr, create=True); K.planes_ptr(mf, create=True)
    fn = lambda: K.corr_lookup(lv, 4, coords, corr, flow, mf.ch(126, 128), planes_only=True)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    algo = B * 11.9e6          # 11.9 MB per pair-iteration
    print(json.dumps({"kernel": os.environ.get("ACCFLOW_LOOKUP", "sep"), "l2_fetch": os.environ.get("L2_FETCH"), "precision": prec, "pairs": B, "us": round(us, 1),
                      "algorithmic_GBps": round(algo / us / 1e3, 1)}), flush=True)
