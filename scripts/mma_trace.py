"""GPU: where the MMA-issuing thread of conv_tc_kernel spends its time (ACCFLOW_TC_DEBUG=16 trace of CTA 0)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ACCFLOW_TC_DEBUG"] = str(16 | int(os.environ.get("TRACE_EXTRA", "0")))
import numpy as np
import torch
from accflow_b200 import _lib as L
from accflow_b200.engine import Kernels, PackedConv, View

torch.set_grad_enabled(False)
K = Kernels(torch.device("cuda:0"), os.environ.get("TRACE_PREC", "fp16x2"))
B, h, w = 8, 64, 64
for name, cins, cout, kh, kw in (("gru_zr 1x5 384->256", [128, 128, 128], 256, 1, 5), ("convc2 3x3 256->192", [256], 192, 3, 3),
                                 ("convc1 1x1 324->256", [324], 256, 1, 1)):
    g = torch.Generator().manual_seed(0)
    srcs = [View(torch.randn(B, h, w, c, generator=g).cuda()) for c in cins]
    wt = (torch.randn(cout, sum(cins), kh, kw, generator=g) * 0.05).cuda()
    pc = PackedConv([wt], [torch.zeros(cout).cuda()], 1, (kh // 2, kw // 2))
    out = View(torch.empty(B, h, w, cout, device="cuda"))
    for _ in range(5):
        K.conv(pc, srcs, out, act=L.ACT_RELU)
    torch.cuda.synchronize()
    n = 3 * 1024
    buf = (C.c_longlong * n)()
    L.call("accflow_tc_debug_trace", C.cast(buf, C.c_void_p), n)
    t = np.array(buf[:], dtype=np.int64).reshape(-1, 3)
    t = t[(t[:, 0] > 0)]
    # drop anything after the first non-monotonic stamp (stale entries of an earlier, longer launch)
    k = 1
    while k < len(t) and t[k, 0] > t[k - 1, 0]:
        k += 1
    t = t[:k]
    issue = t[:, 1] - t[:, 0]
    commit = t[:, 2] - t[:, 1]
    gap = t[1:, 0] - t[:-1, 2]
    period = t[1:, 0] - t[:-1, 0]
    q = lambda a: [int(np.percentile(a, p)) for p in (10, 50, 90)]
    print(json.dumps({"conv": name, "tiles_traced": int(len(t)), "clk_p10_p50_p90": {
        "barriers_passed->last_mma_issued": q(issue), "commit": q(commit),
        "commit->next_barriers_passed": q(gap), "period": q(period)}}))
