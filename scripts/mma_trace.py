"""GPU: where the MMA-issuing thread of conv_tc_kernel spends its time (ACCFLOW_TC_DEBUG=16 trace of CTA 0).
Three clock64 stamps per weight tile: ring barriers passed, last MMA issued, commit issued.  Reported per layer:
the per-weight-tile percentiles and the time budget of the whole CTA (issue / commit / gaps inside a tile / gaps at
tile boundaries, i.e. the accumulator hand-over), so that `tensor work / span` can be compared with ncu's tensor %."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ACCFLOW_TC_DEBUG"] = str(16 | int(os.environ.get("TRACE_EXTRA", "0")))
import numpy as np
import torch
from accflow_b200 import _lib as L
from accflow_b200.engine import Kernels, PackedConv, View

torch.set_grad_enabled(False)
K = Kernels(torch.device("cuda:0"), os.environ.get("TRACE_PREC", "fp16x2"))
B, h, w = int(os.environ.get("TRACE_PAIRS", "18")), 64, 64
g = torch.Generator().manual_seed(0)
mk = lambda c: View(torch.randn(B, h, w, c, generator=g).cuda())
wz = lambda cin, cout, kh, kw: PackedConv([(torch.randn(cout, cin, kh, kw, generator=g) * 0.03).cuda()], [torch.zeros(cout).cuda()], 1, (kh // 2, kw // 2))


def run_zr():
    hid, mf, rh, z, pre = mk(128), mk(128), mk(128), mk(128), mk(256)
    for v in (hid, mf, rh):
        K.ensure_planes(v)
    pc = wz(256, 256, 1, 5)
    return lambda: K.conv(pc, [hid, mf], epilogue=L.EPI_GRU_ZR, h=hid, z=z, out2=rh, planes_only=True, pre_add=pre), 20


def run_q():
    hid, mf, rh, z, pre = mk(128), mk(128), mk(128), mk(128), mk(128)
    for v in (hid, mf, rh):
        K.ensure_planes(v)
    pc = wz(256, 128, 5, 1)
    return lambda: K.conv(pc, [rh, mf], epilogue=L.EPI_GRU_Q, h=hid, z=z, pre_add=pre), 20


def run_plain(cin, cout, kh, kw):
    def f():
        x, out = mk(cin), View(torch.empty(B, h, w, cout, device="cuda"))
        K.ensure_planes(x); K.planes_ptr(out, create=True)
        pc = wz(cin, cout, kh, kw)
        nk = -(-cin // 64) * kh * kw
        return (lambda: K.conv(pc, [x], out, act=L.ACT_RELU, planes_only=True)), nk * (2 if cout > 128 else 1) // (2 if cout > 128 else 1)
    return f


cases = {"gru_zr 1x5 256->256": run_zr, "gru_q 5x1 256->128": run_q, "convc2 3x3 256->192": run_plain(256, 192, 3, 3),
         "convc1 1x1 324->256": run_plain(324, 256, 1, 1), "fh1 3x3 128->256": run_plain(128, 256, 3, 3)}
for name, mkcase in cases.items():
    fn, per_tile = mkcase()
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    n = 3 * 1024
    buf = (C.c_longlong * n)()
    L.call("accflow_tc_debug_trace", C.cast(buf, C.c_void_p), n)
    t = np.array(buf[:], dtype=np.int64).reshape(-1, 3)
    t = t[(t[:, 0] > 0)]
    k = 1          # drop anything after the first non-monotonic stamp (stale entries of an earlier, longer launch)
    while k < len(t) and t[k, 0] > t[k - 1, 0]:
        k += 1
    t = t[:k]
    issue, commit = t[:, 1] - t[:, 0], t[:, 2] - t[:, 1]
    gap = t[1:, 0] - t[:-1, 2]
    period = t[1:, 0] - t[:-1, 0]
    boundary = (np.arange(1, len(t)) % per_tile) == 0          # gap that precedes the first weight tile of a CTA tile
    span = int(t[-1, 2] - t[0, 0])
    q = lambda a: [int(np.percentile(a, p)) for p in (10, 50, 90)] if len(a) else []
    print(json.dumps({"conv": name, "pairs": B, "weight_tiles_traced": int(len(t)), "weight_tiles_per_cta_tile": per_tile,
                      "clk_p10_p50_p90": {"barriers_passed->last_mma_issued": q(issue), "commit": q(commit),
                                          "gap_inside_tile": q(gap[~boundary]), "gap_at_tile_boundary": q(gap[boundary]),
                                          "period": q(period)},
                      "budget_clk": {"span": span, "issue": int(issue.sum()), "commit": int(commit.sum()),
                                     "gaps_inside_tiles": int(gap[~boundary].sum()), "gaps_at_tile_boundaries": int(gap[boundary].sum())},
                      "mean_period": float(period.mean())}), flush=True)
