"""One launch of each kernel of interest at bench shape between cudaProfilerStart/Stop, for
  ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/X python scripts/ncu_targets.py <targets>
Targets: gru_zr gru_q convc2 enc1 stem convc1 lookup corr_gemm attn agg"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from accflow_b200 import _lib as L
from accflow_b200.engine import FlowEstimatorEngine, Kernels, PackedConv, PlanesOnly, View

torch.set_grad_enabled(False)
targets = sys.argv[1:] or ["gru_zr", "convc2", "lookup", "corr_gemm"]
K = Kernels(torch.device("cuda:0"), os.environ.get("NCU_PREC", "fp16x2"))
B, h, w = int(os.environ.get("NCU_PAIRS", "18")), 64, 64
g = torch.Generator().manual_seed(0)
mk = lambda b, hh, ww, c: View(torch.randn(b, hh, ww, c, generator=g).cuda())
wz = lambda cin, cout, kh, kw: PackedConv([(torch.randn(cout, cin, kh, kw, generator=g) * 0.03).cuda()], [torch.zeros(cout).cuda()], 1, (kh // 2, kw // 2))
cudart = torch.cuda.cudart()


def profiled(fn, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    cudart.cudaProfilerStart()
    fn()
    torch.cuda.synchronize()
    cudart.cudaProfilerStop()


for t in targets:
    if t in ("gru_zr", "gru_q"):
        hid, mf, rh, z = mk(B, h, w, 128), mk(B, h, w, 128), mk(B, h, w, 128), mk(B, h, w, 128)
        for v in (hid, mf, rh):
            K.ensure_planes(v)
        if t == "gru_zr":
            pc, pre = wz(256, 256, 1, 5), mk(B, h, w, 256)
            profiled(lambda: K.conv(pc, [hid, mf], epilogue=L.EPI_GRU_ZR, h=hid, z=z, out2=rh, planes_only=True, pre_add=pre))
        else:
            pc, pre = wz(256, 128, 5, 1), mk(B, h, w, 128)
            profiled(lambda: K.conv(pc, [rh, mf], epilogue=L.EPI_GRU_Q, h=hid, z=z, pre_add=pre))
    elif t == "convc2":
        x, out = mk(B, h, w, 256), View(torch.empty(B, h, w, 192, device="cuda"))
        K.ensure_planes(x); K.planes_ptr(out, create=True)
        pc = wz(256, 192, 3, 3)
        profiled(lambda: K.conv(pc, [x], out, act=L.ACT_RELU, planes_only=True))
    elif t == "convc1":
        x, out = mk(B, h, w, 324), View(torch.empty(B, h, w, 256, device="cuda"))
        K.ensure_planes(x); K.planes_ptr(out, create=True)
        pc = wz(324, 256, 1, 1)
        profiled(lambda: K.conv(pc, [x], out, act=L.ACT_RELU, planes_only=True))
    elif t == "enc1":
        x, out = mk(9, 256, 256, 64), View(torch.empty(9, 256, 256, 64, device="cuda"))
        K.ensure_planes(x)
        pc = wz(64, 64, 3, 3)
        profiled(lambda: K.conv(pc, [x], out, emit_planes=False))
    elif t == "lookup":
        P = h * w
        lv = [torch.randn(B * P, (h >> l) * (w >> l), device="cuda") for l in range(4)]
        coords = (torch.rand(B, P, 2, device="cuda") * 8 - 4) + torch.stack(torch.meshgrid(torch.arange(w), torch.arange(h), indexing="xy"), -1).reshape(1, P, 2).float().cuda()
        corr, flow, mfv = View(torch.empty(B, h, w, 324, device="cuda")), torch.empty(B, P, 2, device="cuda"), View(torch.empty(B, h, w, 128, device="cuda"))
        K.planes_ptr(corr, create=True); K.planes_ptr(mfv, create=True)
        profiled(lambda: K.corr_lookup(lv, 4, coords, corr, flow, mfv.ch(126, 128), planes_only=True))
    elif t == "corr_gemm":
        eng = FlowEstimatorEngine.__new__(FlowEstimatorEngine)
        eng.k = K
        f1, f2 = mk(8, h, w, 256), mk(8, h, w, 256)
        profiled(lambda: (K.wrote(f1), eng.corr_pyramid(f1, f2, "ncu.corr")))
    elif t in ("attn", "agg"):
        eng = FlowEstimatorEngine.__new__(FlowEstimatorEngine)
        eng.k, eng.gma, eng.gamma, eng.qk_scale = K, True, 0.5, 128 ** -0.5
        eng.to_qk, eng.to_v = wz(128, 256, 1, 1), wz(128, 128, 1, 1)
        inp, mf, out = mk(4, h, w, 128), mk(4, h, w, 128), View(torch.zeros(4, h, w, 128, device="cuda"))
        K.planes_ptr(out, create=True)
        attn = eng.attention(inp, "ncu.att")
        if t == "attn":
            profiled(lambda: eng.attention(inp, "ncu.att"))
        else:
            profiled(lambda: eng.aggregate(attn, mf, out, "ncu.att"))
    print("profiled", t, flush=True)
