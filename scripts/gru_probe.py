"""GPU experiment: what bounds the GRU convolutions (and a plain 3x3) of conv_tc_kernel?
Times each variant with parts of the kernel disabled (ACCFLOW_TC_DEBUG: 1 = no TMA loads, 2 = no MMAs; results are
garbage with any bit set): dbg32 = the epilogue alone (no main loop), dbg3 = epilogue + the main loop's barrier skeleton,
dbg2 = + operand traffic, dbg0 = everything."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from accflow_b200 import _lib as L
from accflow_b200.engine import Kernels, PackedConv, View

torch.set_grad_enabled(False)
K = Kernels(torch.device("cuda:0"), os.environ.get("PROBE_PREC", "fp16x2"))
h = w = 64
_w = torch.randn(4096, 4096, device="cuda")
for _ in range(60):
    _w @ _w
torch.cuda.synchronize()


def timeit(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


for B in (int(x) for x in os.environ.get("PROBE_PAIRS", "18,27").split(",")):
    g = torch.Generator().manual_seed(0)
    mk = lambda c: View(torch.randn(B, h, w, c, generator=g).cuda())
    hid, inp, mf, rh, z = mk(128), mk(128), mk(128), mk(128), mk(128)
    pre_zr, pre_q = mk(256), mk(128)
    x256 = mk(256)
    out192 = View(torch.empty(B, h, w, 192, device="cuda"))
    for v in (hid, inp, mf, rh, x256):
        K.ensure_planes(v)
    K.planes_ptr(out192, create=True)
    wz = lambda cin, cout, kh, kw: PackedConv([(torch.randn(cout, cin, kh, kw, generator=g) * 0.03).cuda()],
                                                [torch.zeros(cout).cuda()], 1, (kh // 2, kw // 2))
    variants = {
        "zr 1x5 384->256 (reference form)": lambda: K.conv(zr384, [hid, inp, mf], epilogue=L.EPI_GRU_ZR, h=hid, z=z, out2=rh, planes_only=True),
        "zr 1x5 256->256 + pre_add": lambda: K.conv(zr256, [hid, mf], epilogue=L.EPI_GRU_ZR, h=hid, z=z, out2=rh, planes_only=True, pre_add=pre_zr),
        "zr 1x5 256->256 no pre_add": lambda: K.conv(zr256, [hid, mf], epilogue=L.EPI_GRU_ZR, h=hid, z=z, out2=rh, planes_only=True),
        "q 5x1 384->128 (reference form)": lambda: K.conv(q384, [rh, inp, mf], epilogue=L.EPI_GRU_Q, h=hid, z=z),
        "q 5x1 256->128 + pre_add": lambda: K.conv(q256, [rh, mf], epilogue=L.EPI_GRU_Q, h=hid, z=z, pre_add=pre_q),
        "convc2 3x3 256->192 relu planes-only": lambda: K.conv(c2, [x256], out192, act=L.ACT_RELU, planes_only=True),
        "plain 1x5 256->256 relu planes-only": lambda: K.conv(zr256, [hid, mf], pre_zr_out, act=L.ACT_RELU, planes_only=True),
        "convf1 1x1 98->128 relu planes-only": lambda: K.conv(f1, [x98], out128, act=L.ACT_RELU, planes_only=True),
        "convc1 1x1 324->256 relu planes-only": lambda: K.conv(c1, [x324], pre_zr_out, act=L.ACT_RELU, planes_only=True),
        "enc1 3x3 64->64 @256x256 x9 fp32 out": lambda: K.conv(e1, [x64], out64, emit_planes=False),
    }
    if os.environ.get("PROBE_ONLY"):
        variants = {k: v for k, v in variants.items() if os.environ["PROBE_ONLY"] in k}
    zr384, zr256 = wz(384, 256, 1, 5), wz(256, 256, 1, 5)
    q384, q256 = wz(384, 128, 5, 1), wz(256, 128, 5, 1)
    c2 = wz(256, 192, 3, 3)
    f1, c1, e1 = wz(98, 128, 1, 1), wz(324, 256, 1, 1), wz(64, 64, 3, 3)
    x98, x324 = View(torch.randn(B, h, w, 104, generator=g).cuda()).ch(0, 98), mk(324)
    x64, out64 = View(torch.randn(9, 256, 256, 64, generator=g).cuda()), View(torch.empty(9, 256, 256, 64, device="cuda"))
    out128 = mk(128)
    for v in (x98, x324, x64):
        K.ensure_planes(v)
    K.planes_ptr(out128, create=True)
    pre_zr_out = mk(256)
    K.planes_ptr(pre_zr_out, create=True)
    for name, fn in variants.items():
        row = {"pairs": B, "conv": name}
        for dbg in (0, 2, 1, 3, 32):
            os.environ["ACCFLOW_TC_DEBUG"] = str(dbg)
            try:
                row[f"us_dbg{dbg}"] = round(timeit(fn), 1)
            except Exception as e:      # noqa: BLE001
                row[f"us_dbg{dbg}"] = str(e)[:60]
        os.environ["ACCFLOW_TC_DEBUG"] = "0"
        print(json.dumps(row), flush=True)
