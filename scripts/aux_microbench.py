"""GPU microbenchmark of the non-GEMM kernels at bench shapes (CUDA-event timed, warm L2 -> lower bound)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from accflow_b200 import _lib as L
from accflow_b200.engine import Kernels, PackedConv, View

torch.set_grad_enabled(False)
K = Kernels(torch.device("cuda:0"), "fp16x2")
B, h, w = int(os.environ.get("MB_PAIRS", "8")), 64, 64
P = h * w
dev = "cuda"

ONCE = os.environ.get("MB_ONCE") == "1"   # under ncu: one launch per kernel, no timing loop

def timeit(name, fn, bytes_moved, n=20):
    if ONCE:
        fn(); torch.cuda.synchronize(); print(json.dumps({"kernel": name, "MB": round(bytes_moved / 1e6, 1)})); return
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    z.record(); torch.cuda.synchronize()
    us = a.elapsed_time(z) * 1e3 / n
    print(json.dumps({"kernel": name, "us": round(us, 1), "GB/s": round(bytes_moved / us / 1e3, 1), "MB": round(bytes_moved / 1e6, 1)}))

# lookup
lv = [torch.randn(B * P, (h >> l) * (w >> l), device=dev) for l in range(4)]
coords = torch.rand(B, P, 2, device=dev) * 60 + 2
corr = View(torch.empty(B, h, w, 324, device=dev)); flow = torch.empty(B, P, 2, device=dev); mf = View(torch.empty(B, h, w, 128, device=dev))
K.planes_ptr(corr, create=True); K.planes_ptr(mf, create=True)
timeit("corr_lookup (+planes)", lambda: K.corr_lookup(lv, 4, coords, corr, flow, mf.ch(126, 128)), B * P * (324 * 8 + 4 * 144 * 4))
K2 = Kernels(torch.device("cuda:0"), "fp32")
timeit("corr_lookup (fp32 only)", lambda: K2.corr_lookup(lv, 4, coords, corr, flow, mf.ch(126, 128)), B * P * (324 * 4 + 4 * 144 * 4))
# smallcout
x = View(torch.randn(B, h, w, 256, device=dev)); wt = torch.randn(2, 256, 3, 3, device=dev) * 0.02
pc = PackedConv([wt], [torch.zeros(2, device=dev)], 1, (1, 1)); out = View(torch.empty(B, h, w, 2, device=dev))
timeit("conv3x3_smallcout 256->2", lambda: K.conv_smallcout(pc, x, out), B * P * 256 * 4)
# split planes
y = View(torch.randn(B, h, w, 256, device=dev))
def split():
    K.wrote(y); K.ensure_planes(y)
K.ensure_planes(y)
timeit("split_planes 256ch", split, B * P * 256 * 8)
# flow patch
fl = torch.randn(B, P, 2, device=dev); patch = View(torch.empty(B, h, w, 104, device=dev)); pl = K.planes_ptr(patch, create=True)
timeit("flow_patch", lambda: L.call("accflow_flow_patch_f32", fl.data_ptr(), B, h, w, patch.ptr, 104, pl[0], pl[1], pl[2], 2, None), B * P * 104 * 8)
# stem patch (4 images 512x512)
img = torch.randn(4, 3, 512, 512, device=dev); pp = torch.empty(3, 4, 256, 256, 152, device=dev, dtype=torch.bfloat16)
timeit("stem_patch 4x512x512", lambda: L.call("accflow_stem_patch_planes", img.data_ptr(), 4, 512, 512, pp.data_ptr(), 152, 4 * 65536 * 152, 2, None), 4 * 65536 * 152 * 4)
# instnorm 12 x 256x256x64
t = View(torch.randn(12, 256, 256, 64, device=dev))
timeit("instnorm 12x256x256x64 (3 kernels)", lambda: K.instnorm(t, True, None, False, t), 12 * 65536 * 64 * 4 * 3)
# corr pool
l1, l2, l3 = (torch.empty(B * P, (h >> l) * (w >> l), device=dev) for l in (1, 2, 3))
timeit("corr_pool", lambda: L.call("accflow_corr_pool_f32", lv[0].data_ptr(), B * P, h, w, l1.data_ptr(), l2.data_ptr(), l3.data_ptr(), None), B * P * P * 4 * 1.33)
timeit("corr_pool levels 2-3 from level 1 (fused path)", lambda: L.call("accflow_corr_pool_f32", l1.data_ptr(), B * P, h // 2, w // 2, l2.data_ptr(), l3.data_ptr(), None, None), B * P * (P // 4) * 4 * 1.3125)

# convex upsample (flow + 576-channel mask -> 8x flow), once per pair
cflow = torch.randn(B, P, 2, device=dev); mask = torch.randn(B, h, w, 576, device=dev); up = torch.empty(B, 2, 8 * h, 8 * w, device=dev)
timeit("convex_upsample", lambda: L.call("accflow_convex_upsample_f32", cflow.data_ptr(), 2, 1, mask.data_ptr(), 576, B, h, w, up.data_ptr(), None),
       B * P * (576 * 4 + 8 + 128 * 4))
# warp + occlusion (getOcc): binary mask and error map
c1 = torch.randn(B, h, w, 128, device=dev); c2 = torch.randn(B, h, w, 128, device=dev); wf = torch.randn(B, h, w, 2, device=dev) * 3
occ = torch.empty(B, h, w, 1, device=dev); emap = torch.empty(B, h, w, 128, device=dev)
timeit("warp_occ (binary)", lambda: L.call("accflow_warp_occ_f32", c1.data_ptr(), 128, c2.data_ptr(), 128, wf.data_ptr(), B, h, w, 128, occ.data_ptr(), 1, None, 0, None),
       B * P * (128 * 4 * 2 + 8 + 4))
timeit("warp_occ (emap)", lambda: L.call("accflow_warp_occ_f32", c1.data_ptr(), 128, c2.data_ptr(), 128, wf.data_ptr(), B, h, w, 128, None, 0, emap.data_ptr(), 128, None),
       B * P * (128 * 4 * 3 + 8))
# correlation volume GEMM (fmap1 . fmap2^T / sqrt(D)), level 0 of the pyramid
f1 = View(torch.randn(B, h, w, 256, device=dev)); f2 = View(torch.randn(B, h, w, 256, device=dev))
vol = View(lv[0].view(B, h, w, P))
K.ensure_planes(f1); K.ensure_planes(f2)
timeit("corr_gemm (fp16x2) [GFLOP in MB field]", lambda: K.gemm_nt("mb.corr", f1, f2, vol, alpha=1.0 / 16), B * 2.0 * P * P * 256 / 1e3)
