"""Per-kernel parity (CUDA path through the C ABI vs oracle / golden fixtures). Needs a GPU."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.golden import cases

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def K():
    from accflow_b200.engine import Kernels
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return Kernels(torch.device("cuda:0"))


# tolerance (relative to the output scale ~1) per arithmetic mode of the conv/GEMM kernels
PRECISIONS = {"fp32": 2e-5, "bf16x3": 2e-5, "fp16x2": 2e-5, "bf16": 6e-2, "fp16": 8e-3}


@pytest.fixture(scope="module", params=list(PRECISIONS))
def KP(request):
    from accflow_b200.engine import Kernels
    return Kernels(torch.device("cuda:0"), request.param), PRECISIONS[request.param]


def dev(t):
    return t.cuda()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def maxdiff(a, b):
    a = a.detach().float().cpu()
    b = torch.as_tensor(b).float()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max())


def test_library_loaded_and_counts_launches(K):
    from accflow_b200 import _lib as L
    L.call("accflow_launch_count", 1)
    x = torch.zeros(4, device="cuda")
    L.call("accflow_axpy_f32", x.data_ptr(), x.data_ptr(), 1.0, 4, None)
    assert L.call("accflow_launch_count", 0) == 1


# ------------------------------------------------------------------ generic convolution
CONV_CASES = [
    # (B, cins, H, W, cout, kh, kw, stride, pad_h, pad_w)
    (2, [128], 16, 16, 256, 3, 3, 1, 1, 1),
    (1, [324], 20, 18, 256, 1, 1, 1, 0, 0),
    (2, [128, 128, 128], 16, 24, 256, 1, 5, 1, 0, 2),
    (2, [128, 256], 24, 16, 128, 5, 1, 1, 2, 0),
    (1, [128, 128, 1], 10, 12, 256, 3, 3, 1, 1, 1),      # AccPlus conv1.0: cat[df, f, o]
    (1, [256], 17, 19, 2, 3, 3, 1, 1, 1),                # flow head conv2 (cout 2), ragged map
    (2, [64], 32, 32, 96, 3, 3, 2, 1, 1),                # encoder stride-2 block
    (2, [64], 32, 32, 96, 1, 1, 2, 0, 0),                # encoder downsample
    (1, [128], 9, 7, 27, 3, 3, 1, 1, 1),                 # ZeroConv2d
    (1, [130], 8, 8, 126, 3, 3, 1, 1, 1),                # cin not a multiple of 4 (scalar loader)
    (1, [64], 40, 36, 64, 3, 3, 1, 1, 1),                # encoder layer1 shape, several ragged 8x16 tiles
    (2, [96], 17, 33, 96, 3, 3, 1, 1, 1),                # odd map, cout 96
    (1, [128, 128, 128, 128], 24, 40, 256, 5, 1, 1, 2, 0),   # GMA GRU (4 sources), vertical taps
    (1, [128, 128, 128, 128], 24, 40, 128, 1, 5, 1, 0, 2),   # GMA GRU q conv, horizontal taps (x-major tiles)
    (1, [256], 33, 47, 192, 3, 3, 1, 1, 1),              # convc2: two 96-wide N tiles
    (2, [64], 200, 200, 64, 3, 3, 1, 1, 1),              # enough narrow tiles for two sub-tiles per CTA tile (ragged 8x32)
    (1, [128], 280, 300, 64, 1, 5, 1, 0, 2),             # same, horizontal taps (32x8 tiles)
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_matches_torch(KP, case):
    K, tol = KP
    from accflow_b200 import _lib as L
    from accflow_b200.engine import PackedConv, View
    B, cins, H, W, cout, kh, kw, stride, ph, pw = case
    g = torch.Generator().manual_seed(CONV_CASES.index(case))
    xs = [torch.randn(B, c, H, W, generator=g) for c in cins]
    w = torch.randn(cout, sum(cins), kh, kw, generator=g) / math.sqrt(sum(cins) * kh * kw)
    b = torch.randn(cout, generator=g)
    ref = torch.relu(F.conv2d(torch.cat(xs, 1), w, b, stride=stride, padding=(ph, pw)))
    pc = PackedConv([dev(w)], [dev(b)], stride, (ph, pw))
    srcs = [View(dev(nhwc(x))) for x in xs]
    out = torch.empty(B, ref.shape[2], ref.shape[3], cout, device="cuda")
    K.conv(pc, srcs, View(out), act=L.ACT_RELU)
    torch.cuda.synchronize()
    assert maxdiff(out.permute(0, 3, 1, 2), ref) < tol


def test_conv2d_tile_order_does_not_change_results():
    """accflow_conv_desc.tile_order only changes the order in which a launch walks its output tiles (L2 locality):
    forward, reverse and the alternating default give bit-identical outputs (two N tiles, ragged map, several waves)."""
    from accflow_b200 import _lib as L
    from accflow_b200.engine import Kernels, PackedConv, View
    K = Kernels(torch.device("cuda:0"), "fp16x2")
    g = torch.Generator().manual_seed(77)
    x = View(torch.randn(3, 70, 90, 128, generator=g).cuda())
    w = torch.randn(256, 128, 3, 3, generator=g) / math.sqrt(128 * 9)
    b = torch.randn(256, generator=g)
    pc = PackedConv([dev(w)], [dev(b)], 1, (1, 1))
    outs = []
    for order in (L.TILES_FORWARD, L.TILES_REVERSE, L.TILES_AUTO, L.TILES_AUTO):
        out = torch.zeros(3, 70, 90, 256, device="cuda")
        K.conv(pc, [x], View(out), act=L.ACT_RELU, tile_order=order)
        torch.cuda.synchronize()
        outs.append(out)
    ref = torch.relu(F.conv2d(x.t.permute(0, 3, 1, 2).cpu(), w, b, padding=1))
    assert maxdiff(outs[0].permute(0, 3, 1, 2), ref) < 2e-5
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


def test_conv2d_channel_slices_and_residual(KP):
    K, tol = KP
    """Sources / destinations that are channel slices of wider buffers (the cat-free layout)."""
    from accflow_b200 import _lib as L
    from accflow_b200.engine import PackedConv, View
    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 12, 14
    wide = torch.randn(B, H, W, 200, generator=g).cuda()
    res = torch.randn(B, H, W, 64, generator=g).cuda()
    w = torch.randn(64, 96, 3, 3, generator=g) * 0.05
    b = torch.randn(64, generator=g)
    gam, beta = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g)
    rm, rv = torch.randn(64, generator=g) * 0.1, torch.rand(64, generator=g) + 0.5
    x = wide[..., 40:136].permute(0, 3, 1, 2).cpu()
    y = F.conv2d(x, w, b, padding=1)
    y = F.batch_norm(y, rm, rv, gam, beta, False, 0.0, 1e-5)
    ref = torch.relu(res.cpu().permute(0, 3, 1, 2) + torch.relu(y))
    pc = PackedConv([dev(w)], [dev(b)], 1, (1, 1), bn=(dev(gam), dev(beta), dev(rm), dev(rv), 1e-5))
    outw = torch.zeros(B, H, W, 100, device="cuda")
    K.conv(pc, [View(wide).ch(40, 136)], View(outw).ch(20, 84), act=L.ACT_RELU, residual=View(res), post_relu=True)
    torch.cuda.synchronize()
    assert maxdiff(outw[..., 20:84].permute(0, 3, 1, 2), ref) < tol
    assert float(outw[..., :20].abs().max()) == 0 and float(outw[..., 84:].abs().max()) == 0


def test_gru_epilogues(KP):
    """SepConvGRU half-step (raft/update.py:45-52) through the fused ZR / Q epilogues."""
    K, tol = KP
    from accflow_b200 import _lib as L
    from accflow_b200.engine import PackedConv, View
    g = torch.Generator().manual_seed(4)
    B, H, W = 2, 12, 16
    h = torch.tanh(torch.randn(B, 128, H, W, generator=g))
    x = torch.randn(B, 256, H, W, generator=g)
    ws = [torch.randn(128, 384, 1, 5, generator=g) * 0.03 for _ in range(3)]
    bs = [torch.randn(128, generator=g) * 0.1 for _ in range(3)]
    hx = torch.cat([h, x], 1)
    z = torch.sigmoid(F.conv2d(hx, ws[0], bs[0], padding=(0, 2)))
    r = torch.sigmoid(F.conv2d(hx, ws[1], bs[1], padding=(0, 2)))
    q = torch.tanh(F.conv2d(torch.cat([r * h, x], 1), ws[2], bs[2], padding=(0, 2)))
    ref = (1 - z) * h + z * q
    zr = PackedConv([dev(ws[0]), dev(ws[1])], [dev(bs[0]), dev(bs[1])], 1, (0, 2))
    qc = PackedConv([dev(ws[2])], [dev(bs[2])], 1, (0, 2))
    hv, xv = View(dev(nhwc(h))), View(dev(nhwc(x)))
    zb, rh = View(torch.empty(B, H, W, 128, device="cuda")), View(torch.empty(B, H, W, 128, device="cuda"))
    K.conv(zr, [hv, xv], epilogue=L.EPI_GRU_ZR, h=hv, z=zb, out2=rh)
    K.conv(qc, [rh, xv], epilogue=L.EPI_GRU_Q, h=hv, z=zb)
    torch.cuda.synchronize()
    assert maxdiff(hv.t.permute(0, 3, 1, 2), ref) < tol


@pytest.mark.parametrize("shared", [False, True])
def test_gru_hoisted_input_term(KP, shared):
    """The GRU's constant `inp` columns applied once and passed as the pre-activation addend (pre_add) give the
    same half-step as the reference's single convolution over cat[h, inp, mf] (raft/update.py:45-52).
    ``shared``: samples s and s + B/2 are pairs with the same first frame and read one term (pre_mod)."""
    K, tol = KP
    from accflow_b200 import _lib as L
    from accflow_b200.engine import PackedConv, View
    g = torch.Generator().manual_seed(41)
    B, H, W = (4, 20, 12) if shared else (2, 20, 12)
    h = torch.tanh(torch.randn(B, 128, H, W, generator=g))
    inp = torch.relu(torch.randn(B, 128, H, W, generator=g))
    if shared:
        inp[B // 2:] = inp[:B // 2]
    mf = torch.randn(B, 128, H, W, generator=g)
    ws = [torch.randn(128, 384, 5, 1, generator=g) * 0.03 for _ in range(3)]
    bs = [torch.randn(128, generator=g) * 0.1 for _ in range(3)]
    hx = torch.cat([h, inp, mf], 1)
    z = torch.sigmoid(F.conv2d(hx, ws[0], bs[0], padding=(2, 0)))
    r = torch.sigmoid(F.conv2d(hx, ws[1], bs[1], padding=(2, 0)))
    q = torch.tanh(F.conv2d(torch.cat([r * h, inp, mf], 1), ws[2], bs[2], padding=(2, 0)))
    ref = (1 - z) * h + z * q
    rest = lambda w: dev(torch.cat([w[:, :128], w[:, 256:]], 1))
    only = lambda w: dev(w[:, 128:256])
    zr = PackedConv([rest(ws[0]), rest(ws[1])], [dev(bs[0]), dev(bs[1])], 1, (2, 0))
    qc = PackedConv([rest(ws[2])], [dev(bs[2])], 1, (2, 0))
    zr_i = PackedConv([only(ws[0]), only(ws[1])], [None, None], 1, (2, 0))
    q_i = PackedConv([only(ws[2])], [None], 1, (2, 0))
    nb, mod = (B // 2, B // 2) if shared else (B, 0)
    hv, iv, mv = View(dev(nhwc(h))), View(dev(nhwc(inp[:nb]))), View(dev(nhwc(mf)))
    pre_zr, pre_q = View(torch.empty(nb, H, W, 256, device="cuda")), View(torch.empty(nb, H, W, 128, device="cuda"))
    K.conv(zr_i, [iv], pre_zr, emit_planes=False)
    K.conv(q_i, [iv], pre_q, emit_planes=False)
    zb, rh = View(torch.empty(B, H, W, 128, device="cuda")), View(torch.empty(B, H, W, 128, device="cuda"))
    K.conv(zr, [hv, mv], epilogue=L.EPI_GRU_ZR, h=hv, z=zb, out2=rh, pre_add=pre_zr, pre_mod=mod)
    K.conv(qc, [rh, mv], epilogue=L.EPI_GRU_Q, h=hv, z=zb, pre_add=pre_q, pre_mod=mod)
    torch.cuda.synchronize()
    assert maxdiff(hv.t.permute(0, 3, 1, 2), ref) < tol


def test_fp16x2_operands_saturate_instead_of_nan():
    """An activation beyond the fp16 range (|v| > 65504) must not turn into hi = inf, lo = -inf -> NaN: the operand
    split saturates.  In-range pixels of the same launch stay fp32-class."""
    from accflow_b200 import _lib as L
    from accflow_b200.engine import Kernels, PackedConv, View
    K2 = Kernels(torch.device("cuda:0"), "fp16x2")
    g = torch.Generator().manual_seed(43)
    x = torch.randn(1, 64, 16, 16, generator=g)
    x[0, 3, 2, 2] = 3.0e5
    x[0, 7, 12, 9] = -1.0e6
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    out = View(torch.empty(1, 16, 16, 64, device="cuda"))
    K2.conv(PackedConv([dev(w)], [None], 1, (1, 1)), [View(dev(nhwc(x)))], out)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out.t).all())
    ref = F.conv2d(x.clamp(-65504, 65504), w, padding=1)
    assert maxdiff(out.t.permute(0, 3, 1, 2), ref) < 2e-5 * float(ref.abs().max())


@pytest.mark.parametrize("mod", [0, 2])
def test_pre_add_store_epilogue(KP, mod):
    """pre_add on the plain store epilogue: out = relu(conv(x) + b + pre); ``mod``: sample s reads pre[s % mod]."""
    K, tol = KP
    from accflow_b200 import _lib as L
    from accflow_b200.engine import PackedConv, View
    g = torch.Generator().manual_seed(42)
    nb = 4 if mod else 2
    x = torch.randn(nb, 96, 11, 13, generator=g)
    pre = torch.randn(mod or nb, 64, 11, 13, generator=g)
    w = torch.randn(64, 96, 3, 3, generator=g) * 0.05
    b = torch.randn(64, generator=g)
    ref = torch.relu(F.conv2d(x, w, b, padding=1) + (pre.repeat(nb // mod, 1, 1, 1) if mod else pre))
    out = View(torch.empty(nb, 11, 13, 64, device="cuda"))
    K.conv(PackedConv([dev(w)], [dev(b)], 1, (1, 1)), [View(dev(nhwc(x)))], out, act=L.ACT_RELU,
           pre_add=View(dev(nhwc(pre))), pre_mod=mod)
    torch.cuda.synchronize()
    assert maxdiff(out.t.permute(0, 3, 1, 2), ref) < tol


@pytest.mark.parametrize("cfg", [(3, 2, 64, True, 40, 56), (2, 1, 128, False, 18, 21)])
def test_conv_smallc(K, cfg):
    from accflow_b200 import _lib as L
    from accflow_b200.engine import PackedConv, View
    cin, stride, cout, is_nchw, H, W = cfg
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 7, 7, generator=g) * 0.1
    b = torch.randn(cout, generator=g)
    ref = torch.relu(F.conv2d(x, w, b, stride=stride, padding=3))
    pc = PackedConv([dev(w)], [dev(b)], stride, (3, 3))
    pc.w = dev(w).permute(2, 3, 1, 0).reshape(cin * 49, cout).contiguous()
    xin = dev(x) if is_nchw else dev(nhwc(x))
    out = torch.empty(2, ref.shape[2], ref.shape[3], cout, device="cuda")
    K.conv_smallc(xin.data_ptr(), is_nchw, 2, cin, H, W, pc, L.ACT_RELU, View(out))
    torch.cuda.synchronize()
    assert maxdiff(out.permute(0, 3, 1, 2), ref) < 2e-5


@pytest.mark.parametrize("c,hw", [(64, (40, 52)), (96, (33, 20)), (128, (16, 16))])
def test_instance_norm(K, c, hw):
    from accflow_b200.engine import View
    from oracle import ops
    g = torch.Generator().manual_seed(6)
    x = torch.randn(3, c, *hw, generator=g) * 2 + 3.0           # large mean: cancellation check
    res = torch.randn(3, c, *hw, generator=g)
    ref = torch.relu(res + torch.relu(ops.instance_norm(x)))
    xv = View(dev(nhwc(x)))
    K.instnorm(xv, True, View(dev(nhwc(res))), True, xv)
    torch.cuda.synchronize()
    assert maxdiff(xv.t.permute(0, 3, 1, 2), ref) < 2e-5


# ------------------------------------------------------------------ correlation
def test_corr_pyramid_and_lookup(KP, golden):
    K, tol = KP
    from accflow_b200 import _lib as L
    from accflow_b200.engine import FlowEstimatorEngine, View
    g, _ = golden
    f1, f2, coords = cases.corr_case()
    eng = FlowEstimatorEngine.__new__(FlowEstimatorEngine)
    eng.k = K
    lv = eng.corr_pyramid(View(dev(nhwc(f1))), View(dev(nhwc(f2))), "t")
    torch.cuda.synchronize()
    B, _, h, w = f1.shape
    for i, t in enumerate(lv):
        ref = g[f"corr.pyr{i}"]
        assert maxdiff(t.reshape(ref.shape), ref) < tol, i
    out = torch.empty(B, h, w, 324, device="cuda")
    flow = torch.empty(B, h * w, 2, device="cuda")
    c = dev(coords.permute(0, 2, 3, 1).reshape(B, h * w, 2).contiguous())
    L.call("accflow_corr_lookup_f32", lv[0].data_ptr(), lv[1].data_ptr(), lv[2].data_ptr(), lv[3].data_ptr(), B, h, w, 4,
           c.data_ptr(), out.data_ptr(), 324, flow.data_ptr(), None, 0, None, 0, 0, None, 0, 0, 1, None)
    torch.cuda.synchronize()
    assert maxdiff(out.permute(0, 3, 1, 2), g["corr.lookup"]) < tol
    from oracle import ops
    assert maxdiff(flow.view(B, h, w, 2).permute(0, 3, 1, 2), coords - ops.coords_grid(B, h, w)) < 1e-6


@pytest.mark.parametrize("hw", [(16, 32), (64, 64), (24, 96)])
def test_corr_lookup_bench_shaped_maps(hw):
    """CorrBlock.__call__ (raft/corr.py:24-45) on maps whose widths are multiples of 32 (the 512x512 / 1024x1024
    configurations) vs the oracle, coordinates far outside the map and exactly integer coordinates included."""
    from accflow_b200 import _lib as L
    from oracle import ops
    h, w = hw
    B = 2
    g = torch.Generator().manual_seed(61)
    f1 = torch.randn(B, 32, h, w, generator=g)
    f2 = torch.randn(B, 32, h, w, generator=g)
    pyr = ops.corr_pyramid(f1, f2)
    coords = ops.coords_grid(B, h, w) + torch.randn(B, 2, h, w, generator=g) * 12.0
    coords[0, :, 0, 0] = torch.tensor([-7.3, 2.2])               # far outside: all taps of level 0 are zero padding
    coords[1, :, h - 1, w - 1] = torch.tensor([w + 30.5, h + 9.25])
    coords[0, :, 1, 1] = torch.tensor([3.0, 4.0])                # integer coordinates (weights exactly 1 / 0)
    ref = ops.corr_lookup(pyr, coords)
    lv = [t.reshape(B * h * w, -1).cuda().contiguous() for t in pyr]
    out = torch.empty(B, h, w, 324, device="cuda")
    flow = torch.empty(B, h * w, 2, device="cuda")
    c = dev(coords.permute(0, 2, 3, 1).reshape(B, h * w, 2).contiguous())
    L.call("accflow_corr_lookup_f32", lv[0].data_ptr(), lv[1].data_ptr(), lv[2].data_ptr(), lv[3].data_ptr(), B, h, w, 4,
           c.data_ptr(), out.data_ptr(), 324, flow.data_ptr(), None, 0, None, 0, 0, None, 0, 0, 1, None)
    torch.cuda.synchronize()
    assert maxdiff(out.permute(0, 3, 1, 2), ref) < 2e-5


@pytest.mark.parametrize("fmt", [2, 4, 1, 3])
def test_corr_lookup_operand_planes(fmt):
    """The lookup's operand planes (what convc1 reads, raft/update.py:89-90) are the split of the very fp32 values the
    same launch writes: fp16x2 hi = fp16(v), lo = fp16((v - hi) * 2^11); fp16; bf16; bf16x3.  Also the flow tail of the
    motion features (cat([out, flow]), raft/update.py:96-97) in both forms."""
    from accflow_b200 import _lib as L
    h, w, B = 32, 64, 2
    P = h * w
    g = torch.Generator().manual_seed(67)
    lv = [torch.randn(B * P, (h >> l) * (w >> l), generator=g).cuda() for l in range(4)]
    grid = torch.stack(torch.meshgrid(torch.arange(w), torch.arange(h), indexing="xy"), -1).reshape(1, P, 2).float()
    c = (grid + torch.randn(B, P, 2, generator=g) * 6.0).cuda().contiguous()
    npl = 1 if fmt == 4 else fmt
    out = torch.empty(B, P, 324, device="cuda")
    pitch = 328
    pl = torch.zeros(npl, B, P, pitch, device="cuda", dtype=torch.bfloat16)
    mf = torch.zeros(B, P, 128, device="cuda")
    tpl = torch.zeros(npl, B, P, 128, device="cuda", dtype=torch.bfloat16)
    flow = torch.empty(B, P, 2, device="cuda")
    L.call("accflow_corr_lookup_f32", lv[0].data_ptr(), lv[1].data_ptr(), lv[2].data_ptr(), lv[3].data_ptr(), B, h, w, 4,
           c.data_ptr(), out.data_ptr(), 324, flow.data_ptr(), mf.data_ptr() + 4 * 126, 128, pl.data_ptr(), pitch,
           B * P * pitch, tpl.data_ptr() + 2 * 126, 128, B * P * 128, fmt, None)
    torch.cuda.synchronize()

    def split(v):
        if fmt in (2, 4):
            v = v.clamp(-65504.0, 65504.0)
            hi = v.half()
            return [hi] if fmt == 4 else [hi, ((v - hi.float()) * 2048.0).half()]
        p0 = v.bfloat16()
        if fmt == 1:
            return [p0]
        r1 = v - p0.float()
        p1 = r1.bfloat16()
        return [p0, p1, (r1 - p1.float()).bfloat16()]

    view = (lambda t: t.view(torch.float16)) if fmt in (2, 4) else (lambda t: t)
    for i, ref in enumerate(split(out)):
        assert torch.equal(view(pl[i])[..., :324], ref), (fmt, i)
    assert torch.equal(mf[..., 126:], flow)
    for i, ref in enumerate(split(flow)):
        assert torch.equal(view(tpl[i])[..., 126:], ref), (fmt, i)


@pytest.mark.parametrize("hw", [(32, 32), (16, 64), (64, 64)])
def test_corr_volume_with_fused_first_level(hw):
    """CorrBlock.__init__ (raft/corr.py:8-22, 47-55) through the engine path whose GEMM epilogue emits level 1:
    volume = f1 . f2^T / sqrt(D); levels 1..3 = successive 2x2 means.  Checked against torch in every
    tensor-core mode at map sizes that take the fused path (w % 32 == 0)."""
    from accflow_b200.engine import FlowEstimatorEngine, Kernels, View
    h, w = hw
    B, D = 2, 256
    g = torch.Generator().manual_seed(31)
    f1 = torch.randn(B, h, w, D, generator=g)
    f2 = torch.randn(B, h, w, D, generator=g)
    vol = torch.einsum("bpc,bqc->bpq", f1.reshape(B, h * w, D), f2.reshape(B, h * w, D)) / math.sqrt(D)
    ref = [vol.reshape(B * h * w, 1, h, w)]
    for _ in range(3):
        ref.append(F.avg_pool2d(ref[-1], 2, stride=2))
    for prec, tol in (("fp16x2", 3e-5), ("bf16", 0.1)):
        eng = FlowEstimatorEngine.__new__(FlowEstimatorEngine)
        eng.k = Kernels(torch.device("cuda:0"), prec)
        lv = eng.corr_pyramid(View(f1.cuda()), View(f2.cuda()), "t.corr")
        torch.cuda.synchronize()
        for l in range(4):
            assert maxdiff(lv[l].reshape(ref[l].shape), ref[l]) < tol * 4, (prec, l)


@pytest.mark.parametrize("hw", [(16, 16), (17, 16), (24, 40)])
def test_gma_attention_and_aggregate(KP, hw):
    """Attention.forward + Aggregate.forward (gma/modules.py:54-76, 102-115) at kernel level, every arithmetic mode:
    tensor-core modes take the fused path (row statistics pass, softmax written once as operand planes, transposed
    v planes, aggregation GEMM with the gamma-residual epilogue); ragged P (17x16 = 272 columns) covers the masked
    tail of the last N tile."""
    K, tol = KP
    from accflow_b200.engine import FlowEstimatorEngine, PackedConv, View
    from oracle import flow_oracle as fo
    h, w = hw
    B = 2
    g = torch.Generator().manual_seed(51)
    inp = torch.relu(torch.randn(B, 128, h, w, generator=g))
    mf = torch.randn(B, 128, h, w, generator=g)
    sd = {"att.to_qk.weight": torch.randn(256, 128, 1, 1, generator=g) * 0.35,
          "agg.to_v.weight": torch.randn(128, 128, 1, 1, generator=g) * 0.1, "agg.gamma": torch.tensor([0.5])}
    attn_ref = fo.gma_attention(sd, "att.", inp)
    ref = fo.gma_aggregate(sd, "agg.", attn_ref, mf)
    assert float(attn_ref.max()) > 0.05                      # a peaked softmax, not a uniform one
    eng = FlowEstimatorEngine.__new__(FlowEstimatorEngine)
    eng.k, eng.gma = K, True
    eng.to_qk = PackedConv([dev(sd["att.to_qk.weight"])], [None], 1, (0, 0))
    eng.to_v = PackedConv([dev(sd["agg.to_v.weight"])], [None], 1, (0, 0))
    eng.gamma, eng.qk_scale = 0.5, 128 ** -0.5
    inpv, mfv = View(dev(nhwc(inp))), View(dev(nhwc(mf)))
    out = View(torch.zeros(B, h, w, 128, device="cuda"))
    if K.tc:
        K.planes_ptr(out, create=True)                       # as in the engine: the GRU convs read mf_global's planes only
    for _ in range(2):
        attn = eng.attention(inpv, f"tatt{h}x{w}")
        eng.aggregate(attn, mfv, out, f"tatt{h}x{w}")
    torch.cuda.synchronize()
    P = h * w
    if attn[0] == "planes":
        _, ptr, pitch, pstride = attn
        pl = K._ws[(f"tatt{h}x{w}.attn_pl", "bf16", K.nplanes, B, P, (P + 7) // 8 * 8)]
        def unsplit(t):
            if K.precision == "fp16x2":
                return t.view(torch.float16)[0].float() + t.view(torch.float16)[1].float() / 2048.0
            if K.precision == "fp16":
                return t.view(torch.float16)[0].float()
            return t.float().sum(0)
        got = unsplit(pl)
        assert maxdiff(got[:, :, :P], attn_ref) < tol
        rec = unsplit(K._planes[out.t.data_ptr()])
        assert maxdiff(rec[..., :128].permute(0, 3, 1, 2), ref) < tol * 4
    else:
        assert maxdiff(attn[1], attn_ref) < tol
        assert maxdiff(out.t.permute(0, 3, 1, 2), ref) < tol * 4


def test_convex_upsample(golden):
    from accflow_b200 import ops as P
    g, _ = golden
    flow, mask = cases.upsample_case()
    assert maxdiff(P.convex_upsample(dev(flow), dev(mask)), g["upsample.out"]) < 1e-5


def test_upflow8_vs_torch():
    """networks/utils.py:91-93: 8 * F.interpolate(flow, 8x, 'bilinear', align_corners=True)."""
    from accflow_b200.networks.utils import upflow8
    g = torch.Generator().manual_seed(5)
    for shape in ((2, 2, 16, 24), (1, 2, 64, 64), (1, 3, 1, 5)):
        fl = torch.randn(*shape, generator=g) * 3
        ref = 8 * F.interpolate(fl, size=(8 * shape[2], 8 * shape[3]), mode="bilinear", align_corners=True)
        assert maxdiff(upflow8(dev(fl)), ref) < 2e-5, shape


def test_backwarp_downflow_occ(golden):
    from accflow_b200 import ops as P
    g, _ = golden
    img, flow = cases.warp_case()
    assert maxdiff(P.backwarp(dev(img), dev(flow)), g["warp.out"]) < 2e-6
    (fl,) = cases.downflow_case()
    assert maxdiff(P.downflow8(dev(fl)), g["downflow.out"]) < 2e-6
    oflow, c1, c2 = cases.occ_case()
    assert maxdiff(P.get_occ(dev(oflow), dev(c1), dev(c2)), g["occ.binary"]) == 0
    assert maxdiff(P.get_occ(dev(oflow), dev(c1), dev(c2), binary=False), g["occ.emap"]) < 2e-6


def test_deform_conv(KP, golden):
    """torchvision deform_conv2d (AccFlow_.py:83,104) = modulated gather + GEMM, in every arithmetic mode."""
    K, tol = KP
    from accflow_b200 import _lib as L
    from accflow_b200.engine import PackedConv, View
    g, _ = golden
    x, off, mask, wgt, bias = cases.dcn_case()
    n, c, h, w = x.shape
    # kernel takes mask *logits* (sigmoid fused): invert the fixture's probabilities
    logits = torch.log(mask / (1 - mask))
    om = torch.zeros(n, h, w, 28)
    om[..., :18] = nhwc(off)
    om[..., 18:27] = nhwc(logits)
    col = torch.empty(n, h * w, 9 * c, device="cuda")
    xv = dev(nhwc(x))
    omv = dev(om)
    L.call("accflow_deform_gather_f32", xv.data_ptr(), c, omv.data_ptr(), 28, n, h, w, c, col.data_ptr(), None)
    pc = PackedConv([dev(wgt).permute(0, 2, 3, 1).reshape(wgt.shape[0], -1, 1, 1)], [dev(bias)], 1, (0, 0))
    out = torch.empty(n, h, w, wgt.shape[0], device="cuda")
    K.conv(pc, [View(col.view(n, h, w, 9 * c))], View(out))
    torch.cuda.synchronize()
    scale = float(np.abs(g["dcn.out"]).max())
    assert maxdiff(out.permute(0, 3, 1, 2), g["dcn.out"]) < tol * max(1.0, scale)


def test_conv3x3_smallcout(K):
    """Flow-head style 3x3 conv with 2 / 1 output channels on the bandwidth kernel."""
    from accflow_b200 import _lib as L
    from accflow_b200.engine import PackedConv, View
    g = torch.Generator().manual_seed(9)
    for cout, act, tact in ((2, L.ACT_NONE, lambda t: t), (1, L.ACT_SIGMOID, torch.sigmoid)):
        x = torch.randn(2, 256, 13, 11, generator=g)
        w = torch.randn(cout, 256, 3, 3, generator=g) * 0.03
        b = torch.randn(cout, generator=g)
        ref = tact(F.conv2d(x, w, b, padding=1))
        out = torch.empty(2, 13, 11, cout, device="cuda")
        K.conv_smallcout(PackedConv([dev(w)], [dev(b)], 1, (1, 1)), View(dev(nhwc(x))), View(out), act=act)
        torch.cuda.synchronize()
        assert maxdiff(out.permute(0, 3, 1, 2), ref) < 1e-5


def test_conv3x3_smallcout_all_modes(KP):
    """The same op in every arithmetic mode (tensor-core modes: 1x1 conv with 9*cout outputs + tap sum),
    including the fused accumulate (coords1 += delta_flow, raft/raft.py:136) and ragged map sizes."""
    K, tol = KP
    from accflow_b200 import _lib as L
    from accflow_b200.engine import PackedConv, View
    g = torch.Generator().manual_seed(19)
    for cout, act, tact, hw in ((2, L.ACT_NONE, lambda t: t, (13, 11)), (1, L.ACT_SIGMOID, torch.sigmoid, (9, 20)),
                                (2, L.ACT_NONE, lambda t: t, (16, 16))):
        x = torch.randn(2, 256, *hw, generator=g)
        w = torch.randn(cout, 256, 3, 3, generator=g) * 0.02
        b = torch.randn(cout, generator=g)
        ref = tact(F.conv2d(x, w, b, padding=1))
        out = torch.empty(2, *hw, cout, device="cuda")
        acc0 = torch.randn(2, *hw, cout, generator=g)
        acc = acc0.cuda()
        K.conv_smallcout(PackedConv([dev(w)], [dev(b)], 1, (1, 1)), View(dev(nhwc(x))), View(out), act=act,
                         accum=acc, accum_ld=cout)
        torch.cuda.synchronize()
        assert maxdiff(out.permute(0, 3, 1, 2), ref) < tol
        assert maxdiff(acc.permute(0, 3, 1, 2), acc0.permute(0, 3, 1, 2) + ref) < tol


def test_flow_conv7_all_modes(KP):
    """7x7 2->128 flow conv (raft/update.py:85,92): patch gather (planes only) + K=98 1x1 conv, planes-only output
    consumed by a following tensor-core conv."""
    K, tol = KP
    from accflow_b200 import _lib as L
    from accflow_b200.engine import PackedConv, View
    g = torch.Generator().manual_seed(23)
    B, h, w = 2, 12, 20
    flow = torch.randn(B, 2, h, w, generator=g) * 3
    w7 = torch.randn(128, 2, 7, 7, generator=g) * 0.1
    b7 = torch.randn(128, generator=g) * 0.1
    w3 = torch.randn(64, 128, 3, 3, generator=g) * 0.03
    b3 = torch.randn(64, generator=g) * 0.1
    mid = torch.relu(F.conv2d(flow, w7, b7, padding=3))
    ref = torch.relu(F.conv2d(mid, w3, b3, padding=1))
    pc7, pc3 = PackedConv([dev(w7)], [dev(b7)], 1, (3, 3)), PackedConv([dev(w3)], [dev(b3)], 1, (1, 1))
    fl = dev(flow.permute(0, 2, 3, 1).reshape(B, h * w, 2).contiguous())
    flo1 = View(torch.empty(B, h, w, 128, device="cuda"))
    out = View(torch.empty(B, h, w, 64, device="cuda"))
    for it in range(2):      # second pass: the planes of flo1 exist, so its fp32 copy is skipped
        K.flow_conv7("t7", fl, B, h, w, pc7, flo1, planes_only=True)
        K.conv(pc3, [flo1], out, act=L.ACT_RELU)
    torch.cuda.synchronize()
    assert maxdiff(out.t.permute(0, 3, 1, 2), ref) < tol * 3


def test_softmax_and_transpose(K):
    from accflow_b200.engine import View
    g = torch.Generator().manual_seed(8)
    x = torch.randn(37, 1000, generator=g) * 4
    xd = dev(x).contiguous()
    K.softmax_rows(xd, 37, 1000)
    assert maxdiff(xd, torch.softmax(x, -1)) < 1e-6
    t = torch.randn(2, 5, 7, 44, generator=g)
    out = torch.zeros(2, 44, 36, device="cuda")
    K.transpose(View(dev(t)), out, 36)
    torch.cuda.synchronize()
    assert maxdiff(out[:, :, :35], t.reshape(2, 35, 44).transpose(1, 2)) == 0


def test_error_reporting():
    from accflow_b200 import _lib as L
    with pytest.raises(L.AccflowError, match="multiples of 8"):
        L.call("accflow_downflow8_f32", 1, 1, 100, 64, 1, None)
