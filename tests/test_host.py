"""CPU-only checks: C-ABI exports, module surface, data generator, no-fallback behaviour."""
import os
import re

import pytest
import torch

from tests.golden import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports every prototype in include/*.h."""
    from accflow_b200 import _lib as L
    lib = L.load()
    hdr = open(os.path.join(ROOT, "include", "accflow_b200.h")).read()
    declared = set(re.findall(r"ACCFLOW_API\s+(?:int|long long)\s+(accflow_\w+)\s*\(", hdr))
    assert len(declared) >= 18
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.accflow_abi_version() == L.ABI_VERSION == 5
    assert lib.accflow_instnorm_chunks(4096) == 4


def test_conv_desc_layout_matches_header():
    """ctypes mirror of accflow_conv_desc has the C struct's size (LP64)."""
    import ctypes
    from accflow_b200 import _lib as L
    # 4 ptr + 4 int + 4 int + 1 int + 3 int | ptr + i64 | 7 int + float (+pad) | 2 ptr | 3 int (+pad) | ptr + 3 int ...
    assert ctypes.sizeof(L.ConvDesc) % 8 == 0
    assert L.ConvDesc.weight.offset % 8 == 0 and L.ConvDesc.out.offset % 8 == 0


def test_argument_validation_without_gpu():
    import ctypes
    from accflow_b200 import _lib as L
    with pytest.raises(L.AccflowError, match="nsrc"):
        L.call("accflow_conv2d_f32", ctypes.byref(L.ConvDesc()), None)
    with pytest.raises(L.AccflowError, match="bad arguments"):
        L.call("accflow_blend_f32", None, None, None, 1, 10, 128, None, None)


@pytest.mark.parametrize("kind", ["raft", "gma", "acc+raft", "acc+gma"])
def test_module_state_dict_contract(golden, kind):
    from accflow_b200.networks import build_flow_estimator
    from accflow_b200.networks.AccFlow_ import AccFlow
    _, meta = golden
    m = build_flow_estimator(kind)
    if kind.startswith("acc"):
        m = AccFlow(m)
    got = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()]
    assert got == meta["keys"][kind]
    m.load_state_dict(cases.weights(kind))                      # strict
    torch.nn.DataParallel(m).load_state_dict({"module." + k: v for k, v in cases.weights(kind).items()})
    assert m.ofe.hidden_dim == 128 if kind.startswith("acc") else m.hidden_dim == 128


def test_reference_like_default_init():
    from accflow_b200.networks import build_flow_estimator
    from accflow_b200.networks.AccFlow_ import AccFlow
    torch.manual_seed(0)
    m = AccFlow(build_flow_estimator("acc|gma"))
    sd = m.state_dict()
    assert float(sd["accplus.conv2.4.conv.weight"].abs().max()) == 0          # ZeroConv2d
    assert float(sd["ofe.update_block.aggregator.gamma"]) == 0
    assert float(sd["ofe.cnet.norm1.running_var"].min()) == 1
    assert m.mixed_precision is True and m.ofe.args.corr_radius == 4


def test_no_cpu_fallback():
    from accflow_b200.networks import build_flow_estimator
    m = build_flow_estimator("raft")
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 128, 128), torch.zeros(1, 3, 128, 128))


def test_synthetic_clip_ground_truth_is_consistent():
    """Frame i equals frame 0 shifted by the stated long-range flow (exact integer shifts)."""
    from accflow_b200.data import make_clip
    clip = make_clip(0, size=128)
    assert len(clip["imgs"]) == 7 and len(clip["bflows"]) == 5
    i0, i6 = clip["u8"][0], clip["u8"][6]
    dx, dy = (int(v) for v in clip["fflows"][-1][0, :, 0, 0])
    ys = slice(max(0, -dy), 128 - max(0, dy))
    xs = slice(max(0, -dx), 128 - max(0, dx))
    yd = slice(max(0, dy), 128 - max(0, -dy))
    xd = slice(max(0, dx), 128 - max(0, -dx))
    assert torch.equal(i0[:, ys, xs], i6[:, yd, xd])
    assert float(clip["imgs"][0].min()) >= -1 and float(clip["imgs"][0].max()) <= 1


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under accflow_b200/ may reference it."""
    pkg = os.path.join(ROOT, "accflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_cvo_record_helpers():
    """uint16 fixed-point flows (data/dataset.py:65-67) and the preprocess split (test_cvo.py:32-50)."""
    from accflow_b200.data import decode_cvo_flow_u16, encode_cvo_flow_u16, make_batch, preprocess
    raw = torch.tensor([0, 32768, 32768 + 128, 65535], dtype=torch.int32)
    assert decode_cvo_flow_u16(raw).tolist() == [-256.0, 0.0, 1.0, (65535 - 32768) / 128.0]
    f = torch.randn(2, 10, 8, 8) * 20
    assert float((decode_cvo_flow_u16(encode_cvo_flow_u16(f)) - f).abs().max()) <= 0.5 / 128 + 1e-4
    b = make_batch([0, 1], size=128)
    rec = {"imgs": torch.cat([(t + 1) * 127.5 for t in b["imgs"]], 1), "bflows": torch.cat(b["bflows"], 1),
           "fflows": torch.cat(b["fflows"], 1)}
    out = preprocess(rec)
    assert len(out["imgs"]) == 7 and len(out["bflows"]) == 5 and out["imgs"][0].shape == (2, 3, 128, 128)
    assert float((out["imgs"][3] - b["imgs"][3]).abs().max()) < 1e-6
    with pytest.raises(ValueError):
        preprocess({"other": torch.zeros(1, 3, 8, 8)})


def test_packed_conv_1x1_rewrites_are_the_same_filter():
    """Host-side weight rewrites used by the tensor-core path (no GPU): a 3x3 conv with few outputs as a 1x1 conv
    with 9*cout outputs + a 9-tap shifted sum (PackedConv.as_taps1x1 + accflow_tapsum3x3_f32), and a KxK conv as a
    1x1 conv over im2col'd input (PackedConv.as_1x1, flow / stem patch kernels)."""
    import torch.nn.functional as F
    from accflow_b200.engine import PackedConv
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 16, 9, 11, generator=g)
    w = torch.randn(2, 16, 3, 3, generator=g)
    b = torch.randn(2, generator=g)
    ref = F.conv2d(x, w, b, padding=1)
    pc = PackedConv([w], [b], 1, (1, 1))
    t = F.conv2d(x, pc.as_taps1x1().w_oihw)                   # (2, 9*cout, H, W), channel = tap*cout + o
    tp = F.pad(t, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for tap in range(9):
        ky, kx = divmod(tap, 3)
        out += tp[:, tap * 2:tap * 2 + 2, ky:ky + 9, kx:kx + 11]
    out += pc.shift.view(1, 2, 1, 1)
    assert float((out - ref).abs().max()) < 5e-5              # fp32 summation order only (values ~ +-20)
    # KxK conv == 1x1 conv over patches ordered (ky, kx, cin)
    w7 = torch.randn(8, 2, 7, 7, generator=g)
    f = torch.randn(1, 2, 10, 12, generator=g)
    ref7 = F.conv2d(f, w7, None, padding=3)
    cols = F.unfold(f, 7, padding=3).view(1, 2, 49, 10 * 12).permute(0, 2, 1, 3).reshape(1, 98, 10, 12)   # (tap, cin) order
    pc7 = PackedConv([w7], [None], 1, (3, 3))
    assert float((F.conv2d(cols, pc7.as_1x1().w_oihw) - ref7).abs().max()) < 5e-5


def test_graft_entry_build_runs():
    """The driver's build check: compiles (or finds up to date) the library, loads it and imports the package."""
    import __graft_entry__ as g
    g.build()
