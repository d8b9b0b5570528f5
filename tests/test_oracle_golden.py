"""Pins the oracle (oracle/) against the reference's own outputs (tests/golden/golden_v1.npz,
made by tests/golden/make_golden.py from the unmodified reference).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import flow_oracle as fo
from oracle import ops
from tests.golden import cases

torch.set_grad_enabled(False)


def close(a, ref, atol, name="", rel=False):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    assert a.shape == ref.shape, (name, a.shape, ref.shape)
    err = float(np.max(np.abs(a - ref))) if a.size else 0.0
    if rel and a.size:          # tolerance relative to the tensor's scale
        atol = atol * max(1.0, float(np.max(np.abs(ref))))
    assert err <= atol, f"{name}: max-abs {err:.3e} > {atol:.1e}"


def test_input_checksums(golden):
    """RNG drift guard: the regenerated inputs are the ones the fixtures were made from."""
    _, meta = golden
    cs = meta["checksums"]
    assert cases.checksum(*cases.corr_case()) == pytest.approx(cs["corr"], rel=1e-12)
    assert cases.checksum(*cases.dcn_case()) == pytest.approx(cs["dcn"], rel=1e-12)
    assert cases.checksum(*cases.pair_case()) == pytest.approx(cs["raft.pair"], rel=1e-12)
    assert cases.checksum(*cases.clip_case()) == pytest.approx(cs["acc+raft.clip"], rel=1e-12)


def test_corr_pyramid_and_lookup(golden):
    g, _ = golden
    f1, f2, coords = cases.corr_case()
    pyr = ops.corr_pyramid(f1, f2)
    assert [tuple(p.shape[-2:]) for p in pyr] == [(20, 18), (10, 9), (5, 4), (2, 2)]
    for i, p in enumerate(pyr):
        close(p, g[f"corr.pyr{i}"], 2e-6, f"pyr{i}")
    close(ops.corr_lookup(pyr, coords), g["corr.lookup"], 5e-6, "lookup")


def test_convex_upsample(golden):
    g, _ = golden
    close(ops.convex_upsample(*cases.upsample_case()), g["upsample.out"], 1e-5, "upsample")


def test_backwarp_downflow_occ(golden):
    g, _ = golden
    close(ops.backwarp(*cases.warp_case()), g["warp.out"], 2e-6, "warp")
    close(ops.downflow8(*cases.downflow_case()), g["downflow.out"], 2e-6, "downflow8")
    flow, c1, c2 = cases.occ_case()
    close(ops.get_occ(flow, c1, c2), g["occ.binary"], 0, "occ")
    assert 0.05 < float(g["occ.binary"].mean()) < 0.95       # both branches exercised
    close(ops.get_occ(flow, c1, c2, binary=False), g["occ.emap"], 2e-6, "emap")


def test_deform_conv(golden):
    g, _ = golden
    close(ops.deform_conv2d(*cases.dcn_case()), g["dcn.out"], 1e-5, "dcn")


def test_metrics(golden):
    g, _ = golden
    bflow, fflow, pred = cases.metric_case()
    occ_bw, occ_fw = ops.calc_occ_mask(bflow, fflow)
    close(occ_bw, g["metric.occ_bw"], 0, "occ_bw")
    close(occ_fw, g["metric.occ_fw"], 0, "occ_fw")
    for name, v in zip(("all", "occ", "vis"), ops.cal_epe(pred, bflow, occ_bw)):
        close(v, g[f"metric.epe_{name}"], 1e-6, name)


def test_gma_small(golden):
    g, _ = golden
    sd = cases.weights("gma")
    inp, mf = cases.gma_case()
    attn = fo.gma_attention(sd, "att.", inp)
    close(attn, g["gma_small.attn"][:, 0], 1e-6, "attn")
    close(fo.gma_aggregate(sd, "update_block.aggregator.", attn, mf), g["gma_small.agg"], 1e-5, "agg")


def test_acc_modules(golden):
    g, _ = golden
    sd = cases.weights("acc+raft")
    d = cases.acc_modules_case()
    close(fo.acc_plus(sd, d["df"], d["f"], d["o"], d["c"]), g["acc.accplus"], 2e-5, "accplus")
    close(fo.blending(sd, d["f1"], d["f2"], d["emap"]), g["acc.blending"], 1e-5, "blending")
    close(fo.flow_encoder(sd, d["flows"]), g["acc.flow_encoder"], 1e-5, "flow_encoder")
    small, full = fo.flow_decoder(sd, d["f1"])
    close(small, g["acc.dec_small"], 1e-5, "dec_small")
    close(full, g["acc.dec_full"], 1e-4, "dec_full")


@pytest.mark.parametrize("kind", ["raft", "gma"])
def test_pair_end_to_end(golden, kind):
    g, _ = golden
    sd = cases.weights(kind)
    i1, i2, finit = cases.pair_case()
    tr = {}
    flow = fo.flow_estimator(sd, i1, i2, 12, finit, trace=tr)
    close(tr["fmap1"], g[f"{kind}.fmap1"], 2e-5, "fmap1")
    close(tr["corr"][0], g[f"{kind}.corr0"], 5e-6, "corr0", rel=True)
    close(tr["net"][0], g[f"{kind}.net0"], 2e-4, "net0")
    close(tr["delta"][0], g[f"{kind}.delta0"], 1e-4, "delta0")
    close(tr["delta"][11], g[f"{kind}.delta11"], 3e-4, "delta11")
    close(tr["up_mask"], g[f"{kind}.mask11"], 1e-3, "mask11")
    if kind == "gma":
        close(tr["attn"][0, :, ::8], g["gma.attn"], 1e-6, "attn")
    # north-star bar: flows within 1e-3 px max-abs in fp32
    close(flow, g[f"{kind}.flow_up"], 1e-3, "flow_up")
    close(fo.flow_estimator(sd, i1, i2, 3), g[f"{kind}.flow_up_noinit_it3"], 1e-3, "it3")


@pytest.mark.parametrize("kind", ["acc+raft", "acc+gma"])
def test_clip_end_to_end(golden, kind):
    g, _ = golden
    sd = cases.weights(kind)
    flows = fo.accflow_forward(sd, cases.clip_case())
    assert len(flows) == 2
    for i, f in enumerate(flows):
        close(f, g[f"{kind}.flow{i}"], 1e-3, f"flow{i}")


def test_state_dict_contract(golden):
    """Key names, order, shapes and dtypes equal the reference's."""
    _, meta = golden
    for kind, ref in meta["keys"].items():
        sd = cases.weights(kind)
        got = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()]
        assert got == ref, kind
