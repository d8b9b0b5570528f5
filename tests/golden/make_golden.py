"""Generate tests/golden/golden_v1.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

The reference is imported from where it lies; nothing is copied.  ``calc_occ_mask`` and
``cal_epe`` live in test_cvo.py whose body runs at import, so their FunctionDef nodes are
compiled out of the parsed source instead of importing the script.
"""
from __future__ import annotations

import ast
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("ACCFLOW_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from tests.golden import cases  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def main():
    import torchvision
    from networks import build_flow_estimator
    from networks.AccFlow_ import AccFlow, AccPlus, Blending, FlowDecoder, FlowEncoder, downflow8, getOcc
    from networks.gma.modules import Aggregate, Attention
    from networks.raft.corr import CorrBlock
    from networks.utils import backwarp

    out = {}
    meta = {"torch": torch.__version__, "torchvision": torchvision.__version__, "checksums": {}, "keys": {}}
    torch.set_grad_enabled(False)

    # ---- op-level known-answer vectors -------------------------------------------------
    f1, f2, coords = cases.corr_case()
    meta["checksums"]["corr"] = cases.checksum(f1, f2, coords)
    cb = CorrBlock(f1, f2, radius=4)
    for i, lvl in enumerate(cb.corr_pyramid):
        out[f"corr.pyr{i}"] = _np(lvl)
    out["corr.lookup"] = _np(cb(coords))

    flow, mask = cases.upsample_case()
    meta["checksums"]["upsample"] = cases.checksum(flow, mask)
    raft = build_flow_estimator("raft").eval()
    out["upsample.out"] = _np(raft.upsample_flow(flow, mask))

    img, wflow = cases.warp_case()
    meta["checksums"]["warp"] = cases.checksum(img, wflow)
    out["warp.out"] = _np(backwarp(img, wflow))

    (dflow,) = cases.downflow_case()
    meta["checksums"]["downflow"] = cases.checksum(dflow)
    out["downflow.out"] = _np(downflow8(dflow))

    oflow, c1, c2 = cases.occ_case()
    meta["checksums"]["occ"] = cases.checksum(oflow, c1, c2)
    out["occ.binary"] = _np(getOcc(oflow, c1, c2))
    out["occ.emap"] = _np(getOcc(oflow, c1, c2, binary=False))

    x, off, dmask, wgt, bias = cases.dcn_case()
    meta["checksums"]["dcn"] = cases.checksum(x, off, dmask, wgt, bias)
    out["dcn.out"] = _np(torchvision.ops.deform_conv2d(x, off, wgt, bias, padding=1, mask=dmask))

    src = open(os.path.join(REF, "test_cvo.py")).read()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("calc_occ_mask", "cal_epe")]
    ns = {"torch": torch, "backwarp": backwarp}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "test_cvo.py", "exec"), ns)
    bflow, fflow, pred = cases.metric_case()
    meta["checksums"]["metric"] = cases.checksum(bflow, fflow, pred)
    occ_bw, occ_fw = ns["calc_occ_mask"](bflow, fflow)
    out["metric.occ_bw"], out["metric.occ_fw"] = _np(occ_bw), _np(occ_fw)
    for name, v in zip(("all", "occ", "vis"), ns["cal_epe"](pred, bflow, occ_bw)):
        out[f"metric.epe_{name}"] = _np(v)

    # ---- module-level vectors with the seeded state dicts -----------------------------
    for kind in ("raft", "gma", "acc+raft", "acc+gma"):
        sd = cases.weights(kind)
        ofe = build_flow_estimator(kind).eval()
        model = AccFlow(ofe).eval() if kind.startswith("acc") else ofe
        ref_sd = model.state_dict()
        assert list(ref_sd.keys()) == list(sd.keys()), kind
        meta["keys"][kind] = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in ref_sd.items()]
        model.load_state_dict(sd)

        if not kind.startswith("acc"):
            i1, i2, finit = cases.pair_case()
            meta["checksums"][f"{kind}.pair"] = cases.checksum(i1, i2, finit)
            cap = {"corr": [], "net": [], "delta": [], "mask": []}
            def on_fnet(m, a, o):
                cap["fmaps"] = o

            def on_update(m, a, o):
                cap["corr"].append(a[2]); cap["net"].append(o[0])
                cap["mask"].append(o[1]); cap["delta"].append(o[2])

            def on_att(m, a, o):
                cap["attn"] = o

            hooks = [model.fnet.register_forward_hook(on_fnet),
                     model.update_block.register_forward_hook(on_update)]
            if kind == "gma":
                hooks.append(model.att.register_forward_hook(on_att))
            flow_up = model(i1, i2, iters=12, flow_init=finit)
            for h in hooks:
                h.remove()
            out[f"{kind}.flow_up"] = _np(flow_up)
            out[f"{kind}.fmap1"] = _np(cap["fmaps"][0])
            out[f"{kind}.corr0"] = _np(cap["corr"][0])
            out[f"{kind}.net0"] = _np(cap["net"][0])
            out[f"{kind}.delta0"] = _np(cap["delta"][0])
            out[f"{kind}.delta11"] = _np(cap["delta"][11])
            out[f"{kind}.mask11"] = _np(cap["mask"][11])
            if kind == "gma":
                out["gma.attn"] = _np(cap["attn"][0, 0, :, ::8])       # every 8th column
            out[f"{kind}.flow_up_noinit_it3"] = _np(model(i1, i2, iters=3))
        else:
            imgs = cases.clip_case()
            meta["checksums"][f"{kind}.clip"] = cases.checksum(*imgs)
            flows = model(images=imgs, test_mode=False)
            for i, f in enumerate(flows):
                out[f"{kind}.flow{i}"] = _np(f)
            if kind == "acc+raft":
                d = cases.acc_modules_case()
                meta["checksums"]["acc_modules"] = cases.checksum(*d.values())
                out["acc.accplus"] = _np(model.accplus(d["df"], d["f"], d["o"], d["c"]))
                out["acc.blending"] = _np(model.blending(d["f1"], d["f2"], d["emap"]))
                out["acc.flow_encoder"] = _np(model.flow_encoder(d["flows"]))
                small, full = model.flow_decoder(d["f1"])
                out["acc.dec_small"], out["acc.dec_full"] = _np(small), _np(full)

    # GMA attention / aggregate on a small map
    sd = cases.weights("gma")
    gma = build_flow_estimator("gma").eval()
    gma.load_state_dict(sd)
    inp, mf = cases.gma_case()
    meta["checksums"]["gma_small"] = cases.checksum(inp, mf)
    attn = gma.att(inp)
    out["gma_small.attn"] = _np(attn)
    out["gma_small.agg"] = _np(gma.update_block.aggregator(attn, mf))

    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    with open(os.path.join(HERE, "golden_v1.json"), "w") as fh:
        json.dump(meta, fh, indent=0)
    total = sum(v.nbytes for v in out.values())
    print(f"wrote {len(out)} arrays, {total / 1e6:.2f} MB raw")


if __name__ == "__main__":
    main()
