"""GPU: flow parity of every arithmetic mode against the reference-generated golden fixtures
and the CPU oracle.  Prints one JSON line per (case, precision)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from accflow_b200.data import make_batch
from accflow_b200.networks import build_flow_estimator
from accflow_b200.networks.AccFlow_ import AccFlow
from oracle import flow_oracle as fo, ops
from tests.golden import cases

torch.set_grad_enabled(False)
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "golden_v1.npz"))
big = "--big" in sys.argv


def build(kind, precision):
    m = build_flow_estimator(kind)
    if kind.startswith("acc"):
        m = AccFlow(m)
        m.ofe.precision = precision
    else:
        m.precision = precision
    m.load_state_dict(cases.weights(kind))
    return m.cuda().eval()


def md(a, b):
    return float((a.detach().float().cpu() - torch.as_tensor(b).float()).abs().max())


oracle_cache = {}
for precision in ("fp32", "bf16x3", "fp16x2", "bf16"):
    for kind in ("raft", "gma"):
        m = build(kind, precision)
        i1, i2, finit = cases.pair_case()
        out = m(i1.cuda(), i2.cuda(), iters=12, flow_init=finit.cuda())
        print(json.dumps({"case": f"{kind} pair 128x128 vs reference golden", "precision": precision,
                          "max_abs_px": md(out, g[f"{kind}.flow_up"]), "flow_max_px": float(np.abs(g[f'{kind}.flow_up']).max())}))
    for kind in ("acc+raft", "acc+gma"):
        m = build(kind, precision)
        flows = m(images=[t.cuda() for t in cases.clip_case()])
        print(json.dumps({"case": f"{kind} 4-frame clip 128x128 vs reference golden", "precision": precision,
                          "max_abs_px": max(md(f, g[f"{kind}.flow{i}"]) for i, f in enumerate(flows))}))
    if big:
        for kind, size in (("acc+raft", 512), ("acc+gma", 256)):
            batch = make_batch([21], size=size)
            sd = cases.weights(kind)
            key = (kind, size)
            if key not in oracle_cache:
                oracle_cache[key] = fo.accflow_forward(sd, batch["imgs"])
            ref = oracle_cache[key]
            m = build(kind, precision)
            out = m(images=[t.cuda() for t in batch["imgs"]])
            bflow, fflow = batch["bflows"][-1], batch["fflows"][-1]
            occ, _ = ops.calc_occ_mask(bflow, fflow)
            e_ref = torch.stack(ops.cal_epe(ref[-1], bflow, occ))
            e_out = torch.stack(ops.cal_epe(out[-1].cpu(), bflow, occ))
            print(json.dumps({"case": f"{kind} 7-frame clip {size}x{size} vs oracle", "precision": precision,
                              "max_abs_px": max(md(a, b) for a, b in zip(out, ref)),
                              "epe_delta_px": float((e_ref - e_out).abs().max()),
                              "flow_max_px": float(max(r.abs().max() for r in ref))}))
