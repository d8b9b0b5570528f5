"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/ops.py header; parity PINNED by
tests/golden/).

Functional CPU fp32 restatement of the reference networks, driven by a plain state_dict
(no nn.Module).  Each function cites the reference code it follows.  On CPU the reference's
autocast regions are disabled (torch.cuda.amp.autocast self-disables without CUDA), so the
reference path restated here is pure fp32.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import ops

SD = Dict[str, torch.Tensor]


def _conv(sd: SD, name: str, x, stride=1, pad=0):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=pad)


# --------------------------------------------------------------------------- encoders
def basic_encoder(sd: SD, pfx: str, x: torch.Tensor, norm: str) -> torch.Tensor:
    """BasicEncoder.forward, raft/extractor.py:201-225 (+ ResidualBlock :54-63).
    ``x`` is already the batch-concatenated input (the list handling at :204-207 is a cat)."""
    def nrm(t, name):
        if norm == "instance":
            return ops.instance_norm(t)
        if norm == "batch":
            return ops.batch_norm_eval(t, sd[name + ".weight"], sd[name + ".bias"],
                                       sd[name + ".running_mean"], sd[name + ".running_var"])
        return t

    x = torch.relu(nrm(_conv(sd, pfx + "conv1", x, 2, 3), pfx + "norm1"))
    for stage, stride in ((1, 1), (2, 2), (3, 2)):
        for blk in (0, 1):
            p = f"{pfx}layer{stage}.{blk}."
            s = stride if blk == 0 else 1
            y = torch.relu(nrm(_conv(sd, p + "conv1", x, s, 1), p + "norm1"))
            y = torch.relu(nrm(_conv(sd, p + "conv2", y, 1, 1), p + "norm2"))
            if s != 1:
                x = nrm(_conv(sd, p + "downsample.0", x, s, 0), p + "norm3")
            x = torch.relu(x + y)
    return _conv(sd, pfx + "conv2", x)


# --------------------------------------------------------------------------- update block
def motion_encoder(sd: SD, pfx: str, flow, corr):
    """BasicMotionEncoder.forward, raft/update.py:89-97."""
    cor = torch.relu(_conv(sd, pfx + "convc1", corr))
    cor = torch.relu(_conv(sd, pfx + "convc2", cor, pad=1))
    flo = torch.relu(_conv(sd, pfx + "convf1", flow, pad=3))
    flo = torch.relu(_conv(sd, pfx + "convf2", flo, pad=1))
    out = torch.relu(_conv(sd, pfx + "conv", torch.cat([cor, flo], 1), pad=1))
    return torch.cat([out, flow], 1)


def sep_conv_gru(sd: SD, pfx: str, h, x):
    """SepConvGRU.forward, raft/update.py:45-60."""
    for tag, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([h, x], 1)
        z = torch.sigmoid(_conv(sd, f"{pfx}convz{tag}", hx, pad=pad))
        r = torch.sigmoid(_conv(sd, f"{pfx}convr{tag}", hx, pad=pad))
        q = torch.tanh(_conv(sd, f"{pfx}convq{tag}", torch.cat([r * h, x], 1), pad=pad))
        h = (1 - z) * h + z * q
    return h


def flow_head(sd: SD, pfx: str, net):
    """FlowHead.forward, raft/update.py:13-14."""
    return _conv(sd, pfx + "conv2", torch.relu(_conv(sd, pfx + "conv1", net, pad=1)), pad=1)


def mask_head(sd: SD, pfx: str, net, scale: float):
    """mask Sequential, raft/update.py:122-125,135 (scale .25) / AccFlow_.py:21-25 (scale 1)."""
    return scale * _conv(sd, pfx + "2", torch.relu(_conv(sd, pfx + "0", net, pad=1)))


def gma_attention(sd: SD, pfx: str, inp):
    """Attention.forward, gma/modules.py:54-76 with heads=1 and both position flags False."""
    b, c, h, w = inp.shape
    qk = F.conv2d(inp, sd[pfx + "to_qk.weight"])
    q, k = qk.chunk(2, dim=1)
    q = q.reshape(b, -1, h * w).transpose(1, 2) * (q.shape[1] ** -0.5)
    k = k.reshape(b, -1, h * w)
    return torch.softmax(q @ k, dim=-1)                       # (b, hw, hw)


def gma_aggregate(sd: SD, pfx: str, attn, fmap):
    """Aggregate.forward, gma/modules.py:102-115 (project is None: dim == inner_dim)."""
    b, c, h, w = fmap.shape
    v = F.conv2d(fmap, sd[pfx + "to_v.weight"]).reshape(b, c, h * w).transpose(1, 2)
    out = (attn @ v).transpose(1, 2).reshape(b, c, h, w)
    return fmap + sd[pfx + "gamma"] * out


def flow_estimator(sd: SD, image1, image2, iters: int = 12, flow_init=None, pfx: str = "",
                   gma: Optional[bool] = None, trace: Optional[dict] = None):
    """RAFT.forward (raft/raft.py:94-146) / RAFTGMA.forward (gma/gma.py:70-125).

    Returns flow_up (B,2,H,W).  ``trace`` (optional dict) receives stage-boundary tensors for
    per-kernel parity tests.
    """
    if gma is None:
        gma = (pfx + "att.to_qk.weight") in sd
    b = image1.shape[0]
    fmaps = basic_encoder(sd, pfx + "fnet.", torch.cat([image1, image2], 0), "instance")
    fmap1, fmap2 = fmaps[:b], fmaps[b:]
    pyramid = ops.corr_pyramid(fmap1, fmap2)
    cnet = basic_encoder(sd, pfx + "cnet.", image1, "batch")
    net = torch.tanh(cnet[:, :128])
    inp = torch.relu(cnet[:, 128:])
    attn = gma_attention(sd, pfx + "att.", inp) if gma else None
    h, w = image1.shape[-2] // 8, image1.shape[-1] // 8
    coords0 = ops.coords_grid(b, h, w, image1.device)
    coords1 = coords0.clone()
    if flow_init is not None:
        coords1 = coords1 + flow_init
    if trace is not None:
        trace.update(fmap1=fmap1, fmap2=fmap2, pyramid=pyramid, net0=net, inp=inp, attn=attn,
                     corr=[], net=[], delta=[], mf=[])
    ub = pfx + "update_block."
    for _ in range(iters):
        corr = ops.corr_lookup(pyramid, coords1)
        flow = coords1 - coords0
        mf = motion_encoder(sd, ub + "encoder.", flow, corr)
        if gma:
            mfg = gma_aggregate(sd, ub + "aggregator.", attn, mf)
            x = torch.cat([inp, mf, mfg], 1)
        else:
            x = torch.cat([inp, mf], 1)
        net = sep_conv_gru(sd, ub + "gru.", net, x)
        delta = flow_head(sd, ub + "flow_head.", net)
        coords1 = coords1 + delta
        if trace is not None:
            trace["corr"].append(corr); trace["net"].append(net)
            trace["delta"].append(delta); trace["mf"].append(mf)
    # the reference recomputes mask + upsample every iteration and keeps the last
    # (raft.py:139-146); only the final one is observable.
    up_mask = mask_head(sd, ub + "mask.", net, 0.25)
    flow_lr = coords1 - coords0
    if trace is not None:
        trace.update(up_mask=up_mask, flow_lr=flow_lr)
    return ops.convex_upsample(flow_lr, up_mask)


# --------------------------------------------------------------------------- accumulation
def flow_encoder(sd: SD, x):
    """FlowEncoder.forward, AccFlow_.py:56-65 (batch-concatenated input)."""
    x = torch.relu(_conv(sd, "flow_encoder.conv1", x, pad=3))
    x = torch.relu(_conv(sd, "flow_encoder.conv2", x, pad=1))
    return _conv(sd, "flow_encoder.conv3", x)


def acc_plus(sd: SD, df, f, o, c, trace: Optional[dict] = None):
    """AccPlus.forward, AccFlow_.py:97-109; ZeroConv2d networks/modules.py:94-97."""
    p = "accplus."
    x = _conv(sd, p + "conv1.2", torch.relu(_conv(sd, p + "conv1.0", torch.cat([df, f, o], 1), pad=1)), pad=1)
    x = torch.relu(_conv(sd, p + "conv2.0", torch.cat([x, c], 1), pad=1))
    x = torch.relu(_conv(sd, p + "conv2.2", x, pad=1))
    x = _conv(sd, p + "conv2.4.conv", x, pad=1) * torch.exp(sd[p + "conv2.4.scale"] * 3)
    off, m = x[:, :18], torch.sigmoid(x[:, 18:])
    f_ = ops.deform_conv2d(f, off, m, sd[p + "dconv.weight"], sd[p + "dconv.bias"])
    if trace is not None:
        trace.update(dcn_off=off, dcn_mask=m, dcn_out=f_)
    x = _conv(sd, p + "conv3.2", torch.relu(_conv(sd, p + "conv3.0", torch.cat([f_, df, o], 1), pad=1)), pad=1)
    x = torch.relu(_conv(sd, p + "conv4.0", torch.cat([x, c, f_, df], 1), pad=1))
    x = torch.relu(_conv(sd, p + "conv4.2", x, pad=1))
    return _conv(sd, p + "conv4.4", x)


def blending(sd: SD, f1, f2, emap):
    """Blending.forward, AccFlow_.py:122-124."""
    m = torch.relu(_conv(sd, "blending.mask.0", emap))
    m = torch.sigmoid(_conv(sd, "blending.mask.2", m, pad=1))
    return f1 * m + (1 - m) * f2


def flow_decoder(sd: SD, x):
    """FlowDecoder.forward, AccFlow_.py:40-45 (no 0.25 on the mask)."""
    small = flow_head_generic(sd, "flow_decoder.flow.", x)
    mask = mask_head(sd, "flow_decoder.mask.", x, 1.0)
    return small, ops.convex_upsample(small, mask)


def flow_head_generic(sd: SD, pfx: str, x):
    return _conv(sd, pfx + "2", torch.relu(_conv(sd, pfx + "0", x, pad=1)), pad=1)


def acc_iter(sd: SD, i1, i2, i_n, f2n, iters: int = 12, trace: Optional[dict] = None, flow_init=None):
    """AccFlow.iter, AccFlow_.py:177-201.  ``flow_init`` (not in the reference's iter; the estimators' own argument,
    raft/raft.py:123-124) feeds the warm-start mode of accflow_forward."""
    if f2n is None:
        flows = flow_estimator(sd, torch.cat([i1, i1, i2]), torch.cat([i2, i_n, i_n]), iters, pfx="ofe.")
        dflow, flow_ini, f2n = ops.downflow8(flows).chunk(3)
    else:
        flows = flow_estimator(sd, torch.cat([i1, i1]), torch.cat([i2, i_n]), iters, flow_init, pfx="ofe.")
        dflow, flow_ini = ops.downflow8(flows).chunk(2)
    if trace is not None:
        trace["dflow_small"] = dflow
    b = i1.shape[0]
    enc = flow_encoder(sd, torch.cat([flow_ini, dflow, f2n], 0))
    f_ini, df, f = enc[:b], enc[b:2 * b], enc[2 * b:]
    ctx = basic_encoder(sd, "context.", torch.cat([i1, i2, i_n], 0), "none")
    c1, c2, cn = ctx[:b], ctx[b:2 * b], ctx[2 * b:]
    o = ops.get_occ(dflow, c1, c2)
    f_acc = acc_plus(sd, df, f, o, c1, trace)
    emap = ops.get_occ(flow_ini, c1, cn, binary=False)
    f_fuse = blending(sd, f_ini, f_acc, emap)
    if trace is not None:
        trace.update(dflow=dflow, flow_ini=flow_ini, f2n=f2n, f_ini=f_ini, df=df, f=f, c1=c1, c2=c2,
                     cn=cn, o=o, f_acc=f_acc, emap=emap, f_fuse=f_fuse)
    return flow_decoder(sd, f_fuse)


def accflow_forward(sd: SD, images: List[torch.Tensor], iters: int = 12, warm_start: bool = False,
                    warm_iters: Optional[int] = None) -> List[torch.Tensor]:
    """AccFlow.forward, AccFlow_.py:157-175: [F(2->0), F(3->0), ..., F(n-1->0)].

    ``warm_start`` is the mode the reference lists as a TODO (README.md:11) and this repo defines:
    from the second step on the estimator starts pair (i -> i-1) from the previous step's F(i-1 -> i-2) and pair
    (i -> 0) from the previous accumulated F(i-1 -> 0), both at 1/8 resolution, for ``warm_iters`` iterations."""
    flow = None
    outs = []
    prev_dflow = None
    for i in range(2, len(images)):
        tr = {}
        if warm_start and flow is not None:
            flow, flow_up = acc_iter(sd, images[i], images[i - 1], images[0], flow, warm_iters or iters, tr,
                                     flow_init=torch.cat([prev_dflow, flow]))
        else:
            flow, flow_up = acc_iter(sd, images[i], images[i - 1], images[0], flow, iters, tr)
        prev_dflow = tr["dflow_small"]
        outs.append(flow_up)
    return outs
