"""ORACLE — TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product package).

"As executed" restatement of the reference's GPU path, for timing the north-star denominator
("the reference's own PyTorch CUDA path") on the GPU box.  The reference is Python and cannot
travel to the GPU box (tier rule: nothing may read /root/reference at run time there, and its
sources may not be copied into this repo), so this module restates its *execution schedule*
with the very same library calls the reference makes — unlike ``oracle/ops.py``, which spells
every sampling op out as index arithmetic for readability:

* ``F.grid_sample`` / ``F.avg_pool2d`` / ``F.unfold`` / ``F.interpolate`` / ``torch.matmul`` /
  ``F.instance_norm`` / ``F.batch_norm`` / ``torchvision.ops.deform_conv2d`` / cuDNN ``conv2d``;
* fp16 autocast in the regions the reference opens (raft/raft.py:107,115,132; gma/gma.py:83,91,
  110; AccFlow_.py:191), fp32 correlation volume (raft.py:110-111) and lookup;
* the dead work the reference really performs: mask head + convex upsample on *every* GRU
  iteration (raft.py:139-146), all three encoders re-run on every accumulation step
  (AccFlow_.py:184,188,193), the sampling grids rebuilt per level and per iteration
  (raft/corr.py:32-38), ``backwarp`` building its mesh on the CPU and uploading it
  (networks/utils.py:106-113), GMA's CPU-built ``delta`` uploaded per level (gma/corr.py:34-37).

Faithfulness is pinned by ``tests/test_eager_ref_faithful.py`` (runs where /root/reference
exists): same outputs as the reference on CPU to 1e-5 px and the same ATen-op histogram
(``TorchDispatchMode``) for a RAFT pair, a GMA pair and an AccFlow clip.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import flow_oracle as fo

SD = Dict[str, torch.Tensor]


def _autocast(x: torch.Tensor, enabled: bool):
    """torch.cuda.amp.autocast(enabled=...) of the reference: fp16 on CUDA, a no-op on the CPU."""
    return torch.autocast("cuda", dtype=torch.float16, enabled=enabled and x.is_cuda)


# --------------------------------------------------------------------------- library-backed helpers
def _pixel_sampler(img, coords):
    """raft/utils/utils.py:66-80: split, normalise with (size-1), cat, F.grid_sample(align_corners=True)."""
    H, W = img.shape[-2:]
    xg, yg = coords.split([1, 1], dim=-1)
    xg = 2 * xg / (W - 1) - 1
    yg = 2 * yg / (H - 1) - 1
    return F.grid_sample(img, torch.cat([xg, yg], dim=-1), align_corners=True)


def _coords_grid(b, h, w, device, via_cpu=False):
    """raft/utils/utils.py:83-87 (device-side) / gma/gma.py:51-52 (built on the CPU, then moved)."""
    dev = torch.device("cpu") if via_cpu else device
    ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing="ij")
    g = torch.stack([xs, ys], dim=0).float()[None].repeat(b, 1, 1, 1)
    return g.to(device) if via_cpu else g


class _CorrVolume:
    """CorrBlock (raft/corr.py:8-55; gma/corr.py:8-58 differs only in where ``delta`` is built)."""

    def __init__(self, fmap1, fmap2, levels=4, radius=4, delta_on_cpu=False):
        b, d, h, w = fmap1.shape
        vol = torch.matmul(fmap1.view(b, d, h * w).transpose(1, 2), fmap2.view(b, d, h * w))
        vol = vol.view(b, h, w, 1, h, w) / torch.sqrt(torch.tensor(d).float())
        vol = vol.reshape(b * h * w, 1, h, w)
        self.levels, self.radius, self.delta_on_cpu = [vol], radius, delta_on_cpu
        for _ in range(levels - 1):
            vol = F.avg_pool2d(vol, 2, stride=2)
            self.levels.append(vol)

    def __call__(self, coords):
        r = self.radius
        coords = coords.permute(0, 2, 3, 1)
        b, h, w, _ = coords.shape
        outs = []
        for i, vol in enumerate(self.levels):
            if self.delta_on_cpu:
                dx = torch.linspace(-r, r, 2 * r + 1)
                dy = torch.linspace(-r, r, 2 * r + 1)
                delta = torch.stack(torch.meshgrid(dy, dx, indexing="ij"), dim=-1).to(coords.device)
            else:
                dx = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
                dy = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
                delta = torch.stack(torch.meshgrid(dy, dx, indexing="ij"), dim=-1)
            centroid = coords.reshape(b * h * w, 1, 1, 2) / 2 ** i
            window = centroid + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
            outs.append(_pixel_sampler(vol, window).view(b, h, w, -1))
        return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def _convex_upsample(flow, mask):
    """raft/raft.py:81-92: softmax over 9, F.unfold of 8*flow, weighted sum, pixel shuffle."""
    n, _, h, w = flow.shape
    m = torch.softmax(mask.view(n, 1, 9, 8, 8, h, w), dim=2)
    nb = F.unfold(8 * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
    up = torch.sum(m * nb, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(n, 2, 8 * h, 8 * w)


def _backwarp(img, flow):
    """networks/utils.py:96-124: the mesh is built on the CPU and uploaded on every call."""
    n, _, h, w = img.shape
    xx = torch.arange(0, w).view(1, -1).repeat(h, 1).view(1, 1, h, w).repeat(n, 1, 1, 1)
    yy = torch.arange(0, h).view(-1, 1).repeat(1, w).view(1, 1, h, w).repeat(n, 1, 1, 1)
    grid = torch.cat((xx, yy), 1).float()
    if img.is_cuda:
        grid = grid.cuda()
    v = grid + flow
    v[:, 0] = 2.0 * v[:, 0] / max(w - 1, 1) - 1.0
    v[:, 1] = 2.0 * v[:, 1] / max(h - 1, 1) - 1.0
    return F.grid_sample(img, v.permute(0, 2, 3, 1), mode="bilinear", padding_mode="zeros", align_corners=True)


def _get_occ(flow12, i1, i2, binary=True):
    """AccFlow_.py:127-135."""
    e = torch.abs(i1 - _backwarp(i2, flow12))
    if binary:
        e = torch.mean(e, dim=1, keepdim=True)
        return torch.where(e <= 1.0, torch.ones_like(e), torch.zeros_like(e))
    return e


def _downflow8(flow):
    """AccFlow_.py:138-142."""
    h, w = flow.shape[-2:]
    return F.interpolate(flow, size=(h // 8, w // 8), mode="bilinear", align_corners=True) / 8


def _encoder(sd: SD, pfx: str, images: List[torch.Tensor], norm: str):
    """BasicEncoder.forward (raft/extractor.py:201-225): list -> cat -> net -> split."""
    def nrm(t, name):
        if norm == "instance":
            return F.instance_norm(t, eps=1e-5)
        if norm == "batch":
            return F.batch_norm(t, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                                sd[name + ".bias"], False, 0.1, 1e-5)
        return t

    bdim = images[0].shape[0]
    x = torch.cat(images, dim=0) if len(images) > 1 else images[0]
    x = torch.relu(nrm(fo._conv(sd, pfx + "conv1", x, 2, 3), pfx + "norm1"))
    for stage, stride in ((1, 1), (2, 2), (3, 2)):
        for blk in (0, 1):
            p = f"{pfx}layer{stage}.{blk}."
            s = stride if blk == 0 else 1
            y = torch.relu(nrm(fo._conv(sd, p + "conv1", x, s, 1), p + "norm1"))
            y = torch.relu(nrm(fo._conv(sd, p + "conv2", y, 1, 1), p + "norm2"))
            if s != 1:
                x = nrm(fo._conv(sd, p + "downsample.0", x, s, 0), p + "norm3")
            x = torch.relu(x + y)
    x = fo._conv(sd, pfx + "conv2", x)
    return torch.split(x, [bdim] * len(images), dim=0) if len(images) > 1 else x


# --------------------------------------------------------------------------- pair estimator
def flow_estimator(sd: SD, image1, image2, iters: int = 12, flow_init=None, pfx: str = "", gma: Optional[bool] = None,
                   mixed_precision: bool = True):
    """RAFT.forward / RAFTGMA.forward exactly as the reference schedules them."""
    if gma is None:
        gma = (pfx + "att.to_qk.weight") in sd
    image1, image2 = image1.contiguous(), image2.contiguous()
    with _autocast(image1, mixed_precision):
        fmap1, fmap2 = _encoder(sd, pfx + "fnet.", [image1, image2], "instance")
    corr_fn = _CorrVolume(fmap1.float(), fmap2.float(), delta_on_cpu=gma)
    ub = pfx + "update_block."
    with _autocast(image1, mixed_precision):
        cnet = _encoder(sd, pfx + "cnet.", [image1], "batch")
        net, inp = torch.split(cnet, [128, 128], dim=1)
        net, inp = torch.tanh(net), torch.relu(inp)
        if gma:
            b, c, h, w = inp.shape                                            # gma/modules.py:54-76
            q, k = F.conv2d(inp, sd[pfx + "att.to_qk.weight"]).chunk(2, dim=1)
            q = q.reshape(b, 1, c, h * w).transpose(2, 3) * (c ** -0.5)
            k = k.reshape(b, 1, c, h * w).transpose(2, 3)
            attn = torch.einsum("bhid,bhjd->bhij", q, k).softmax(dim=-1)
    b, _, H, W = image1.shape
    coords0 = _coords_grid(b, H // 8, W // 8, image1.device, via_cpu=gma)
    coords1 = _coords_grid(b, H // 8, W // 8, image1.device, via_cpu=gma)
    if flow_init is not None:
        coords1 = coords1 + flow_init
    flow_up = None
    for _ in range(iters):
        corr = corr_fn(coords1)
        flow = coords1 - coords0
        with _autocast(image1, mixed_precision):
            mf = fo.motion_encoder(sd, ub + "encoder.", flow, corr)
            if gma:                                                           # gma/modules.py:102-115
                bb, cc, hh, ww = mf.shape
                v = F.conv2d(mf, sd[ub + "aggregator.to_v.weight"]).reshape(bb, 1, cc, hh * ww).transpose(2, 3)
                agg = torch.einsum("bhij,bhjd->bhid", attn, v).transpose(2, 3).reshape(bb, cc, hh, ww)
                mfg = mf + sd[ub + "aggregator.gamma"] * agg
                x = torch.cat([inp, mf, mfg], dim=1)
            else:
                x = torch.cat([inp, mf], dim=1)
            net = fo.sep_conv_gru(sd, ub + "gru.", net, x)
            delta = fo.flow_head(sd, ub + "flow_head.", net)
            up_mask = fo.mask_head(sd, ub + "mask.", net, 0.25)               # every iteration (raft.py:133-135)
        coords1 = coords1 + delta
        flow_up = _convex_upsample(coords1 - coords0, up_mask)                # every iteration (raft.py:139-146)
    return flow_up


# --------------------------------------------------------------------------- accumulation
def _acc_plus(sd: SD, df, f, o, c):
    """AccPlus.forward (AccFlow_.py:97-109) with torchvision's deform_conv2d."""
    from torchvision.ops import deform_conv2d
    p = "accplus."
    x = fo._conv(sd, p + "conv1.2", torch.relu(fo._conv(sd, p + "conv1.0", torch.cat([df, f, o], 1), pad=1)), pad=1)
    x = torch.relu(fo._conv(sd, p + "conv2.0", torch.cat([x, c], 1), pad=1))
    x = torch.relu(fo._conv(sd, p + "conv2.2", x, pad=1))
    x = fo._conv(sd, p + "conv2.4.conv", x, pad=1) * torch.exp(sd[p + "conv2.4.scale"] * 3)
    off, m = torch.split(x, [18, 9], dim=1)
    f_ = deform_conv2d(f, off, sd[p + "dconv.weight"], sd[p + "dconv.bias"], stride=1, padding=1, mask=torch.sigmoid(m))
    x = fo._conv(sd, p + "conv3.2", torch.relu(fo._conv(sd, p + "conv3.0", torch.cat([f_, df, o], 1), pad=1)), pad=1)
    x = torch.relu(fo._conv(sd, p + "conv4.0", torch.cat([x, c, f_, df], 1), pad=1))
    x = torch.relu(fo._conv(sd, p + "conv4.2", x, pad=1))
    return fo._conv(sd, p + "conv4.4", x)


def acc_iter(sd: SD, i1, i2, i_n, f2n, iters: int = 12, mixed_precision: bool = True):
    """AccFlow.iter (AccFlow_.py:177-201)."""
    if f2n is None:
        flows = flow_estimator(sd, torch.cat([i1, i1, i2]), torch.cat([i2, i_n, i_n]), iters, pfx="ofe.",
                               mixed_precision=mixed_precision)
        dflow, flow_ini, f2n = _downflow8(flows).chunk(3)
    else:
        flows = flow_estimator(sd, torch.cat([i1, i1]), torch.cat([i2, i_n]), iters, pfx="ofe.",
                               mixed_precision=mixed_precision)
        dflow, flow_ini = _downflow8(flows).chunk(2)
    b = i1.shape[0]
    with _autocast(i1, mixed_precision):
        f_ini, df, f = torch.split(fo.flow_encoder(sd, torch.cat([flow_ini, dflow, f2n], dim=0)), b, dim=0)
        c1, c2, cn = _encoder(sd, "context.", [i1, i2, i_n], "none")
        o = _get_occ(dflow, c1, c2)
        f_acc = _acc_plus(sd, df, f, o, c1)
        emap = _get_occ(flow_ini, c1, cn, binary=False)
        m = torch.sigmoid(fo._conv(sd, "blending.mask.2", torch.relu(fo._conv(sd, "blending.mask.0", emap)), pad=1))
        f_fuse = f_ini * m + (1 - m) * f_acc
        small = fo.flow_head_generic(sd, "flow_decoder.flow.", f_fuse)
        mask = fo._conv(sd, "flow_decoder.mask.2", torch.relu(fo._conv(sd, "flow_decoder.mask.0", f_fuse, pad=1)))
        out = _convex_upsample(small, mask)
    return small.float(), out.float()


def accflow_forward(sd: SD, images: List[torch.Tensor], iters: int = 12, mixed_precision: bool = True):
    """AccFlow.forward (AccFlow_.py:157-175)."""
    flow, outs = None, []
    for i in range(2, len(images)):
        flow, up = acc_iter(sd, images[i], images[i - 1], images[0], flow, iters, mixed_precision)
        outs.append(up)
    return outs


def configure_like_test_cvo(fp32_exact: bool = False):
    """test_cvo.py:115 sets cudnn.benchmark; ``fp32_exact`` pins both TF32 switches off."""
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = not fp32_exact
    torch.backends.cuda.matmul.allow_tf32 = False if fp32_exact else torch.backends.cuda.matmul.allow_tf32
