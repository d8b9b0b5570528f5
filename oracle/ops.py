"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU fp32 restatement of the tensor helpers on AccFlow's flow-estimation + accumulation path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this package, and only as the checker / reported CPU baseline.

Every sampling op is restated as explicit index arithmetic (floor, four corners, bounds
test) rather than by calling ``grid_sample`` / ``unfold`` / ``deform_conv2d``, so that the
conventions the CUDA kernels must reproduce are written down once, in the open:

* parity is PINNED: ``tests/golden/make_golden.py`` imports the reference from
  ``/root/reference`` in the build container, runs it on seeded inputs/weights and commits
  the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function
  here against those fixtures (and, when ``/root/reference`` is present, live).
* third-party arithmetic not vendored in the reference: torchvision ``deform_conv2d``
  (pinned 0.16.1 in environment.yml:160; golden generated with 0.26.0), ATen ``grid_sample``
  / ``avg_pool2d`` / ``interpolate`` (torch 2.1.1 pinned; golden generated with 2.11.0).

Dense convolutions / matmuls use torch's fp32 CPU kernels (the "plain fp32 reference" for a
floating-point kernel).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- normalisation
def instance_norm(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.InstanceNorm2d(affine=False): per (n,c) plane, biased variance
    (raft/extractor.py:35-38, :150-151)."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def batch_norm_eval(x, weight, bias, running_mean, running_var, eps: float = 1e-5):
    """nn.BatchNorm2d in eval mode (raft/extractor.py:29-33; test_cvo.py:14 ``.eval()``)."""
    inv = weight / torch.sqrt(running_var + eps)
    return x * inv.view(1, -1, 1, 1) + (bias - running_mean * inv).view(1, -1, 1, 1)


# --------------------------------------------------------------------------- correlation
def corr_volume(fmap1: torch.Tensor, fmap2: torch.Tensor) -> torch.Tensor:
    """All-pairs correlation, raft/corr.py:47-55.  Returns (B*h*w, 1, h, w) fp32."""
    b, d, h, w = fmap1.shape
    f1 = fmap1.reshape(b, d, h * w)
    f2 = fmap2.reshape(b, d, h * w)
    corr = torch.einsum("bdp,bdq->bpq", f1, f2) / math.sqrt(d)
    return corr.reshape(b * h * w, 1, h, w)


def avg_pool2(x: torch.Tensor) -> torch.Tensor:
    """F.avg_pool2d(x, 2, stride=2): floor on odd sizes (raft/corr.py:21)."""
    h2, w2 = x.shape[-2] // 2, x.shape[-1] // 2
    x = x[..., : 2 * h2, : 2 * w2]
    return (x[..., 0::2, 0::2] + x[..., 0::2, 1::2] + x[..., 1::2, 0::2] + x[..., 1::2, 1::2]) / 4


def corr_pyramid(fmap1, fmap2, num_levels: int = 4):
    """raft/corr.py:8-22."""
    pyr = [corr_volume(fmap1, fmap2)]
    for _ in range(num_levels - 1):
        pyr.append(avg_pool2(pyr[-1]))
    return pyr


def _sample_zeros(img: torch.Tensor, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Bilinear sample with zero padding at *pixel* coordinates.

    img (N,C,H,W); x,y (N,...) pixel coords -> (N,C,...).  Mirrors what
    ``bilinear_sampler`` + ``F.grid_sample(align_corners=True)`` evaluate
    (raft/utils/utils.py:66-80): normalise to [-1,1] with (size-1), un-normalise, floor,
    four corners, each corner contributes only when it lies inside the image.
    """
    n, c, h, w = img.shape
    xn = 2 * x / (w - 1) - 1
    yn = 2 * y / (h - 1) - 1
    ix = ((xn + 1) / 2) * (w - 1)
    iy = ((yn + 1) / 2) * (h - 1)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    fx = ix - x0
    fy = iy - y0
    out_shape = x.shape[1:]
    flat = img.reshape(n, c, h * w)
    res = torch.zeros((n, c) + tuple(out_shape), dtype=img.dtype, device=img.device)
    for dy, dx, wgt in ((0, 0, (1 - fx) * (1 - fy)), (0, 1, fx * (1 - fy)),
                        (1, 0, (1 - fx) * fy), (1, 1, fx * fy)):
        xi = x0 + dx
        yi = y0 + dy
        ok = (xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)
        lin = (yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)).long().reshape(n, 1, -1)
        v = torch.gather(flat, 2, lin.expand(n, c, -1)).reshape((n, c) + tuple(out_shape))
        res = res + v * (wgt * ok).unsqueeze(1)
    return res


def corr_lookup(pyramid, coords: torch.Tensor, radius: int = 4) -> torch.Tensor:
    """CorrBlock.__call__, raft/corr.py:24-45.

    coords (B,2,h,w) with channel 0 = x, 1 = y.  Output (B, L*(2r+1)^2, h, w); channel
    ``l*(2r+1)^2 + a*(2r+1) + b`` samples level l at (x/2^l + a - r, y/2^l + b - r)
    — the *first* window index moves x.
    """
    b, _, h, w = coords.shape
    n = b * h * w
    k = 2 * radius + 1
    d = torch.arange(-radius, radius + 1, dtype=coords.dtype, device=coords.device)
    cx = coords[:, 0].reshape(n, 1, 1)
    cy = coords[:, 1].reshape(n, 1, 1)
    outs = []
    for lvl, corr in enumerate(pyramid):
        xs = cx / 2 ** lvl + d.view(1, k, 1)       # index a -> x offset
        ys = cy / 2 ** lvl + d.view(1, 1, k)       # index b -> y offset
        xs, ys = torch.broadcast_tensors(xs, ys)
        s = _sample_zeros(corr, xs, ys)            # (n,1,k,k)
        outs.append(s.reshape(b, h, w, k * k))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous()


# --------------------------------------------------------------------------- flow helpers
def coords_grid(b: int, h: int, w: int, device=None) -> torch.Tensor:
    """raft/utils/utils.py:83-87: channel 0 = x (column index), channel 1 = y."""
    ys, xs = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(b, 1, 1, 1)


def convex_upsample(flow: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """RAFT.upsample_flow, raft/raft.py:81-92 (same code gma.py:57-68, AccFlow_.py:27-38).

    out[n,c,8y+i,8x+j] = sum_k softmax_k(mask[n, k*64+i*8+j, y, x]) * 8*flow[n,c,y+ky-1,x+kx-1]
    with k = ky*3+kx and zero padding of the flow.
    """
    n, _, h, w = flow.shape
    m = mask.reshape(n, 9, 8, 8, h, w)
    m = torch.softmax(m, dim=1)
    fp = F.pad(8 * flow, (1, 1, 1, 1))
    out = torch.zeros(n, 2, 8, 8, h, w, dtype=flow.dtype, device=flow.device)
    for ky in range(3):
        for kx in range(3):
            nb = fp[:, :, ky:ky + h, kx:kx + w]                  # (n,2,h,w)
            out = out + m[:, ky * 3 + kx].unsqueeze(1) * nb.reshape(n, 2, 1, 1, h, w)
    return out.permute(0, 1, 4, 2, 5, 3).reshape(n, 2, 8 * h, 8 * w)


def backwarp(img: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """networks/utils.py:96-124: sample img at (x+flow_x, y+flow_y), zeros, align_corners."""
    n, _, h, w = img.shape
    g = coords_grid(n, h, w, img.device)
    x = g[:, 0] + flow[:, 0]
    y = g[:, 1] + flow[:, 1]
    # reference normalises with max(size-1, 1) — identical for size >= 2
    return _sample_zeros(img, x, y)


def resize_bilinear_ac(x: torch.Tensor, oh: int, ow: int) -> torch.Tensor:
    """F.interpolate(mode='bilinear', align_corners=True) restated (AccFlow_.py:141)."""
    n, c, h, w = x.shape
    sy = (h - 1) / (oh - 1) if oh > 1 else 0.0
    sx = (w - 1) / (ow - 1) if ow > 1 else 0.0
    ys = torch.arange(oh, dtype=torch.float32, device=x.device) * torch.tensor(sy, dtype=torch.float32, device=x.device)
    xs = torch.arange(ow, dtype=torch.float32, device=x.device) * torch.tensor(sx, dtype=torch.float32, device=x.device)
    y0 = ys.floor().long().clamp(max=h - 1)
    x0 = xs.floor().long().clamp(max=w - 1)
    y1 = (y0 + 1).clamp(max=h - 1)
    x1 = (x0 + 1).clamp(max=w - 1)
    ly = (ys - y0).view(1, 1, oh, 1)
    lx = (xs - x0).view(1, 1, 1, ow)
    top = x[:, :, y0][:, :, :, x0] * (1 - lx) + x[:, :, y0][:, :, :, x1] * lx
    bot = x[:, :, y1][:, :, :, x0] * (1 - lx) + x[:, :, y1][:, :, :, x1] * lx
    return top * (1 - ly) + bot * ly


def downflow8(flow: torch.Tensor) -> torch.Tensor:
    """AccFlow_.py:138-142: align_corners bilinear resize to 1/8 size, then /8."""
    h, w = flow.shape[-2:]
    assert h % 8 == 0 and w % 8 == 0
    return resize_bilinear_ac(flow, h // 8, w // 8) / 8


def get_occ(flow12, i1, i2, binary: bool = True):
    """getOcc, AccFlow_.py:127-135."""
    e = torch.abs(i1 - backwarp(i2, flow12))
    if binary:
        e = e.mean(dim=1, keepdim=True)
        return (e <= 1.0).to(i1.dtype)
    return e


# --------------------------------------------------------------------------- deformable conv
def deform_conv2d(x, offset, mask, weight, bias):
    """torchvision.ops.deform_conv2d, 3x3 / stride 1 / pad 1 / dil 1 / 1 offset group
    (call site AccFlow_.py:83,104).

    offset (N,18,H,W): channel 2k = dy, 2k+1 = dx for tap k = ki*3+kj; sample position
    (y-1+ki+dy, x-1+kj+dx); a tap whose centre is <= -1 or >= size is dropped, otherwise
    corners outside the image contribute 0; sample is multiplied by mask[:,k]; then the
    ordinary weight contraction over (cin, tap) and + bias  (SURVEY.md §4 table).
    """
    n, c, h, w = x.shape
    g = coords_grid(n, h, w, x.device)
    cols = []
    flat = x.reshape(n, c, h * w)
    for k in range(9):
        ki, kj = divmod(k, 3)
        py = g[:, 1] - 1 + ki + offset[:, 2 * k]
        px = g[:, 0] - 1 + kj + offset[:, 2 * k + 1]
        valid = (py > -1) & (py < h) & (px > -1) & (px < w)
        y0 = torch.floor(py)
        x0 = torch.floor(px)
        fy = py - y0
        fx = px - x0
        acc = torch.zeros(n, c, h, w, dtype=x.dtype, device=x.device)
        for dy, dx, wgt in ((0, 0, (1 - fy) * (1 - fx)), (0, 1, (1 - fy) * fx),
                            (1, 0, fy * (1 - fx)), (1, 1, fy * fx)):
            yi = y0 + dy
            xi = x0 + dx
            ok = valid & (yi >= 0) & (yi <= h - 1) & (xi >= 0) & (xi <= w - 1)
            lin = (yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)).long().reshape(n, 1, -1)
            v = torch.gather(flat, 2, lin.expand(n, c, -1)).reshape(n, c, h, w)
            acc = acc + v * (wgt * ok).unsqueeze(1)
        cols.append(acc * mask[:, k:k + 1])
    col = torch.stack(cols, dim=2)                       # (n, c, 9, h, w)
    wmat = weight.reshape(weight.shape[0], c * 9)        # (cout, c*9), tap fastest
    out = torch.einsum("ok,nkp->nop", wmat, col.reshape(n, c * 9, h * w))
    return out.reshape(n, -1, h, w) + bias.view(1, -1, 1, 1)


# --------------------------------------------------------------------------- metrics
def calc_occ_mask(bflow, fflow):
    """test_cvo.py:53-78 (the script body runs at import, so it is restated, not imported)."""
    def length(v):
        return torch.sqrt(torch.sum(v ** 2, dim=1, keepdim=True))

    mag = length(fflow) + length(bflow)
    diff_fw = fflow + backwarp(bflow, fflow)
    diff_bw = bflow + backwarp(fflow, bflow)
    thresh = 0.01 * mag + 0.5
    return (length(diff_bw) > thresh).float(), (length(diff_fw) > thresh).float()


def cal_epe(pred, label, occ_mask):
    """test_cvo.py:81-101 -> (epe_all, epe_occ, epe_vis), each (N,)."""
    diff = torch.sqrt(torch.sum((pred - label) ** 2, dim=1, keepdim=True))
    epe_all = diff.mean(dim=(1, 2, 3))
    epe_occ = (diff * occ_mask).sum(dim=(1, 2, 3)) / occ_mask.sum(dim=(1, 2, 3))
    epe_vis = (diff * (1 - occ_mask)).sum(dim=(1, 2, 3)) / (1 - occ_mask).sum(dim=(1, 2, 3))
    return epe_all, epe_occ, epe_vis
