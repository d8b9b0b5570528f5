"""NCHW-tensor front ends of individual kernels (the reference's free functions).

These exist for API parity (``backwarp``, ``upsample_flow``, ``downflow8``, ``getOcc``) and
for per-kernel parity tests; the engines call the C ABI directly on NHWC workspaces.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

F32 = torch.float32


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _cuda(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("accflow_b200.ops: CUDA tensors only (no CPU path)")
    return t.to(F32).contiguous()


def nhwc(t: torch.Tensor) -> torch.Tensor:
    return _cuda(t).permute(0, 2, 3, 1).contiguous()


def nchw(t: torch.Tensor) -> torch.Tensor:
    return t.permute(0, 3, 1, 2).contiguous()


def coords_grid(batch, ht, wd, device="cuda"):
    out = torch.empty(batch, ht * wd, 2, device=device, dtype=F32)
    with torch.cuda.device(out.device):
        L.call("accflow_coords_init_f32", None, batch, ht, wd, out.data_ptr(), _s())
    return out.view(batch, ht, wd, 2).permute(0, 3, 1, 2).contiguous()


def convex_upsample(flow: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    n, _, h, w = flow.shape
    f, m = nhwc(flow), nhwc(mask)
    out = torch.empty(n, 2, 8 * h, 8 * w, device=f.device, dtype=F32)
    with torch.cuda.device(f.device):
        L.call("accflow_convex_upsample_f32", f.data_ptr(), 2, 0, m.data_ptr(), 576, n, h, w, out.data_ptr(), _s())
    return out


def backwarp(image: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    image, flow = _cuda(image), _cuda(flow)
    n, c, h, w = image.shape
    out = torch.empty_like(image)
    with torch.cuda.device(image.device):
        L.call("accflow_backwarp_nchw_f32", image.data_ptr(), flow.data_ptr(), n, c, h, w, out.data_ptr(), _s())
    return out


def downflow8(flow: torch.Tensor) -> torch.Tensor:
    flow = _cuda(flow)
    n, _, H, W = flow.shape
    out = torch.empty(n, H // 8, W // 8, 2, device=flow.device, dtype=F32)
    with torch.cuda.device(flow.device):
        L.call("accflow_downflow8_f32", flow.data_ptr(), n, H, W, out.data_ptr(), _s())
    return nchw(out)


def upflow8(flow: torch.Tensor, mode: str = "bilinear") -> torch.Tensor:
    """8 * F.interpolate(flow, 8x, bilinear, align_corners=True) (networks/utils.py:91-93)."""
    if mode != "bilinear":
        raise ValueError("upflow8: only the reference's default mode 'bilinear' exists here")
    flow = _cuda(flow)
    n, c, h, w = flow.shape
    out = torch.empty(n, c, 8 * h, 8 * w, device=flow.device, dtype=F32)
    with torch.cuda.device(flow.device):
        L.call("accflow_upflow8_f32", flow.data_ptr(), n, c, h, w, out.data_ptr(), _s())
    return out


def get_occ(flow12, i1, i2, binary=True):
    c1, c2 = nhwc(i1), nhwc(i2)
    fl = nhwc(flow12)
    n, h, w, c = c1.shape
    with torch.cuda.device(c1.device):
        if binary:
            occ = torch.empty(n, h, w, 1, device=c1.device, dtype=F32)
            L.call("accflow_warp_occ_f32", c1.data_ptr(), c, c2.data_ptr(), c, fl.data_ptr(), n, h, w, c,
                   occ.data_ptr(), 1, None, 0, _s())
            return nchw(occ)
        emap = torch.empty_like(c1)
        L.call("accflow_warp_occ_f32", c1.data_ptr(), c, c2.data_ptr(), c, fl.data_ptr(), n, h, w, c, None, 0,
               emap.data_ptr(), c, _s())
        return nchw(emap)
