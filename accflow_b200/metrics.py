"""Evaluation metrics of test_cvo.py (:53-101) on the device: bidirectional occlusion mask and
EPE all / occ / vis.  The two full-resolution backwarps run in the CUDA kernel
(accflow_backwarp_nchw_f32); the remaining reductions are a handful of elementwise torch ops
on (N,1,H,W) maps (metric glue, not the hot path)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .ops import backwarp


def clip_epe(pred: torch.Tensor, bflow: torch.Tensor, fflow: torch.Tensor) -> torch.Tensor:
    """Fused device path of ``cal_epe(pred, bflow, calc_occ_mask(bflow, fflow)[0])``:
    -> (N, 3) tensor of per-clip (epe_all, epe_occ, epe_vis) from one kernel pass."""
    pred, bflow, fflow = (t.to(torch.float32).contiguous() for t in (pred, bflow, fflow))
    if not pred.is_cuda:
        raise RuntimeError("accflow_b200.metrics: CUDA tensors only (no CPU path)")
    n, _, h, w = pred.shape
    partial = torch.empty(n * ((h * w + 255) // 256) * 3, device=pred.device, dtype=torch.float32)
    out = torch.empty(n, 3, device=pred.device, dtype=torch.float32)
    with torch.cuda.device(pred.device):
        L.call("accflow_epe_metrics_f32", pred.data_ptr(), bflow.data_ptr(), fflow.data_ptr(), n, h, w,
               partial.data_ptr(), out.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    return out


def _length(v: torch.Tensor) -> torch.Tensor:
    return torch.sqrt(torch.sum(v * v, dim=1, keepdim=True))


def calc_occ_mask(bflow: torch.Tensor, fflow: torch.Tensor):
    """-> (occ_bw, occ_fw) binary (N,1,H,W); 1 = occluded (test_cvo.py:53-78)."""
    mag = _length(fflow) + _length(bflow)
    thresh = 0.01 * mag + 0.5
    diff_fw = fflow + backwarp(bflow, fflow)
    diff_bw = bflow + backwarp(fflow, bflow)
    return (_length(diff_bw) > thresh).float(), (_length(diff_fw) > thresh).float()


def cal_epe(pred: torch.Tensor, label: torch.Tensor, occ_mask: torch.Tensor):
    """-> (epe_all, epe_occ, epe_vis), each (N,) (test_cvo.py:81-101)."""
    diff = _length(pred - label)
    epe_all = diff.mean(dim=(1, 2, 3))
    epe_occ = (diff * occ_mask).sum(dim=(1, 2, 3)) / occ_mask.sum(dim=(1, 2, 3))
    vis = 1 - occ_mask
    epe_vis = (diff * vis).sum(dim=(1, 2, 3)) / vis.sum(dim=(1, 2, 3))
    return epe_all, epe_occ, epe_vis
