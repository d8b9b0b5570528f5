"""Evaluation metrics of test_cvo.py (:53-101) on the device: bidirectional occlusion mask and
EPE all / occ / vis.  The two full-resolution backwarps run in the CUDA kernel
(accflow_backwarp_nchw_f32); the remaining reductions are a handful of elementwise torch ops
on (N,1,H,W) maps (metric glue, not the hot path)."""
from __future__ import annotations

import torch

from .ops import backwarp


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
