"""Synthetic CVO-shaped clips (the CVO LMDB and checkpoints are not available offline).

Shapes follow the reference loader / evaluation script (data/README.md:8-23,
test_cvo.py:32-50, :139-141): a clip is 7 RGB frames of HxW in [-1,1] (= 2*(u8/255)-1), with
long-range backward flows F(i->0) and forward flows F(0->i), i = 2..6.

Generator: a smooth random texture is translated by an integer-rounded
per-clip velocity, so the ground-truth long-range flows are exact constants and the border
strip that leaves the frame gives a non-empty occlusion mask.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F


def make_clip(clip_id: int, size: int = 512, frames: int = 7, seed: int = 1234) -> Dict[str, object]:
    """Returns {'imgs': [frames x (1,3,S,S) fp32], 'bflows': [...], 'fflows': [...], 'u8': (frames,3,S,S) uint8}."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + clip_id)
    margin = size // 8
    tex_lr = torch.rand(1, 3, (size + 2 * margin) // 8, (size + 2 * margin) // 8, generator=g) * 255.0
    tex = F.interpolate(tex_lr, size=(size + 2 * margin, size + 2 * margin), mode="bicubic",
                        align_corners=False).clamp(0, 255).round()
    vel = (torch.rand(2, generator=g) * 8.0 - 4.0) * (size / 512.0)       # px / frame, (dx, dy)
    if float(vel.abs().max()) < 0.25:
        vel = vel + 1.0
    u8 = []
    shifts = []
    for k in range(frames):
        dx = int(round(float(vel[0]) * k))
        dy = int(round(float(vel[1]) * k))
        dx = max(-margin, min(margin, dx))
        dy = max(-margin, min(margin, dy))
        shifts.append((dx, dy))
        u8.append(tex[0, :, margin + dy: margin + dy + size, margin + dx: margin + dx + size])
    u8 = torch.stack(u8).to(torch.uint8)
    imgs = [(2.0 * (u8[k].float() / 255.0) - 1.0)[None].contiguous() for k in range(frames)]
    # frame k shows texture shifted by +shift_k, i.e. content moves by -(shift_k) in image space:
    # a point at x in frame 0 is at x - (s_i - s_0) in frame i.
    bflows: List[torch.Tensor] = []
    fflows: List[torch.Tensor] = []
    for i in range(2, frames):
        dx, dy = shifts[i][0] - shifts[0][0], shifts[i][1] - shifts[0][1]
        f0i = torch.tensor([-dx, -dy], dtype=torch.float32).view(1, 2, 1, 1).expand(1, 2, size, size)
        fflows.append(f0i.contiguous())
        bflows.append((-f0i).contiguous())
    return {"imgs": imgs, "bflows": bflows, "fflows": fflows, "u8": u8, "shifts": shifts}


def make_batch(clip_ids, size: int = 512, frames: int = 7, seed: int = 1234):
    """Batch several clips along dim 0 (the reference evaluates with batch 10, test_cvo.py:114)."""
    clips = [make_clip(c, size, frames, seed) for c in clip_ids]
    imgs = [torch.cat([c["imgs"][k] for c in clips], 0) for k in range(frames)]
    bflows = [torch.cat([c["bflows"][k] for c in clips], 0) for k in range(frames - 2)]
    fflows = [torch.cat([c["fflows"][k] for c in clips], 0) for k in range(frames - 2)]
    return {"imgs": imgs, "bflows": bflows, "fflows": fflows}


def decode_cvo_flow_u16(raw: torch.Tensor) -> torch.Tensor:
    """CVO stores flows as uint16 fixed point (data/dataset.py:65-67): flow = (v - 2**15) / 128.
    ``raw``: integer tensor (..., H, W, C) or (..., C, H, W) as read from the LMDB record; returns fp32."""
    return (raw.to(torch.float32) - 32768.0) / 128.0


def encode_cvo_flow_u16(flow: torch.Tensor) -> torch.Tensor:
    """Inverse of :func:`decode_cvo_flow_u16` (round to nearest, clamp to the uint16 range); int32 container."""
    return torch.clamp(torch.round(flow * 128.0 + 32768.0), 0, 65535).to(torch.int32)


def preprocess(batch: Dict[str, torch.Tensor], device=None) -> Dict[str, List[torch.Tensor]]:
    """test_cvo.py:32-50: 'imgs' (B,21,H,W) uint8-valued -> 7 x (B,3,H,W) in [-1,1]; '*flows' (B,10,H,W) ->
    5 x (B,2,H,W).  Same assertions as the reference."""
    out = {}
    for key, value in batch.items():
        if device is not None:
            value = value.to(device)
        if "flow" in key:
            value = list(value.split(2, dim=1))
            assert len(value) in (5, 6), len(value)
        elif "imgs" in key:
            value = list((2 * (value / 255.0) - 1).split(3, dim=1))
            assert len(value) == 7, len(value)
        else:
            raise ValueError(key)
        out[key] = value
    return out
