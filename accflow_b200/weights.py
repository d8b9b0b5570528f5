"""Seeded, de-degenerated weight factory.

There are no checkpoints offline, so every test / benchmark uses random weights.  The
reference's *default* init hides bugs (SURVEY.md §7 "hard part 3"): ZeroConv2d outputs 0
(deformable offsets never exercised), ``Aggregate.gamma`` is 0, BatchNorm running stats are
(0, 1).  This factory draws every tensor of the state-dict contract (``spec.py``) from a CPU
``torch.Generator`` so that the oracle, the reference (when importable) and the CUDA path are
all loaded from the *same* dictionary.

Scales are chosen so that activations stay O(1) through 12 recurrent iterations and so that
the per-iteration flow update is a fraction of a 1/8-res pixel (a chaotic recurrence would
make a 1e-3 px parity bar meaningless).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import spec as S


def make_state_dict(kind: str, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Return an ordered state_dict for ``kind`` in {'raft','gma','acc+raft','acc+gma'}."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + 17)
    sd: Dict[str, torch.Tensor] = {}
    for e in S.entries_for(kind):
        if e.alias_of is not None:
            sd[e.name] = sd[e.alias_of]
            continue
        shp = e.shape
        if e.role == S.CONV_W:
            fan_in = shp[1] * shp[2] * shp[3]
            gain = 1.0
            # heads that feed the recurrence / the deformable sampler get smaller gains
            if e.name.endswith("flow_head.conv2.weight"):
                gain = 0.05
            elif e.name.endswith("flow_decoder.flow.2.weight"):
                gain = 0.25
            t = torch.randn(shp, generator=g) * (gain * math.sqrt(1.6 / fan_in))
        elif e.role == S.CONV_B:
            t = (torch.rand(shp, generator=g) - 0.5) * 0.1
        elif e.role == S.ZCONV_W:
            fan_in = shp[1] * shp[2] * shp[3]
            t = torch.randn(shp, generator=g) * (1.2 * math.sqrt(1.0 / fan_in))
        elif e.role == S.ZCONV_B:
            t = (torch.rand(shp, generator=g) - 0.5) * 0.4
        elif e.role == S.ZSCALE:
            t = (torch.rand(shp, generator=g) - 0.5) * 0.2
        elif e.role == S.BN_W:
            t = 0.6 + 0.8 * torch.rand(shp, generator=g)
        elif e.role == S.BN_B:
            t = (torch.rand(shp, generator=g) - 0.5) * 0.4
        elif e.role == S.BN_RM:
            t = torch.randn(shp, generator=g) * 0.2
        elif e.role == S.BN_RV:
            t = 0.5 + torch.rand(shp, generator=g)
        elif e.role == S.BN_NBT:
            t = torch.tensor(100, dtype=torch.int64)
        elif e.role == S.GAMMA:
            t = torch.full(shp, 0.5)
        elif e.role == S.EMB:
            t = torch.randn(shp, generator=g)
        elif e.role == S.RELIND:
            n = shp[0]
            t = torch.arange(n).view(1, -1) - torch.arange(n).view(-1, 1) + n - 1
        else:  # pragma: no cover
            raise ValueError(e.role)
        if t.is_floating_point():
            t = t.to(dtype)
        sd[e.name] = t.contiguous()
    return sd
