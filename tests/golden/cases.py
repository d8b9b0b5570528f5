"""Seeded inputs for the golden fixtures (shared by make_golden.py and the parity tests).

Only *outputs* of the reference are stored in golden_v1.npz; inputs and weights are
regenerated from these seeds on both sides (a float64 checksum of every input is stored so
RNG drift is detected instead of silently comparing different problems).
"""
from __future__ import annotations

import torch

from accflow_b200.data import make_clip
from accflow_b200.weights import make_state_dict


def _g(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def corr_case():
    g = _g(11)
    b, d, h, w = 1, 16, 20, 18            # odd pyramid sizes: 20x18, 10x9, 5x4, 2x2
    f1 = torch.randn(b, d, h, w, generator=g)
    f2 = torch.randn(b, d, h, w, generator=g)
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    grid = torch.stack([xs, ys], 0).float()[None]
    coords = grid + torch.randn(b, 2, h, w, generator=g) * 6.0   # many taps leave the map
    coords[0, :, 0, 0] = torch.tensor([-7.3, 2.2])               # fully outside
    coords[0, :, 1, 1] = torch.tensor([3.0, 4.0])                # integer coords
    return f1, f2, coords


def upsample_case():
    g = _g(12)
    return torch.randn(2, 2, 6, 5, generator=g) * 3, torch.randn(2, 576, 6, 5, generator=g) * 2


def warp_case():
    g = _g(13)
    img = torch.randn(2, 4, 9, 11, generator=g)
    flow = torch.randn(2, 2, 9, 11, generator=g) * 3
    flow[0, :, 0, 0] = torch.tensor([-30.0, 1.0])
    flow[1, :, 3, 3] = torch.tensor([2.0, -1.0])
    return img, flow


def downflow_case():
    return (torch.randn(2, 2, 64, 48, generator=_g(14)) * 5,)


def occ_case():
    g = _g(15)
    c1 = torch.randn(2, 8, 10, 12, generator=g) * 1.2
    c2 = torch.randn(2, 8, 10, 12, generator=g) * 1.2
    flow = torch.randn(2, 2, 10, 12, generator=g) * 2
    return flow, c1, c2


def dcn_case():
    g = _g(16)
    x = torch.randn(2, 8, 7, 9, generator=g)
    off = torch.randn(2, 18, 7, 9, generator=g) * 2.5
    off[0, :, 0, 0] = -1.0                                         # exactly on the -1 border
    off[0, :, 6, 8] = 1.0                                          # exactly on the +size border
    mask = torch.rand(2, 9, 7, 9, generator=g)
    wgt = torch.randn(6, 8, 3, 3, generator=g) * 0.3
    bias = torch.randn(6, generator=g)
    return x, off, mask, wgt, bias


def metric_case():
    g = _g(17)
    bflow = torch.randn(2, 2, 24, 20, generator=g) * 2 + 3
    fflow = -bflow + torch.randn(2, 2, 24, 20, generator=g) * 0.6
    pred = bflow + torch.randn(2, 2, 24, 20, generator=g) * 0.4
    return bflow, fflow, pred


def gma_case():
    g = _g(18)
    inp = torch.relu(torch.randn(2, 128, 6, 5, generator=g))
    mf = torch.randn(2, 128, 6, 5, generator=g)
    return inp, mf


def acc_modules_case():
    g = _g(19)
    b, c, h, w = 2, 128, 8, 10
    t = lambda *s, k=1.0: torch.randn(*s, generator=g) * k
    return dict(df=t(b, c, h, w), f=t(b, c, h, w), c=t(b, c, h, w), o=(torch.rand(b, 1, h, w, generator=g) > 0.3).float(),
                emap=t(b, c, h, w).abs(), f1=t(b, c, h, w), f2=t(b, c, h, w), flows=t(3 * b, 2, h, w, k=2.0))


def pair_case(size=128):
    clip = make_clip(3, size=size)
    flow_init = torch.randn(1, 2, size // 8, size // 8, generator=_g(20)) * 1.5
    return clip["imgs"][2], clip["imgs"][0], flow_init


def clip_case(size=128, frames=4):
    return make_clip(5, size=size, frames=7)["imgs"][:frames]


WEIGHT_SEED = {"raft": 1, "gma": 1, "acc+raft": 2, "acc+gma": 2}


def weights(kind):
    return make_state_dict(kind, seed=WEIGHT_SEED[kind])


def checksum(*tensors) -> float:
    return float(sum(t.double().abs().sum() for t in tensors))
