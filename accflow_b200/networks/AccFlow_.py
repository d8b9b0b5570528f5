"""AccFlow backward accumulation — same surface as networks/AccFlow_.py:145-201."""
from __future__ import annotations

import torch
from torch import nn

from .. import spec as S
from . import _tree
from ._estimator import FlowEstimatorBase


def downflow8(flow, mode="bilinear"):
    """(N,2,H,W) -> (N,2,H/8,W/8): align_corners bilinear resize, then /8 (AccFlow_.py:138-142)."""
    from ..ops import downflow8 as _d8
    assert mode == "bilinear"
    return _d8(flow)


def getOcc(F12, I1, I2, binary=True):
    """|I1 - backwarp(I2, F12)|; binary: channel mean <= 1 -> 1 else 0 (AccFlow_.py:127-135)."""
    from ..ops import get_occ
    return get_occ(F12, I1, I2, binary)


class AccFlow(nn.Module):
    def __init__(self, ofe: nn.Module):
        super().__init__()
        if not isinstance(ofe, FlowEstimatorBase):
            raise TypeError("AccFlow expects a flow estimator built by accflow_b200.networks.build_flow_estimator")
        self.ofe = ofe
        self.hidden_channel = 128
        self.mixed_precision = True
        self.iters = 12          # additive: the reference hard-codes the ofe default (AccFlow_.py:184,188)
        # additive (the reference README's TODO "Add warmstart mode"): chain the previous step's flows as flow_init of
        # the next step's estimator calls (raft/raft.py:123-124); warm_iters = iterations of the warm-started calls
        self.warm_start = False
        self.warm_iters = None
        ofe_keys = ("ofe.",)
        _tree.populate(self, [e for e in S.accflow_entries("gma" if ofe._GMA else "raft")
                              if not e.name.startswith(ofe_keys)])
        self._engines = {}          # shared with DataParallel replicas (shallow __dict__ copy)
        _tree.register_source(self)

    def engine(self, device=None):
        from ..engine import AccFlowEngine
        device = next(self.parameters()).device if device is None else device
        if device.type != "cuda":
            raise RuntimeError("accflow_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        precision = self.ofe.precision
        src = _tree.source_of(self)          # DataParallel replica -> the module it was made from (see _estimator.engine)
        sig = (_tree.signature(src), precision)
        key = (device, precision)
        hit = self._engines.get(key)
        if hit is None or hit[0] != sig:
            hit = (sig, AccFlowEngine(dict(src.state_dict()), device, self.ofe._GMA, precision))
            self._engines[key] = hit
        return hit[1]

    @torch.no_grad()
    def iter(self, I1, I2, In, F2n):
        """input: I1, I2, IN; F2N (1/8 size) -> F1N_small (1/8 size), F1N."""
        return self.engine(I1.device if I1.is_cuda else None).iter(I1, I2, In, F2n, self.iters)

    @torch.no_grad()
    def forward(self, images, test_mode=False):
        """[I0, I1, ..., In] -> [F(2->0), F(3->0), ..., F(n->0)] (``test_mode`` is ignored, as in the reference)."""
        images = list(images)
        eng = self.engine(images[0].device if images[0].is_cuda else None)
        return eng.forward(images, self.iters, graph=self.ofe.use_cuda_graph, warm_start=self.warm_start,
                           warm_iters=self.warm_iters)
