"""oracle/eager_ref.py (the "reference as executed" proxy timed as the north-star denominator on the GPU box)
must do what the reference does: same outputs and the same ATen compute-op histogram.  Runs only where the
reference checkout exists (the build container); CPU only."""
import collections
import os
import sys

import pytest
import torch
from torch.utils._python_dispatch import TorchDispatchMode

from tests.golden import cases

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
torch.set_grad_enabled(False)

# layout-only ops (einops / view / cat-split bookkeeping differ legitimately between an nn.Module tree and a
# functional restatement) are not part of the comparison; everything that launches arithmetic is.
LAYOUT = {"view", "_unsafe_view", "reshape", "permute", "transpose", "t", "expand", "slice", "select", "split",
          "split_with_sizes", "chunk", "unsqueeze", "squeeze", "detach", "alias", "clone", "contiguous", "_to_copy",
          "as_strided", "unbind", "empty", "empty_like", "empty_strided", "zeros", "ones", "lift_fresh", "copy_",
          "unfold", "scalar_tensor", "full", "new_empty", "zeros_like", "ones_like", "fill_", "zero_", "repeat",
          "arange", "stack", "cat"}


class OpCount(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.counts = collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = func.overloadpacket.__name__
        if name.endswith("_") and not name.startswith("_"):
            name = name[:-1]                 # in-place variant: the same kernel (nn.ReLU(inplace=True) vs torch.relu)
        if name not in LAYOUT:
            self.counts[name] += 1
        return func(*args, **(kwargs or {}))


def _reference_models(kind):
    sys.path.insert(0, REF)
    try:
        from networks import build_flow_estimator
        from networks.AccFlow_ import AccFlow
    finally:
        sys.path.remove(REF)
    ofe = build_flow_estimator(kind.split("+")[-1])
    model = AccFlow(ofe) if kind.startswith("acc") else ofe
    model.load_state_dict(cases.weights(kind))
    return model.eval()


def _count(fn):
    with OpCount() as oc:
        out = fn()
    return out, oc.counts


@pytest.mark.parametrize("kind", ["raft", "gma"])
def test_pair_same_outputs_and_ops(kind):
    from oracle import eager_ref as er
    ref = _reference_models(kind)
    sd = cases.weights(kind)
    i1, i2, finit = cases.pair_case()
    want, ops_ref = _count(lambda: ref(i1, i2, iters=3, flow_init=finit))
    got, ops_new = _count(lambda: er.flow_estimator(sd, i1, i2, 3, finit))
    assert float((want - got).abs().max()) < 1e-5
    assert ops_ref == ops_new, {k: (ops_ref[k], ops_new[k]) for k in set(ops_ref) | set(ops_new) if ops_ref[k] != ops_new[k]}


@pytest.mark.parametrize("kind", ["acc+raft", "acc+gma"])
def test_clip_same_outputs_and_ops(kind):
    from oracle import eager_ref as er
    ref = _reference_models(kind)
    sd = cases.weights(kind)
    imgs = cases.clip_case()

    def run_ref():
        # the reference hard-codes 12 iterations per pair (AccFlow_.py:184,188)
        return ref(images=imgs, test_mode=False)

    want, ops_ref = _count(run_ref)
    got, ops_new = _count(lambda: er.accflow_forward(sd, imgs, 12))
    for a, b in zip(want, got):
        assert float((a - b).abs().max()) < 1e-5
    assert ops_ref == ops_new, {k: (ops_ref[k], ops_new[k]) for k in set(ops_ref) | set(ops_new) if ops_ref[k] != ops_new[k]}
