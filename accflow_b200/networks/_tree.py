"""Builds the nn.Module parameter tree of a network from the declarative table in spec.py.

The modules created here only *hold* parameters (so that state_dict()/load_state_dict(),
.cuda(), .eval(), nn.DataParallel and optimisers see exactly the reference's names and
shapes); the arithmetic lives in the CUDA kernels driven by accflow_b200.engine.
"""
from __future__ import annotations

import math
from typing import Iterable

import torch
from torch import nn

from .. import spec as S


class Params(nn.Module):
    """Anonymous parameter container (one per dotted-name component)."""

    def extra_repr(self):
        return ", ".join(f"{n}{tuple(p.shape)}" for n, p in self._parameters.items())


def _walk(root: nn.Module, parts):
    node = root
    for name in parts:
        child = node._modules.get(name)
        if child is None:
            child = Params()
            node.add_module(name, child)
        node = child
    return node


def _default_init(e: S.Entry, encoder_style: bool) -> torch.Tensor:
    """Same families of initial values as the reference's constructors (kaiming fan_out for
    the encoders, raft/extractor.py:183-194; torch defaults elsewhere; zeros for ZeroConv2d,
    gamma, networks/modules.py:89-92, gma/modules.py:95)."""
    shp = e.shape
    if e.role == S.CONV_W:
        t = torch.empty(shp)
        if encoder_style:
            nn.init.kaiming_normal_(t, mode="fan_out", nonlinearity="relu")
        else:
            nn.init.kaiming_uniform_(t, a=math.sqrt(5))
        return t
    if e.role == S.CONV_B:
        fan_in = 1  # bias bound needs the matching weight's fan_in; resolved by caller
        return torch.zeros(shp)
    if e.role in (S.ZCONV_W, S.ZCONV_B, S.ZSCALE, S.GAMMA, S.BN_B, S.BN_RM):
        return torch.zeros(shp)
    if e.role in (S.BN_W, S.BN_RV):
        return torch.ones(shp)
    if e.role == S.BN_NBT:
        return torch.tensor(0, dtype=torch.long)
    if e.role == S.EMB:
        return torch.randn(shp)
    if e.role == S.RELIND:
        n = shp[0]
        return torch.arange(n).view(1, -1) - torch.arange(n).view(-1, 1) + n - 1
    raise ValueError(e.role)


def populate(root: nn.Module, entries: Iterable[S.Entry]) -> None:
    last_fan_in = 1
    for e in entries:
        *path, leaf = e.name.split(".")
        if e.alias_of is not None:
            # the reference registers one norm module under two names (norm3 / downsample.1)
            src = _walk(root, e.alias_of.split(".")[:-1])
            parent = _walk(root, path[:-1])
            if path[-1] not in parent._modules:
                parent.add_module(path[-1], src)
            continue
        node = _walk(root, path)
        enc = any(t in e.name for t in ("fnet.", "cnet.", "context."))
        val = _default_init(e, enc)
        if e.role == S.CONV_W:
            last_fan_in = e.shape[1] * e.shape[2] * e.shape[3]
        elif e.role == S.CONV_B:
            bound = 1.0 / math.sqrt(last_fan_in)
            val = torch.empty(e.shape).uniform_(-bound, bound)
        if e.buffer:
            node.register_buffer(leaf, val)
        else:
            node.register_parameter(leaf, nn.Parameter(val))


# nn.DataParallel re-creates the replicas (fresh broadcast copies of every parameter) on every forward, so a replica
# cannot validate a cached engine against its own tensors.  A replica's __dict__ is a shallow copy of its source's
# (Module._replicate_for_data_parallel), so an id stored at construction survives replication and leads back to the
# source module, whose parameter versions are the truth.
import weakref

_SOURCES: "weakref.WeakValueDictionary[int, nn.Module]" = weakref.WeakValueDictionary()


def register_source(module: nn.Module) -> None:
    module._src_id = id(module)
    _SOURCES[module._src_id] = module


def source_of(module: nn.Module) -> nn.Module:
    """The module a DataParallel replica was made from (the module itself when it is not a replica)."""
    if getattr(module, "_is_replica", False):
        src = _SOURCES.get(getattr(module, "_src_id", None))
        if src is not None:
            return src
    return module


def signature(module: nn.Module):
    """Cheap change detector for packed weights: (data_ptr, _version) of every tensor."""
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))
