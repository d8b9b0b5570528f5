"""Declarative parameter tables for the flow path.

The drop-in contract with the reference is the *state_dict*: key names, shapes and dtypes
(SURVEY.md §8b "Parameter contract").  Instead of re-declaring a class per sub-network, the
whole contract is one table of ``Entry`` rows that is used three ways:

* ``accflow_b200.networks`` builds the ``nn.Module`` tree from it (so ``state_dict()`` /
  ``load_state_dict()`` round-trip with reference checkpoints, incl. the ``module.`` prefix),
* ``accflow_b200.weights`` draws seeded, de-degenerated test weights from it,
* ``accflow_b200.engine`` walks it to repack OIHW weights into the kernels' layouts.

Reference layouts followed (names/shapes only):
  RAFT            networks/raft/raft.py:25-65, raft/extractor.py:137-199, raft/update.py:79-125
  RAFTGMA         networks/gma/gma.py:14-41, gma/modules.py:34-100, gma/update.py:112-125
  AccFlow         networks/AccFlow_.py:13-155, networks/modules.py:81-97
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List, Tuple

# roles: how the seeded factory / module builder treats an entry
CONV_W, CONV_B = "conv_w", "conv_b"
BN_W, BN_B, BN_RM, BN_RV, BN_NBT = "bn_w", "bn_b", "bn_rm", "bn_rv", "bn_nbt"
GAMMA, ZSCALE, EMB, RELIND = "gamma", "zero_scale", "emb", "rel_ind"
ZCONV_W, ZCONV_B = "zconv_w", "zconv_b"


@dataclass(frozen=True)
class Entry:
    name: str
    shape: Tuple[int, ...]
    role: str
    alias_of: str | None = None  # shared tensor (reference registers norm3 twice)
    buffer: bool = False


def _conv(name: str, cout: int, cin: int, kh: int, kw: int, bias: bool = True,
          roles=(CONV_W, CONV_B)) -> Iterator[Entry]:
    yield Entry(f"{name}.weight", (cout, cin, kh, kw), roles[0])
    if bias:
        yield Entry(f"{name}.bias", (cout,), roles[1])


def _bn(name: str, c: int, alias: str | None = None) -> Iterator[Entry]:
    for suffix, role, shape, buf in (("weight", BN_W, (c,), False), ("bias", BN_B, (c,), False),
                                     ("running_mean", BN_RM, (c,), True),
                                     ("running_var", BN_RV, (c,), True),
                                     ("num_batches_tracked", BN_NBT, (), True)):
        yield Entry(f"{name}.{suffix}", shape, role,
                    alias_of=None if alias is None else f"{alias}.{suffix}", buffer=buf)


def encoder_entries(prefix: str, out_dim: int, norm: str) -> List[Entry]:
    """7x7/s2 stem, 3 stages x 2 residual blocks (64, 96/s2, 128/s2), 1x1 head.

    ``norm``: 'instance' (no parameters), 'batch' (affine + running stats) or 'none'.
    Key order follows the reference's module registration order so that
    ``list(state_dict())`` is identical, not merely the key set.
    """
    out: List[Entry] = []
    bn = norm == "batch"
    if bn:
        out += _bn(f"{prefix}norm1", 64)
    out += _conv(f"{prefix}conv1", 64, 3, 7, 7)
    cin = 64
    for stage, (dim, stride) in enumerate(((64, 1), (96, 2), (128, 2)), start=1):
        for blk in (0, 1):
            p = f"{prefix}layer{stage}.{blk}."
            s = stride if blk == 0 else 1
            out += _conv(p + "conv1", dim, cin, 3, 3)
            out += _conv(p + "conv2", dim, dim, 3, 3)
            if bn:
                out += _bn(p + "norm1", dim)
                out += _bn(p + "norm2", dim)
                if s != 1:
                    out += _bn(p + "norm3", dim)
            if s != 1:
                out += _conv(p + "downsample.0", dim, cin, 1, 1)
                if bn:
                    out += _bn(p + "downsample.1", dim, alias=p + "norm3")
            cin = dim
    out += _conv(f"{prefix}conv2", out_dim, 128, 1, 1)
    return out


def update_block_entries(prefix: str, gma: bool) -> List[Entry]:
    out: List[Entry] = []
    e = prefix + "encoder."
    out += _conv(e + "convc1", 256, 324, 1, 1)
    out += _conv(e + "convc2", 192, 256, 3, 3)
    out += _conv(e + "convf1", 128, 2, 7, 7)
    out += _conv(e + "convf2", 64, 128, 3, 3)
    out += _conv(e + "conv", 126, 256, 3, 3)
    gin = 128 + (384 if gma else 256)
    g = prefix + "gru."
    for tag, kh, kw in (("1", 1, 5), ("2", 5, 1)):
        for gate in "zrq":
            out += _conv(f"{g}conv{gate}{tag}", 128, gin, kh, kw)
    out += _conv(prefix + "flow_head.conv1", 256, 128, 3, 3)
    out += _conv(prefix + "flow_head.conv2", 2, 256, 3, 3)
    out += _conv(prefix + "mask.0", 256, 128, 3, 3)
    out += _conv(prefix + "mask.2", 576, 256, 1, 1)
    if gma:
        out.append(Entry(prefix + "aggregator.gamma", (1,), GAMMA))
        out += _conv(prefix + "aggregator.to_v", 128, 128, 1, 1, bias=False)
    return out


def raft_entries(prefix: str = "") -> List[Entry]:
    return (encoder_entries(prefix + "fnet.", 256, "instance")
            + encoder_entries(prefix + "cnet.", 256, "batch")
            + update_block_entries(prefix + "update_block.", gma=False))


def gma_entries(prefix: str = "", max_pos: int = 160) -> List[Entry]:
    out = (encoder_entries(prefix + "fnet.", 256, "instance")
           + encoder_entries(prefix + "cnet.", 256, "batch")
           + update_block_entries(prefix + "update_block.", gma=True))
    out += _conv(prefix + "att.to_qk", 256, 128, 1, 1, bias=False)
    out.append(Entry(prefix + "att.pos_emb.rel_ind", (max_pos, max_pos), RELIND, buffer=True))
    out.append(Entry(prefix + "att.pos_emb.rel_height.weight", (2 * max_pos - 1, 128), EMB))
    out.append(Entry(prefix + "att.pos_emb.rel_width.weight", (2 * max_pos - 1, 128), EMB))
    return out


def accflow_entries(ofe: str) -> List[Entry]:
    """AccFlow(ofe): ofe.* first, then the accumulation sub-networks (AccFlow_.py:146-154)."""
    c = 128
    out = raft_entries("ofe.") if ofe == "raft" else gma_entries("ofe.")
    out += _conv("flow_encoder.conv1", c, 2, 7, 7)
    out += _conv("flow_encoder.conv2", 2 * c, c, 3, 3)
    out += _conv("flow_encoder.conv3", c, 2 * c, 1, 1)
    out += _conv("flow_decoder.flow.0", 2 * c, c, 3, 3)
    out += _conv("flow_decoder.flow.2", 2, 2 * c, 3, 3)
    out += _conv("flow_decoder.mask.0", 2 * c, c, 3, 3)
    out += _conv("flow_decoder.mask.2", 576, 2 * c, 1, 1)
    out += encoder_entries("context.", c, "none")
    out += _conv("accplus.conv1.0", 2 * c, 2 * c + 1, 3, 3)
    out += _conv("accplus.conv1.2", c, 2 * c, 3, 3)
    out += _conv("accplus.conv2.0", 2 * c, 2 * c, 3, 3)
    out += _conv("accplus.conv2.2", c, 2 * c, 3, 3)
    out.append(Entry("accplus.conv2.4.scale", (1, 27, 1, 1), ZSCALE))
    out += _conv("accplus.conv2.4.conv", 27, c, 3, 3, roles=(ZCONV_W, ZCONV_B))
    out += _conv("accplus.dconv", c, c, 3, 3)
    out += _conv("accplus.conv3.0", 2 * c, 2 * c + 1, 3, 3)
    out += _conv("accplus.conv3.2", c, 2 * c, 3, 3)
    out += _conv("accplus.conv4.0", 2 * c, 4 * c, 3, 3)
    out += _conv("accplus.conv4.2", c, 2 * c, 3, 3)
    out += _conv("accplus.conv4.4", c, c, 1, 1)
    out += _conv("blending.mask.0", 2 * c, c, 1, 1)
    out += _conv("blending.mask.2", 1, 2 * c, 3, 3)
    return out


def entries_for(kind: str) -> List[Entry]:
    """kind in {'raft','gma','acc+raft','acc+gma'}."""
    k = kind.lower()
    if k.startswith("acc"):
        return accflow_entries("gma" if "gma" in k else "raft")
    return gma_entries() if "gma" in k else raft_entries()
