"""Tensor helpers with the reference's names (networks/utils.py:66-124), backed by the kernels."""
from ..ops import backwarp, coords_grid  # noqa: F401


def upflow8(flow, mode="bilinear"):
    raise NotImplementedError("upflow8 is only used by the small RAFT model, which is out of scope")
