"""Tensor helpers with the reference's names (networks/utils.py:66-124), backed by the kernels."""
from ..ops import backwarp, coords_grid, upflow8  # noqa: F401
