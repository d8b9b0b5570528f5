"""RAFT-GMA — same constructor / forward / state_dict as networks/gma/gma.py:14-125."""
from .._estimator import FlowEstimatorBase


class RAFTGMA(FlowEstimatorBase):
    _GMA = True

    def __init__(self, args):
        if getattr(args, "position_only", False) or getattr(args, "position_and_content", False):
            raise NotImplementedError("positional attention is disabled in build_flow_estimator (networks/__init__.py:14-19)")
        if getattr(args, "num_heads", 1) != 1:
            raise NotImplementedError("num_heads is fixed to 1 on this path")
        super().__init__(args)
