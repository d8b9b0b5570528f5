"""RAFT (large model) — same constructor / forward / state_dict as networks/raft/raft.py:24-146."""
from .._estimator import FlowEstimatorBase


class RAFT(FlowEstimatorBase):
    _GMA = False

    def __init__(self, args):
        if getattr(args, "small", False):
            raise NotImplementedError("small RAFT is never constructed on the AccFlow path (SURVEY.md §2 #4)")
        if "alternate_corr" not in args:
            args.alternate_corr = False
        super().__init__(args)
