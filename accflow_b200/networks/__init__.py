"""Drop-in for the reference's ``networks`` package on the flow path (networks/__init__.py:4-23)."""
import argparse


def build_flow_estimator(name):
    """'raft' in name -> RAFT(small=False, mixed_precision=True); 'gma' -> RAFTGMA(num_heads=1, ...)."""
    lowered = name.lower()
    if "raft" in lowered:
        from .raft.raft import RAFT
        return RAFT(argparse.Namespace(small=False, mixed_precision=True))
    if "gma" in lowered:
        from .gma.gma import RAFTGMA
        return RAFTGMA(argparse.Namespace(num_heads=1, mixed_precision=True, position_only=False,
                                          position_and_content=False))
    raise NotImplementedError("not supported yet..")
