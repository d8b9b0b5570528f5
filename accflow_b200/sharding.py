"""Clip-parallel sharding across the GPUs of one box.

Clips are independent, so the pair path needs no collective: rank r evaluates its own clips
with its own weight replica.  The only exchange is the metric gather at the end — three
floats per clip (test_cvo.py:151-159) — done with one ``all_gather_into_tensor``.  This
replaces the reference's ``nn.DataParallel`` (test_cvo.py:18,26), which re-broadcasts ~47 MB
of parameters on every forward.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_clip_ids(n_clips: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: rank r evaluates clips r, r+world, r+2*world, ..."""
    return list(range(rank, n_clips, world))


def gather_clip_metrics(local: torch.Tensor, n_clips: int, rank: int, world: int) -> torch.Tensor:
    """local: (len(shard), 3) per-clip (epe_all, epe_occ, epe_vis) of this rank's shard.
    Returns the (n_clips, 3) table in clip-id order on every rank."""
    if world == 1:
        return local
    per_rank = (n_clips + world - 1) // world
    padded = torch.full((per_rank, local.shape[1]), float("nan"), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty(world * per_rank, local.shape[1], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    table = torch.empty(n_clips, local.shape[1], dtype=local.dtype, device=local.device)
    for r in range(world):
        ids = shard_clip_ids(n_clips, r, world)
        table[ids] = out[r * per_rank: r * per_rank + len(ids)]
    return table
