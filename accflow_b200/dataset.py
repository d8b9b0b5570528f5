"""Evaluation data loader with the surface of the reference's ``data/dataset.py`` (``fetch_valid_dataloader``,
:146-161; ``CVO.__getitem__`` record layout, :72-108) over synthetic CVO-shaped clips.

The CVO LMDB (``data/dataset.py:23-69``: pyarrow-serialised uint8 frames, uint16 fixed-point flows decoded as
``(v - 2**15) / 128``) is not available offline and needs ``lmdb`` + the removed ``pyarrow.deserialize``; this
module yields records of exactly the same keys / shapes / value ranges from ``accflow_b200.data.make_clip`` so that
the evaluation driver (``accflow_b200/eval_cvo.py``) and the reference's own ``test_cvo.py`` run unchanged on it:

    record = {"imgs": (21, H, W) float32 in [0, 255] (7 RGB frames, channel-concatenated),
              "fflows": (10, H, W) float32 (F(0->2) .. F(0->6)), "bflows": (10, H, W) (F(2->0) .. F(6->0))}
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader, Dataset

from .data import decode_cvo_flow_u16, encode_cvo_flow_u16, make_clip

ALL_KEYS = ["fflows", "bflows"]        # data/dataset.py:73 also lists delta_* keys, never requested by test_cvo.py


class SyntheticCVO(Dataset):
    """``CVO(keys, split, is_training=False)`` of the reference (data/dataset.py:72-108) on synthetic clips."""

    def __init__(self, keys: Sequence[str] = None, split: str = "clean", n_clips: int = None, size: int = None,
                 seed: int = 1234):
        keys = list(ALL_KEYS) if keys is None else [k.lower() for k in keys]
        for k in keys:
            assert k in ALL_KEYS, f"Invalid key value: {k}"
        assert split in ("clean", "final"), split
        self.keys, self.split, self.seed = keys, split, seed
        self.n_clips = int(os.environ.get("ACCFLOW_CVO_CLIPS", "20")) if n_clips is None else n_clips
        self.size = int(os.environ.get("ACCFLOW_CVO_SIZE", "512")) if size is None else size

    def __len__(self):
        return self.n_clips

    def __getitem__(self, index: int):
        clip = make_clip(index, size=self.size, frames=7, seed=self.seed)
        u8 = clip["u8"].float()                                       # (7,3,H,W), integer-valued 0..255
        if self.split == "final":                                     # the "final" pass of CVO is the blurred rendering
            u8 = F.avg_pool2d(F.pad(u8, (1, 1, 1, 1), mode="replicate"), 3, stride=1).round()
        out = {"imgs": u8.reshape(21, self.size, self.size).contiguous()}
        for k in self.keys:
            flows = torch.cat([f[0] for f in clip[k]], 0)             # (10,H,W)
            # through the on-disk representation (uint16 fixed point, data/dataset.py:65-67), as a real record would be
            out[k] = decode_cvo_flow_u16(encode_cvo_flow_u16(flows))
        return out


def fetch_valid_dataloader(keys: List[str], split: str = "clean", batch: int = 1, **synthetic):
    """Same call and return shape as data/dataset.py:146-161: (DataLoader, dataset), no shuffling, no drop_last."""
    if "+" in split:
        dataset = SyntheticCVO(keys, "clean", **synthetic) + SyntheticCVO(keys, "final", **synthetic)
    else:
        dataset = SyntheticCVO(keys, split, **synthetic)
    loader = DataLoader(dataset, batch_size=batch, pin_memory=torch.cuda.is_available(), shuffle=False, num_workers=0,
                        drop_last=False)
    return loader, dataset
