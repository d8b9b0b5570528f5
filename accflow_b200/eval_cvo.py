"""Evaluation driver: the caller of the hot path, same command line and printed result as the reference's
``test_cvo.py`` (:106-166), on the CUDA engine.

    python -m accflow_b200.eval_cvo -d clean -acc acc -ofe raft --acc_ckpt ckpt.pth
    torchrun --nproc-per-node 8 -m accflow_b200.eval_cvo -d clean -acc acc -ofe gma --acc_ckpt ckpt.pth

Differences from the reference script, all additive:
* one process per GPU (``torchrun``): batches are dealt round-robin to the ranks, the per-clip EPE triplets are
  gathered with ONE collective at the end (the reference wraps the model in ``nn.DataParallel`` and re-broadcasts
  47 MB of parameters on every forward, test_cvo.py:18,26);
* occlusion mask + EPE all / occ / vis are one fused kernel on the device (``metrics.clip_epe``);
* ``--size/--clips/--precision/--warm-start`` select the synthetic data shape and the arithmetic mode.
Checkpoints may carry the ``module.`` prefix the reference's DataParallel-wrapped training writes (train_acc.py:109).
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import Optional, Sequence

import torch

END = 6            # CVO-6: frames 0..6 (test_cvo.py:116)
BATCH = 10         # test_cvo.py:114


def _strip(sd):
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def build_model(acc: str, ofe: str, acc_ckpt: Optional[str], ofe_ckpt: Optional[str], device, precision=None):
    """build_acc_model / build_ofe_model of test_cvo.py:11-29 without the DataParallel wrapper."""
    from .networks import build_flow_estimator
    from .networks.AccFlow_ import AccFlow
    name = acc + "|" + ofe
    est = build_flow_estimator(name)
    if "acc" in name:
        model = AccFlow(est)
        ckpt = acc_ckpt
    else:
        model, ckpt = est, ofe_ckpt
    if ckpt is None:
        raise SystemExit("a checkpoint is required (--acc_ckpt for -acc acc, --ofe_ckpt for -acc direct)")
    model.load_state_dict(_strip(torch.load(ckpt, map_location="cpu")))
    model = model.to(device).eval()
    if precision:
        (model.ofe if "acc" in name else model).precision = precision
    return name, model


def main(argv: Optional[Sequence[str]] = None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", "-d", type=str, choices=["clean", "final"], required=True)
    ap.add_argument("--acc", "-acc", type=str, choices=["acc", "direct"], required=True)
    ap.add_argument("--acc_ckpt", type=str, default=None)
    ap.add_argument("--ofe", "-ofe", type=str, choices=["raft", "gma"], required=True)
    ap.add_argument("--ofe_ckpt", type=str, default=None)
    ap.add_argument("--size", type=int, default=None, help="synthetic clip size (default ACCFLOW_CVO_SIZE or 512)")
    ap.add_argument("--clips", type=int, default=None, help="number of synthetic clips (default ACCFLOW_CVO_CLIPS or 20)")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16x3", "fp16x2", "bf16", "fp16"])
    ap.add_argument("--warm-start", action="store_true")
    ap.add_argument("--out-dir", default=".")
    args = ap.parse_args(argv)

    import torch.distributed as dist
    from . import metrics
    from .data import preprocess
    from .dataset import fetch_valid_dataloader

    if not torch.cuda.is_available():
        raise RuntimeError("accflow_b200.eval_cvo needs a CUDA device (sm_100a); there is no CPU path")
    torch.set_grad_enabled(False)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)

    name, model = build_model(args.acc, args.ofe, args.acc_ckpt, args.ofe_ckpt, dev, args.precision)
    if "acc" in name:
        model.warm_start = bool(args.warm_start)
    kw = {k: v for k, v in (("size", args.size), ("n_clips", args.clips)) if v is not None}
    loader, dst = fetch_valid_dataloader(keys=["fflows", "bflows"], split=args.data, batch=args.batch, **kw)

    mine, ids = [], []
    first = 0
    for index, data in enumerate(loader):
        nb = int(data["imgs"].shape[0])
        if index % world == rank:                                    # batches dealt round-robin to the ranks
            data = preprocess(data, dev)
            imgs = data["imgs"][: END + 1]
            bflows, fflows = data["bflows"][: END - 1], data["fflows"][: END - 1]
            if "acc" in name:
                fn0 = model(images=imgs, test_mode=False)[-1]
            else:
                fn0 = model(imgs[-1], imgs[0])
            mine.append(metrics.clip_epe(fn0, bflows[-1], fflows[-1]))          # (nb, 3): all, occ, vis
            ids.extend(range(first, first + nb))
        first += nb
    n_clips = first
    local_tab = torch.cat(mine) if mine else torch.empty(0, 3, device=dev)
    if world > 1:
        # one collective: every rank contributes a (n_clips, 3) table that is NaN except for its own clips
        full = torch.full((n_clips, 3), float("nan"), device=dev)
        if ids:
            full[torch.tensor(ids, device=dev)] = local_tab
        parts = [torch.empty_like(full) for _ in range(world)]
        dist.all_gather(parts, full)
        stacked = torch.stack(parts)
        table = torch.where(torch.isnan(stacked), torch.zeros((), device=dev), stacked).sum(0)
    else:
        table = local_tab
    avg_all, avg_occ, avg_vis = (float(table[:, j].mean()) for j in range(3))
    if rank == 0:
        print("Finish".center(50, "="))
        print("AVG EPE %s: " % name)
        print("all:%.4f vis:%.4f occ:%.4f" % (avg_all, avg_vis, avg_occ))
        with open(os.path.join(args.out_dir, "test_result_%s_E%d.txt" % (args.data, END)), "a+") as f:
            f.write("AVG EPE %s: \n" % name)
            f.write("all:%.4f vis:%.4f occ:%.4f \n\n" % (avg_all, avg_vis, avg_occ))
    if world > 1:
        dist.barrier()
    return {"all": avg_all, "vis": avg_vis, "occ": avg_occ, "per_clip": table.cpu()}


if __name__ == "__main__":
    main(sys.argv[1:])
