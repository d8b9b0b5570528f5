"""world_size-2 gloo test (CPU) of the clip sharding + metric gather used for N > 1 GPUs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from accflow_b200.sharding import gather_clip_metrics, shard_clip_ids


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = shard_clip_ids(n_clips, rank, world)
    # per-clip "metrics" are a deterministic function of the clip id
    local = torch.tensor([[i + 0.25, 10.0 * i, -float(i)] for i in ids], dtype=torch.float32).reshape(len(ids), 3)
    table = gather_clip_metrics(local, n_clips, rank, world)
    q.put((rank, ids, table))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_is_a_partition():
    for n, w in ((64, 8), (7, 2), (5, 4), (3, 8)):
        shards = [shard_clip_ids(n, r, w) for r in range(w)]
        assert sorted(i for s in shards for i in s) == list(range(n))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def test_two_rank_metric_gather():
    world, n_clips = 2, 7                       # ragged: rank 0 owns 4 clips, rank 1 owns 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]   # spawn re-imports torch: slow on a loaded box
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    expect = torch.tensor([[i + 0.25, 10.0 * i, -float(i)] for i in range(n_clips)])
    for rank, ids, table in results:
        assert ids == list(range(rank, n_clips, world))
        assert torch.equal(table, expect)
