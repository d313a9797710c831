"""CPU, world_size 2 over gloo: the N>1 plumbing bench.py relies on (scene sharding, max-over-ranks timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sparse2dense_b200 import sharding


def test_scene_partition_is_disjoint_and_complete():
    for gb in (1, 4, 7, 32):
        for ws in (1, 2, 3, 8):
            parts = [sharding.scenes_of_rank(gb, r, ws) for r in range(ws)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(gb))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert sharding.scene_seeds(1, 4, 0) == [1000, 1001, 1002, 1003]
    assert sharding.scene_seeds(1, 4, 1) == [1004, 1005, 1006, 1007]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ms = sharding.max_over_ranks([10.0 + rank, 5.0 - rank])           # per-rank event times -> max
        n = sharding.sum_over_ranks([len(sharding.scenes_of_rank(9, rank, world))])
        dist.barrier()
        out.put((rank, ms, n))
    finally:
        dist.destroy_process_group()


def test_max_over_ranks_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, n in res:
        assert ms == [11.0, 5.0]          # max over ranks of (10+r, 5-r)
        assert n == [9.0]                 # all scenes accounted for exactly once
