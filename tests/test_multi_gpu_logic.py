"""CPU tier: the N>1 host logic (contiguous sharding + all-gather of packed output rows) with world_size 2 over gloo.
The per-rank compute is replaced by a deterministic stand-in because the product has no CPU compute path."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from optimization_dynamics_b200.device import shard_range, all_gather_rows


def _worker(rank, world, port, B, width, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(B, rank, world)
    rows = torch.arange(lo, hi, dtype=torch.float64)[:, None] * 10.0 + torch.arange(width, dtype=torch.float64)[None, :]
    gathered = all_gather_rows(rows, B)
    expect = torch.arange(B, dtype=torch.float64)[:, None] * 10.0 + torch.arange(width, dtype=torch.float64)[None, :]
    q.put((rank, bool(torch.equal(gathered, expect))))
    dist.destroy_process_group()


def _run(B, width, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, width, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_even_shards_all_gather_into_tensor():
    _run(4096, 44, 29611)


def test_ragged_shards_are_padded_and_compacted():
    _run(4097, 44, 29612)
