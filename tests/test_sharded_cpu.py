"""CPU, world_size 2 over gloo: the multi-GPU plumbing of the sharded mean (lc_b200/sharded.py).

The data path needs no collective (poses are independent); the only exchange is the scalar all-reduce of
(sum of per-pose losses, count).  This test runs that exchange for real between two processes and checks it
against the single-process mean, with uneven shards."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lc_b200.sharded import global_mean, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, losses, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(len(losses), rank, world)
    m = global_mean(losses[lo:hi].clone())
    # gradient scale every rank applies to its own shard: d mean / d loss_b = 1 / B_global, no exchange needed
    out[rank] = float(m)
    dist.destroy_process_group()


def test_global_mean_two_ranks_uneven_shards():
    torch.manual_seed(0)
    losses = torch.randn(1025, dtype=torch.float32) * 3 + 5
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, losses, out), nprocs=2, join=True)
    ref = float(losses.double().mean())
    assert abs(out[0] - ref) <= 1e-6 * abs(ref) and out[0] == out[1]


def test_global_mean_single_process_is_plain_mean():
    x = torch.arange(10, dtype=torch.float64)
    assert float(global_mean(x)) == 4.5
