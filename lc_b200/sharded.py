"""Batch sharding of the LC operators across the GPUs of one box.

Every pose is independent (the robust statistics of ``Loss_cov_mixed`` are per sample over N,
``cov_mixed.py:30-36``), so the batch is cut into contiguous slices, one per rank, with NO data-path
collective.  The only exchange is the scalar the callers take the mean of (``losses.py:334,386``):
one 16-byte ``all_reduce(SUM)`` of (sum of losses, count) over NCCL/NVLink.  Gradients need no
exchange: each rank scales its own slice by ``1 / B_global``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_bounds(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [start, end) of a batch of B poses owned by `rank`."""
    base, rem = divmod(B, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def global_mean(local_loss: Tensor, group: Optional[dist.ProcessGroup] = None) -> Tensor:
    """Mean over the GLOBAL batch of per-pose losses held shard-wise: one scalar all-reduce."""
    acc = torch.stack((local_loss.sum().to(torch.float64), local_loss.new_full((), float(local_loss.numel()), dtype=torch.float64)))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return (acc[0] / acc[1]).to(local_loss.dtype)


def global_mean_from_sums(loss_sum: Tensor, group: Optional[dist.ProcessGroup] = None) -> Tensor:
    """`loss_sum` = the (2,) fp64 [sum of losses, pose count] buffer the kernels accumulate into (lc_args.loss_sum).
    One in-place all-reduce; returns the global mean as a 0-dim tensor (no host synchronisation)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM, group=group)
    return loss_sum[0] / loss_sum[1]


def sharded_mean_loss(K, pose, pts3d, pts2d, inv_std, valid, bbox_3d, *, global_batch: int,
                      group: Optional[dist.ProcessGroup] = None, **kwargs):
    """Mean LC loss over the global batch + this rank's input gradients of that mean, one launch + one
    scalar all-reduce.  The tensors passed in are THIS RANK'S shard.  Returns (mean_loss, grads dict)."""
    from .cov_mixed import loss_fwd_bwd
    acc = torch.zeros(2, dtype=torch.float64, device=pts3d.device)
    out = loss_fwd_bwd(K, pose, pts3d, pts2d, inv_std, valid, bbox_3d, grad_scale=1.0 / float(global_batch),
                       loss_sum=acc, **kwargs)
    return global_mean_from_sums(acc, group).to(pts3d.dtype), out
