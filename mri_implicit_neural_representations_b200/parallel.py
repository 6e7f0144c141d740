"""Multi-GPU host logic (one process per GPU, torch.distributed).

The path shards in two ways:
  * independent fits (different slices / samples): no collective at all;
  * coordinate data-parallel inside one fit: replicated weights, every rank takes its own rows of each global
    grid-order batch, ONE all-reduce(mean) of the flat fp32 gradient buffer per step, then the same fused Adam on
    every rank (identical inputs -> identical replicas, bit for bit).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rows(global_start: int, global_bs: int, n_rows: int, rank: int, world: int):
    """Rows of the global grid-order batch [global_start, global_start+global_bs) (clipped to n_rows) owned by
    `rank`: contiguous, balanced to within one row.  Returns (start, count)."""
    end = min(global_start + global_bs, n_rows)
    n = max(end - global_start, 0)
    base, extra = divmod(n, world)
    count = base + (1 if rank < extra else 0)
    start = global_start + rank * base + min(rank, extra)
    return start, count


def allreduce_mean_(flat: torch.Tensor, weight: float = 1.0, group=None) -> torch.Tensor:
    """In-place weighted mean over ranks of a flat gradient buffer.  `weight` = this rank's share of the global
    batch (rows_local / rows_global * world) so that ragged shards still give the exact global-mean gradient."""
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    if weight != 1.0:
        flat.mul_(weight)
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    return flat
