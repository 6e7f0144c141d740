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


class PeerGradExchange:
    """Symmetric (peer-mapped) gradient buffers for the optimiser-fused data-parallel exchange (`inr_adam_step_peers`).

    Every rank allocates 2 x [n_params] fp32 (double-buffered by step parity) plus a flag array in
    torch.distributed symmetric memory; after the rendezvous each rank holds device pointers to all peers' copies.
    `grads(parity)` is the local buffer `inr_grad_step` writes; `ChainEngine.adam_step_peers(ex)` runs the optimiser
    kernel that waits for all ranks' flags, sums the peers' gradients over NVLink and applies Adam -- there is no
    separate all-reduce kernel.  Raises if symmetric memory is unavailable (callers fall back to NCCL)."""

    FLAG_WORDS = 64

    def __init__(self, n_params: int, device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("the fused exchange supports at most 8 ranks (one NVSwitch domain)")
        self.n = (int(n_params) + 3) // 4 * 4
        total = 2 * self.n + self.FLAG_WORDS
        self.buf = symm_mem.empty(total, dtype=torch.float32, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm_mem.rendezvous(self.buf, group)
        dist.barrier(group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self._grad_ptrs = [(C.c_void_p * self.world)(*[p + 4 * par * self.n for p in ptrs]) for par in (0, 1)]
        self._flag_ptrs = (C.c_void_p * self.world)(*[p + 8 * self.n for p in ptrs])

    # The exchange has ONE barrier per step (entry of the optimiser kernel) and no exit barrier: a fast rank's next
    # backward pass may start while a slow rank still gathers.  That is safe only because consecutive steps use
    # different buffers -- the parity MUST alternate with every optimiser step (parity = step & 1), which is why it is an
    # explicit argument everywhere and never a default.

    def grads(self, parity: int) -> torch.Tensor:
        """This rank's gradient buffer for the optimiser step of the given parity (step & 1)."""
        return self.buf[(parity & 1) * self.n:((parity & 1) + 1) * self.n]

    def pointers(self, parity: int):
        return self._grad_ptrs[parity & 1], self._flag_ptrs


# ---------------------------------------------------------------------------------------------------------------
# Ring models, one per GPU (reference src/train_variations/train_clustering.py:57-59,
# 173-189, 209-225): `no_models` independent networks, model i is fitted to the k-space samples whose distance to the
# centre lies in [radii[i], radii[i+1]].  The rings share nothing while training -> ring index -> rank, no data-path
# collective; validation assembles the slice from the per-ring predictions with ONE all-reduce(sum) of [N, out] rows.

def owned_rings(no_models: int, rank: int, world: int):
    """Rings fitted by `rank`: round-robin, so 4 rings on 4 (or 8) GPUs = one ring model per GPU, on 2 GPUs two each,
    on one GPU all of them in turn."""
    return [i for i in range(int(no_models)) if i % max(int(world), 1) == int(rank)]


def ring_jitter(rng, radii, no_models: int):
    """The training-time ring limits of one batch (reference :175-176): every ring is widened by |N(0, 0.05)| on both
    sides, inner radius clipped at 0.  Draws for ALL rings in ring order -- every rank consumes the same stream, so
    the limits of ring i do not depend on how the rings are spread over GPUs."""
    out = []
    for i in range(int(no_models)):
        r0 = max(0.0, float(radii[i]) - abs(float(rng.normal(0, 0.05))))
        r1 = float(radii[i + 1]) + abs(float(rng.normal(0, 0.05)))
        out.append((r0, r1))
    return out


def ring_writer_masks(dist: torch.Tensor, radii, no_models: int):
    """Validation (reference :218-232): ring i predicts the rows with radii[i] <= dist <= radii[i+1] and writes them
    into the batch in ring order, so a row on a shared edge ends up with the OUTER ring's value.  Returns, per ring,
    the bool mask of rows whose final value comes from that ring (disjoint; rows in no ring stay zero)."""
    inside = [(dist >= float(radii[i])) & (dist <= float(radii[i + 1])) for i in range(int(no_models))]
    masks = []
    later = torch.zeros_like(inside[0])
    for i in reversed(range(int(no_models))):
        masks.append(inside[i] & ~later)
        later = later | inside[i]
    return masks[::-1]


def combine_ring_outputs(partial: torch.Tensor, group=None) -> torch.Tensor:
    """Sum over ranks of the per-rank partial reconstructions (each row non-zero on exactly one rank)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial
