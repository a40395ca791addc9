"""Multi-GPU: the batch shards embarrassingly, one process per GPU, no data-path collective
inside the solve; one gather of solutions / gradients at the end.

Instances never interact -- the reference itself maps them over a host ThreadPool inside diffcp
(call site ``src/cvxpylayers/interfaces/diffcp_if.py:365``) -- so rank r owns the contiguous batch
range ``[r*B/G, (r+1)*B/G)`` and runs its own engine handle on the replicated structure.
Works with the ``nccl`` backend on GPUs and ``gloo`` on CPU (tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition of ``range(B)``; the first ``B % world`` ranks get one extra."""
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(B: int, world: int) -> list[int]:
    return [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]


def gather_rows(local: torch.Tensor, B: int, dst: int | None = 0, group=None) -> torch.Tensor | None:
    """Gather row-sharded ``local[B_r, ...]`` into the full ``[B, ...]`` tensor.

    ``dst=None`` -> all-gather (every rank gets the result); otherwise only ``dst`` does.
    Uneven shards are padded to the largest shard for the collective and trimmed afterwards.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(B, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {sizes[rank]}")
    mx = max(sizes)
    pad = local
    if local.shape[0] != mx:
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    pad = pad.contiguous()
    if dst is None:
        out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, pad, group=group)
        parts = [out[r * mx: r * mx + sizes[r]] for r in range(world)]
        return torch.cat(parts, dim=0) if any(s != mx for s in sizes) else out
    if rank == dst:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.gather(pad, bufs, dst=dst, group=group)
        return torch.cat([bufs[r][: sizes[r]] for r in range(world)], dim=0)
    dist.gather(pad, None, dst=dst, group=group)
    return None
