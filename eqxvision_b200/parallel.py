"""Data-parallel plumbing for the one place the path shards: the batch dimension.

Every image is independent in inference (BatchNorm uses running statistics; `axis_name="batch"` has no
numeric effect), so N GPUs = N replicas of the weights, a contiguous split of the batch, and - only
when the caller wants the gathered output - ONE all-gather of the logits.
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is used as the launcher/collective; no
collective sits inside the forward pass.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous, balanced split of range(n): the first n % world ranks get one extra item"""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return None
    return dist


def shard(images, rank: Optional[int] = None, world: Optional[int] = None):
    """this rank's slice of a [B, ...] batch"""
    dist = _dist()
    if rank is None:
        rank = dist.get_rank() if dist else 0
    if world is None:
        world = dist.get_world_size() if dist else 1
    lo, hi = shard_bounds(images.shape[0], rank, world)
    return images[lo:hi]


def all_gather_rows(local: torch.Tensor, total_rows: int) -> torch.Tensor:
    """concatenate every rank's [rows_r, ...] block in rank order (uneven blocks are padded for the
    collective and trimmed afterwards). Single process: returns `local`."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(total_rows, r, world) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], 0)


def data_parallel_forward(forward: Callable[[torch.Tensor], torch.Tensor], images, gather: bool = True):
    """run `forward` (e.g. `vmap(net, axis_name="batch")` bound to its keys) on this rank's shard of
    `images`; with `gather` every rank returns the logits of the whole batch."""
    local = forward(shard(images))
    return all_gather_rows(local, images.shape[0]) if gather else local
