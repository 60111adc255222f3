"""Image sharding across the GPUs of one box.

Every image's solve is independent (SR_single_class.py:83-127 is a sequential loop), so the path
shards with no collective: rank r owns a contiguous block of images.  The only exchange is the final
gather of the thresholded masks to rank 0 (NCCL for CUDA tensors, gloo in the CPU tests).
torch.distributed is plumbing only.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int):
    """Contiguous blocks of ceil(n/world) items; trailing ranks may be short or empty."""
    per = -(-n_items // world_size)
    start = min(rank * per, n_items)
    return start, min(start + per, n_items)


def gather_masks(local: torch.Tensor, n_total: int, group=None):
    """All ranks pass their [n_local, ...] block (same trailing shape / dtype); rank 0 gets
    [n_total, ...] in image order, other ranks get None."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local[:n_total]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-n_total // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat(bufs, dim=0)[:n_total]
