"""Batch-of-latents data parallelism for the sampling path (new functionality: the reference's samplers
assert a single GPU, lumina_next_t2i/sample.py:339).

Every (prompt, noise) pair is an independent ODE solve, so the path shards with NO collective inside the
solve: one process per GPU, weights replicated, each rank integrates its own contiguous slice of the
samples, and a single all-gather returns the final latents.  Works with any
torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) slice of `total` independent samples owned by `rank`."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world: int) -> List[int]:
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def all_gather_latents(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """Gather per-rank final latents [n_local, ...] into [total, ...] in sample order (the one collective of
    the path).  Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} samples, expected {sizes[rank]}")
    cap = max(sizes)
    buf = local.new_zeros((cap,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    out = local.new_empty((world * cap,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    chunks = [out[r * cap: r * cap + sizes[r]] for r in range(world)]
    return torch.cat(chunks, dim=0)


def sample_sharded(sample_one, total: int, group=None) -> torch.Tensor:
    """Run `sample_one(global_index) -> latent tensor` for this rank's slice and gather all results."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(total, rank, world)
    outs = [sample_one(i) for i in range(lo, hi)]
    if outs:
        local = torch.stack(outs, dim=0)
    else:   # a rank may own nothing when total < world; it still joins the collective
        probe = sample_one(0)
        local = probe.new_zeros((0,) + tuple(probe.shape))
    return all_gather_latents(local, total, group)
