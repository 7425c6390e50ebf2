"""Data-parallel plumbing for the path (torch.distributed is plumbing, not product).

The minibatch shards by columns (samples are independent inside the vector field; the only coupling in the
reference is the solver's global RMS norm).  Implemented mode: independent controllers -- every
rank integrates its shard with its own step sequence and the gradients are averaged (the loss is a mean over
the global batch: logitcrossentropy agg=mean, experiments/mnist_node.jl:135).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_columns(B: int, rank: int, world: int) -> tuple[int, int]:
    """Columns [lo, hi) of a global batch of B owned by `rank` (contiguous, sizes differ by at most 1)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def average_gradients_(grads, world: int) -> None:
    """Sum over ranks (NCCL all-reduce on GPU tensors, gloo on CPU) then divide by the number of ranks."""
    if world <= 1:
        return
    for g in grads:
        if g.numel():
            dist.all_reduce(g)
            g.div_(world)


def max_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def exchange_ipc_handles(mine: bytes, world: int) -> list[bytes]:
    """All-gather the 64-byte CUDA IPC handles of the ranks' exchange buffers (rnde_dist_export ->
    rnde_dist_import).  Plumbing only: a collective on host bytes."""
    out = [None] * world
    dist.all_gather_object(out, mine)
    return out
