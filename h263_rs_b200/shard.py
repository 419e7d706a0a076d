"""Sharding of independent streams over the GPUs of one node.

Streams are fully independent (one H263State per stream, no shared state), so the path shards
by stream with NO data-path collective: rank r owns the streams s with s % world == r, its own
context, pinned staging and CUDA stream.  torch.distributed is used only to line the ranks up
for timing (barrier) and to combine per-rank measurements (max of times, sum of units) -- with
the nccl backend on GPUs, with gloo in the CPU tests.
"""
import numpy as np


def shard_streams(n_streams, world, rank):
    """Global stream ids owned by `rank`: s % world == rank (round-robin keeps the per-rank
    mix of stream types even when stream properties vary with the id)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    return np.arange(rank, n_streams, world, dtype=np.int64)


def owner_of(stream, world):
    return int(stream) % world


def local_slot(stream, world):
    """Slot of a global stream id inside its owner's context."""
    return int(stream) // world


def reduce_max(dist, value, device="cpu"):
    """Max over ranks of a scalar (step time: the job is as slow as its slowest rank)."""
    if dist is None:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(dist, value, device="cpu"):
    """Sum over ranks of a scalar (units processed: whole-job throughput = sum / max time)."""
    if dist is None:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_to_rank0(dist, array, world, rank):
    """Concatenates equally-shaped per-rank uint64 arrays on rank 0 in rank order (used to put
    per-stream checksums of a sharded run back into global stream order for verification)."""
    array = np.ascontiguousarray(array, dtype=np.uint64)
    if dist is None:
        return [array]
    import torch

    mine = torch.from_numpy(array.view(np.int64).copy())
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    if rank != 0:
        return None
    return [p.numpy().view(np.uint64) for p in parts]


def interleave_shards(parts, n_streams):
    """Inverse of shard_streams on rank 0: parts[r][k] belongs to global stream r + k * world."""
    world = len(parts)
    first = parts[0]
    out = np.zeros((n_streams,) + first.shape[1:], dtype=first.dtype)
    for r, p in enumerate(parts):
        ids = shard_streams(n_streams, world, r)
        out[ids] = p[: len(ids)]
    return out
