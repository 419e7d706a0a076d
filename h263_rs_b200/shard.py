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


def gather_to_rank0(dist, array, world, rank, device="cpu"):
    """Concatenates equally-shaped per-rank uint64 arrays on rank 0 in rank order (used to put
    per-stream checksums / parity verdicts of a sharded run back into global order for verification).
    `device` = where the collective's tensors must live ("cuda" with the nccl backend, "cpu" with gloo)."""
    array = np.ascontiguousarray(array, dtype=np.uint64)
    if dist is None:
        return [array]
    import torch

    mine = torch.from_numpy(array.view(np.int64).copy()).to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)  # every backend has all_gather; the volume is a few words
    if rank != 0:
        return None
    return [p.cpu().numpy().view(np.uint64) for p in parts]


def interleave_shards(parts, n_streams):
    """Inverse of shard_streams on rank 0: parts[r][k] belongs to global stream r + k * world."""
    world = len(parts)
    first = parts[0]
    out = np.zeros((n_streams,) + first.shape[1:], dtype=first.dtype)
    for r, p in enumerate(parts):
        ids = shard_streams(n_streams, world, r)
        out[ids] = p[: len(ids)]
    return out


# ---- host placement ---------------------------------------------------------------------------
# The path never exchanges data between GPUs, but every rank moves ~420 MB of RGBA per step into
# pinned host memory and runs its own parser threads.  On a two-socket box that traffic should stay
# on the socket the GPU hangs off: a rank whose threads and pinned buffers sit on the other socket
# pays the inter-socket link on every byte.

def _parse_cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index):
    """NUMA node of a CUDA device from sysfs, or None when the platform does not tell."""
    try:
        import torch

        p = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read())
        return node if node >= 0 else None
    except Exception:
        return None


def node_cpus(node):
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            return _parse_cpulist(f.read())
    except Exception:
        return []


def plan_cpu_slices(nodes, allowed, cpus_of_node):
    """CPU set per local rank.  `nodes[r]` = NUMA node of rank r's GPU (None = unknown), `allowed`
    = CPUs this job may use, `cpus_of_node(n)` = CPUs of node n.  Ranks on the same node split its
    allowed CPUs evenly; ranks without a node (or whose node has no allowed CPU) split what is left
    of an even share of everything.  Pure function: tested on CPU."""
    world = len(nodes)
    allowed = sorted(allowed)
    out = [None] * world
    by_node = {}
    for r, n in enumerate(nodes):
        cpus = [c for c in cpus_of_node(n) if c in set(allowed)] if n is not None else []
        if cpus:
            by_node.setdefault(n, (cpus, []))[1].append(r)
    for n, (cpus, ranks) in by_node.items():
        k = len(ranks)
        for i, r in enumerate(ranks):
            share = cpus[i * len(cpus) // k:(i + 1) * len(cpus) // k]
            out[r] = share or cpus
    for r in range(world):
        if out[r] is None:
            share = allowed[r * len(allowed) // world:(r + 1) * len(allowed) // world]
            out[r] = share or allowed
    return out


def bind_rank_to_gpu_node(local_rank, local_world):
    """Pins the calling thread (and the threads and pinned allocations it makes afterwards) to this
    rank's share of the CPUs of its GPU's NUMA node.  Returns a dict describing what was done."""
    import os

    if not hasattr(os, "sched_setaffinity"):
        return {"numa_node": None, "cpus": None}
    allowed = sorted(os.sched_getaffinity(0))
    nodes = [gpu_numa_node(r) for r in range(local_world)]
    slices = plan_cpu_slices(nodes, allowed, node_cpus)
    mine = slices[local_rank]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return {"numa_node": nodes[local_rank], "cpus": None}
    return {"numa_node": nodes[local_rank], "cpus": len(mine)}
