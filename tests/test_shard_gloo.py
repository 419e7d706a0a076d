"""N > 1 host logic on CPU: two gloo ranks shard a set of streams, parse their shards with the
product front end (host code of libh263cu.so, no GPU involved) and rank 0 checks that the
gathered per-stream side-info checksums equal an unsharded parse; plus the timing reductions
bench.py uses."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from h263_rs_b200 import shard  # noqa: E402

N_STREAMS, N_PICS = 10, 3


def stream_checksums(ids):
    """Per stream: FNV-style checksum over the side info of all its pictures."""
    from h263_rs_b200 import frontend, synth

    out = np.zeros(len(ids), np.uint64)
    for k, s in enumerate(ids):
        ps = frontend.Parser(1)
        h = np.uint64(1469598103934665603)
        for pk in synth.make_stream(176, 144, N_PICS, int(s), mv_mode=int(s) % 3):
            pic, mbs, ev = ps.parse_picture(pk)
            for arr in (mbs.view(np.uint8).reshape(-1), np.asarray(ev).view(np.uint8).reshape(-1)):
                h = np.uint64((int(h) * 1099511628211 + int(arr.astype(np.uint64).sum()) * 31 + arr.size) % (1 << 64))
        out[k] = h
    return out


def _worker(rank, world, port, ret):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = shard.shard_streams(N_STREAMS, world, rank)
        sums = stream_checksums(ids)
        # equal shapes for gather: pad to the largest shard
        pad = np.zeros((N_STREAMS + world - 1) // world, np.uint64)
        pad[: len(sums)] = sums
        parts = shard.gather_to_rank0(dist, pad, world, rank)
        t_max = shard.reduce_max(dist, 10.0 + rank)
        units = shard.reduce_sum(dist, float(len(ids)))
        dist.barrier()
        if rank == 0:
            ret["sums"] = shard.interleave_shards(parts, N_STREAMS)
            ret["t_max"], ret["units"] = t_max, units
    finally:
        dist.destroy_process_group()


def test_shard_streams_partition():
    for world in (1, 2, 3, 4, 8):
        seen = np.concatenate([shard.shard_streams(37, world, r) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(37))
        for s in range(37):
            r = shard.owner_of(s, world)
            assert shard.shard_streams(37, world, r)[shard.local_slot(s, world)] == s
    with pytest.raises(ValueError):
        shard.shard_streams(4, 2, 2)


def test_two_gloo_ranks_reproduce_the_unsharded_parse():
    import torch.multiprocessing as mp

    world, port = 2, 29533 + os.getpid() % 200
    mgr = mp.get_context("spawn").Manager()  # no fork: the parser library keeps worker threads
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    expect = stream_checksums(np.arange(N_STREAMS))
    assert np.array_equal(np.asarray(ret["sums"]), expect)
    assert ret["t_max"] == 11.0 and ret["units"] == float(N_STREAMS)


def test_plan_cpu_slices():
    """Ranks split the CPUs of their GPU's NUMA node; ranks without node information share evenly."""
    from h263_rs_b200 import shard

    cpus = {0: list(range(0, 16)), 1: list(range(16, 32))}
    plan = shard.plan_cpu_slices([0, 0, 1, 1], range(32), lambda n: cpus[n])
    assert plan == [list(range(0, 8)), list(range(8, 16)), list(range(16, 24)), list(range(24, 32))]
    # all GPUs on node 1, job restricted to a subset of the CPUs
    plan = shard.plan_cpu_slices([1, 1], [4, 5, 20, 21, 22, 23], lambda n: cpus[n])
    assert plan == [[20, 21], [22, 23]]
    # unknown placement: an even share of everything, never empty
    plan = shard.plan_cpu_slices([None, None, None], range(4), lambda n: [])
    assert all(plan) and sorted(sum(plan, [])) == [0, 1, 2, 3]
    plan = shard.plan_cpu_slices([None] * 8, range(2), lambda n: [])
    assert all(plan)
    # a node none of whose CPUs is allowed falls back to the even share
    plan = shard.plan_cpu_slices([0, 1], range(16, 32), lambda n: cpus[n])
    assert plan[1] == list(range(16, 32)) and plan[0] == list(range(16, 24))
    assert shard._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
