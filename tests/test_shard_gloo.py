"""The N > 1 path on CPU: two gloo ranks each assemble their shard (with the oracle standing in for the GPU) and merge
the counter vectors exactly as bench.py does across GPUs; the merged STAT must equal the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pandaseq_b200 as pb
from pandaseq_b200.shard import dist_merge_counters, merge_counters, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import datasets
    import oracle_lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = datasets.cfg1(n)
    a, b = shard_range(n, rank, world)
    got = oracle_lib.assemble("port", pb.make_config("simple_bayesian"), batch.slice(a, b), want_seq=False)
    merged = dist_merge_counters(torch.from_numpy(got["counters"].copy()), dist)
    gathered = [torch.zeros(pb.PB_NCOUNTERS, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(got["counters"].copy()))
    if rank == 0:
        np.savez(out_path, merged=merged.numpy(), parts=np.stack([g.numpy() for g in gathered]))
    dist.destroy_process_group()


def test_two_rank_stat_merge(built, tmp_path):
    import datasets
    import oracle_lib
    n, world = 3001, 2
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
    z = np.load(out)
    whole = oracle_lib.assemble("port", pb.make_config("simple_bayesian"), datasets.cfg1(n), want_seq=False)["counters"]
    assert np.array_equal(z["merged"], whole)
    assert np.array_equal(merge_counters(z["parts"]), whole)
    assert z["parts"][:, pb.C_COUNT].tolist() == [1500, 1501]


def test_shard_ranges_tile_the_batch():
    for n in (0, 1, 7, 1000, 10_000_019):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_merge_matches_c_helper(built):
    import ctypes as C
    rng = np.random.default_rng(3)
    a, b = rng.integers(0, 1000, pb.PB_NCOUNTERS), rng.integers(0, 1000, pb.PB_NCOUNTERS)
    want = merge_counters([a, b])
    dst = a.astype(np.int64).copy()
    src = b.astype(np.int64).copy()
    pb.lib().pb_counters_merge(dst.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p))
    assert np.array_equal(dst, want)
