"""CPU: the N > 1 host logic — shard bounds and the single count all-reduce — over gloo, world size 2 and 3.
Each rank computes its shard's partial histogram with the ORACLE (there is no GPU here); what is under
test is the sharding arithmetic and the reduction plumbing that bench.py / the executors use on NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from learningorchestra_b200.sharding import all_shard_bounds, allreduce_counts, shard_bounds

SEED = 20260921


def test_shard_bounds_cover_and_align():
    for total in (0, 1, 31, 32, 33, 1000, 100_000_000, 12_345_677):
        for world in (1, 2, 3, 4, 8):
            b = all_shard_bounds(total, world)
            assert b[0][0] == 0 and b[-1][1] == total
            for (a0, a1), (b0, b1) in zip(b[:-1], b[1:]):
                assert a1 == b0 and a0 <= a1
            for (x0, _x1) in b[1:]:
                assert x0 % 32 == 0
    assert shard_bounds(100_000_000, 8, 3) == (37_500_000, 50_000_000)
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_rows, ncols, nbins, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cport
    r0, r1 = shard_bounds(total_rows, world, rank)
    lo = np.full(ncols, -1000.0, np.float32)
    hi = np.full(ncols, 1000.0, np.float32)
    part, _ = cport.synth_project_cast_hist(1, SEED, r0, r1 - r0, -1000.0, 1000.0, list(range(ncols)), nbins, lo, hi)
    t = torch.from_numpy(part.view(np.int64).copy())
    allreduce_counts(t)
    if rank == 0:
        ret.put(t.numpy().view(np.uint64).copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_counts_allreduce_over_gloo(built, world):
    from oracle import cport
    total_rows, ncols, nbins = 250_007, 4, 256
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_rows, ncols, nbins, ret)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        got = ret.get(timeout=240)                       # never block forever if a worker died
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    lo = np.full(ncols, -1000.0, np.float32)
    hi = np.full(ncols, 1000.0, np.float32)
    exp, _ = cport.synth_project_cast_hist(1, SEED, 0, total_rows, -1000.0, 1000.0, list(range(ncols)), nbins, lo, hi)
    np.testing.assert_array_equal(got, exp)
