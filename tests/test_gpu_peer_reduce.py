"""GPU: the peer-memory histogram merge (sharding.PeerReduce) — world size 2 and 3 as separate processes.
On a one-GPU box the ranks share device 0 (CUDA IPC works between processes on the same device), so the
protocol (IPC mapping, system-scope REDs into the root's matrix, double-buffered flags) is exercised by the
regular `-m gpu` run; `bench.py --merge p2p --gpus N` runs it across NVLink."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
SEED = 20260921


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_rows, ncols, nbins, steps, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    dev = rank % ndev
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from learningorchestra_b200.engine import Engine
    from learningorchestra_b200.sharding import PeerReduce, shard_bounds
    eng = Engine(dev)
    r0, r1 = shard_bounds(total_rows, world, rank)
    table = eng.table("f64", r1 - r0, ncols).fill_synthetic(1, SEED, row_offset=r0)
    lo = np.full(ncols, -1000.0, np.float32)
    hi = np.full(ncols, 1000.0, np.float32)
    pr = PeerReduce(eng, ncols, nbins, timeout_ms=5000)
    results = []
    for _ in range(steps):
        pr.before_kernel()
        eng.project_cast_hist(table, range(ncols), nbins, lo, hi, counts=pr.counts_for_step(), peer_counts=True)
        pr.after_kernel()
        if rank == 0:
            results.append(pr.result_numpy())
    eng.sync()
    timed_out = pr.timed_out()
    dist.barrier()
    if rank == 0:
        ret.put((results, timed_out))
    else:
        assert timed_out == 0
    pr.close()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_reduce_equals_oracle(built, world):
    from oracle import cport
    total_rows, ncols, nbins, steps = 700_001, 4, 256, 5
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_rows, ncols, nbins, steps, ret)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        results, timed_out = ret.get(timeout=240)       # never block forever if a worker died
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert timed_out == 0
    lo = np.full(ncols, -1000.0, np.float32)
    hi = np.full(ncols, 1000.0, np.float32)
    exp, _ = cport.synth_project_cast_hist(1, SEED, 0, total_rows, -1000.0, 1000.0, list(range(ncols)), nbins, lo, hi)
    assert len(results) == steps
    for r in results:                       # every step (both buffers, re-zeroed in between) gives the exact merge
        np.testing.assert_array_equal(r, exp)
