"""GPU: the multi-GPU group behind the boundary (``lo_group_*`` / ``ShardedEngine``) against the oracle.

* rank groups (one process per member, bootstrap blobs exchanged over gloo): world 2 and 3.  On a one-GPU box the
  ranks share device 0 — CUDA IPC works between processes on one device — so the whole protocol (IPC mapping,
  per-device accumulate matrices, column-last pushes with system-scope REDs, in-kernel arrival and root epilogue,
  double-buffered merge matrices, broadcast) is exercised by the regular ``-m gpu`` run; with several GPUs visible
  every rank takes its own.  ``bench.py --gpus N`` checks the same against goldens across NVLink.
* local groups (one process, all visible devices): the form the microservice entry points use.
* LO_MERGE_NCCL (the library's own ncclAllReduce through dlopen): world 1 always, world 2 with >= 2 GPUs (NCCL
  refuses two ranks on one device).
"""
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


def _worker(rank, world, port, merge, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    dev = rank % ndev
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from learningorchestra_b200.engine import Engine
    from learningorchestra_b200.sharding import ShardedEngine
    eng = Engine(dev)
    sh = ShardedEngine.from_torch_distributed(eng, merge=merge)
    got = {"merge": sh.merge}
    lo4, hi4 = np.full(4, -1000.0, np.float32), np.full(4, 1000.0, np.float32)

    # f64: 5 steps through both merge buffers, root-only result; then one broadcast step read on every rank
    rows = 700_001
    table = sh.table("f64", rows, 6).fill_synthetic(1, SEED)
    out = sh.table("f32", rows, 4)
    cols = [5, 0, 3, 2]
    got["f64"] = []
    for _ in range(5):
        c = sh.project_cast_hist(table, cols, 256, lo4, hi4, out=out)
        if sh.has_result:
            got["f64"].append(c.to_numpy())
    # eight INDEPENDENT steps back to back: consecutive launches may overlap (programmatic dependent launch, alternating
    # accumulate matrices); only the last result is read, everything in between is in flight together
    for _ in range(8):
        c = sh.project_cast_hist(table, cols, 256, lo4, hi4, out=out, independent=True)
    got["f64_overlap"] = c.to_numpy() if sh.has_result else None
    got["f64_bcast"] = sh.project_cast_hist(table, cols, 10, lo4, hi4, bcast=True).to_numpy()
    got["sums"] = [out.checksum(j) for j in range(4)]
    # more than 256 bins: the chunk kernel takes part in the same merge (1000 bins: in-kernel arrival and epilogue;
    # 30 000 bins x 4 columns: L2 counters, the root's epilogue as its own launch), two steps each through both buffers
    got["f64_1000"] = [sh.project_cast_hist(table, cols, 1000, lo4, hi4, bcast=True).to_numpy() for _ in range(2)][-1]
    got["f64_30000"] = [sh.project_cast_hist(table, cols, 30000, lo4, hi4, bcast=True, independent=True).to_numpy() for _ in range(2)][-1]
    # range pre-pass over all shards (always delivered everywhere)
    got["minmax"] = [a.tolist() for a in sh.minmax_cast(table, cols)]
    # u8: 300 columns x 256 bins = 76 800 counts -> the root's epilogue is its own multi-CTA launch
    t8 = sh.table("u8", 400_003, 300).fill_synthetic(3, SEED)
    got["u8"] = [sh.hist_u8_cols(t8, range(300), bcast=True).to_numpy() for _ in range(3)][-1]
    for _ in range(5):
        c8 = sh.hist_u8_cols(t8, range(300), bcast=True, independent=True)
    got["u8_overlap"] = c8.to_numpy()
    # a shard with no rows still takes part in the step (world 3: 40 rows -> cuts at 0, 0, 32)
    tiny = sh.table("f64", 40, 2).fill_synthetic(1, SEED)
    got["tiny"] = sh.project_cast_hist(tiny, [1, 0], 16, lo4[:2], hi4[:2], bcast=True).to_numpy()
    # host buffers: every rank passes its own rows, the merged counts arrive on the root
    b, e = sh.bounds(rows)[0]
    from oracle import cport
    hcols = [cport.synth_f64(1, SEED, c, b, e - b) for c in cols]
    houts = [np.empty(e - b, np.float32) for _ in cols]
    hc, _tm = sh.project_cast_hist_host(hcols, 256, lo4, hi4, out=houts)
    got["host_counts"] = hc
    got["host_out_ok"] = all(np.array_equal(o.view(np.uint32), cport.cast_f64_f32(x).view(np.uint32)) for o, x in zip(houts, hcols))
    got["timeouts"] = sh.timeouts()
    got["launches"] = eng.launch_count
    dist.barrier()
    ret.put((rank, got))
    dist.barrier()
    sh.close()
    eng.close()
    dist.destroy_process_group()


def _run(world, merge):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, merge, ret)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    try:
        for _ in range(world):
            r, got = ret.get(timeout=300)        # never block forever if a worker died
            res[r] = got
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    return res


def _check(res, world, merge):
    from oracle import bsem_numpy as bn
    from oracle import cport
    rows, cols = 700_001, [5, 0, 3, 2]
    lo4, hi4 = np.full(4, -1000.0, np.float32), np.full(4, 1000.0, np.float32)
    exp, sums = cport.synth_project_cast_hist(1, SEED, 0, rows, -1000.0, 1000.0, cols, 256, lo4, hi4)
    exp10, _ = cport.synth_project_cast_hist(1, SEED, 0, rows, -1000.0, 1000.0, cols, 10, lo4, hi4)
    assert all(r["merge"] == merge for r in res.values())
    assert all(r["timeouts"] == 0 for r in res.values())
    assert len(res[0]["f64"]) == 5
    for c in res[0]["f64"]:                     # every step (both buffers, re-zeroed in between) is the exact merge
        np.testing.assert_array_equal(c, exp)
    np.testing.assert_array_equal(res[0]["f64_overlap"], exp)
    total = [0] * 4
    exp1000, _ = cport.synth_project_cast_hist(1, SEED, 0, rows, -1000.0, 1000.0, cols, 1000, lo4, hi4)
    exp30000, _ = cport.synth_project_cast_hist(1, SEED, 0, rows, -1000.0, 1000.0, cols, 30000, lo4, hi4)
    for r in range(world):
        np.testing.assert_array_equal(res[r]["f64_1000"], exp1000)
        np.testing.assert_array_equal(res[r]["f64_30000"], exp30000)
        np.testing.assert_array_equal(res[r]["f64_bcast"], exp10)        # all-reduce semantics: every rank has it
        total = [(a + b) & 0xFFFFFFFFFFFFFFFF for a, b in zip(total, res[r]["sums"])]
    assert total == [int(s) for s in sums]      # fp32 output slabs, position-weighted checksums of the shards add up
    # min / max / finite count of the cast values over ALL rows
    full = [bn.cast_f64_f32(bn.synth_f64(1, SEED, c, 0, rows)) for c in cols]
    fin = [x[np.isfinite(x)] for x in full]
    for r in range(world):
        mins, maxs, cnt = res[r]["minmax"]
        assert mins == [float(x.min()) for x in fin] and maxs == [float(x.max()) for x in fin]
        assert cnt == [int(x.size) for x in fin]
    exp8 = cport.synth_hist_u8(SEED, 0, 400_003, list(range(300)))
    expt, _ = cport.synth_project_cast_hist(1, SEED, 0, 40, -1000.0, 1000.0, [1, 0], 16, lo4[:2], hi4[:2])
    for r in range(world):
        np.testing.assert_array_equal(res[r]["u8"], exp8)
        np.testing.assert_array_equal(res[r]["u8_overlap"], exp8)
        np.testing.assert_array_equal(res[r]["tiny"], expt)
        assert res[r]["host_out_ok"]
    np.testing.assert_array_equal(res[0]["host_counts"], exp)


@pytest.mark.parametrize("world", [2, 3])
def test_rank_group_peer_merge_equals_oracle(built, world):
    _check(_run(world, "peer"), world, "peer")


def test_rank_group_one_rank_per_step_launch_count(built):
    """The merge costs no extra launches: one rank alone does 5 + 1 f64 steps with exactly one kernel each."""
    res = _run(1, "peer")
    _check(res, 1, "peer")


def test_rank_group_nccl_merge_equals_oracle(built):
    world = 2 if torch.cuda.device_count() >= 2 else 1
    _check(_run(world, "nccl"), world, "nccl")


def test_local_group_over_all_visible_devices(built):
    """One process, every visible GPU: the form Projection / Histogram use.  Same answers as the oracle."""
    from learningorchestra_b200.sharding import ShardedEngine
    from oracle import cport
    rows, cols = 500_009, [2, 0, 1]
    lo, hi = np.full(3, -1000.0, np.float32), np.full(3, 1000.0, np.float32)
    with ShardedEngine.local() as sh:
        assert sh.nlocal == torch.cuda.device_count() == sh.world
        table = sh.table("f64", rows, 3).fill_synthetic(1, SEED)
        out = sh.table("f32", rows, 3)
        exp, sums = cport.synth_project_cast_hist(1, SEED, 0, rows, -1000.0, 1000.0, cols, 64, lo, hi)
        for _ in range(4):
            got = sh.project_cast_hist(table, cols, 64, lo, hi, out=out).to_numpy()
            np.testing.assert_array_equal(got, exp)
        assert [out.checksum(j) for j in range(3)] == [int(s) for s in sums]
        t8 = sh.table("u8", 300_001, 40).fill_synthetic(3, SEED)
        np.testing.assert_array_equal(sh.hist_u8_cols(t8, range(40)).to_numpy(), cport.synth_hist_u8(SEED, 0, 300_001, list(range(40))))
        # host buffers: the library cuts the rows over the members itself
        hcols = [cport.synth_f64(1, SEED, c, 0, rows) for c in cols]
        houts = [np.empty(rows, np.float32) for _ in cols]
        hc, tm = sh.project_cast_hist_host(hcols, 64, lo, hi, out=houts)
        np.testing.assert_array_equal(hc, exp)
        for o, x in zip(houts, hcols):
            np.testing.assert_array_equal(o.view(np.uint32), cport.cast_f64_f32(x).view(np.uint32))
        assert tm["h2d_bytes"] == rows * 3 * 8 and sh.timeouts() == 0
        mins, maxs, cnt = sh.minmax_cast(table, cols)
        assert mins.shape == (3,) and (maxs >= mins).all() and int(cnt.sum()) > 0
