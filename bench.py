#!/usr/bin/env python
"""bench.py — rows/s of the fused projection + fp64->fp32 cast + 256-bin histogram hot path.

Workload (BASELINE.json configs[2]/[3], SURVEY.md §8d "S100"): synthetic columnar table,
100 000 000 rows x 32 fp64 columns, K = 32 projected columns (a fixed permutation), fp32 output
table written, 256-bin histogram of every projected column over [-1000, 1000].  With N GPUs the
rows are range-sharded (rank r owns rows [r*R/N, (r+1)*R/N)) and the per-GPU partial histograms
are merged by ONE NCCL all-reduce per step (strong scaling: the table size is fixed).

One "step" = one pass of the hot path over the (rank's shard of the) table:
    zero counts -> fused kernel -> (N > 1) all-reduce of the uint64 count matrix.

Output: ONE JSON line on rank 0 (see README / DESIGN.md §6 for the keys).

    python bench.py                       # 1 GPU, defaults
    torchrun ... bench.py --gpus 8        # one rank per GPU
    python bench.py --impl reference      # CPU arm: the oracle port on all host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 20260921
GEN_LO, GEN_HI = -1000.0, 1000.0
NBINS = 256
METRIC = "rows/sec project+cast+histogram 100M\u00d732 fp64\u2192fp32; HBM GB/s vs peak @1/2/4/8 GPU"   # BASELINE.json "metric"
try:
    METRIC = json.loads((ROOT / "BASELINE.json").read_text())["metric"]
except Exception:
    pass


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def projected_columns(ncols: int) -> list[int]:
    """K = C, a fixed non-identity permutation ('all columns, arbitrary permutation')."""
    return [(7 * j + 3) % ncols for j in range(ncols)] if ncols % 7 else list(range(ncols))[::-1]


def peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md 'clocks line').  The process is started
    before the warm-up (nvidia-smi needs ~100 ms to come up); every line is stamped on receipt and only the
    samples that fall inside [mark_start, mark_end] are summarised."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc, self.thread = device, [], None, None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_ready(self, wait_s: float = 2.0):
        """Block until nvidia-smi is actually producing lines.  Must be called BEFORE the barrier that precedes
        the timed region: only rank 0 samples, and waiting after the barrier would let the other ranks start
        their timed steps and then sit in the merge waiting for rank 0."""
        deadline = time.time() + wait_s
        while self.proc is not None and not self.rows and time.time() < deadline:
            time.sleep(0.01)

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.1]
        used = inside if inside else [r for _t, r in self.rows[-3:]]
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in used:
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
            except Exception:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm),
                "samples_inside_timed_region": len(inside), "reasons": sorted(reasons)}


# ======================================================================================================
# CPU arm: the oracle port (oracle/bsem.c, OpenMP, all host cores) on a bounded sample of the workload
# ======================================================================================================
def cpu_pass_setup(sample_rows: int, ncols: int):
    from learningorchestra_b200.build import build_oracle
    build_oracle()
    from oracle import cport
    cport.use_all_cores()
    cols = [cport.synth_f64(0, SEED, c, 0, sample_rows, GEN_LO, GEN_HI) for c in range(ncols)]
    proj = [cols[c] for c in projected_columns(ncols)]
    lo = np.full(ncols, GEN_LO, np.float32)
    hi = np.full(ncols, GEN_HI, np.float32)
    outs = [np.empty(sample_rows, dtype=np.float32) for _ in range(ncols)]

    import ctypes as C
    L = cport.lib()
    in_p = (C.POINTER(C.c_double) * ncols)(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in proj])
    out_p = (C.POINTER(C.c_float) * ncols)(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in outs])
    counts = np.zeros((ncols, NBINS), dtype=np.uint64)

    def one_pass():
        L.oracle_project_cast_hist(in_p, C.c_int64(sample_rows), C.c_int(ncols), out_p, C.c_int(NBINS),
                                   lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p),
                                   counts.ctypes.data_as(C.c_void_p))
        return counts

    return one_pass, cport.num_threads(), (proj, outs)


def run_cpu_baseline(ncols: int, sample_rows: int, passes: int = 3) -> dict:
    one_pass, threads, _keep = cpu_pass_setup(sample_rows, ncols)
    one_pass()
    best = float("inf")
    for _ in range(passes):
        t0 = time.perf_counter()
        one_pass()
        best = min(best, time.perf_counter() - t0)
    return {"value": sample_rows / best, "unit": "rows/s", "cores": threads, "kind": "port",
            "sample": f"{sample_rows} rows x {ncols} cols of the same synthetic table, host-resident, "
                      f"oracle/bsem.c (gcc -O2 -fopenmp), best of {passes} passes; the reference's own "
                      "PySpark+MongoDB path cannot run here (no JVM/pyspark/pymongo/mongod)"}


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncols, sample_rows = args.cols, args.cpu_rows
    one_pass, threads, _keep = cpu_pass_setup(sample_rows, ncols)
    for _ in range(args.warmup):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_pass()
    dt = time.perf_counter() - t0
    value = sample_rows * args.steps / dt
    sample = (f"each step = {sample_rows} rows x {ncols} cols (bounded sample of the {args.rows}-row table), "
              "host-resident columns, oracle port oracle/bsem.c with OpenMP on all host cores")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64->f32", "data": "synthetic",
        "config": {"workload": f"fused project+cast+{NBINS}-bin histogram, {args.rows} x {ncols} fp64 -> fp32, K={ncols}",
                   "note": "reference PySpark/MongoDB stack is not runnable offline; this is the CPU oracle port"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ======================================================================================================
# GPU arm
# ======================================================================================================
def _claim_stdout() -> int:
    """The contract is ONE JSON line on stdout.  Libraries talk on fd 1 too (NCCL prints its version banner there
    when NCCL_DEBUG=VERSION is in the environment), so fd 1 is pointed at stderr for the whole run and the JSON line
    is written to a private duplicate of the real stdout at the end."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_stdout: int, line: dict) -> None:
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


def run_gpu(args) -> None:
    real_stdout = _claim_stdout()
    import torch
    import torch.distributed as dist

    from learningorchestra_b200.build import build_native
    from learningorchestra_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        build_native()                      # no-op when the in-tree .so is newer than its sources
    if world > 1:
        dist.barrier()                      # nobody loads libloexec.so while rank 0 might be rewriting it

    eng = Engine(local_rank)
    ncols, total_rows = args.cols, args.rows
    from learningorchestra_b200.sharding import allreduce_counts, shard_bounds
    r_begin, r_end = shard_bounds(total_rows, world, rank)
    nrows = r_end - r_begin
    cols = projected_columns(ncols)
    k = len(cols)
    # a non-default stream: libloexec launches on exactly the stream it is handed (NULL would mean its
    # own), and torch.cuda.Event / NCCL then see the same stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    table = eng.table("f64", nrows, ncols).fill_synthetic(0, SEED, row_offset=r_begin, lo=GEN_LO, hi=GEN_HI, stream=stream)
    out = eng.table("f32", nrows, k)
    counts_t = torch.zeros(k * NBINS, dtype=torch.int64, device="cuda")      # uint64 bit patterns; sums are identical
    counts = eng.wrap_counts(k, NBINS, counts_t.data_ptr(), keepalive=counts_t)
    lo = np.full(k, GEN_LO, np.float32)
    hi = np.full(k, GEN_HI, np.float32)
    torch.cuda.synchronize()

    # N > 1: "p2p" fuses the merge into the kernel's flush (system-scope REDs into rank 0's matrix over NVLink);
    # "nccl" is local counts + one all-reduce.  "auto" = p2p when the CUDA-IPC setup succeeds on every rank and a
    # probe step completes without a flag time-out, else nccl — both are GPU paths, the line says which ran.
    peer, merge_note = None, "nccl"
    if world > 1 and args.merge in ("p2p", "auto"):
        from learningorchestra_b200.sharding import PeerReduce
        ok = torch.ones(1, device="cuda")
        try:
            peer = PeerReduce(eng, k, NBINS)
        except Exception as exc:          # e.g. IPC not permitted in this container
            log(f"[rank {rank}] peer-memory merge unavailable: {exc!r}")
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok[0]) == 0.0:
            if args.merge == "p2p":
                raise SystemExit("--merge p2p requested but the peer-memory setup failed")
            if peer is not None:
                peer.close()
            peer = None
        merge_note = "p2p" if peer is not None else "nccl (p2p setup failed)"

    kev = []   # (start, end) events around the fused kernel only, for the roofline

    def step(record: bool):
        if peer is not None:
            # merge fused into the kernel's flush: system-scope REDs into the root's matrix over NVLink
            peer.before_kernel(stream)
            dst, is_peer = peer.counts_for_step(), True
        else:
            counts.zero(stream)
            dst, is_peer = counts, False
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        eng.project_cast_hist(table, cols, NBINS, lo, hi, out=out, counts=dst, stream=stream, peer_counts=is_peer)
        if record:
            e1.record(stream)
            kev.append((e0, e1))
        if peer is not None:
            peer.after_kernel(stream)
        elif world > 1:
            allreduce_counts(counts_t)          # ONE ncclAllReduce of k*nbins int64 over NVLink, same stream

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    if peer is not None:
        bad = torch.tensor([float(peer.timed_out(stream))], device="cuda")
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if float(bad[0]) > 0:
            if args.merge == "p2p":
                raise SystemExit("peer-memory merge timed out during warm-up")
            log(f"[rank {rank}] peer-memory merge timed out in warm-up; using the NCCL all-reduce")
            peer.close()
            peer, merge_note = None, "nccl (p2p timed out in warm-up)"
            for _ in range(args.warmup):
                step(False)
            torch.cuda.synchronize()
    sampler.wait_ready()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_start()
    launches0 = eng.launch_count
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for _ in range(args.steps):
        step(True)
    t_end.record(stream)
    torch.cuda.synchronize()
    sampler.mark_end()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    kernel_ms = [a.elapsed_time(b) for a, b in kev]
    t = torch.tensor([elapsed_ms, sum(kernel_ms) / len(kernel_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, kernel_ms_avg = float(t[0]), float(t[1])
    if peer is not None:
        assert peer.timed_out(stream) == 0, "peer-memory merge timed out"
        final_counts = peer.result_numpy(stream)
    else:
        final_counts = counts.to_numpy(stream)
    if final_counts is not None:
        total_counted = int(final_counts.sum())
        assert total_counted == total_rows * k, f"histogram lost rows: {total_counted} != {total_rows * k}"

    # ---- end to end: host buffers in, host buffers out, through the plugin-facing C-ABI call ----------
    e2e = None
    e2e_ready = False
    if not args.no_e2e:
        import ctypes as C
        import psutil
        from learningorchestra_b200 import _native as N
        avail = psutil.virtual_memory().available
        per_row = 12 * k
        budget_rows = int(avail * 0.45 / world / per_row)
        e2e_rows = min(nrows, args.e2e_rows if args.e2e_rows else nrows, budget_rows)
        e2e_rows = max(61440, e2e_rows // 61440 * 61440) if e2e_rows >= 61440 else e2e_rows
        # set-up (pinned host buffers, staging) can fail on a box with little free RAM: every rank reports, and e2e
        # is skipped on ALL ranks together rather than leaving some of them waiting in a collective
        setup_error = None
        try:
            hin = eng.pinned_empty((k, e2e_rows), np.float64)
            hout = eng.pinned_empty((k, e2e_rows), np.float32)
            for j in range(k):   # host inputs = the projected columns of this rank's shard (device -> pinned host, untimed)
                N.check(eng._lib.lo_table_download_col(eng._ctx, table._h, cols[j], 0, hin[j].ctypes.data_as(C.c_void_p), e2e_rows))
            in_cols = [hin[j] for j in range(k)]
            out_cols = [hout[j] for j in range(k)]
            eng.project_cast_hist_host(in_cols, NBINS, lo, hi, out=out_cols)      # warm-up (allocates staging)
        except Exception as exc:          # noqa: BLE001
            setup_error = repr(exc)
            log(f"[rank {rank}] e2e set-up failed: {setup_error}")
        okf = torch.tensor([0.0 if setup_error else 1.0], device="cuda")
        if world > 1:
            dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        e2e_ready = float(okf[0]) == 1.0
        if not e2e_ready:
            e2e = {"value": None, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "error": setup_error or "set-up failed on another rank"}
    if e2e_ready:
        if world > 1:
            dist.barrier()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        l0 = eng.launch_count
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            c_host, tm = eng.project_cast_hist_host(in_cols, NBINS, lo, hi, out=out_cols)
            if world > 1:
                ct = torch.from_numpy(c_host.view(np.int64)).cuda()
                dist.all_reduce(ct)
                c_host = ct.cpu().numpy()
        dt = time.perf_counter() - t0
        e2e_launches = eng.launch_count - l0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        rr = torch.tensor([e2e_rows], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(rr, op=dist.ReduceOp.SUM)
        e2e = {"value": float(rr[0]) * e2e_steps / float(tt[0]), "unit": "rows/s",
               "h2d_bytes_per_step": int(tm["h2d_bytes"]) * world, "d2h_bytes_per_step": int(tm["d2h_bytes"]) * world,
               "rows_per_step": int(float(rr[0])), "steps": e2e_steps, "launches": e2e_launches,
               "api": "Engine.project_cast_hist_host -> lo_project_cast_hist_host (pinned host buffers, "
                      "chunked H2D / kernel / D2H on three streams)"}

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = 12.0 * k * nrows                       # 8 B read + 4 B written per projected element
        achieved = alg_bytes / (kernel_ms_avg * 1e-3) / 1e9
        cpu = run_cpu_baseline(ncols, args.cpu_rows) if world == 1 and not args.no_cpu else None
        line = {
            "metric": METRIC, "value": total_rows * args.steps / (elapsed_ms * 1e-3), "unit": "rows/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64->f32",
            "data": "synthetic",
            "config": {"workload": f"fused project+cast+{NBINS}-bin histogram, {total_rows} x {ncols} fp64 -> fp32, "
                                   f"K={k} (permutation), columnar, range [{GEN_LO}, {GEN_HI}]",
                       "rows": total_rows, "cols": ncols, "k": k, "nbins": NBINS, "rows_per_gpu": nrows,
                       "merge": merge_note if world > 1 else None,
                       "parallelism": (f"row-range shards x{world}, " + (
                           f"one NCCL all-reduce of {k}x{NBINS} uint64 per step" if peer is None else
                           "merge fused into the kernel flush: system-scope RED.64 into rank 0's matrix over NVLink (CUDA IPC)"))
                                      if world > 1 else "single GPU",
                       "l2": f"inputs larger than L2: {nrows * ncols * 8 / 1e9:.1f} GB read + "
                             f"{nrows * k * 4 / 1e9:.1f} GB written per GPU per step (L2 = 126 MB), no flush needed"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "lo::k_project_cast_hist<1,true,true>",
                         "kernel_ms_avg": kernel_ms_avg, "algorithmic_bytes_per_launch": alg_bytes},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        tr = ROOT / "profiles" / "traffic.json"
        if tr.exists():
            try:
                t100 = json.loads(tr.read_text()).get("k_project_cast_hist_bytes_per_launch")   # ncu, 100M-row launch
                line["roofline"]["traffic"] = t100 * nrows / 100_000_000 if t100 else None
                line["roofline"]["traffic_source"] = "ncu --set full dram__bytes_read+write of one 100M x 32 launch (profiles/), scaled to this launch's rows"
            except Exception:
                pass
        _emit(real_stdout, line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--cols", type=int, default=32)
    ap.add_argument("--cpu-rows", type=int, default=8_000_000, help="rows of the bounded CPU sample")
    ap.add_argument("--e2e-rows", type=int, default=0, help="cap on e2e rows per rank (0 = whole shard if RAM allows)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--merge", default="auto", choices=["auto", "nccl", "p2p"],
                    help="N > 1: how partial histograms are merged (NCCL all-reduce, or fused peer-memory REDs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
