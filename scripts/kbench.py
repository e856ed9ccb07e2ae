"""Kernel-variant timing (CUDA events, data resident, inputs >> L2).  Diagnostic, not the bench contract."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from learningorchestra_b200.engine import Engine

def timeit(fn, stream, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in evs)
    return ts[len(ts) // 2], ts[0]

def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    ncols = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    kind = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    eng = Engine(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    t = eng.table("f64", rows, ncols).fill_synthetic(kind, 20260921, stream=stream)
    out = eng.table("f32", rows, ncols)
    cols = [(7 * j + 3) % ncols for j in range(ncols)]
    counts = eng.counts(ncols, 256)
    lo, hi = np.full(ncols, -1000, np.float32), np.full(ncols, 1000, np.float32)
    res = {}
    def rep(name, med, best, nbytes):
        res[name] = {"ms_med": med, "ms_best": best, "GBs_med": nbytes / med / 1e6, "GBs_best": nbytes / best / 1e6}
        print(f"{name:28s} med {med:8.3f} ms  best {best:8.3f} ms  {nbytes/med/1e6:8.1f} GB/s (best {nbytes/best/1e6:8.1f})", flush=True)
    m, b = timeit(lambda: eng.project_cast(t, cols, out=out, stream=stream), stream); rep("project+cast", m, b, 12.0 * rows * ncols)
    m, b = timeit(lambda: eng.project_cast_hist(t, cols, 256, lo, hi, out=out, counts=counts, stream=stream), stream); rep("project+cast+hist256", m, b, 12.0 * rows * ncols)
    m, b = timeit(lambda: eng.project_cast_hist(t, cols, 256, lo, hi, out=None, counts=counts, stream=stream), stream); rep("hist256 only", m, b, 8.0 * rows * ncols)
    m, b = timeit(lambda: eng.project_cast_hist(t, cols, 10, lo, hi, out=out, counts=counts, stream=stream), stream); rep("project+cast+hist10", m, b, 12.0 * rows * ncols)
    out64 = None
    if rows * ncols * 8 * 2 < 150e9:
        out.free(); out64 = eng.table("f64", rows, ncols)
        m, b = timeit(lambda: eng.project_cast(t, cols, out=out64, stream=stream), stream); rep("project f64 copy", m, b, 16.0 * rows * ncols)
        a = torch.empty(rows * ncols // 2, dtype=torch.float64, device="cuda"); bb = torch.empty_like(a)
        m, b = timeit(lambda: bb.copy_(a), stream); rep("torch copy_ (peak probe)", m, b, 16.0 * a.numel())
    # u8
    tu = eng.table("u8", 1_000_000, 784).fill_synthetic(3, 20260921, stream=stream)
    c8 = eng.counts(784, 256)
    m, b = timeit(lambda: eng.hist_u8_cols(tu, range(784), counts=c8, stream=stream), stream); rep("hist_u8 1Mx784", m, b, 784e6)
    tu2 = eng.table("u8", 8_000_000, 784).fill_synthetic(3, 20260921, stream=stream)
    m, b = timeit(lambda: eng.hist_u8_cols(tu2, range(784), counts=c8, stream=stream), stream); rep("hist_u8 8Mx784", m, b, 8 * 784e6)
    rng = np.random.default_rng(3)
    dense = eng.table("u8", 4_000_000, 128)
    blk = rng.integers(0, 256, 4_000_000, dtype=np.uint8)
    for c in range(128):
        dense.upload(c, np.roll(blk, c * 977))
    c128 = eng.counts(128, 256)
    m, b = timeit(lambda: eng.hist_u8_cols(dense, range(128), counts=c128, stream=stream), stream); rep("hist_u8 dense random 4Mx128", m, b, 128 * 4e6)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/kbench.json").write_text(json.dumps(res, indent=1))

main()
