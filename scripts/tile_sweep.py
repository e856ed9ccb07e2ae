"""Tapered-tail sweep of the fused kernel (CUDA events, data resident): which (tail_batches, tail_waves) of
make_tilemap() is fastest at the shard sizes of the 1/2/4/8-GPU runs.  Diagnostic; writes gpurun_out/tile_sweep.json."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from learningorchestra_b200.engine import Engine


def timeit(fn, stream, iters=12, warm=3):
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
    sizes = [int(x) for x in sys.argv[1:]] or [12_500_000, 25_000_000, 50_000_000, 100_000_000]
    eng = Engine(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    ncols = 32
    cols = [(7 * j + 3) % ncols for j in range(ncols)]
    lo, hi = np.full(ncols, -1000, np.float32), np.full(ncols, 1000, np.float32)
    res = []
    for rows in sizes:
        t = eng.table("f64", rows, ncols).fill_synthetic(0, 20260921, stream=stream)
        out = eng.table("f32", rows, ncols)
        counts = eng.counts(ncols, 256)
        for tb, tw in [(0, 1), (2, 0.5), (2, 1), (4, 0.5), (4, 1), (4, 1.5), (4, 2), (6, 1), (6, 2), (8, 1), (8, 2)]:
            os.environ["LOEXEC_TAIL_BATCHES"] = str(tb); os.environ["LOEXEC_TAIL_WAVES"] = str(tw)
            med, best = timeit(lambda: eng.project_cast_hist(t, cols, 256, lo, hi, out=out, counts=counts, stream=stream), stream)
            gbs = 12.0 * rows * ncols / med / 1e6
            res.append({"rows": rows, "tail_batches": tb, "tail_waves": tw, "ms_med": med, "ms_best": best, "GBs_med": gbs})
            print(f"rows {rows:>11d} tail_batches {tb} waves {tw:<4} med {med:7.4f} ms best {best:7.4f}  {gbs:7.1f} GB/s", flush=True)
        t.free(); out.free(); counts.free()
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/tile_sweep.json").write_text(json.dumps(res, indent=1))


main()
