"""k_project_cast_hist_bins (more than 256 bins) against the 256-bin tile kernel on the same resident table: CUDA events,
median of 10 launches, GB/s of algorithmic bytes (12 B per element with the fp32 output).  Writes gpurun_out/bins_bench.json."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from learningorchestra_b200.engine import Engine

eng = Engine(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
rows, k = 25_000_000, 16
t = eng.table("f64", rows, k).fill_synthetic(1, 20260921, stream=stream)
out = eng.table("f32", rows, k)
lo, hi = np.full(k, -1000, np.float32), np.full(k, 1000, np.float32)
res = []
for nbins in (256, 257, 1000, 4096, 16384, 16385, 57344, 65536):
    c = eng.counts(k, nbins)
    fn = lambda: eng.project_cast_hist(t, range(k), nbins, lo, hi, out=out, counts=c, stream=stream)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); evs.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(x.elapsed_time(y) for x, y in evs)[5]
    c.zero(stream); fn()
    total = int(c.to_numpy(stream).sum())
    res.append({"nbins": nbins, "ms": ms, "GBs": rows * k * 12 / ms / 1e6, "rows_per_s": rows / ms * 1e3, "counted": total})
    print(res[-1], flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bins_bench.json").write_text(json.dumps({"table": f"{rows} x {k} f64 -> f32", "results": res}, indent=1))
