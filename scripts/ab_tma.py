"""Interleaved A/B of the LDG-pipelined and TMA-staged fused kernels (same process, alternating)."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from learningorchestra_b200.engine import Engine
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
eng = Engine(0); stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
t = eng.table("f64", rows, 32).fill_synthetic(0, 20260921, stream=stream); out = eng.table("f32", rows, 32)
cols = [(7 * j + 3) % 32 for j in range(32)]; counts = eng.counts(32, 256)
lo, hi = np.full(32, -1000, np.float32), np.full(32, 1000, np.float32)
def run(kind):
    if kind == "fused": eng.project_cast_hist(t, cols, 256, lo, hi, out=out, counts=counts, stream=stream)
    else: eng.project_cast(t, cols, out=out, stream=stream)
res = {}
for kind in ("fused", "cast"):
    times = {0: [], 1: []}
    for rep in range(12):
        for tma in (0, 1):
            eng.set_tma(bool(tma))
            for _ in range(2): run(kind)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(5): run(kind)
            b.record(stream); torch.cuda.synchronize()
            times[tma].append(a.elapsed_time(b) / 5)
    res[kind] = {("tma" if k else "ldg"): {"median_ms": float(np.median(v)), "min_ms": float(min(v))} for k, v in times.items()}
    print(kind, res[kind], flush=True)
Path("gpurun_out").mkdir(exist_ok=True); Path("gpurun_out/ab_tma_%d.json" % rows).write_text(json.dumps(res, indent=1))
