"""A/B of two builds of libloexec.so on the same box, interleaved: each (library, workload) is timed in a fresh
subprocess (LOEXEC_LIB selects the build).  usage: ab_libs.py libA.so libB.so [rows]"""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
CHILD = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
from learningorchestra_b200.engine import Engine
rows = int(sys.argv[1])
eng = Engine(0); st = torch.cuda.Stream(); torch.cuda.set_stream(st)
t = eng.table("f64", rows, 32).fill_synthetic(0, 20260921, stream=st); out = eng.table("f32", rows, 32)
cols = [(7 * j + 3) %% 32 for j in range(32)]; c = eng.counts(32, 256)
lo, hi = np.full(32, -1000, np.float32), np.full(32, 1000, np.float32)
fn = lambda: eng.project_cast_hist(t, cols, 256, lo, hi, out=out, counts=c, stream=st)
for _ in range(3): fn()
torch.cuda.synchronize(); ev = []
for _ in range(15):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st); fn(); b.record(st); ev.append((a, b))
torch.cuda.synchronize(); ts = sorted(x.elapsed_time(y) for x, y in ev)
print(json.dumps({"ms_med": ts[len(ts) // 2], "ms_best": ts[0]}))
''' % str(ROOT)
libs = [a for a in sys.argv[1:] if a.endswith(".so")]
rows = next((int(a) for a in sys.argv[1:] if a.isdigit()), 100_000_000)
res = {l: [] for l in libs}
for rep in range(4):
    for l in libs:
        out = subprocess.run([sys.executable, "-c", CHILD, str(rows)], env=dict(os.environ, LOEXEC_LIB=l, LOEXEC_LAX="1"), capture_output=True, text=True)
        line = [x for x in out.stdout.splitlines() if x.startswith("{")]
        if not line:
            print("FAILED", l, out.stderr[-500:]); continue
        res[l].append(json.loads(line[-1])["ms_med"])
        print(rep, l, res[l][-1], flush=True)
summary = {l: {"runs_ms": v, "median_ms": sorted(v)[len(v) // 2] if v else None} for l, v in res.items()}
print(json.dumps(summary))
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"ab_libs_{rows}.json").write_text(json.dumps(summary, indent=1))
