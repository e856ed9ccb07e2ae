"""e2e (host buffers) throughput probe: chunk size x CPU affinity.  Diagnostic."""
import os, sys, time, subprocess, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from learningorchestra_b200.engine import Engine
    import ctypes as C
    from learningorchestra_b200 import _native as N
    rows, k = 30_000_000, 32
    eng = Engine(0)
    t = eng.table("f64", rows, k).fill_synthetic(0, 1)
    hin = eng.pinned_empty((k, rows), np.float64); hout = eng.pinned_empty((k, rows), np.float32)
    for j in range(k):
        N.check(eng._lib.lo_table_download_col(eng._ctx, t._h, j, 0, hin[j].ctypes.data_as(C.c_void_p), rows))
    lo, hi = np.full(k, -1000, np.float32), np.full(k, 1000, np.float32)
    ins, outs = [hin[j] for j in range(k)], [hout[j] for j in range(k)]
    eng.project_cast_hist_host(ins, 256, lo, hi, out=outs)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); eng.project_cast_hist_host(ins, 256, lo, hi, out=outs); best = min(best, time.perf_counter() - t0)
    t0 = time.perf_counter(); eng.project_cast_hist_host(ins, 256, lo, hi, out=None); h_only = time.perf_counter() - t0
    print(json.dumps({"rows_s": rows / best, "h2d_GBs": rows * k * 8 / best / 1e9, "hist_only_h2d_GBs": rows * k * 8 / h_only / 1e9}))
    eng.close()
else:
    for aff in ("0-31", "32-63", None):
        for mb in (64, 256, 1024):
            env = dict(os.environ, LOEXEC_CHUNK_MB=str(mb))
            cmd = ([ "taskset", "-c", aff] if aff else []) + [sys.executable, __file__, "child"]
            r = subprocess.run(cmd, env=env, capture_output=True, text=True)
            print(aff, mb, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
