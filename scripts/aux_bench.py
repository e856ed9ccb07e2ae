"""The SURVEY 8(f) kernels — decimal parser, numeric / text group-by — through their host entry points: whole-call rate
(H2D and D2H included, pageable and pinned input) and the kernels' own device time (lo_host_timing.kernel_ms, two events on
the call's stream) against their algorithmic bytes.  Diagnostic; writes gpurun_out/aux_bench.json."""
import ctypes as C
import json
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from learningorchestra_b200 import _native as N
from learningorchestra_b200.engine import Engine

eng = Engine(0); rng = np.random.default_rng(1); res = {}
PEAK = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6591.9) \
    if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else 6591.9


def best(fn, n=5):
    fn(); runs = []
    for _ in range(n):
        t0 = time.perf_counter(); k = fn(); runs.append((time.perf_counter() - t0, k))
    return min(runs)


def pinned_copy(a):
    p = eng.pinned_empty(a.shape, a.dtype); p[...] = a
    return p


# ---- parser: 8 M cells like "-1234.567890" -------------------------------------------------------------------------
ncell = 8_000_000
cells = np.char.mod("%.6f", rng.uniform(-1e4, 1e4, ncell))
enc = [c.encode() for c in cells.tolist()]
offs = np.zeros(ncell + 1, np.int64); np.cumsum([len(b) for b in enc], out=offs[1:])
chars = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8).copy()
out_v = np.zeros(ncell); out_s = np.zeros(ncell, np.uint8)


def parse(ch, of, ov, os_):
    tm = N.HostTiming()
    N.check(eng._lib.lo_parse_number_host(eng._ctx, ch.ctypes.data_as(C.c_void_p), of.ctypes.data_as(C.c_void_p), ncell,
                                          ov.ctypes.data_as(C.c_void_p), os_.ctypes.data_as(C.c_void_p), C.byref(tm)))
    return tm.kernel_ms


t, kms = best(lambda: parse(chars, offs, out_v, out_s))
assert np.array_equal(out_v[:200_000], np.array([float(c) for c in cells[:200_000].tolist()]))
pc, po, pv, ps = pinned_copy(chars), pinned_copy(offs), eng.pinned_empty(ncell, np.float64), eng.pinned_empty(ncell, np.uint8)
tp, kmsp = best(lambda: parse(pc, po, pv, ps))
abytes = int(offs[-1]) + (ncell + 1) * 8 + ncell * 9            # text + offsets read, value + status written
res["parse_number"] = {"cells": ncell, "text_bytes": int(offs[-1]), "algorithmic_bytes": abytes,
                       "call_s_pageable": t, "Mcells_per_s_pageable": ncell / t / 1e6,
                       "call_s_pinned": tp, "Mcells_per_s_pinned": ncell / tp / 1e6,
                       "kernel_ms": min(kms, kmsp), "kernel_Mcells_per_s": ncell / min(kms, kmsp) / 1e3,
                       "kernel_GBs": abytes / min(kms, kmsp) / 1e6, "kernel_frac_of_hbm_peak": abytes / min(kms, kmsp) / 1e6 / PEAK}
t0 = time.perf_counter(); _ = [float(c) for c in cells[:1_000_000].tolist()]; t_py = time.perf_counter() - t0
res["python_float_loop"] = {"Mcells_per_s": 1.0 / t_py}

# ---- numeric group-by ------------------------------------------------------------------------------------------------
nrow = 20_000_000
for name, x in [("10 keys", rng.integers(0, 10, nrow).astype(np.float64)), ("1M keys", rng.integers(0, 1_000_000, nrow).astype(np.float64)),
                ("all distinct", rng.permutation(nrow).astype(np.float64))]:
    cap = nrow
    keys = np.empty(cap); counts = np.empty(cap, np.uint64); nd = C.c_int64()

    def vc(src):
        tm = N.HostTiming()
        N.check(eng._lib.lo_value_counts_f64_host(eng._ctx, src.ctypes.data_as(C.c_void_p), nrow, keys.ctypes.data_as(C.c_void_p),
                                                  counts.ctypes.data_as(C.c_void_p), cap, C.byref(nd), C.byref(tm)))
        return tm.kernel_ms
    t, kms = best(lambda: vc(x), 3)
    assert int(counts[:nd.value].sum()) == nrow
    xp = pinned_copy(x)
    tp, kmsp = best(lambda: vc(xp), 3)
    k = min(kms, kmsp)
    res["value_counts_f64 " + name] = {"rows": nrow, "groups": int(nd.value), "call_s_pageable": t, "Mrows_per_s_pageable": nrow / t / 1e6,
                                       "call_s_pinned": tp, "Mrows_per_s_pinned": nrow / tp / 1e6, "kernel_ms": k,
                                       "kernel_Mrows_per_s": nrow / k / 1e3, "kernel_GBs_of_8B_keys": nrow * 8 / k / 1e6}

# ---- text group-by: 10 M cells, 1000 distinct strings of 4-12 bytes ----------------------------------------------------
nstr = 10_000_000
vocab = [("k%d" % i).encode() + b"x" * int(rng.integers(2, 10)) for i in range(1000)]
pick = rng.integers(0, 1000, nstr)
lens = np.array([len(v) for v in vocab], np.int64)[pick]
soffs = np.zeros(nstr + 1, np.int64); np.cumsum(lens, out=soffs[1:])
schars = np.frombuffer(b"".join(vocab[i] for i in pick.tolist()), dtype=np.uint8).copy()
rep = np.empty(1 << 16, np.int64); cnt = np.empty(1 << 16, np.uint64); nds = C.c_int64()


def vcs(ch, of):
    tm = N.HostTiming()
    N.check(eng._lib.lo_value_counts_str_host(eng._ctx, ch.ctypes.data_as(C.c_void_p), of.ctypes.data_as(C.c_void_p), nstr,
                                              rep.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), 1 << 16, C.byref(nds), C.byref(tm)))
    return tm.kernel_ms


t, kms = best(lambda: vcs(schars, soffs), 3)
assert nds.value == 1000 and int(cnt[:1000].sum()) == nstr
sp, so = pinned_copy(schars), pinned_copy(soffs)
tp, kmsp = best(lambda: vcs(sp, so), 3)
k = min(kms, kmsp); sb = int(soffs[-1]) + (nstr + 1) * 8
res["value_counts_str 1000 keys"] = {"rows": nstr, "text_bytes": int(soffs[-1]), "call_s_pageable": t, "Mrows_per_s_pageable": nstr / t / 1e6,
                                     "call_s_pinned": tp, "Mrows_per_s_pinned": nstr / tp / 1e6, "kernel_ms": k,
                                     "kernel_Mrows_per_s": nstr / k / 1e3, "kernel_GBs_text_plus_offsets": sb / k / 1e6}
print(json.dumps(res, indent=1))
Path("gpurun_out").mkdir(exist_ok=True); Path("gpurun_out/aux_bench.json").write_text(json.dumps(res, indent=1))
