"""GPU soak: broad randomized validation, written to gpurun_out/soak.json (summarised into profiles/).
  1. exhaustive (all 2^32 fp32 patterns) fast-divide == IEEE-divide self-test for many random (lo, hi, nbins);
  2. k_parse_number vs CPython float() on millions of random cells;
  3. fused kernel vs the streaming C oracle on random shapes / column picks / ranges / bin counts (1 .. 65 536: the
     tile kernel and the wide-bin chunk kernel);
  4. byte-histogram kernel vs the streaming C oracle on random MNIST-shaped tables.
Bounded by --seconds."""
import argparse, json, math, random, struct, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from learningorchestra_b200.build import build_all
build_all()
from learningorchestra_b200.engine import Engine
from oracle import cport

ap = argparse.ArgumentParser(); ap.add_argument("--seconds", type=float, default=120); ap.add_argument("--seed", type=int, default=1)
a = ap.parse_args()
rng = random.Random(a.seed); eng = Engine(0); t_end = time.time() + a.seconds
res = {"fastdiv": {"triples": 0, "fast_path_used": 0, "mismatching_triples": 0},
       "parse": {"cells": 0, "mismatches": 0}, "fused": {"cases": 0, "rows": 0, "mismatches": 0, "cases_above_256_bins": 0},
       "bytes": {"cases": 0, "rows": 0, "mismatches": 0}}
# 1 ------------------------------------------------------------------------------------------------------
while time.time() < t_end - a.seconds * 0.6:
    kind = rng.random()
    if kind < 0.4:
        lo = rng.uniform(-1e4, 1e4); hi = lo + rng.uniform(1e-3, 1e4)
    elif kind < 0.7:
        lo = rng.uniform(-1, 1) * 10 ** rng.randint(-30, 30); hi = lo + abs(lo) * rng.uniform(1e-6, 10) + 10 ** rng.randint(-35, 30)
    else:
        lo = float(rng.randint(-1000, 1000)); hi = lo + float(rng.randint(1, 100000))
    nb = rng.choice([1, 2, 3, 7, 10, 16, 100, 255, 256, rng.randint(1, 256), 1000, rng.randint(257, 65536)])
    lo32, hi32 = float(np.float32(lo)), float(np.float32(hi))
    if not (hi32 > lo32 and math.isfinite(hi32 - lo32)):
        continue
    try:
        used, bad = eng.selftest_fastdiv(lo32, hi32, nb)
    except Exception:
        continue
    res["fastdiv"]["triples"] += 1; res["fastdiv"]["fast_path_used"] += int(used)
    if used and bad:
        res["fastdiv"]["mismatching_triples"] += 1
        res["fastdiv"].setdefault("examples", []).append([lo32, hi32, nb, bad])
# 2 ------------------------------------------------------------------------------------------------------
while time.time() < t_end - a.seconds * 0.4:
    cells = []
    for _ in range(400_000):
        k = rng.random(); b = rng.getrandbits(64)
        if k < 0.2: b = (b & 0x800FFFFFFFFFFFFF) | (rng.randint(0, 3) << 52)
        elif k < 0.3: b = (b & 0x800FFFFFFFFFFFFF) | (rng.randint(0x7FB, 0x7FE) << 52)
        v = struct.unpack("<d", struct.pack("<Q", b))[0]
        if v != v or v in (float("inf"), float("-inf")):
            v = rng.uniform(-1e9, 1e9)
        s = "%.*e" % (rng.choice([0, 5, 14, 15, 16, 17, 18, 19, 20, 24, 40]), v) if rng.random() < 0.7 else repr(v)
        if rng.random() < 0.1: s = s.replace("e", "E")
        if rng.random() < 0.05: s = " " + s + "\t"
        cells.append(s)
    vals, st = eng.parse_number_host(cells)
    exp = np.array([float(c) for c in cells])
    ok = (vals.view(np.uint64) == exp.view(np.uint64))
    exp_int = np.array([1 if (math.isfinite(x) and x.is_integer()) else 0 for x in exp], dtype=np.uint8)
    res["parse"]["cells"] += len(cells); res["parse"]["mismatches"] += int((~ok).sum() + (st != exp_int).sum())
# 3 ------------------------------------------------------------------------------------------------------
while time.time() < t_end - a.seconds * 0.12:
    nrows = rng.choice([rng.randint(1, 5000), rng.randint(50_000, 70_000), rng.randint(100_000, 3_000_000)])
    ncols = rng.randint(1, 40); k = rng.randint(1, 12); cols = [rng.randrange(ncols) for _ in range(k)]
    nb = rng.choice([1, 2, 10, 64, 255, 256, rng.randint(1, 256), 257, rng.randint(257, 20000), rng.randint(257, 65536)]); kind = rng.choice([0, 1, 2]); seed = rng.getrandbits(40); row0 = rng.randint(0, 10 ** 12)
    lo = np.array([rng.uniform(-1200, 0) for _ in range(k)], np.float32); hi = lo + np.array([rng.uniform(1, 2400) for _ in range(k)], np.float32)
    t = eng.table("f64", nrows, ncols).fill_synthetic(kind, seed, row_offset=row0)
    out = eng.table("f32", nrows, k)
    dc = eng.project_cast_hist(t, cols, nb, lo, hi, out=out); got = dc.to_numpy(); dc.free()
    exp, sums = cport.synth_project_cast_hist(kind, seed, row0, nrows, -1000.0, 1000.0, cols, nb, lo, hi)
    bad = int(not np.array_equal(got, exp)) + sum(int(out.checksum(j, row0) != int(sums[j])) for j in range(k))
    res["fused"]["cases"] += 1; res["fused"]["rows"] += nrows * k; res["fused"]["mismatches"] += bad
    res["fused"]["cases_above_256_bins"] += int(nb > 256)
    if bad: res["fused"].setdefault("examples", []).append([nrows, ncols, cols, nb, kind, seed, row0])
    t.free(); out.free()
# 4 ------------------------------------------------------------------------------------------------------
while time.time() < t_end:
    nrows = rng.choice([rng.randint(1, 5000), rng.randint(30_000, 70_000), rng.randint(100_000, 2_000_000)])
    ncols = rng.randint(1, 400); k = rng.randint(1, min(ncols, 300)); cols = [rng.randrange(ncols) for _ in range(k)]
    seed = rng.getrandbits(40); row0 = rng.randint(0, 10 ** 9)
    t = eng.table("u8", nrows, ncols).fill_synthetic(3, seed, row_offset=row0)
    dc = eng.hist_u8_cols(t, cols); got = dc.to_numpy(); dc.free()
    exp = cport.synth_hist_u8(seed, row0, nrows, cols)
    bad = int(not np.array_equal(got, exp))
    res["bytes"]["cases"] += 1; res["bytes"]["rows"] += nrows * k; res["bytes"]["mismatches"] += bad
    if bad: res["bytes"].setdefault("examples", []).append([nrows, ncols, cols, seed, row0])
    t.free()
eng.close()
Path("gpurun_out").mkdir(exist_ok=True); Path("gpurun_out/soak.json").write_text(json.dumps(res, indent=1)); print(json.dumps(res))
