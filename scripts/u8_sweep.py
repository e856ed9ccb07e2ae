"""k_hist_u8_cols variants (LOEXEC_U8_MODE) on MNIST-shaped and dense random byte tables: CUDA events, data resident.
Diagnostic; writes gpurun_out/u8_sweep.json."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from learningorchestra_b200.engine import Engine


def timeit(fn, stream, iters=20, warm=3):
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
    eng = Engine(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    tables = {}
    tables["mnist_1Mx784"] = (eng.table("u8", 1_000_000, 784).fill_synthetic(3, 20260921, stream=stream), 784)
    tables["mnist_8Mx784"] = (eng.table("u8", 8_000_000, 784).fill_synthetic(3, 20260921, stream=stream), 784)
    rng = np.random.default_rng(3)
    dense = eng.table("u8", 4_000_000, 128)
    blk = rng.integers(0, 256, 4_000_000, dtype=np.uint8)
    for c in range(128):
        dense.upload(c, np.roll(blk, c * 977))
    tables["dense_random_4Mx128"] = (dense, 128)
    res = []
    ref = {}
    for mode in (4, 2, 5, 6):
        os.environ["LOEXEC_U8_MODE"] = str(mode)
        for name, (t, k) in tables.items():
            c = eng.counts(k, 256)
            med, best = timeit(lambda: eng.hist_u8_cols(t, range(k), counts=c, stream=stream), stream)
            c.zero(stream); eng.hist_u8_cols(t, range(k), counts=c, stream=stream)
            got = c.to_numpy(stream)
            if name not in ref: ref[name] = got
            same = bool(np.array_equal(got, ref[name]))
            gbs = t.nrows * k / med / 1e6
            res.append({"mode": mode, "table": name, "ms_med": med, "ms_best": best, "GBs_med": gbs, "same_counts_as_mode4": same})
            print(f"mode {mode} {name:22s} med {med:7.4f} ms best {best:7.4f}  {gbs:7.1f} GB/s  same={same}", flush=True)
            c.free()
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/u8_sweep.json").write_text(json.dumps(res, indent=1))


main()
