"""k_hist_u8_cols variants (LOEXEC_U8_MODE) on MNIST-shaped and dense random byte tables: CUDA events, data resident.
Each mode runs in its own process (a variant that traps must not take the others down).  Diagnostic; writes
gpurun_out/u8_sweep.json."""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def child(mode: int):
    import numpy as np
    import torch
    from learningorchestra_b200.engine import Engine
    os.environ["LOEXEC_U8_MODE"] = str(mode)

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

    eng = Engine(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    tables = {}
    only = os.environ.get("U8_SWEEP_TABLES", "")          # substring filter (for one-table ncu captures)
    if only in "mnist_1Mx784":
        tables["mnist_1Mx784"] = (eng.table("u8", 1_000_000, 784).fill_synthetic(3, 20260921, stream=stream), 784)
    if only in "mnist_8Mx784":
        tables["mnist_8Mx784"] = (eng.table("u8", 8_000_000, 784).fill_synthetic(3, 20260921, stream=stream), 784)
    if only in "dense_random_4Mx128":
        rng = np.random.default_rng(3)
        dense = eng.table("u8", 4_000_000, 128)
        blk = rng.integers(0, 256, 4_000_000, dtype=np.uint8)
        for c in range(128):
            dense.upload(c, np.roll(blk, c * 977))
        tables["dense_random_4Mx128"] = (dense, 128)
    res = []
    for name, (t, k) in tables.items():
        c = eng.counts(k, 256)
        med, best = timeit(lambda: eng.hist_u8_cols(t, range(k), counts=c, stream=stream), stream)
        c.zero(stream); eng.hist_u8_cols(t, range(k), counts=c, stream=stream)
        got = c.to_numpy(stream)
        res.append({"mode": mode, "table": name, "ms_med": med, "ms_best": best, "GBs_med": t.nrows * k / med / 1e6,
                    "checksum": int((got.astype(np.uint64) * (np.arange(got.size, dtype=np.uint64).reshape(got.shape) + np.uint64(1))).sum() & np.uint64(0xFFFFFFFFFFFF))})
    print("RESULT " + json.dumps(res), flush=True)


def main():
    modes = [int(a) for a in sys.argv[1:]] or [4, 5, 6, 7, 8, 9, 10, 11, 12]
    allres, ref = [], {}
    for mode in modes:
        out = subprocess.run([sys.executable, __file__, "--child", str(mode)], capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(f"mode {mode} FAILED: {out.stderr.strip().splitlines()[-1] if out.stderr.strip() else out.returncode}", flush=True)
            continue
        for r in json.loads(line[-1][7:]):
            ref.setdefault(r["table"], r["checksum"])
            r["same_counts_as_first_mode"] = r["checksum"] == ref[r["table"]]
            allres.append(r)
            print(f"mode {r['mode']} {r['table']:22s} med {r['ms_med']:7.4f} ms best {r['ms_best']:7.4f}  {r['GBs_med']:7.1f} GB/s  same={r['same_counts_as_first_mode']}", flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "u8_sweep.json").write_text(json.dumps(allres, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    else:
        main()
