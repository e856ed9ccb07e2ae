"""Oracle-made goldens for the three bench workloads (bench.py compares EVERY run with them, at every N).

    python tests/golden/make_bench_goldens.py        # ~ a minute on 8 cores; writes tests/golden/bench_goldens.json

For each workload the streaming C oracle (oracle/bsem.c: oracle_synth_project_cast_hist / oracle_synth_hist_u8)
regenerates the counter-based synthetic table row by row — nothing is materialised — and records

  s100  100 000 000 x 32 fp64, K = 32: the 32 x 256 uint64 count matrix and the 32 position-weighted checksums
        of the fp32 output slabs (sum over rows of bits(x[r]) * (2 r + 1) mod 2^64, r = GLOBAL row number, so the
        checksums of row shards simply add up);
  s10   10 000 000 x 16 fp64, K = 16: the 16 output checksums;
  m     1 000 000 x 784 uint8: the 784 x 256 count matrix.

Counts are stored as flat lists; checksums as decimal strings (they exceed 2^53).
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402  (SEED, ranges and the projected-column permutation are the bench's own)
from learningorchestra_b200.build import build_oracle  # noqa: E402

build_oracle()
from oracle import cport  # noqa: E402

cport.use_all_cores()
out = {}
for name in ("s100", "s10"):
    W = bench.WORKLOADS[name]
    cols = bench.workload_columns(name, W["cols"])
    k = len(cols)
    counts, sums = cport.synth_project_cast_hist(0, bench.SEED, 0, W["rows"], bench.GEN_LO, bench.GEN_HI, cols, bench.NBINS,
                                                 [bench.GEN_LO] * k, [bench.GEN_HI] * k)
    assert int(counts.sum()) == W["rows"] * k
    out[name] = {"rows": W["rows"], "cols": W["cols"], "seed": bench.SEED, "k": k, "nbins": bench.NBINS,
                 "projected_columns": cols, "checksums": [str(int(x)) for x in sums]}
    if name == "s100":
        out[name]["counts"] = [int(x) for x in counts.reshape(-1)]
W = bench.WORKLOADS["m"]
c8 = cport.synth_hist_u8(bench.SEED, 0, W["rows"], list(range(W["cols"])))
assert int(c8.sum()) == W["rows"] * W["cols"]
out["m"] = {"rows": W["rows"], "cols": W["cols"], "seed": bench.SEED, "k": W["cols"], "nbins": 256,
            "counts": [int(x) for x in c8.reshape(-1)]}
(ROOT / "tests" / "golden" / "bench_goldens.json").write_text(json.dumps(out, separators=(",", ":")) + "\n")
print("written", {k: list(v) for k, v in out.items()})
