"""Regenerates the known-answer fixtures in this directory (committed; the tests only read the JSON).

cast_f64_f32_kat.json  — fp64 -> fp32 round-to-nearest-even known answers, written as bit
                         patterns.  The expected outputs are LITERALS below (IEEE-754 facts), not computed.
bin_kat.json           — hand-derived bin indices for i = min((int)((x-lo)/w), nbins-1), w = (hi-lo)/nbins.
synth_prefix.json      — first values of the counter-based generator, frozen from oracle/bsem.c.
"""
import json
import struct
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))


def f64_bits(x: float) -> str:
    return f"{struct.unpack('<Q', struct.pack('<d', x))[0]:016x}"


CASTS = [  # (input double or raw bits, expected fp32 bits)
    (0.1, 0x3DCCCCCD), (16777217.0, 0x4B800000), (1e39, 0x7F800000), (-1e39, 0xFF800000), (-1e-46, 0x80000000),
    (1e-46, 0x00000000), (3.4028235677973366e38, 0x7F800000), (3.4028234663852886e38, 0x7F7FFFFF),
    (1e-40, 0x000116C2), (-0.0, 0x80000000), (0.0, 0x00000000), (1.0 + 2.0 ** -24, 0x3F800000),
    (1.0 + 3 * 2.0 ** -24, 0x3F800002), (1.0 + 2.0 ** -24 + 2.0 ** -52, 0x3F800001), (7.25, 0x40E80000),
    (float("inf"), 0x7F800000), (float("-inf"), 0xFF800000), (2.0 ** -149, 0x00000001), (2.0 ** -150, 0x00000000),
    (2.0 ** -150 + 2.0 ** -200, 0x00000001), (1000.0, 0x447A0000), (-1000.0, 0xC47A0000),
    ("7ff8000000000000", 0x7FC00000), ("fff4000000000001", 0x7FC00000), ("7ff0000000000001", 0x7FC00000),
]

BINS = [
    {"lo": 0.0, "hi": 10.0, "nbins": 10,
     "x": [0.0, 0.5, 1.0, 9.999999, 10.0, -0.0, 10.000001, -1e-7, 5.0, 2.9999998],
     "bin": [0, 0, 1, 9, 9, 0, -1, -1, 5, 2]},
    {"lo": -1000.0, "hi": 1000.0, "nbins": 256,   # w = 7.8125 exactly
     # 992.18744 is one fp32 ulp below the last edge, but (x - lo) rounds up to 1992.1875 in fp32, so it
     # lands in bin 255; 992.18738 (two ulps below) subtracts exactly and stays in bin 254
     "x": [-1000.0, 1000.0, -992.1875, -992.18756, 0.0, 7.8125, 999.99994, 992.1875, 992.18744, 992.18738],
     "bin": [0, 255, 1, 0, 128, 129, 255, 255, 255, 254]},
    {"lo": 0.0, "hi": 1.0, "nbins": 4, "x": [0.25, 0.24999999, 0.5, 0.75, 1.0, 2.0], "bin": [1, 0, 2, 3, 3, -1]},
    {"lo": 1.0, "hi": 2.0, "nbins": 1, "x": [1.0, 1.5, 2.0, 0.99999994, 2.0000002], "bin": [0, 0, 0, -1, -1]},
    {"lo": 0.0, "hi": 255.0, "nbins": 255, "x": [0.0, 1.0, 254.0, 254.99998, 255.0, 100.5], "bin": [0, 1, 254, 254, 254, 100]},
]


def main():
    (HERE / "cast_f64_f32_kat.json").write_text(json.dumps({
        "in_f64_bits": [c if isinstance(c, str) else f64_bits(c) for c, _ in CASTS],
        "out_f32_bits": [f"{o:08x}" for _, o in CASTS]}, indent=1))
    nan = float("nan")
    for case in BINS:
        case["x"] = list(case["x"])
    BINS[0]["x"].append(nan); BINS[0]["bin"].append(-1)
    js = json.dumps({"cases": BINS}, indent=1).replace("NaN", "NaN")
    (HERE / "bin_kat.json").write_text(js)
    from learningorchestra_b200.build import build_oracle
    build_oracle()
    from oracle import cport
    import numpy as np
    seed, col, row0, lo, hi = 20260921, 5, 1009 * 3 - 8, -1000.0, 1000.0
    x = cport.synth_f64(1, seed, col, row0, 24, lo, hi)
    u = cport.synth_u8(seed, 4 * 28 + 9, row0, 48)
    (HERE / "synth_prefix.json").write_text(json.dumps({
        "seed": seed, "col": col, "row0": row0, "lo": lo, "hi": hi,
        "f64_bits": [f"{int(v):016x}" for v in x.view(np.uint64)], "u8_col": 4 * 28 + 9, "u8": u.tolist()}, indent=1))


if __name__ == "__main__":
    main()
