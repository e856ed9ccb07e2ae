"""CPU: the two oracle restatements (C and numpy) agree with each other and with known answers."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import bsem_numpy as bn
from oracle import cport

GOLD = Path(__file__).resolve().parent / "golden"
SEED = 20260921


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    return built


def _bits(a):
    return a.view({4: np.uint32, 8: np.uint64, 1: np.uint8}[a.dtype.itemsize])


def test_known_answer_casts():
    """SURVEY.md §8c known answers (numpy-verified) + canonical NaN policy."""
    kat = json.loads((GOLD / "cast_f64_f32_kat.json").read_text())
    x = np.array([int(h, 16) for h in kat["in_f64_bits"]], dtype=np.uint64).view(np.float64)
    want = np.array([int(h, 16) for h in kat["out_f32_bits"]], dtype=np.uint32)
    np.testing.assert_array_equal(_bits(bn.cast_f64_f32(x)), want)
    np.testing.assert_array_equal(_bits(cport.cast_f64_f32(x)), want)


def test_known_answer_bins():
    """Hand-derived bin indices for the §8c formula (edges, closed last bin, skips)."""
    kat = json.loads((GOLD / "bin_kat.json").read_text())
    for case in kat["cases"]:
        x = np.array(case["x"], dtype=np.float32)
        idx = bn.bin_index_f32(x, case["lo"], case["hi"], case["nbins"])
        assert idx.tolist() == case["bin"], case
        h = cport.hist_f32(x, case["lo"], case["hi"], case["nbins"])
        exp = np.bincount(np.array([b for b in case["bin"] if b >= 0], dtype=np.int64), minlength=case["nbins"])
        np.testing.assert_array_equal(h, exp.astype(np.uint64))


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_generators_agree(kind):
    for col in (0, 1, 7, 31, 1008, 1009):
        for row0 in (0, 1009 * 20 - 5, 2 ** 33 + 17):
            a = cport.synth_f64(kind, SEED, col, row0, 5000)
            b = bn.synth_f64(kind, SEED, col, row0, 5000)
            np.testing.assert_array_equal(_bits(a), _bits(b))


def test_generator_golden_prefix():
    g = json.loads((GOLD / "synth_prefix.json").read_text())
    x = cport.synth_f64(1, g["seed"], g["col"], g["row0"], len(g["f64_bits"]), g["lo"], g["hi"])
    assert [f"{int(v):016x}" for v in _bits(x)] == g["f64_bits"]
    u = cport.synth_u8(g["seed"], g["u8_col"], g["row0"], len(g["u8"]))
    assert u.tolist() == g["u8"]


def test_u8_generator_agrees_and_is_mnist_shaped():
    cols = [0, 27, 4 * 28 + 4, 400, 783, 784 + 400]
    for c in cols:
        np.testing.assert_array_equal(cport.synth_u8(SEED, c, 12345, 20000), bn.synth_u8(SEED, c, 12345, 20000))
    t = bn.synth_table_u8(SEED, 784, 0, 2000)
    zeros = (t == 0).mean()
    assert 0.75 < zeros < 0.85
    assert (t[0] == 0).all() and (t[783] == 0).all()


def test_cast_and_hist_agree_between_c_and_numpy():
    x = np.concatenate([bn.synth_f64(1, SEED + c, c, 0, 60000) for c in range(4)])
    f_c, f_n = cport.cast_f64_f32(x), bn.cast_f64_f32(x)
    np.testing.assert_array_equal(_bits(f_c), _bits(f_n))
    for lo, hi, nb in [(-1000, 1000, 256), (-1000, 1000, 10), (0, 1, 7), (-0.5, 999.25, 255), (1e-30, 1e30, 3)]:
        np.testing.assert_array_equal(cport.hist_f32(f_c, lo, hi, nb), bn.hist_f32(f_n, lo, hi, nb))


def test_fused_c_oracle_matches_numpy_pipeline():
    table = bn.synth_table_f64(2, SEED, 6, 500, 70001)
    cols = [5, 0, 0, 3]
    lo = np.array([-1000, -10, 0, 100], np.float32)
    hi = np.array([1000, 10, 500, 101], np.float32)
    outs, counts = cport.project_cast_hist([table[c] for c in cols], 64, lo, hi)
    exp_out, exp_counts = bn.project_cast_hist(table, cols, 64, lo, hi)
    np.testing.assert_array_equal(counts, exp_counts)
    for j in range(4):
        np.testing.assert_array_equal(_bits(outs[j]), _bits(exp_out[j]))
    # streaming variant (regenerates rows, nothing materialised) and checksums
    c2, sums = cport.synth_project_cast_hist(2, SEED, 500, 70001, -1000.0, 1000.0, cols, 64, lo, hi)
    np.testing.assert_array_equal(c2, exp_counts)
    for j in range(4):
        assert int(sums[j]) == bn.checksum(exp_out[j], 500) == cport.checksum(exp_out[j], 500)


def test_u8_value_counts_are_group_counts():
    """For byte columns the 256-bin histogram IS $group/$sum:1 (histogram_image/histogram.py:31-36)."""
    from collections import Counter
    t = bn.synth_table_u8(SEED, 150, 0, 30011)
    got = cport.hist_u8_cols([t[c] for c in range(150)])
    np.testing.assert_array_equal(got, bn.hist_u8_cols(t, range(150)))
    np.testing.assert_array_equal(got, cport.synth_hist_u8(SEED, 0, 30011, list(range(150))))
    for c in (0, 4 * 28 + 5, 149):
        cnt = Counter(t[c].tolist())
        assert {v: int(n) for v, n in enumerate(got[c]) if n} == dict(cnt)


def test_empty_and_single_row():
    assert cport.cast_f64_f32(np.array([], dtype=np.float64)).shape == (0,)
    assert cport.hist_f32(np.array([], dtype=np.float32), 0, 1, 4).sum() == 0
    np.testing.assert_array_equal(bn.hist_f32(np.array([1.0], np.float32), 0, 1, 4), [0, 0, 0, 1])
    np.testing.assert_array_equal(cport.hist_f32(np.array([1.0], np.float32), 0, 1, 4), [0, 0, 0, 1])


def test_linearity_of_counts():
    x = bn.cast_f64_f32(bn.synth_f64(1, SEED, 3, 0, 100000))
    whole = cport.hist_f32(x, -1000, 1000, 256)
    parts = sum(cport.hist_f32(x[a:b], -1000, 1000, 256) for a, b in [(0, 1), (1, 33333), (33333, 100000)])
    np.testing.assert_array_equal(whole, parts)


def test_cast_has_a_third_independent_witness():
    """fp64 -> fp32 RNE as computed by torch's CPU kernels (another code base, another compiler) agrees bit for bit
    with both oracle restatements on the generator's special values and on random doubles (NaNs canonicalised)."""
    import torch
    rng = np.random.default_rng(17)
    x = np.concatenate([bn.special_values(-1000.0, 1000.0), rng.standard_normal(50_000) * 10.0 ** rng.integers(-45, 40, 50_000),
                        rng.integers(0, 2 ** 63, 50_000, dtype=np.uint64).view(np.float64)])
    t = torch.from_numpy(x.copy()).to(torch.float32).numpy()
    tb = t.view(np.uint32).copy()
    tb[np.isnan(t)] = 0x7FC00000
    np.testing.assert_array_equal(tb, _bits(bn.cast_f64_f32(x)))
    np.testing.assert_array_equal(tb, _bits(cport.cast_f64_f32(x)))


@pytest.mark.parametrize("nbins", [257, 1000, 16385, 65536])
def test_wide_histograms_agree_between_the_two_restatements(nbins):
    """Above 256 bins (the chunk kernel's territory on the GPU) the C and numpy restatements still agree bit for bit, the
    streaming variant included, and every in-range finite value lands in exactly one bin."""
    cols = [3, 0]
    lo = np.array([-1000.0, -3.5], dtype=np.float32); hi = np.array([7.25, 1000.0], dtype=np.float32)
    table = bn.synth_table_f64(1, SEED + 5, 4, 250, 40_001)
    _outs, counts = cport.project_cast_hist([table[c] for c in cols], nbins, lo, hi)
    exp_out, exp_counts = bn.project_cast_hist(table, cols, nbins, lo, hi)
    np.testing.assert_array_equal(counts, exp_counts)
    c2, _sums = cport.synth_project_cast_hist(1, SEED + 5, 250, 40_001, -1000.0, 1000.0, cols, nbins, lo, hi)
    np.testing.assert_array_equal(c2, exp_counts)
    for j in range(2):
        f = exp_out[j]
        assert int(counts[j].sum()) == int(((f >= lo[j]) & (f <= hi[j])).sum())
