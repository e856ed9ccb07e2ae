"""GPU parity at BASELINE.json's full sizes (configs S10, S100, M — SURVEY.md §8d).

Inputs are produced on the device by the counter-based generator (bit-identical to the oracle's,
see test_gpu_parity.test_device_generator_matches_oracle); the oracle regenerates the same rows on the
host cores and streams them through its C restatement, so nothing of size 25 GB crosses PCIe.
"""
import numpy as np
import pytest

from oracle import bsem_numpy as bn
from oracle import cport

pytestmark = pytest.mark.gpu
SEED = 20260921


def _bits(a):
    return a.view({4: np.uint32, 8: np.uint64, 1: np.uint8}[a.dtype.itemsize])


def test_s10_projection_cast_bit_exact(engine):
    """10M x 16 fp64 -> fp32, K = 16 (permutation) and K = 4: every output bit vs the oracle."""
    nrows, ncols = 10_000_000, 16
    t = engine.table("f64", nrows, ncols).fill_synthetic(1, SEED)
    for cols in ([(5 * j + 3) % ncols for j in range(ncols)], [15, 2, 2, 9]):
        out = engine.project_cast(t, cols)
        for j, c in enumerate(cols):
            exp = cport.cast_f64_f32(cport.synth_f64(1, SEED, c, 0, nrows))
            got = out.to_numpy(j)
            assert np.array_equal(_bits(got), _bits(exp)), f"column {j} (source {c}) differs"
            assert out.checksum(j) == cport.checksum(exp)
        out.free()
    t.free()


def test_s100_fused_counts_and_checksums(engine):
    """100M x 32, fused project + cast + 256-bin histogram: uint64 counts and the checksum of every fp32
    output slab equal the streaming oracle's; sampled windows are compared bit for bit; linearity holds."""
    nrows, ncols, nbins = 100_000_000, 32, 256
    cols = [(7 * j + 3) % ncols for j in range(ncols)]
    lo = np.full(ncols, -1000.0, np.float32)
    hi = np.full(ncols, 1000.0, np.float32)
    t = engine.table("f64", nrows, ncols).fill_synthetic(1, SEED)
    out = engine.table("f32", nrows, ncols)
    counts = engine.project_cast_hist(t, cols, nbins, lo, hi, out=out).to_numpy()
    exp_counts, exp_sums = cport.synth_project_cast_hist(1, SEED, 0, nrows, -1000.0, 1000.0, cols, nbins, lo, hi)
    assert np.array_equal(counts, exp_counts)
    for j in range(ncols):
        assert out.checksum(j) == int(exp_sums[j]), f"fp32 slab {j} checksum differs"
    # every row is either counted or one of the generator's NaN / out-of-range specials
    assert (counts.sum(axis=1) <= nrows).all() and (counts.sum(axis=1) >= nrows - nrows // 1009 - 1).all()
    # sampled windows, bit for bit (starts chosen to straddle tile boundaries and the ragged end)
    for r0 in (0, 61_440 - 100, 49_999_871, nrows - 70_001):
        for j in (0, 13, 31):
            exp = cport.cast_f64_f32(cport.synth_f64(1, SEED, cols[j], r0, 70_001))
            assert np.array_equal(_bits(out.to_numpy(j, r0, 70_001)), _bits(exp))
    # histogram only (no projected output) gives the same counts
    only = engine.project_cast_hist(t, cols, nbins, lo, hi).to_numpy()
    assert np.array_equal(only, exp_counts)
    out.free(); t.free()


def test_s100_row_shards_add_up(engine):
    """counts(whole table) == sum of counts over row-range shards generated independently — what the
    8-GPU run relies on (each rank fills rows [r0, r1) with row_offset = r0)."""
    nrows, ncols, nbins = 20_000_000, 32, 256
    cols = list(range(ncols))
    lo, hi = np.full(ncols, -1000.0, np.float32), np.full(ncols, 1000.0, np.float32)
    acc = engine.counts(ncols, nbins)
    bounds = [(nrows * r) // 8 for r in range(9)]
    for r0, r1 in zip(bounds[:-1], bounds[1:]):
        shard = engine.table("f64", r1 - r0, ncols).fill_synthetic(1, SEED, row_offset=r0)
        engine.project_cast_hist(shard, cols, nbins, lo, hi, counts=acc)
        shard.free()
    exp_counts, _ = cport.synth_project_cast_hist(1, SEED, 0, nrows, -1000.0, 1000.0, cols, nbins, lo, hi)
    assert np.array_equal(acc.to_numpy(), exp_counts)
    acc.free()


def test_mnist_shaped_u8_value_counts(engine):
    """1M x 784 uint8: per-column 256-bin counts == $group value counts of the oracle."""
    nrows, ncols = 1_000_000, 784
    t = engine.table("u8", nrows, ncols).fill_synthetic(3, SEED)
    got = engine.hist_u8_cols(t, range(ncols)).to_numpy()
    exp = cport.synth_hist_u8(SEED, 0, nrows, list(range(ncols)))
    assert np.array_equal(got, exp)
    assert (got.sum(axis=1) == nrows).all()
    assert got[0, 0] == nrows and got[783, 0] == nrows          # border pixels are constant 0
    t.free()
