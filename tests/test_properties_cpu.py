"""CPU property tests (hypothesis): host logic and the parser over machine-generated corner cases."""
import math
import struct

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from learningorchestra_b200 import columnar
from learningorchestra_b200.sharding import all_shard_bounds
from oracle import bsem_numpy as bn
from oracle import cport, rsem
from test_parse_cpu import EMPTY, FLOAT, INTEGER, INVALID, parse  # noqa: F401  (fixture)

ALPHABET = "0123456789+-.eE_ \tinfatyINFATY"


@pytest.fixture(scope="module", autouse=True)
def _native_built(built):
    return built


@settings(max_examples=3000, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(cells=st.lists(st.text(alphabet=ALPHABET, min_size=0, max_size=14), min_size=1, max_size=40))
def test_parser_agrees_with_float_on_arbitrary_grammar_soup(parse, cells):
    bits, status = parse(cells)
    for s, b, t in zip(cells, bits, status):
        if s == "":
            assert t == EMPTY
            continue
        try:
            v = float(s)
        except ValueError:
            assert t == INVALID, repr(s)
            continue
        assert t == (INTEGER if (math.isfinite(v) and v.is_integer()) else FLOAT), repr(s)
        if not math.isnan(v):
            assert int(b) == struct.unpack("<Q", struct.pack("<d", v))[0], repr(s)


@settings(max_examples=300, deadline=None)
@given(total=st.integers(0, 10 ** 10), world=st.integers(1, 16))
def test_shard_bounds_partition(total, world):
    b = all_shard_bounds(total, world)
    assert b[0][0] == 0 and b[-1][1] == total
    assert all(a1 == b0 and a0 <= a1 for (a0, a1), (b0, _b1) in zip(b[:-1], b[1:]))
    assert all(x0 % 32 == 0 for x0, _ in b[1:])
    if total >= 32 * world:
        sizes = [e - s for s, e in b]
        assert max(sizes) - min(sizes) <= 65      # every interior cut is rounded down by < 32 rows


_values = st.one_of(st.none(), st.booleans(), st.integers(-2 ** 60, 2 ** 60), st.floats(allow_nan=True, allow_infinity=True),
                    st.text(alphabet="ab1 ", max_size=3))


@settings(max_examples=500, deadline=None)
@given(values=st.lists(_values, max_size=60))
def test_product_group_key_is_the_oracle_group_key(values):
    """The adapter's MongoDB-equality key (product) and the oracle's restatement agree, and dictionary codes
    reproduce the oracle's $group counts."""
    assert [columnar.group_key(v) for v in values] == [rsem.group_key(v) for v in values]
    codes, reps = columnar.dictionary_encode(values)
    counts = np.bincount(codes, minlength=len(reps)) if len(values) else []
    got = [{"_id": r, "count": int(c)} for r, c in zip(reps, counts)]
    exp = rsem.group_counts([{"f": v} for v in values], "f")
    assert rsem.normalise_group_result(got) == rsem.normalise_group_result(exp)


@settings(max_examples=200, deadline=None)
@given(x=st.lists(st.floats(allow_nan=True, allow_infinity=True, width=64), min_size=1, max_size=200),
       lo=st.floats(-1000000.0, 1000000.0, width=32), span=st.floats(0.0009765625, 1000000.0, width=32), nbins=st.integers(1, 256))
def test_c_and_numpy_oracles_agree_on_arbitrary_inputs(x, lo, span, nbins):
    hi = float(np.float32(lo) + np.float32(span))
    if not hi > lo:
        return
    a = np.array(x, dtype=np.float64)
    f_c, f_n = cport.cast_f64_f32(a), bn.cast_f64_f32(a)
    assert np.array_equal(f_c.view(np.uint32), f_n.view(np.uint32))
    w = bn.bin_width(lo, hi, nbins)
    if not (np.isfinite(w) and w > 0):
        return
    assert np.array_equal(cport.hist_f32(f_c, lo, hi, nbins), bn.hist_f32(f_n, lo, hi, nbins))
