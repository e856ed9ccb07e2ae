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


# ---- executors' adapter logic on machine-generated documents (oracle-backed test double for the Engine) ------------
_cell = st.one_of(st.none(), st.integers(-50, 50), st.floats(-5, 5, allow_nan=False).map(lambda x: round(x, 1)),
                  st.sampled_from(["", "a", "b", "1", "1.0", " x"]), st.booleans())
_column_kinds = st.sampled_from(["ints", "floats", "text", "mixed", "sparse"])


def _column(kind, n, draw):
    if kind == "ints":
        return draw(st.lists(st.one_of(st.none(), st.integers(-3, 3)), min_size=n, max_size=n))
    if kind == "floats":
        return draw(st.lists(st.one_of(st.none(), st.sampled_from([0.0, -0.0, 1.5, 2.0, float("nan"), 1e300, -7.25])), min_size=n, max_size=n))
    if kind == "text":
        return draw(st.lists(st.one_of(st.none(), st.sampled_from(["", "a", "b", "ab", "é", "1"])), min_size=n, max_size=n))
    if kind == "sparse":
        return draw(st.lists(st.sampled_from([None, "MISSING", 1, "x"]), min_size=n, max_size=n))
    return draw(st.lists(_cell, min_size=n, max_size=n))


@st.composite
def _collections(draw):
    n = draw(st.integers(0, 30))
    kinds = draw(st.lists(_column_kinds, min_size=1, max_size=4))
    cols = {f"f{i}": _column(kd, n, draw) for i, kd in enumerate(kinds)}
    docs = []
    for r in range(n):
        d = {"_id": r + 1}
        for f, vals in cols.items():
            if vals[r] != "MISSING":            # a missing key groups with null, like in MongoDB
                d[f] = vals[r]
        docs.append(d)
    return list(cols), docs


@settings(max_examples=300, deadline=None)
@given(data=_collections())
def test_histogram_executor_equals_group_semantics_on_arbitrary_documents(data):
    from learningorchestra_b200 import utils
    from learningorchestra_b200.histogram import Histogram
    from oracle_engine import OracleEngine
    fields, docs = data
    db = utils.Database()
    db.insert_one_in_file("d", rsem.dataset_metadata("d", fields))
    db.insert_many_in_file("d", docs)
    job = Histogram(db, utils.HistogramMetadata(db), OracleEngine())
    job.create_file("d", "h", list(fields))
    job.wait()
    meta = db.find_one("h", {"_id": 0})
    assert meta["finished"] is True and meta["fields"] == fields
    results = sorted((x for x in db.find("h", {}) if x["_id"] != 0), key=lambda x: x["_id"])
    everything = db.find("d", {})                # metadata document included, as in the reference's pipeline
    assert [x["_id"] for x in results] == list(range(1, len(fields) + 1))
    for res, f in zip(results, fields):
        assert rsem.normalise_group_result(res[f]) == rsem.normalise_group_result(rsem.group_counts(everything, f)), f


@settings(max_examples=200, deadline=None)
@given(cells=st.lists(st.one_of(st.none(), st.sampled_from(["", "1", "2.5", " 3 ", "1e2", "-0.0", "nan", "1_0"]),
                                st.integers(-5, 5), st.floats(-2, 2, allow_nan=False), st.booleans()), max_size=25),
       poison=st.one_of(st.none(), st.integers(0, 24)))
def test_datatype_number_equals_the_reference_loop_on_arbitrary_cells(cells, poison):
    """The adapter around the GPU parser keeps the reference's per-document semantics, including where it stops."""
    from learningorchestra_b200 import utils
    from learningorchestra_b200.data_type_update import DataType
    from oracle_engine import OracleEngine
    cells = list(cells)
    if poison is not None and poison < len(cells):
        cells[poison] = "abc"
    docs = [{"_id": i + 1, "v": c} for i, c in enumerate(cells)]
    db = utils.Database()
    db.insert_one_in_file("t", rsem.dataset_metadata("t", ["v"]))
    db.insert_many_in_file("t", docs)
    expected, failed = [dict(d) for d in docs], False
    for d in expected:                               # the reference's loop (oracle restatement), stopping at a ValueError
        try:
            changed, v = rsem.convert_value(d["v"], "number")
        except ValueError:
            failed = True
            break
        if changed:
            d["v"] = v
    job = DataType(db, utils.DataTypeMetadata(db), OracleEngine())
    job.convert_existent_file("t", {"v": "number"})
    if failed:
        with pytest.raises(ValueError):
            job.wait()
    else:
        job.wait()
    got = sorted((d for d in db.find("t", {}) if d["_id"] != 0), key=lambda d: d["_id"])
    for g, e in zip(got, expected):
        same = (g["v"] == e["v"] and type(g["v"]) is type(e["v"])) or (isinstance(g["v"], float) and g["v"] != g["v"] and e["v"] != e["v"])
        assert same, (g, e)
    assert db.find_one("t", {"_id": 0})["finished"] is (not failed)


@settings(max_examples=300, deadline=None)
@given(values=st.lists(st.one_of(st.none(), st.integers(-2 ** 62, 2 ** 62), st.integers(-5, 5), st.floats(allow_nan=True),
                                 st.booleans(), st.text(max_size=2)), min_size=64, max_size=120),
       homogeneous=st.sampled_from(["any", "ints", "floats", "text", "nulls"]))
def test_arrow_fast_path_packs_exactly_like_the_loop(values, homogeneous):
    if homogeneous == "ints":
        values = [v if (v is None or (isinstance(v, int) and not isinstance(v, bool))) else 3 for v in values]
    elif homogeneous == "floats":
        values = [v if (v is None or isinstance(v, float)) else (1.5 if not isinstance(v, int) or isinstance(v, bool) else v % 7) for v in values]
    elif homogeneous == "text":
        values = [v if (v is None or isinstance(v, str)) else "x" for v in values]
    elif homogeneous == "nulls":
        values = [None] * len(values)
    fast, slow = columnar.numeric_column(values), columnar._numeric_column_loop(values)
    assert (fast is None) == (slow is None)
    if fast is not None:
        assert fast[2] == slow[2] and np.array_equal(fast[1], slow[1])
        assert np.array_equal(fast[0].view(np.uint64)[fast[1]], slow[0].view(np.uint64)[slow[1]])
        assert np.isnan(fast[0][~fast[1]]).all()


@settings(max_examples=200, deadline=None)
@given(cells=st.lists(st.text(max_size=6), max_size=50), as_bytes=st.booleans())
def test_pack_cells_layout(cells, as_bytes):
    src = [c.encode("utf-8") for c in cells] if as_bytes else cells
    chars, offsets = columnar.pack_cells(src)
    enc = [c.encode("utf-8") for c in cells]
    assert offsets.dtype == np.int64 and offsets.shape == (len(cells) + 1,) and offsets[0] == 0
    for i, b in enumerate(enc):
        assert bytes(chars[offsets[i]:offsets[i + 1]]) == b
