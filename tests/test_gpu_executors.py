"""GPU: the reference-shaped executors (Projection / DataType / Histogram) and REST routes, with the
counting / casting on the device, against (a) golden fixtures produced by the reference's own
histogram.py / data_type_update.py and (b) the oracle."""
import json
from pathlib import Path

import numpy as np
import pytest
from werkzeug.test import Client

from learningorchestra_b200 import server, utils
from learningorchestra_b200.data_type_update import DataType
from learningorchestra_b200.histogram import Histogram
from learningorchestra_b200.projection import Projection
from oracle import bsem_numpy as bn
from oracle import rsem

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _load(name):
    return json.loads((GOLD / name).read_text())


def _titanic_db():
    db = utils.Database()
    g = _load("titanic_shaped_input.json")
    headers, docs = rsem.csv_rows_to_documents(g["headers"], g["rows"])
    db.insert_one_in_file("titanic", rsem.dataset_metadata("titanic", headers))
    db.insert_many_in_file("titanic", docs)
    return db


def _url(name):
    return utils.Database.collection_database_url("mongodb://x", "database", name, "rs")


def test_titanic_histogram_value_counts_match_reference_execution(engine):
    """Config T, R-semantics: string -> number cast, then per-field value counts ($group) on the GPU."""
    db = _titanic_db()
    job = DataType(db, utils.DataTypeMetadata(db), engine)
    job.convert_existent_file("titanic", {"Survived": "number", "Pclass": "number", "Age": "number", "Fare": "number"})
    job.wait()
    gold = _load("reference_histogram.json")
    h = Histogram(db, utils.HistogramMetadata(db), engine)
    h.create_file("titanic", "titanic_hist", list(gold["fields"]))
    h.wait()
    got = sorted(db.find("titanic_hist", {}), key=lambda d: d["_id"])
    ref = sorted(gold["documents"], key=lambda d: d["_id"])
    assert got[0]["finished"] is True and got[0]["type"] == "explore/histogram" and got[0]["fields"] == gold["fields"]
    assert [d["_id"] for d in got] == [d["_id"] for d in ref]
    for mine, theirs, f in zip(got[1:], ref[1:], gold["fields"]):
        assert set(mine) == {f, "_id"}
        assert rsem.normalise_group_result(mine[f]) == rsem.normalise_group_result(theirs[f]), f


def test_number_cast_on_gpu_matches_reference_execution(engine):
    """PATCH /fieldTypes "number": every text cell parsed by k_parse_number == the reference's float()+collapse."""
    import math
    db = _titanic_db()
    gold = _load("reference_datatype_number.json")
    job = DataType(db, utils.DataTypeMetadata(db), engine)
    job.convert_existent_file("titanic", {f: "number" for f in gold["fields"]})
    job.wait()
    assert db.find_one("titanic", {"_id": 0})["finished"] is True
    got = sorted(([d["_id"]] + [d[f] for f in gold["fields"]] for d in db.find("titanic", {}) if d["_id"] != 0))
    assert len(got) == 891
    for a, b in zip(got, gold["rows"]):
        assert all(type(x) is type(y) and x == y for x, y in zip(a, b)), (a, b)
    # the per-value vectors of SURVEY.md §8c, produced by the reference's own converter
    vec = _load("reference_cast_vectors.json")
    vdb = utils.Database()
    vdb.insert_one_in_file("vec", rsem.dataset_metadata("vec", ["v"]))
    vdb.insert_many_in_file("vec", [{"_id": i + 1, "v": v} for i, v in enumerate(vec["in"])])
    job = DataType(vdb, utils.DataTypeMetadata(vdb), engine)
    job.convert_existent_file("vec", {"v": "number"})
    job.wait()
    out = [d["v"] for d in sorted(vdb.find("vec", {}), key=lambda d: d["_id"]) if d["_id"] != 0]
    for s_in, g, want in zip(vec["in"], out, vec["number"]):
        if isinstance(want, dict) and "float" in want:
            w = float(want["float"])
            assert isinstance(g, float) and (g == w or (math.isnan(g) and math.isnan(w))), (s_in, g, want)
        elif isinstance(want, dict):
            assert isinstance(g, int) and g == int(want["int"]), (s_in, g, want)
        else:
            assert g is None and want is None
    # an unparsable cell: earlier documents converted, job fails, finished stays False (reference behaviour)
    bdb = utils.Database()
    bdb.insert_one_in_file("b", rsem.dataset_metadata("b", ["v"]))
    bdb.insert_many_in_file("b", [{"_id": 1, "v": "1.5"}, {"_id": 2, "v": "abc"}, {"_id": 3, "v": "2"}])
    job = DataType(bdb, utils.DataTypeMetadata(bdb), engine)
    job.convert_existent_file("b", {"v": "number"})
    with pytest.raises(ValueError, match="could not convert string to float: 'abc'"):
        job.wait()
    assert [d["v"] for d in sorted(bdb.find("b", {}), key=lambda d: d["_id"]) if d["_id"]] == [1.5, "abc", "2"]
    assert bdb.find_one("b", {"_id": 0})["finished"] is False


def test_gpu_parser_equals_python_float_on_random_text(engine):
    import random
    import struct
    rng = random.Random(99)
    cells = []
    for _ in range(200_000):
        k = rng.random()
        if k < 0.4:
            cells.append(repr(struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64) & 0x7FEFFFFFFFFFFFFF))[0]))
        elif k < 0.6:
            cells.append(f"{rng.uniform(-1e4, 1e4):.{rng.randint(0, 12)}f}")
        elif k < 0.8:
            cells.append(str(rng.randint(-10 ** 12, 10 ** 12)))
        elif k < 0.9:
            d = "".join(rng.choice("0123456789") for _ in range(rng.randint(20, 40)))
            cells.append(d[:7] + "." + d[7:] + f"e{rng.randint(-300, 280)}")
        else:
            cells.append(rng.choice(["", " 5 ", "1_000", "nan", "-inf", "abc", "1e", "0x10", "１２", "+.5", "1.e3", "-0.0", "1e400",
                                     "٣.٥", "\u2003 7\u00a0", "1é", "½", "0." + "0" * 1500 + "25", "9" * 1200, "1" * 1100 + "e-1100"]))
    vals, st = engine.parse_number_host(cells)
    for c, v, t in zip(cells, vals, st):
        if c == "":
            assert t == 2
        else:       # Unicode digits / whitespace and > 1024-byte cells included: exactly what float(c) does
            try:
                w = float(c)
            except ValueError:
                assert t == 3, c
                continue
            assert t == (1 if (w == w and abs(w) != float("inf") and w.is_integer()) else 0), (c, t)
            assert (v == w and np.signbit(v) == np.signbit(w)) or (v != v and w != w), (c, v, w)


def test_byte_table_value_counts_match_reference_execution(engine):
    """Config M bridge: on byte columns the 256-bin histogram IS the reference's $group output."""
    gold = _load("reference_histogram_bytes.json")
    t = bn.synth_table_u8(gold["seed"], gold["ncols"], 0, gold["nrows"])
    names = [f"px{c}" for c in gold["cols"]]
    db = utils.Database()
    db.insert_one_in_file("bytes", rsem.dataset_metadata("bytes", names))
    db.insert_many_in_file("bytes", [dict({f"px{c}": int(t[c, r]) for c in gold["cols"]}, _id=r + 1)
                                     for r in range(gold["nrows"])])
    h = Histogram(db, utils.HistogramMetadata(db), engine)
    h.create_file("bytes", "bytes_hist", list(names))
    h.wait()
    got = sorted((d for d in db.find("bytes_hist", {}) if d["_id"] != 0), key=lambda d: d["_id"])
    ref = sorted((d for d in gold["documents"] if d["_id"] != 0), key=lambda d: d["_id"])
    for mine, theirs, f in zip(got, ref, names):
        assert rsem.normalise_group_result(mine[f]) == rsem.normalise_group_result(theirs[f])
    # and directly through the device-resident kernel: counts[v] for v in 0..255 (+1 null for the metadata doc)
    cols = [np.ascontiguousarray(t[c]) for c in gold["cols"]]
    counts, _ = engine.hist_u8_cols_host(cols)
    for j, (theirs, f) in enumerate(zip(ref, names)):
        want = {g["_id"]: g["count"] for g in theirs[f] if g["_id"] is not None}
        assert {v: int(n) for v, n in enumerate(counts[j]) if n} == want


def test_high_cardinality_field_uses_code_counting_kernel(engine):
    db = _titanic_db()          # PassengerId / Name / Ticket have > 256 distinct strings
    h = Histogram(db, utils.HistogramMetadata(db), engine)
    h.create_file("titanic", "hc", ["PassengerId", "Ticket", "Embarked"])
    h.wait()
    docs = db.find("titanic", {})
    got = sorted((d for d in db.find("hc", {}) if d["_id"] != 0), key=lambda d: d["_id"])
    for mine, f in zip(got, ["PassengerId", "Ticket", "Embarked"]):
        assert rsem.normalise_group_result(mine[f]) == rsem.normalise_group_result(rsem.group_counts(docs, f))
    rng = np.random.default_rng(1)
    codes = rng.integers(0, 70_000, 1_000_003, dtype=np.uint32)
    np.testing.assert_array_equal(engine.value_counts_u32_host(codes, 70_000), np.bincount(codes, minlength=70_000))
    from learningorchestra_b200._native import LoexecError
    with pytest.raises(LoexecError):
        engine.value_counts_u32_host(codes, 69_000)


def test_titanic_flow_project_cast_and_10_bin_histogram(engine):
    """Config T, B-semantics extension: project 4 columns -> fp32 cast -> 10-bin histogram, fused on the GPU."""
    db = _titanic_db()
    c = Client(server.create_app(db, engine, synchronous=True))
    fields = ["Survived", "Pclass", "Age", "Fare"]
    assert c.patch("/fieldTypes", json={"inputDatasetName": "titanic", "types": {f: "number" for f in fields}}).status_code == 200
    r = c.post("/projections", json={"inputDatasetName": "titanic", "outputDatasetName": "t4", "names": fields,
                                     "castTo": "float32", "bins": 10})
    assert r.status_code == 201
    meta = db.find_one("t4", {"_id": 0})
    assert meta["finished"] is True, meta
    rows = sorted((d for d in db.find("titanic", {}) if d["_id"] != 0), key=lambda d: d["_id"])
    out = {d["_id"]: d for d in db.find("t4", {}) if d["_id"] != 0}
    for j, f in enumerate(fields):
        vals = np.array([np.nan if d[f] is None else float(d[f]) for d in rows])
        f32 = bn.cast_f64_f32(vals)
        finite = f32[np.isfinite(f32)]
        lo, hi = finite.min(), finite.max()
        hist = meta["histogram"][f]
        assert hist["bins"] == 10 and hist["range"] == [float(lo), float(hi)]
        assert hist["counts"] == bn.hist_f32(f32, lo, hi, 10).tolist()
        for d, x, raw in zip(rows, f32, vals):
            got = out[d["_id"]][f]
            assert (got is None and np.isnan(raw)) or np.float32(got) == x
    # histogram endpoint with explicit range on the converted collection
    r = c.post("/histograms", json={"inputDatasetName": "titanic", "outputDatasetName": "h10", "names": ["Age", "Fare"],
                                    "bins": 10, "range": [0, 100]})
    assert r.status_code == 201 and db.find_one("h10", {"_id": 0})["finished"] is True
    for d, f in zip(sorted((d for d in db.find("h10", {}) if d["_id"] != 0), key=lambda d: d["_id"]), ["Age", "Fare"]):
        vals = np.array([np.nan if r_[f] is None else float(r_[f]) for r_ in rows])
        assert d[f]["counts"] == bn.hist_f32(bn.cast_f64_f32(vals), 0, 100, 10).tolist()


def test_datatype_float32_extension(engine):
    db = utils.Database()
    vals = [0.1, 16777217, 1e39, None, -1e-46, 7.25, 3]
    db.insert_one_in_file("v", rsem.dataset_metadata("v", ["x"]))
    db.insert_many_in_file("v", [{"_id": i + 1, "x": v} for i, v in enumerate(vals)])
    job = DataType(db, utils.DataTypeMetadata(db), engine)
    job.convert_existent_file("v", {"x": "float32"})
    job.wait()
    got = [d["x"] for d in sorted(db.find("v", {}), key=lambda d: d["_id"]) if d["_id"] != 0]
    exp = bn.cast_f64_f32(np.array([np.nan if v is None else float(v) for v in vals]))
    for g, e, v in zip(got, exp, vals):
        assert (v is None and g is None) or np.float32(g).view(np.uint32) == e.view(np.uint32)


def test_minmax_prepass(engine):
    table = bn.synth_table_f64(1, 20260921, 5, 0, 300_001)
    mn, mx, cnt = engine.minmax_cast_host([table[c] for c in range(5)])
    for c in range(5):
        f = bn.cast_f64_f32(table[c])
        fin = f[np.isfinite(f)]
        assert mn[c] == fin.min() and mx[c] == fin.max() and cnt[c] == fin.size


def test_gpu_hash_group_by_on_numeric_columns(engine):
    from collections import Counter
    rng = np.random.default_rng(11)
    cases = [
        rng.integers(0, 3, 500_003).astype(np.float64),                       # 3 hot keys (warp aggregation path)
        rng.integers(-50_000, 50_000, 700_001).astype(np.float64),           # many keys
        np.round(rng.normal(0, 10, 300_000), 2),
        np.array([0.0, -0.0, np.nan, float.fromhex("0x1.8p1"), np.nan, 1e300, -1e300, np.inf, -np.inf, 5e-324] * 1000),
        np.full(200_000, 7.25),
        np.arange(100_000, dtype=np.float64),                                 # all distinct (> initial capacity)
    ]
    for x in cases:
        keys, counts = engine.value_counts_f64_host(x)
        got = {rsem.group_key(float(k)): int(c) for k, c in zip(keys, counts)}
        exp = Counter(rsem.group_key(float(v)) for v in x)
        assert got == dict(exp)
        assert int(counts.sum()) == x.size and len(keys) == len(exp)


def test_resident_dataset_cache(engine):
    """Binned histograms run on the HBM-resident copy of the dataset: second request = cache hit, same counts;
    a write to the collection invalidates it."""
    engine.resident.clear()
    db = _titanic_db()
    job = DataType(db, utils.DataTypeMetadata(db), engine)
    job.convert_existent_file("titanic", {"Age": "number", "Fare": "number", "Pclass": "number"})
    job.wait()
    rows = sorted((d for d in db.find("titanic", {}) if d["_id"] != 0), key=lambda d: d["_id"])

    def expect(field, nb, lo, hi):
        vals = np.array([np.nan if r[field] is None else float(r[field]) for r in rows])
        return bn.hist_f32(bn.cast_f64_f32(vals), lo, hi, nb).tolist()

    h0, m0 = engine.resident.hits, engine.resident.misses
    for n, (fields, out) in enumerate([(["Age", "Fare"], "r1"), (["Fare"], "r2"), (["Age"], "r3")]):
        h = Histogram(db, utils.HistogramMetadata(db), engine)
        h.create_file("titanic", out, list(fields), bins=16, value_range=[0, 128])
        h.wait()
        docs = sorted((d for d in db.find(out, {}) if d["_id"] != 0), key=lambda d: d["_id"])
        for d, f in zip(docs, fields):
            assert d[f]["counts"] == expect(f, 16, 0, 128)
    assert engine.resident.misses == m0 + 1 and engine.resident.hits == h0 + 2
    # a field that was not resident yet extends the table (miss), keeping the old columns
    h = Histogram(db, utils.HistogramMetadata(db), engine)
    h.create_file("titanic", "r4", ["Pclass"], bins=3)             # range from the device min/max pre-pass
    h.wait()
    d = [x for x in db.find("r4", {}) if x["_id"] == 1][0]
    assert d["Pclass"]["range"] == [1.0, 3.0] and d["Pclass"]["counts"] == expect("Pclass", 3, 1.0, 3.0)
    assert engine.resident.misses == m0 + 2
    # write -> stale -> rebuilt, and the new values are what gets counted
    db.update_one("titanic", {"Age": 127.0}, {"_id": 1})
    rows[0]["Age"] = 127.0
    h = Histogram(db, utils.HistogramMetadata(db), engine)
    h.create_file("titanic", "r5", ["Age"], bins=16, value_range=[0, 128])
    h.wait()
    d = [x for x in db.find("r5", {}) if x["_id"] == 1][0]
    assert d["Age"]["counts"] == expect("Age", 16, 0, 128) and engine.resident.misses == m0 + 3
    engine.resident.clear()


def test_gpu_hash_group_by_on_text_columns(engine):
    from collections import Counter
    import random
    rng = random.Random(5)
    words = ["", " ", "a", "b", "ab", "ba", "male", "female", "S", "C", "Q", "é", "日本", "x" * 300] + [f"T{i}" for i in range(5000)]
    cases = [
        [rng.choice(words[:6]) for _ in range(200_003)],                      # few hot keys (warp aggregation)
        [rng.choice(words) for _ in range(300_001)],                          # thousands of keys, long and non-ASCII cells
        [f"Name{i}, Mr. Given{i % 977}" for i in range(150_000)],             # all distinct (> initial capacity)
        ["same"] * 100_000,
        [""] * 1000,
    ]
    for cells in cases:
        rep, counts = engine.value_counts_str_host(cells)
        got = {cells[int(r)]: int(c) for r, c in zip(rep, counts)}
        assert got == dict(Counter(cells)) and len(rep) == len(got)
    # mixed-type field still goes through the dictionary path and agrees with the oracle
    db = utils.Database()
    vals = ["1", 1, 1.0, None, "x", True, 2.5, "1"] * 50
    db.insert_one_in_file("m", rsem.dataset_metadata("m", ["v", "t"]))
    db.insert_many_in_file("m", [{"_id": i + 1, "v": v, "t": str(v)} for i, v in enumerate(vals)])
    h = Histogram(db, utils.HistogramMetadata(db), engine)
    h.create_file("m", "mh", ["v", "t"])
    h.wait()
    docs = db.find("m", {})
    for d, f in zip(sorted((x for x in db.find("mh", {}) if x["_id"]), key=lambda x: x["_id"]), ["v", "t"]):
        assert rsem.normalise_group_result(d[f]) == rsem.normalise_group_result(rsem.group_counts(docs, f))


def test_builder_front_end_serves_resident_columns(engine):
    from learningorchestra_b200.builder_frontend import file_processor
    engine.resident.clear()
    db = _titanic_db()
    job = DataType(db, utils.DataTypeMetadata(db), engine)
    job.convert_existent_file("titanic", {"Age": "number", "Fare": "number", "Survived": "number"})
    job.wait()
    host = file_processor(db, "titanic")
    dev = file_processor(db, "titanic", engine)
    assert host.equals(dev)
    assert {"Age", "Fare", "Survived"} <= set(engine.resident.ensure(db, "titanic", ["Age"]).fields)
    engine.resident.clear()


# ---- the columnar store on the real engine (tests/test_column_store_cpu.py runs the same flow on the stand-in) ----------
def test_columnar_store_titanic_flow_matches_reference_execution(engine, tmp_path):
    import csv
    import io
    from learningorchestra_b200.column_store import ColumnarDatabase, NumberColumn
    g = json.loads((GOLD / "titanic_shaped_input.json").read_text())
    buf = io.StringIO()
    w = csv.writer(buf, lineterminator="\n")
    w.writerow(g["headers"]); w.writerows(g["rows"])
    path = tmp_path / "titanic.csv"
    path.write_text(buf.getvalue())
    db = ColumnarDatabase()
    assert db.ingest_csv("titanic", str(path)) == 891
    gold = json.loads((GOLD / "reference_datatype_number.json").read_text())
    job = DataType(db, utils.DataTypeMetadata(db), engine=engine)
    job.convert_existent_file("titanic", {f: "number" for f in gold["fields"]})
    job.wait(120)
    assert db.find_one("titanic", {"_id": 0})["finished"] is True
    assert all(isinstance(db.column("titanic", f), NumberColumn) for f in gold["fields"])
    got = [[d["_id"]] + [d[f] for f in gold["fields"]] for d in db.find("titanic", {}) if d["_id"] != 0]
    for a, b in zip(got, gold["rows"]):
        assert all(type(x) is type(y) and x == y for x, y in zip(a, b)), (a, b)
    hg = json.loads((GOLD / "reference_histogram.json").read_text())
    hist = Histogram(db, utils.HistogramMetadata(db), engine=engine)
    hist.create_file("titanic", "titanic_hist", list(hg["fields"]))
    hist.wait(120)
    ours = [d for d in db.find("titanic_hist", {}) if d["_id"] != 0]
    ref = [d for d in hg["documents"] if d["_id"] != 0]
    for mine, theirs, f in zip(ours, ref, hg["fields"]):
        assert rsem.normalise_group_result(mine[f]) == rsem.normalise_group_result(theirs[f]), f
    # binned extension from the HBM-resident copy, constant and all-null columns included (range widened, not an error)
    n = db.nrows("titanic")
    db.set_column("titanic", "Const", NumberColumn(np.full(n, 7.0), np.ones(n, bool), np.ones(n, bool)))
    db.set_column("titanic", "Nulls", NumberColumn(np.full(n, np.nan), np.zeros(n, bool)))
    hist = Histogram(db, utils.HistogramMetadata(db), engine=engine)
    hist.create_file("titanic", "titanic_bins", ["Age", "Const", "Nulls"], bins=10)
    hist.wait(120)
    docs = {list(d)[0]: d[list(d)[0]] for d in db.find("titanic_bins", {}) if d["_id"] != 0}
    age = np.array([r[3] if r[3] is not None else np.nan for r in gold["rows"]], dtype=np.float64)
    lo, hi = bn.auto_range([np.nanmin(age).astype(np.float32)], [np.nanmax(age).astype(np.float32)], [int(np.isfinite(age).sum())])
    assert docs["Age"]["counts"] == bn.hist_f32(bn.cast_f64_f32(age), lo[0], hi[0], 10).tolist()
    assert docs["Const"]["range"] == [6.5, 7.5] and sum(docs["Const"]["counts"]) == n and docs["Const"]["counts"][5] == n
    assert docs["Nulls"]["range"] == [0.0, 1.0] and sum(docs["Nulls"]["counts"]) == 0
    # castTo projection through the resident slabs
    proj = Projection(utils.ProjectionMetadata(db), engine)
    proj.create("titanic", "titanic_f32", ["Fare", "Age"], "mongodb://h/database.titanic?r", "mongodb://h/database.titanic_f32?r", cast_to="float32")
    proj.wait(120)
    fare = np.array([r[4] for r in gold["rows"]], dtype=np.float64)
    gotf = [d["Fare"] for d in db.find("titanic_f32", {}) if d["_id"] != 0]
    assert gotf == [float(x) for x in bn.cast_f64_f32(fare)]


def test_columnar_store_large_text_column_parse(engine):
    """2 M cells straight from Arrow buffers to the GPU parser: values / int-collapse equal CPython's float() on a sample."""
    import pyarrow as pa
    import pyarrow.compute as pc
    from learningorchestra_b200.column_store import ColumnarDatabase, TextColumn
    rng = np.random.default_rng(5)
    n = 2_000_000
    vals = np.where(rng.random(n) < 0.5, np.round(rng.uniform(-1e4, 1e4, n), 3), rng.integers(-10 ** 9, 10 ** 9, n).astype(np.float64))
    text = pc.cast(pa.array(vals), pa.large_string())
    db = ColumnarDatabase()
    db.ingest_columns("big", {"x": TextColumn(text)})
    job = DataType(db, utils.DataTypeMetadata(db), engine=engine)
    job.convert_existent_file("big", {"x": "number"})
    job.wait(300)
    col = db.column("big", "x")
    cells = text.to_pylist()
    for i in rng.integers(0, n, 5000):
        w = float(cells[i])
        assert col.values[i] == w and bool(col.is_int[i]) == w.is_integer(), (cells[i], col.values[i])
    hist = Histogram(db, utils.HistogramMetadata(db), engine=engine)
    hist.create_file("big", "big_h", ["x"])
    hist.wait(300)
    groups = db.find_one("big_h", {"_id": 1})["x"]
    assert sum(g["count"] for g in groups) == n + 1            # + the metadata document under null


def test_builder_front_end_hands_the_frame_over_on_the_device(engine):
    """the load -> filter -> drop-metadata-columns frame as zero-copy views of the resident slabs
    (``__cuda_array_interface__``): a GPU consumer (torch here) reads them in place, no D2H."""
    import torch
    from learningorchestra_b200 import builder_frontend
    db = utils.Database()
    db.insert_one_in_file("d", {"_id": 0, "datasetName": "d", "finished": True, "fields": ["a", "b", "s"], "timeCreated": "t"})
    rng = np.random.default_rng(2)
    a = rng.normal(size=5000)
    for i in range(5000):
        db.insert_one_in_file("d", {"_id": i + 1, "a": float(a[i]), "b": (int(i) if i % 7 else None), "s": f"x{i}"})
    with builder_frontend.device_frame(db, "d", engine) as frame:
        assert frame.fields == ["a", "b"] and frame.nulls["b"] == len([i for i in range(5000) if i % 7 == 0])
        ta = torch.as_tensor(frame.columns["a"][0], device="cuda")
        tb = torch.as_tensor(frame.columns["b"][0], device="cuda")
        assert ta.data_ptr() == frame.columns["a"][0].data_ptr          # in place
        np.testing.assert_array_equal(ta.cpu().numpy(), a)
        got_b = tb.cpu().numpy()
        assert np.isnan(got_b[0]) and got_b[1] == 1.0 and int(np.isnan(got_b).sum()) == frame.nulls["b"]
        host = builder_frontend.file_processor(db, "d", engine)          # the second form: pyarrow.Table on the host
        assert host.column_names == ["a", "b", "s"] and host.num_rows == 5000


def test_rest_binned_histogram_shards_over_every_visible_gpu(built, tmp_path):
    """POST /histograms with ``bins`` through the REST surface with the engine the server itself opens
    (``sharding.open_engine``: every visible GPU; forced through the group path here even on a one-GPU box): the resident
    table is row-sharded over the group's members, the partial histograms are merged inside the kernels, counts == oracle."""
    import torch
    from learningorchestra_b200.column_store import ColumnarDatabase, NumberColumn
    from learningorchestra_b200.sharding import ShardedEngine
    rng = np.random.default_rng(11)
    n = 300_017
    cols = {"a": rng.normal(0, 50, n), "b": np.round(rng.uniform(-5, 5, n), 2), "c": rng.integers(0, 1000, n).astype(np.float64)}
    cols["b"][::13] = np.nan
    with ShardedEngine.local(list(range(torch.cuda.device_count()))) as eng:
        db = ColumnarDatabase()
        db.ingest_columns("d", {k: NumberColumn(v, ~np.isnan(v)) for k, v in cols.items()})
        c = Client(server.create_app(db, eng, synchronous=True))
        r = c.post("/histograms", json={"inputDatasetName": "d", "outputDatasetName": "h", "names": ["a", "b", "c"], "bins": 100})
        assert r.status_code == 201
        assert db.find_one("h", {"_id": 0})["finished"] is True, db.find_one("h", {"_id": 0})
        docs = {list(d)[0]: d[list(d)[0]] for d in db.find("h", {}) if d["_id"] != 0}
        for name, v in cols.items():
            f = bn.cast_f64_f32(v)
            fin = f[np.isfinite(f)]
            lo, hi = bn.auto_range([fin.min()], [fin.max()], [fin.size])
            assert docs[name]["range"] == [float(lo[0]), float(hi[0])]
            assert docs[name]["counts"] == bn.hist_f32(f, lo[0], hi[0], 100).tolist(), name
        # second request with an explicit range: served from the resident shards (cache hit), same group
        r = c.post("/histograms", json={"inputDatasetName": "d", "outputDatasetName": "h2", "names": ["c"], "bins": 10, "range": [0, 1000]})
        assert r.status_code == 201 and eng.resident.hits >= 1
        got = [d for d in db.find("h2", {}) if d["_id"] != 0][0]["c"]["counts"]
        assert got == bn.hist_f32(bn.cast_f64_f32(cols["c"]), np.float32(0), np.float32(1000), 10).tolist()
        # ``bins`` above the tile kernel's 256: same route, the chunk kernel and the same in-kernel merge
        r = c.post("/histograms", json={"inputDatasetName": "d", "outputDatasetName": "h3", "names": ["a", "c"], "bins": 2000,
                                        "range": [-200, 1000]})
        assert r.status_code == 201 and db.find_one("h3", {"_id": 0})["finished"] is True
        docs3 = {list(d)[0]: d[list(d)[0]] for d in db.find("h3", {}) if d["_id"] != 0}
        for name in ("a", "c"):
            assert docs3[name]["counts"] == bn.hist_f32(bn.cast_f64_f32(cols[name]), np.float32(-200), np.float32(1000), 2000).tolist()
        assert eng.timeouts() == 0


def test_download_waits_for_its_own_stream_only(engine):
    """``lo_table_download_col`` orders itself after the stream it is given and nothing else: while a long kernel runs
    on stream A, a download of another table on stream B returns long before that kernel finishes."""
    import threading
    import time
    import torch
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    big = engine.table("f64", 40_000_000, 16).fill_synthetic(0, 1, stream=sa)
    out = engine.table("f32", 40_000_000, 16)
    small = engine.table("f64", 1000, 1).fill_synthetic(0, 2, stream=sb)
    torch.cuda.synchronize()
    lo, hi = np.full(16, -1000, np.float32), np.full(16, 1000, np.float32)
    counts = engine.counts(16, 256)
    done = {}
    for _ in range(40):                                          # ~40 x 1.1 ms queued on stream A
        engine.project_cast_hist(big, range(16), 256, lo, hi, out=out, counts=counts, stream=sa)
    t0 = time.perf_counter()
    got = small.to_numpy(0, stream=sb)
    done["download"] = time.perf_counter() - t0
    engine.sync(sa)
    done["kernels"] = time.perf_counter() - t0
    np.testing.assert_array_equal(got, bn.synth_f64(0, 2, 0, 0, 1000))
    assert done["download"] < 0.5 * done["kernels"], done       # a device-wide sync would have waited for all 40
    for t in (big, out, small):
        t.free()
    counts.free()


def test_concurrent_host_calls_on_one_engine_overlap(engine):
    """``*_host`` calls no longer queue behind one staging set: four threads (the REST services run a thread per job)
    push different columns through one engine at once; every result is the oracle's."""
    import threading
    from oracle import cport
    n = 1_500_000
    jobs = []
    for t in range(4):
        cols = [cport.synth_f64(1, 77 + t, c, 0, n) for c in range(3)]
        outs = [np.empty(n, np.float32) for _ in cols]
        jobs.append((cols, outs))
    lo, hi = np.full(3, -1000.0, np.float32), np.full(3, 1000.0, np.float32)
    results, errors = [None] * 4, []

    def run(i):
        try:
            results[i] = engine.project_cast_hist_host(jobs[i][0], 128, lo, hi, out=jobs[i][1])[0]
        except Exception as exc:      # noqa: BLE001
            errors.append(exc)
    threads = [threading.Thread(target=run, args=(i,)) for i in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for i, (cols, outs) in enumerate(jobs):
        exp_out, exp_counts = cport.project_cast_hist(cols, 128, lo, hi)
        np.testing.assert_array_equal(results[i], exp_counts)
        for o, e in zip(outs, exp_out):
            np.testing.assert_array_equal(o.view(np.uint32), e.view(np.uint32))
