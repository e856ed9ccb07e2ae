"""CPU: the columnar document store (``column_store.ColumnarDatabase``) behind the reference's wrapper API.

The same jobs — CSV ingest, PATCH /fieldTypes, POST /histograms, POST /projections, the paged reader — run on a
collection stored as columns must leave exactly the documents the reference's own files produced (golden fixtures made
by executing them, tests/golden/make_golden.py) and exactly what the document-backed ``Database`` leaves.  The GPU
arithmetic is played by the oracle stand-in (tests/oracle_engine.py); tests/test_gpu_executors.py runs the same
paths on the real engine."""
import io
import json
import math
from pathlib import Path

import numpy as np
import pytest

from learningorchestra_b200 import utils
from learningorchestra_b200.column_store import ColumnarDatabase, NumberColumn, ObjectColumn, TextColumn, column_from_values
from learningorchestra_b200.data_type_update import DataType
from learningorchestra_b200.histogram import Histogram
from learningorchestra_b200.projection import Projection
from oracle import rsem
from oracle_engine import OracleEngine

GOLD = Path(__file__).resolve().parent / "golden"


def _load(name):
    return json.loads((GOLD / name).read_text())


def _same(a, b):
    if isinstance(a, float) and isinstance(b, float) and math.isnan(a) and math.isnan(b):
        return True
    return type(a) is type(b) and a == b


def _csv_text(headers, rows):
    import csv
    buf = io.StringIO()
    w = csv.writer(buf, lineterminator="\n")
    w.writerow(headers)
    w.writerows(rows)
    return buf.getvalue()


def _titanic(tmp_path):
    g = _load("titanic_shaped_input.json")
    path = tmp_path / "titanic.csv"
    path.write_text(_csv_text(g["headers"], g["rows"]))
    db = ColumnarDatabase()
    n = db.ingest_csv("titanic", str(path), url="file://titanic.csv")
    return db, g, n


def test_csv_ingest_produces_the_reference_document_format(tmp_path):
    db, g, n = _titanic(tmp_path)
    headers, docs = rsem.csv_rows_to_documents(g["headers"], g["rows"])       # the oracle's restatement of database.py:110-137
    assert n == len(docs) == 891 and db.has_columns("titanic")
    meta = db.find_one("titanic", {"_id": 0})
    assert meta["finished"] is True and meta["fields"] == headers and meta["datasetName"] == "titanic"
    got = [d for d in db.find("titanic", {}) if d["_id"] != 0]
    assert got == docs                                                        # every cell a str, _id from 1, same keys
    assert all(isinstance(db.column("titanic", h), TextColumn) for h in headers)
    # header sanitising: re.sub(r"\W+", "", name)
    p = tmp_path / "odd.csv"
    p.write_text('a b,c-d!,"e,f"\n1,2,3\n')
    db.ingest_csv("odd", str(p))
    assert db.find_one("odd", {"_id": 0})["fields"] == ["ab", "cd", "ef"]
    assert db.find_one("odd", {"_id": 1}) == {"ab": "1", "cd": "2", "ef": "3", "_id": 1}


def test_paged_reader_only_materialises_the_page(tmp_path):
    db, g, _ = _titanic(tmp_path)
    plain = utils.Database()
    for d in db.find("titanic", {}):
        plain.insert_one_in_file("titanic", d)
    for skip, limit in [(0, 20), (0, 1), (1, 5), (885, 20), (891, 10), (892, 3), (5000, 5)]:
        assert db.find_in_file("titanic", {}, skip, limit) == plain.find_in_file("titanic", {}, skip, limit)
    assert db.find_in_file("titanic", {"Sex": "male"}, 3, 4) == plain.find_in_file("titanic", {"Sex": "male"}, 3, 4)
    assert db.find_one("titanic", {"_id": 77}) == plain.find_one("titanic", {"_id": 77})
    assert db.find("titanic", {"_id": 5000}) == []


def test_number_cast_histogram_and_string_cast_match_the_reference_execution(tmp_path):
    db, g, _ = _titanic(tmp_path)
    eng = OracleEngine()
    gold = _load("reference_datatype_number.json")
    job = DataType(db, utils.DataTypeMetadata(db), engine=eng)
    job.convert_existent_file("titanic", {f: "number" for f in gold["fields"]})
    job.wait(60)
    assert db.find_one("titanic", {"_id": 0})["finished"] is True and db.has_columns("titanic")
    assert all(isinstance(db.column("titanic", f), NumberColumn) for f in gold["fields"])
    got = [[d["_id"]] + [d[f] for f in gold["fields"]] for d in db.find("titanic", {}) if d["_id"] != 0]
    assert len(got) == len(gold["rows"])
    for a, b in zip(got, gold["rows"]):
        assert all(_same(x, y) for x, y in zip(a, b)), (a, b)

    hg = _load("reference_histogram.json")
    hist = Histogram(db, utils.HistogramMetadata(db), engine=eng)
    hist.create_file("titanic", "titanic_hist", list(hg["fields"]))
    hist.wait(60)
    ours = [d for d in db.find("titanic_hist", {}) if d["_id"] != 0]
    ref = [d for d in hg["documents"] if d["_id"] != 0]
    assert [d["_id"] for d in ours] == [d["_id"] for d in ref]
    for mine, theirs, f in zip(ours, ref, hg["fields"]):
        assert rsem.normalise_group_result(mine[f]) == rsem.normalise_group_result(theirs[f]), f

    sg = _load("reference_datatype_string.json")
    job = DataType(db, utils.DataTypeMetadata(db), engine=eng)
    job.convert_existent_file("titanic", {f: "string" for f in sg["fields"]})
    job.wait(60)
    got = [[d["_id"]] + [d[f] for f in sg["fields"]] for d in db.find("titanic", {}) if d["_id"] != 0]
    assert got == sg["rows"]


def test_reference_cast_vectors_through_a_text_column():
    vec = _load("reference_cast_vectors.json")
    import pyarrow as pa
    cells = [v for v in vec["in"]]
    db = ColumnarDatabase()
    db.ingest_columns("vec", {"v": TextColumn(pa.array(cells, type=pa.large_string()))})
    job = DataType(db, utils.DataTypeMetadata(db), engine=OracleEngine())
    job.convert_existent_file("vec", {"v": "number"})
    job.wait(60)
    got = [d["v"] for d in db.find("vec", {}) if d["_id"] != 0]

    def dec(v):
        return float(v["float"]) if isinstance(v, dict) and "float" in v else int(v["int"]) if isinstance(v, dict) else v
    for s, a, b in zip(cells, got, vec["number"]):
        assert _same(a, dec(b)), (s, a, b)          # Unicode digits, > 1 KiB cells, "" -> None, None stays None


def test_bad_cell_leaves_the_collection_as_the_reference_leaves_it():
    import pyarrow as pa
    db = ColumnarDatabase()
    db.ingest_columns("t", {"x": TextColumn(pa.array(["1", "2.5", "abc", "4"], type=pa.large_string()))})
    job = DataType(db, utils.DataTypeMetadata(db), engine=OracleEngine())
    job.convert_existent_file("t", {"x": "number"})
    with pytest.raises(ValueError, match="could not convert string to float: 'abc'"):
        job.wait(60)
    meta = db.find_one("t", {"_id": 0})
    assert meta["finished"] is False and "ValueError" in meta["exception"]
    assert [d["x"] for d in db.find("t", {}) if d["_id"] != 0] == [1, 2.5, "abc", "4"]     # converted up to the bad row


def test_projection_shares_columns_and_row_writes_fall_back_to_documents(tmp_path):
    db, g, _ = _titanic(tmp_path)
    proj = Projection(utils.ProjectionMetadata(db), None)
    fields = ["Name", "Age"]
    proj.create("titanic", "titanic_p", fields, "mongodb://x/database.titanic?r", "mongodb://x/database.titanic_p?r")
    proj.wait(60)
    assert fields == ["Name", "Age", "_id"]                                   # projection.py:42's mutation, kept
    assert db.find_one("titanic_p", {"_id": 0})["finished"] is True
    assert db.column("titanic_p", "Name") is db.column("titanic", "Name")     # zero copy
    want = rsem.select_projection([d for d in db.find("titanic", {})], ["Name", "Age"])
    assert [d for d in db.find("titanic_p", {}) if d["_id"] != 0] == want
    # a row-addressed write turns the collection back into documents, then behaves as the plain store
    db.update_one("titanic_p", {"Age": "99"}, {"_id": 3})
    assert not db.has_columns("titanic_p") and db.find_one("titanic_p", {"_id": 3})["Age"] == "99"
    assert len(db.find("titanic_p", {})) == 892


def test_column_classification():
    assert isinstance(column_from_values(["a", None, ""]), TextColumn)
    c = column_from_values([1, None, 2.5, 3.0])
    assert isinstance(c, NumberColumn) and c.to_pylist() == [1, None, 2.5, 3.0] and [type(v) for v in c.to_pylist()] == [int, type(None), float, float]
    assert isinstance(column_from_values([1, "a"]), ObjectColumn)
    assert isinstance(column_from_values([2 ** 60, 1]), ObjectColumn)
    assert isinstance(column_from_values([True, 1]), ObjectColumn) or column_from_values([True, 1]).to_pylist() == [True, 1]


def test_rest_ingest_route_then_the_three_services(tmp_path):
    """POST /files -> PATCH /fieldTypes -> POST /histograms -> GET /files: the Titanic-shaped flow over HTTP, keys /
    codes / messages of database_api_image/server.py:19-49 and utils.py:78-95."""
    from werkzeug.test import Client
    from learningorchestra_b200 import server
    g = _load("titanic_shaped_input.json")
    path = tmp_path / "t.csv"
    path.write_text(_csv_text(g["headers"], g["rows"]))
    app = server.create_app(None, OracleEngine(), synchronous=True)
    c = Client(app)
    r = c.post("/files", json={"datasetName": "titanic", "datasetURI": f"file://{path}"})
    assert r.status_code == 201 and r.get_json() == {"result": "/api/learningOrchestra/v1/dataset/titanic?query={}&limit=10&skip=0"}
    assert c.post("/files", json={"datasetName": "titanic", "datasetURI": f"file://{path}"}).get_json() == {"result": "duplicated dataset name"}
    r = c.post("/files", json={"datasetName": "nope", "datasetURI": "file:///does/not/exist.csv"})
    assert r.status_code == 406 and r.get_json() == {"result": "invalid url"}
    assert c.patch("/fieldTypes", json={"inputDatasetName": "titanic", "types": {"Age": "number", "Fare": "number"}}).status_code == 200
    assert c.post("/histograms", json={"inputDatasetName": "titanic", "outputDatasetName": "h", "names": ["Age", "Sex"]}).status_code == 201
    page = c.get("/files/titanic?limit=3&skip=0").get_json()["result"]
    assert [d["_id"] for d in page] == [0, 1, 2] and page[0]["finished"] is True and isinstance(page[1]["Fare"], (int, float))
    h = c.get("/files/h?limit=10").get_json()["result"]
    assert [d["_id"] for d in h] == [0, 1, 2] and h[0]["finished"] is True


def test_device_number_column_matches_the_plain_formulation():
    """The packing of a stored number column for the HBM-resident copy (no copy when nothing is null, kind without
    materialising the valid subset) gives exactly what np.where / boolean indexing give."""
    from learningorchestra_b200.table_cache import device_number_column
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 1000):
        for null_rate in (0.0, 0.3, 1.0):
            for ints in (True, False):
                values = np.round(rng.uniform(-50, 50, n)) if ints else rng.uniform(-50, 50, n)
                valid = rng.random(n) >= null_rate
                is_int = (values == np.round(values)) & valid
                slab, v, kind = device_number_column(values, valid, is_int)
                want = np.where(valid, values, np.nan)
                assert v is valid and np.array_equal(np.isnan(slab[~valid]), np.ones(int((~valid).sum()), bool))
                assert np.array_equal(slab[valid], want[valid])
                assert kind == ("int" if is_int[valid].all() else "float")
                if valid.all():
                    assert slab is values          # handed over without a copy
