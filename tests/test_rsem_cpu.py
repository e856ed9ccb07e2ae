"""CPU: R-semantics.  (1) the oracle's restatement (oracle/rsem.py) equals what the REFERENCE'S OWN files
produced (golden fixtures made by tests/golden/make_golden.py); (2) the product's host logic — validators,
routes, the text<->number adapter, the plain select — equals the same goldens; (3) without a GPU the jobs
that need one fail loudly and leave ``finished: False``."""
import json
import math
from pathlib import Path

import pytest
from werkzeug.test import Client

from learningorchestra_b200 import server, utils
from learningorchestra_b200.data_type_update import DataType
from learningorchestra_b200.histogram import Histogram
from learningorchestra_b200.projection import Projection
from oracle import rsem

GOLD = Path(__file__).resolve().parent / "golden"


def _load(name):
    return json.loads((GOLD / name).read_text())


def _titanic_db(db):
    g = _load("titanic_shaped_input.json")
    headers, docs = rsem.csv_rows_to_documents(g["headers"], g["rows"])
    db.insert_one_in_file("titanic", rsem.dataset_metadata("titanic", headers))
    for d in docs:
        db.insert_one_in_file("titanic", d)
    return headers, docs


def _dec(v):
    if isinstance(v, dict) and "float" in v:
        return float(v["float"])
    if isinstance(v, dict) and "int" in v:
        return int(v["int"])
    return v


def _same(a, b):
    if isinstance(a, float) and isinstance(b, float) and math.isnan(a) and math.isnan(b):
        return True
    return type(a) is type(b) and a == b


# ---- (1) oracle restatement vs the reference's own execution --------------------------------------------
def test_titanic_generator_is_frozen():
    g = _load("titanic_shaped_input.json")
    assert g["rows"] == rsem.titanic_shaped_rows() and g["headers"] == rsem.TITANIC_HEADERS


def test_oracle_cast_matches_reference_execution():
    vec = _load("reference_cast_vectors.json")
    for s, want_num, want_str in zip(vec["in"], vec["number"], vec["back_to_string"]):
        changed, got = rsem.convert_value(s, "number")
        assert _same(got, _dec(want_num)), (s, got, want_num)
        _, back = rsem.convert_value(got, "string")
        assert back == want_str
    with pytest.raises(ValueError):
        rsem.convert_value("abc", "number")
    with pytest.raises(ValueError):
        rsem.convert_value("0x10", "number")
    gold = _load("reference_datatype_number.json")
    _, docs = rsem.csv_rows_to_documents(rsem.TITANIC_HEADERS, rsem.titanic_shaped_rows())
    for f in gold["fields"]:
        rsem.convert_field(docs, f, "number")
    got = [[d["_id"]] + [d[f] for f in gold["fields"]] for d in docs]
    assert len(got) == len(gold["rows"]) == 891
    for a, b in zip(got, gold["rows"]):
        assert all(_same(x, y) for x, y in zip(a, b)), (a, b)


def test_oracle_group_counts_match_reference_execution():
    gold = _load("reference_histogram.json")
    db = rsem.MemoryDatabase()
    _titanic_db(db)
    docs = db.collections["titanic"]
    for f in ("Survived", "Pclass", "Age", "Fare"):
        rsem.convert_field(docs, f, "number")
    ours = rsem.histogram_documents(docs, gold["fields"])
    ref_docs = [d for d in gold["documents"] if d["_id"] != 0]
    assert [d["_id"] for d in ours] == [d["_id"] for d in ref_docs] == [1, 2, 3, 4, 5]
    for mine, ref, f in zip(ours, ref_docs, gold["fields"]):
        assert rsem.normalise_group_result(mine[f]) == rsem.normalise_group_result(ref[f])
    # the metadata document inflates the null group: Age has 176 blanks + 1
    age = {rsem.group_key(g["_id"]): g["count"] for g in ref_docs[2]["Age"]}
    assert age[("null",)] == 177
    assert sum(age.values()) == 892


def test_oracle_select_is_keyed_by_id():
    _, docs = rsem.csv_rows_to_documents(rsem.TITANIC_HEADERS, rsem.titanic_shaped_rows())
    docs = [rsem.dataset_metadata("t", rsem.TITANIC_HEADERS)] + docs
    out = rsem.select_projection(docs, ["Fare", "Age"])
    assert len(out) == 891 and list(out[0]) == ["Fare", "Age", "_id"] and out[0]["_id"] == 1


# ---- (2) product host logic vs the goldens ----------------------------------------------------------------
def test_product_datatype_string_branch_matches_reference_execution():
    """number -> string is host-side text formatting (a5); checked from the reference's own converted rows."""
    db = utils.Database()
    gold_n, gold_s = _load("reference_datatype_number.json"), _load("reference_datatype_string.json")
    db.insert_one_in_file("t", rsem.dataset_metadata("t", gold_n["fields"]))
    db.insert_many_in_file("t", [dict(zip(["_id"] + gold_n["fields"], row)) for row in gold_n["rows"]])
    job = DataType(db, utils.DataTypeMetadata(db))
    job.convert_existent_file("t", {"Age": "string", "Survived": "string"})
    job.wait()
    assert db.find_one("t", {"_id": 0})["finished"] is True
    got = sorted([d["_id"], d["Age"], d["Survived"]] for d in db.find("t", {}) if d["_id"] != 0)
    assert got == gold_s["rows"]


def test_product_datatype_number_needs_the_gpu():
    db = utils.Database()
    db.insert_one_in_file("t", rsem.dataset_metadata("t", ["v"]))
    db.insert_one_in_file("t", {"_id": 1, "v": "12"})
    job = DataType(db, utils.DataTypeMetadata(db))          # no engine
    job.convert_existent_file("t", {"v": "number"})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        job.wait()
    meta = db.find_one("t", {"_id": 0})
    assert meta["finished"] is False and "no CPU fallback" in meta["exception"]
    assert db.find_one("t", {"_id": 1})["v"] == "12"       # untouched


def test_product_projection_select_matches_oracle():
    db = utils.Database()
    _, docs = _titanic_db(db)
    job = Projection(utils.ProjectionMetadata(db), None)
    fields = ["Survived", "Pclass", "Age", "Fare"]
    job.create("titanic", "proj", fields, utils.Database.collection_database_url("mongodb://x", "database", "titanic", "rs"),
               utils.Database.collection_database_url("mongodb://x", "database", "proj", "rs"))
    job.wait()
    assert fields[-1] == "_id"                      # projection.py:42 mutates the caller's list
    meta = db.find_one("proj", {"_id": 0})
    assert meta["finished"] is True and meta["fields"] == ["Survived", "Pclass", "Age", "Fare"]
    assert meta["type"] == "transform/projection" and meta["parentDatasetName"] == "titanic"
    got = {d["_id"]: d for d in db.find("proj", {}) if d["_id"] != 0}
    exp = {d["_id"]: d for d in rsem.select_projection(db.find("titanic", {}), ["Survived", "Pclass", "Age", "Fare"])}
    assert got == exp and len(got) == 891


def _client(db, engine=None):
    return Client(server.create_app(db, engine, synchronous=True))


def test_rest_validation_messages_and_codes():
    db = utils.Database()
    _titanic_db(db)
    c = _client(db)
    r = c.post("/projections", json={"inputDatasetName": "titanic", "outputDatasetName": "titanic", "names": ["Age"]})
    assert (r.status_code, r.get_json()) == (409, {"result": "duplicated projection name"})
    r = c.post("/projections", json={"inputDatasetName": "nope", "outputDatasetName": "p", "names": ["Age"]})
    assert (r.status_code, r.get_json()) == (406, {"result": "invalid dataset name"})
    r = c.post("/projections", json={"inputDatasetName": "titanic", "outputDatasetName": "p", "names": []})
    assert (r.status_code, r.get_json()) == (406, {"result": "missing fields"})
    r = c.post("/projections", json={"inputDatasetName": "titanic", "outputDatasetName": "p", "names": ["Nope"]})
    assert (r.status_code, r.get_json()) == (406, {"result": "invalid fields"})
    db.update_one("titanic", {"finished": False}, {"_id": 0})
    r = c.post("/projections", json={"inputDatasetName": "titanic", "outputDatasetName": "p", "names": ["Age"]})
    assert (r.status_code, r.get_json()) == (406, {"result": "unfinished processing in input dataset"})
    r = c.patch("/fieldTypes", json={"inputDatasetName": "titanic", "types": {"Age": "number"}})
    assert (r.status_code, r.get_json()) == (406, {"result": "unfinished processing in input dataset"})
    db.update_one("titanic", {"finished": True}, {"_id": 0})
    r = c.patch("/fieldTypes", json={"inputDatasetName": "titanic", "types": {"Age": "integer"}})
    assert (r.status_code, r.get_json()) == (406, {"result": "invalid fields"})
    r = c.patch("/fieldTypes", json={"inputDatasetName": "titanic", "types": {}})
    assert (r.status_code, r.get_json()) == (406, {"result": "missing fields"})
    r = c.post("/histograms", json={"inputDatasetName": "titanic", "outputDatasetName": "titanic", "names": ["Age"]})
    assert (r.status_code, r.get_json()) == (409, {"result": "duplicated dataset name"})
    r = c.post("/histograms", json={"inputDatasetName": "titanic", "outputDatasetName": "h", "names": ["Nope"]})
    assert (r.status_code, r.get_json()) == (406, {"result": "invalid fields"})


def test_rest_success_bodies_and_reader():
    db = utils.Database()
    _titanic_db(db)
    c = _client(db)
    r = c.post("/projections", json={"inputDatasetName": "titanic", "outputDatasetName": "p1", "names": ["Age", "Fare"]})
    assert r.status_code == 201
    assert r.get_json() == {"result": "/api/learningOrchestra/v1/transform/projection/p1?query={}&limit=20&skip=0"}
    r = c.patch("/fieldTypes", json={"inputDatasetName": "p1", "types": {"Age": "string", "Fare": "string"}})
    assert r.status_code == 200
    assert r.get_json() == {"result": "/api/learningOrchestra/v1/dataset/p1?query={}&limit=20&skip=0"}
    page = c.get("/files/p1?skip=0&limit=3&query={}").get_json()["result"]
    assert [d["_id"] for d in page] == [0, 1, 2] and page[0]["finished"] is True
    assert isinstance(page[1]["Fare"], str)
    assert len(c.get("/files/p1?limit=1000").get_json()["result"]) == 100       # limit capped at 100
    # POST /histograms answers 201 at once (async protocol); with no GPU the job itself must fail loudly
    r = c.post("/histograms", json={"inputDatasetName": "p1", "outputDatasetName": "h1", "names": ["Age"]})
    assert r.status_code == 201
    assert r.get_json() == {"result": "/api/learningOrchestra/v1/explore/histogram/h1?query={}&limit=10&skip=0"}
    meta = db.find_one("h1", {"_id": 0})
    assert meta["type"] == "explore/histogram" and meta["finished"] is False and "no CPU fallback" in meta["exception"]


# ---- (3) no GPU, no result ---------------------------------------------------------------------------------
def test_gpu_jobs_fail_loudly_without_an_engine():
    db = utils.Database()
    _titanic_db(db)
    h = Histogram(db, utils.HistogramMetadata(db), engine=None)
    h.create_file("titanic", "hh", ["Sex"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        h.wait()
    assert db.find_one("hh", {"_id": 0})["finished"] is False
    assert [d for d in db.find("hh", {}) if d["_id"] != 0] == []


def test_builder_front_end_drops_metadata_columns():
    """builder_image/builder.py:172-194: load -> filter(_id != 0) -> drop(metadata fields)."""
    from learningorchestra_b200.builder_frontend import METADATA_FIELDS, file_processor
    db = utils.Database()
    _titanic_db(db)
    gold = _load("reference_datatype_number.json")
    db.update_by_id("titanic", {row[0]: dict(zip(gold["fields"], row[1:])) for row in gold["rows"]})
    t = file_processor(db, "titanic")
    assert t.num_rows == 891 and not set(t.column_names) & set(METADATA_FIELDS)
    assert t.column_names == [f for f in rsem.TITANIC_HEADERS if f not in METADATA_FIELDS]
    import pyarrow as pa
    assert t.schema.field("Survived").type == pa.int64() and t.schema.field("Fare").type == pa.float64()
    assert t.schema.field("Name").type == pa.string()
    ages = t.column("Age").to_pylist()
    assert ages == [row[3] for row in gold["rows"]] and ages.count(None) == 176
