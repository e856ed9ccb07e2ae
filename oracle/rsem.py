"""CPU ORACLE (test infrastructure, NOT product code) — pure-Python restatement of the REFERENCE
("R") semantics of the three hot-path jobs over list-of-dict collections, plus the in-memory
``MemoryDatabase`` that lets the reference's OWN ``data_type_update.py`` / ``histogram.py`` run here
(tests/golden/make_golden.py).  Only ``tests/`` and ``__graft_entry__.smoke()`` may import this.

Restated sites (under /root/reference/microservices):
  * ``projection_image/projection.py:35-46``  load -> filter(_id != 0) -> select(*fields, _id) -> append
    (the arithmetic lives in Spark 2.4.7 + mongo-spark-connector 2.4.2, neither present: restated).
  * ``data_type_handler_image/data_type_update.py:15-45``  per-document cast incl. its dead checks.
  * ``histogram_image/histogram.py:25-44`` + MongoDB 3.6 ``$group``/``$sum:1`` (restated, see
    ``group_key``): numbers compare by value across int/long/double, null and missing fall in one
    group (so the metadata document inflates it by one), NaN groups with NaN, strings bytewise,
    booleans are not numbers.  Output order of ``$group`` is unspecified: compare as multisets.
  * document format ``database_api_image/database.py:110-151``: every CSV cell a ``str``, ``_id`` 1..N,
    metadata document ``_id: 0``.

PARITY: data_type_update.py and histogram.py are pinned by executing the reference files themselves
(golden fixtures under tests/golden/); Spark's select and mongod's $group are third-party and absent,
so those two restatements are UNPINNED assumptions, written down here.
"""
from __future__ import annotations

import copy
import math
import re
from collections import OrderedDict

METADATA_ID = 0


# ---- the table format (database_api_image/database.py:110-151) ------------------------------------
def sanitize_headers(headers):
    return [re.sub(r"\W+", "", h) for h in headers]          # database.py:118-119


def csv_rows_to_documents(headers, rows):
    """database.py:124-137: zip header -> cell (all ``str``), ``_id`` = 1, 2, ..."""
    headers = sanitize_headers(headers)
    docs = []
    for i, row in enumerate(rows, start=1):
        d = {headers[j]: row[j] for j in range(len(headers))}
        d["_id"] = i
        docs.append(d)
    return headers, docs


def dataset_metadata(name, fields, finished=True, type_="dataset/csv", url="file://synthetic"):
    return {"datasetName": name, "url": url, "timeCreated": "2026-09-21T00:00:00-00:00", "_id": METADATA_ID,
            "finished": finished, "fields": list(fields), "type": type_}


# ---- projection (projection.py:38-43) ---------------------------------------------------------------
def select_projection(documents, fields):
    """rows with ``_id != 0``; output keys = requested fields then ``_id``; a missing key reads as None
    (Spark null).  Row identity is ``_id`` (insertion order across partitions is not guaranteed)."""
    out = []
    for d in documents:
        if d.get("_id") == METADATA_ID:
            continue
        row = OrderedDict((f, d.get(f)) for f in fields)
        row["_id"] = d["_id"]
        out.append(dict(row))
    return out


# ---- cast (data_type_update.py:15-45) ---------------------------------------------------------------
def convert_value(value, field_type):
    """Returns (changed, new_value) for one stored value; raises ValueError like ``float()`` does."""
    if field_type == "string":
        if value == str:                      # :23 compares to the type object: always False
            return False, value
        return True, ("" if value is None else str(value))
    if field_type == "number":
        if value == int or value == float or value is None:     # :32-36 (first two never true)
            return False, value
        if value == "":
            return True, None
        v = float(value)
        if v.is_integer():
            v = int(v)
        return True, v
    return False, value                        # unknown type: `values` stays {} -> $set {} (no-op here)


def convert_field(documents, field, field_type):
    """In place, like the reference (it updates the INPUT collection)."""
    for d in documents:
        if d.get("_id") == METADATA_ID:
            continue
        changed, v = convert_value(d[field], field_type)
        if changed:
            d[field] = v
    return documents


# ---- $group / $sum:1 (histogram.py:31-36; MongoDB 3.6.17 semantics, restated) ----------------------
def group_key(value):
    """Canonical key under MongoDB's grouping equality."""
    if value is None:
        return ("null",)
    if isinstance(value, bool):
        return ("bool", value)
    if isinstance(value, (int, float)):
        if isinstance(value, float) and math.isnan(value):
            return ("num", "nan")
        if isinstance(value, float) and math.isinf(value):
            return ("num", "inf" if value > 0 else "-inf")
        if isinstance(value, float) and value.is_integer():
            return ("num", int(value))
        return ("num", value)                  # -0.0 == 0.0 -> both hit ("num", 0) above
    if isinstance(value, str):
        return ("str", value)
    return ("other", repr(value))


def group_counts(documents, field):
    """``[{"_id": value, "count": n}, ...]`` over EVERY document incl. the metadata one (no filter in
    the pipeline): documents lacking the field count under null."""
    groups = OrderedDict()
    for d in documents:
        v = d.get(field)
        k = group_key(v)
        if k not in groups:
            groups[k] = [v, 0]
        groups[k][1] += 1
    return [{"_id": v, "count": n} for v, n in groups.values()]


def normalise_group_result(result):
    """Order-free, type-normalised form for comparing two ``$group`` outputs."""
    return sorted(((group_key(g["_id"]), g["count"]) for g in result), key=repr)


def histogram_documents(documents, fields):
    """histogram.py:25-44: one result document per field, ``_id`` = 1.. in request order."""
    return [{f: group_counts(documents, f), "_id": i} for i, f in enumerate(fields, start=1)]


# ---- in-memory stand-in for the reference's pymongo ``Database`` wrappers ----------------------------
def _matches(doc, query):
    for k, v in query.items():
        if k not in doc:
            if v is None:
                continue
            return False
        if group_key(doc[k]) != group_key(v) and doc[k] != v:
            return False
    return True


class MemoryDatabase:
    """The union of the ``Database`` interfaces in ``*/utils.py`` (find, find_one, aggregate,
    insert_one_in_file, update_one, get_filenames) over ``{collection: [documents]}``."""

    def __init__(self):
        self.collections = OrderedDict()

    def get_filenames(self):
        return list(self.collections)

    def find(self, filename, query):
        return [copy.deepcopy(d) for d in self.collections.get(filename, []) if _matches(d, query)]

    def find_one(self, filename, query):
        for d in self.collections.get(filename, []):
            if _matches(d, query):
                return copy.deepcopy(d)
        return None

    def insert_one_in_file(self, filename, json_object):
        self.collections.setdefault(filename, []).append(copy.deepcopy(json_object))

    def update_one(self, filename, new_value, query):
        for d in self.collections.get(filename, []):
            if _matches(d, query):
                d.update(copy.deepcopy(new_value))
                return

    def aggregate(self, filename, pipeline):
        assert len(pipeline) == 1 and list(pipeline[0]) == ["$group"], "only the reference's $group pipeline"
        spec = pipeline[0]["$group"]
        assert spec["count"] == {"$sum": 1} and spec["_id"].startswith("$")
        return group_counts(self.collections.get(filename, []), spec["_id"][1:])


# ---- Titanic-shaped fixture (the real file is not available offline) ----------------
TITANIC_HEADERS = ["PassengerId", "Survived", "Pclass", "Name", "Sex", "Age", "SibSp", "Parch", "Ticket", "Fare",
                   "Cabin", "Embarked"]


def titanic_shaped_rows(n=891, seed=20260921):
    """Deterministic 891 x 12 all-string table with the real file's quirks: blanks in Age / Cabin /
    Embarked, fractional ages like "0.42", integer-valued floats like "7.0", quoted names with commas."""
    state = seed & 0xFFFFFFFFFFFFFFFF

    def rnd():
        nonlocal state
        state = (state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        return (state >> 33) / float(1 << 31)

    rows = []
    for i in range(1, n + 1):
        survived = "1" if rnd() < 0.38 else "0"
        pclass = "1" if rnd() < 0.24 else ("2" if rnd() < 0.27 else "3")
        sex = "female" if rnd() < 0.35 else "male"
        r = rnd()
        if r < 0.2:
            age = ""
        elif r < 0.23:
            age = ["0.42", "0.67", "0.75", "0.83", "0.92"][int(rnd() * 5)]
        elif r < 0.3:
            age = f"{int(rnd() * 70) + 1}.5"
        else:
            age = str(int(rnd() * 79) + 1)
        fare_cents = int(rnd() * 51233)
        fare = f"{fare_cents / 100:.4f}".rstrip("0").rstrip(".") if rnd() < 0.9 else str(int(rnd() * 30) + 7) + ".0"
        if fare == "":
            fare = "0"
        cabin = "" if rnd() < 0.77 else "ABCDEFG"[int(rnd() * 7)] + str(int(rnd() * 120) + 1)
        embarked = "" if rnd() < 0.003 else ("S" if rnd() < 0.72 else ("C" if rnd() < 0.67 else "Q"))
        name = f"Surname{i % 97}, {'Mrs.' if sex == 'female' else 'Mr.'} Given{i}"
        rows.append([str(i), survived, pclass, name, sex, age, str(int(rnd() * 4)), str(int(rnd() * 3)),
                     f"T{int(rnd() * 99999)}", fare, cabin, embarked])
    return rows
