"""``DataType`` — drop-in for ``data_type_handler_image/data_type_update.py`` (same constructor and
methods, in-place conversion of the input collection, same ``finished`` False -> True protocol).

* ``"string"`` / ``"number"`` (the reference's types): text <-> number conversion of Python objects —
  ``float(str)`` with the integer collapse, ``str(v)`` — is format conversion in the document adapter
  and runs on the host exactly as the reference's lines do (``data_type_update.py:22-43``, dead
  ``== str/int/float`` checks included).  Moving the decimal parse to the GPU is the first "next" row
  (SURVEY.md §8f rank 1); it is not wired in yet and nothing pretends otherwise.
* ``"float32"`` (this build's optional extension): the B-semantics cast of SURVEY.md §0 — numeric values
  go through the sm_100a kernel (fp64 -> fp32 round-to-nearest-even) and are stored back widened.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import columnar
from .utils import record_exception


class DataType:
    METADATA_DOCUMENT_ID = 0
    DOCUMENT_ID_NAME = "_id"
    STRING_TYPE = "string"
    NUMBER_TYPE = "number"
    FLOAT32_TYPE = "float32"

    def __init__(self, database_connector, metadata_handler, engine=None):
        self.database_connector = database_connector
        self.thread_pool = ThreadPoolExecutor()
        self.metadata_handler = metadata_handler
        self.engine = engine
        self.last_job = None

    def field_converter(self, filename, field, field_type):
        documents = columnar.data_rows(self.database_connector.find(filename, {}))
        updates = {}
        if field_type == self.FLOAT32_TYPE:
            updates = self.__gpu_float32(documents, field)
        else:
            for document in documents:
                values = {}
                if field_type == self.STRING_TYPE:
                    if document[field] == str:
                        continue
                    if document[field] is None:
                        values[field] = ""
                    else:
                        values[field] = str(document[field])
                elif field_type == self.NUMBER_TYPE:
                    if document[field] == int or document[field] == float or document[field] is None:
                        continue
                    if document[field] == "":
                        values[field] = None
                    else:
                        values[field] = float(document[field])
                        if values[field].is_integer():
                            values[field] = int(values[field])
                if values:
                    updates[document[self.DOCUMENT_ID_NAME]] = values
        self.database_connector.update_by_id(filename, updates)

    def __gpu_float32(self, documents, field):
        if self.engine is None:
            raise RuntimeError("type 'float32' needs an Engine (there is no CPU fallback)")
        packed = columnar.numeric_column([d.get(field) for d in documents])
        if packed is None:
            raise ValueError(f"field {field!r} is not numeric; convert it to 'number' first")
        col, valid, _kind = packed
        out = np.empty(len(documents), dtype=np.float32)
        self.engine.project_cast_hist_host([col], None, out=[out])
        return {d[self.DOCUMENT_ID_NAME]: {field: float(out[i])} for i, d in enumerate(documents) if valid[i]}

    def convert_existent_file(self, filename, fields_dictionary):
        self.metadata_handler.update_finished_flag(filename, False)
        self.last_job = self.thread_pool.submit(self.field_file_converter, filename, fields_dictionary)

    def field_file_converter(self, filename, fields_dictionary):
        try:
            for field in fields_dictionary:
                self.field_converter(filename, field, fields_dictionary[field])
            self.metadata_handler.update_finished_flag(filename, True)
        except BaseException as exc:
            record_exception(self.database_connector, filename, exc)
            raise

    def wait(self, timeout=None):
        if self.last_job is not None:
            self.last_job.result(timeout)
