"""ctypes loader for the C oracle (``oracle/bsem.c`` -> ``oracle/_build/liboracle.so``).

CPU ORACLE — test infrastructure, NOT product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_LIB_PATH = Path(__file__).resolve().parent / "_build" / "liboracle.so"
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(f"{_LIB_PATH} missing: run __graft_entry__.build() (gcc oracle/bsem.c)")
        L = C.CDLL(str(_LIB_PATH))
        L.oracle_bin_width.restype = C.c_float
        L.oracle_bin_width.argtypes = [C.c_float, C.c_float, C.c_int]
        L.oracle_checksum.restype = C.c_uint64
        L.oracle_checksum.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64]
        L.oracle_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _ptr_array(arrs, ctype):
    return (C.POINTER(ctype) * len(arrs))(*[a.ctypes.data_as(C.POINTER(ctype)) for a in arrs])


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def use_all_cores() -> int:
    """Let OpenMP use every core this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oracle_set_num_threads(C.c_int(n))
    return num_threads()


def cast_f64_f32(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(x.shape, dtype=np.float32)
    lib().oracle_cast_f64_f32(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_int64(x.size))
    return out


def hist_f32(x: np.ndarray, lo: float, hi: float, nbins: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    counts = np.zeros(nbins, dtype=np.uint64)
    lib().oracle_hist_f32(x.ctypes.data_as(C.c_void_p), C.c_int64(x.size), C.c_int(nbins), C.c_float(lo),
                          C.c_float(hi), counts.ctypes.data_as(C.c_void_p))
    return counts


def project_cast_hist(cols, nbins: int = 0, lo=None, hi=None, want_out: bool = True):
    """cols: list of k contiguous float64 arrays (already projected).  Returns (out [k][n] f32 | None, counts [k, nbins] | None)."""
    cols = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
    k, n = len(cols), cols[0].shape[0]
    outs = [np.empty(n, dtype=np.float32) for _ in range(k)] if want_out else None
    counts = np.zeros((k, nbins), dtype=np.uint64) if nbins else np.zeros((k, 1), dtype=np.uint64)
    lo_a = np.ascontiguousarray(lo if lo is not None else np.zeros(k), dtype=np.float32)
    hi_a = np.ascontiguousarray(hi if hi is not None else np.ones(k), dtype=np.float32)
    lib().oracle_project_cast_hist(
        _ptr_array(cols, C.c_double), C.c_int64(n), C.c_int(k),
        _ptr_array(outs, C.c_float) if want_out else None,
        C.c_int(nbins), lo_a.ctypes.data_as(C.c_void_p), hi_a.ctypes.data_as(C.c_void_p),
        counts.ctypes.data_as(C.c_void_p))
    return outs, (counts if nbins else None)


def hist_u8_cols(cols) -> np.ndarray:
    cols = [np.ascontiguousarray(c, dtype=np.uint8) for c in cols]
    k, n = len(cols), cols[0].shape[0]
    counts = np.zeros((k, 256), dtype=np.uint64)
    lib().oracle_hist_u8_cols(_ptr_array(cols, C.c_uint8), C.c_int64(n), C.c_int(k), counts.ctypes.data_as(C.c_void_p))
    return counts


def checksum(col: np.ndarray, row_offset: int = 0) -> int:
    col = np.ascontiguousarray(col)
    return int(lib().oracle_checksum(col.ctypes.data_as(C.c_void_p), col.dtype.itemsize, col.shape[0], row_offset))


def synth_f64(kind: int, seed: int, col: int, row0: int, n: int, lo: float = -1000.0, hi: float = 1000.0) -> np.ndarray:
    out = np.empty(n, dtype=np.float64)
    lib().oracle_synth_f64(C.c_int(kind), C.c_uint64(seed), C.c_int(col), C.c_int64(row0), C.c_int64(n),
                           C.c_double(lo), C.c_double(hi), out.ctypes.data_as(C.c_void_p))
    return out


def synth_u8(seed: int, col: int, row0: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    lib().oracle_synth_u8(C.c_uint64(seed), C.c_int(col), C.c_int64(row0), C.c_int64(n), out.ctypes.data_as(C.c_void_p))
    return out


def synth_project_cast_hist(kind: int, seed: int, row0: int, nrows: int, glo: float, ghi: float, col_idx, nbins: int,
                            lo, hi):
    """Streaming (nothing materialised): counts [k, nbins] and fp32-output checksums [k] for global rows [row0, row0+nrows)."""
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
    k = col_idx.shape[0]
    lo_a = np.ascontiguousarray(lo, dtype=np.float32)
    hi_a = np.ascontiguousarray(hi, dtype=np.float32)
    counts = np.zeros((k, nbins), dtype=np.uint64)
    sums = np.zeros(k, dtype=np.uint64)
    lib().oracle_synth_project_cast_hist(
        C.c_int(kind), C.c_uint64(seed), C.c_int64(row0), C.c_int64(nrows), C.c_double(glo), C.c_double(ghi),
        col_idx.ctypes.data_as(C.c_void_p), C.c_int(k), C.c_int(nbins), lo_a.ctypes.data_as(C.c_void_p),
        hi_a.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), sums.ctypes.data_as(C.c_void_p))
    return counts, sums


def synth_hist_u8(seed: int, row0: int, nrows: int, col_idx) -> np.ndarray:
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
    counts = np.zeros((col_idx.shape[0], 256), dtype=np.uint64)
    lib().oracle_synth_hist_u8(C.c_uint64(seed), C.c_int64(row0), C.c_int64(nrows), col_idx.ctypes.data_as(C.c_void_p),
                               C.c_int(col_idx.shape[0]), counts.ctypes.data_as(C.c_void_p))
    return counts
