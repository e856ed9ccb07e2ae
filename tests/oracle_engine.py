"""TEST DOUBLE: an object with the Engine methods the executors call, computing with the CPU ORACLE.

Lives under tests/ (only tests may touch the oracle) and exists so the adapters' logic — null / missing handling,
type routing, result documents, finished protocol — can be exercised on a CPU-only host with machine-generated
documents.  It is NOT a fallback: nothing in the product imports it, and the GPU tests run the same executors
against the real Engine."""
from collections import Counter

import numpy as np

from learningorchestra_b200 import _native as N
from oracle import bsem_numpy as bn
from oracle import rsem


class OracleEngine:
    def value_counts_f64_host(self, values):
        groups = {}
        for v in np.asarray(values, dtype=np.float64):
            k = rsem.group_key(float(v))
            if k not in groups:
                groups[k] = [0.0 if v == 0 else float(v), 0]
            groups[k][1] += 1
        keys = np.array([g[0] for g in groups.values()], dtype=np.float64)
        return keys, np.array([g[1] for g in groups.values()], dtype=np.uint64)

    def value_counts_str_host(self, cells):
        first, counts = {}, Counter()
        for i, c in enumerate(cells):
            first.setdefault(c, i)
            counts[c] += 1
        return (np.array(list(first.values()), dtype=np.int64), np.array([counts[c] for c in first], dtype=np.uint64))

    def hist_u8_cols_host(self, cols):
        return bn.hist_u8_cols(np.stack(cols), range(len(cols))), {}

    def value_counts_u32_host(self, codes, ncodes):
        return np.bincount(np.asarray(codes, dtype=np.int64), minlength=ncodes).astype(np.uint64)

    def parse_number_host(self, cells):
        vals, st = np.zeros(len(cells)), np.zeros(len(cells), dtype=np.uint8)
        for i, c in enumerate(cells):
            if c == "":
                st[i] = N.LO_NUM_EMPTY
                continue
            try:
                v = float(c)
            except ValueError:
                st[i] = N.LO_NUM_INVALID
                continue
            vals[i] = v
            st[i] = N.LO_NUM_INTEGER if (np.isfinite(v) and v.is_integer()) else N.LO_NUM_FLOAT
        return vals, st

    def parse_number_packed(self, chars, offsets):
        raw = bytes(chars)
        cells = [raw[int(offsets[i]):int(offsets[i + 1])].decode("utf-8") for i in range(len(offsets) - 1)]
        return self.parse_number_host(cells)

    def value_counts_str_packed(self, chars, offsets):
        raw = bytes(chars)
        return self.value_counts_str_host([raw[int(offsets[i]):int(offsets[i + 1])] for i in range(len(offsets) - 1)])

    def minmax_cast_host(self, cols):
        mins, maxs, cnt = [], [], []
        for c in cols:
            f = bn.cast_f64_f32(c)
            fin = f[np.isfinite(f)]
            mins.append(fin.min() if fin.size else 0.0); maxs.append(fin.max() if fin.size else 0.0); cnt.append(fin.size)
        return np.array(mins, np.float32), np.array(maxs, np.float32), np.array(cnt, np.uint64)

    def project_cast_hist_host(self, cols, nbins=None, lo=None, hi=None, out=None):
        k = len(cols)
        counts = None
        if nbins:
            lo = np.broadcast_to(np.asarray(lo, np.float32), (k,)); hi = np.broadcast_to(np.asarray(hi, np.float32), (k,))
            counts = np.zeros((k, nbins), dtype=np.uint64)
        for j, c in enumerate(cols):
            f = bn.cast_f64_f32(c)
            if out is not None:
                out[j][:] = f
            if nbins:
                counts[j] = bn.hist_f32(f, lo[j], hi[j], nbins)
        return counts, {}
