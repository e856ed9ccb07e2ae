/*
 * oracle/bsem.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (learningorchestra_b200 + libloexec.so) never does.
 *
 * What it restates (paths under /root/reference/microservices), in plain C, one IEEE operation
 * per C operator (compiled with -ffp-contract=off, SSE2 scalar math, default rounding mode):
 *
 *   projection  projection_image/projection.py:38-43   out[j] = in[col_idx[j]]   (select by position;
 *               the name -> position mapping and the `_id != 0` filter live in the host adapter)
 *   cast        data_type_handler_image/data_type_update.py:40-43 turns a stored value into a
 *               "number"; the benchmark path (BASELINE.json, SURVEY.md §0 "B-semantics") defines
 *               the numeric cast as IEEE binary64 -> binary32 round-to-nearest-even.
 *   histogram   histogram_image/histogram.py:31-36 counts rows per key with $group/$sum:1.
 *               B-semantics keys a value by its fixed-width bin:
 *                   skip NaN and x outside [lo, hi];  w = (hi - lo) / nbins          (fp32)
 *                   i = (int)((x - lo) / w)  (fp32 subtract, fp32 divide, truncate); i = min(i, nbins-1)
 *               For byte columns key == value, 256 bins: exactly $group's value counts.
 *
 * PARITY UNPINNED for the fp32 cast and the binning: the reference holds no tests, golden
 * vectors or fixtures for this path and defines neither fp32 nor bins; this
 * file IS the definition.  It is cross-checked against an independent numpy restatement
 * (oracle/bsem_numpy.py) and against hand-derived known answers in tests/golden/.
 * The uint8 histogram is pinned through the reference's own histogram.py executed against an
 * in-memory collection (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_SPECIAL_PERIOD 1009
#define ORACLE_NUM_SPECIALS 20

/* ---- element semantics ------------------------------------------------------------------ */
static inline float cast_one(double x) {
    float f = (float)x; /* cvtsd2ss: round to nearest even, overflow -> inf, subnormals kept */
    if (f != f) {
        uint32_t canon = 0x7fc00000u;
        memcpy(&f, &canon, 4);
    }
    return f;
}

static inline int bin_one(float x, float lo, float hi, float w, int nbins) {
    if (!(x >= lo && x <= hi)) return -1; /* NaN fails both comparisons */
    float d = x - lo;
    float t = d / w;
    int i = (int)t; /* truncation toward zero; t is finite and >= 0 here */
    return i < nbins - 1 ? i : nbins - 1;
}

float oracle_bin_width(float lo, float hi, int nbins) {
    float span = hi - lo;
    return span / (float)nbins;
}

/* ---- building blocks on materialised columns ---------------------------------------------- */
void oracle_cast_f64_f32(const double *in, float *out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = cast_one(in[i]);
}

void oracle_hist_f32(const float *x, int64_t n, int nbins, float lo, float hi, uint64_t *counts) {
    const float w = oracle_bin_width(lo, hi, nbins);
    for (int64_t i = 0; i < n; ++i) {
        int b = bin_one(x[i], lo, hi, w, nbins);
        if (b >= 0) counts[b]++;
    }
}

/* projection + cast (+ histogram when nbins > 0) over k already-selected host columns.
 * out_cols may be NULL (histogram only).  counts[k*nbins] is overwritten.  All host threads: rows are cut into
 * 64 Ki-row chunks handed out STATICALLY (chunk c -> thread c mod T, the same map oracle_synth_fill_*_mt uses to
 * first-touch the data, so on a multi-socket host every thread scans pages of its own NUMA node); each thread
 * walks all k columns of its chunk and keeps a private k x nbins count matrix. */
#define ORACLE_CHUNK ((int64_t)1 << 16)
void oracle_project_cast_hist(const double *const *in_cols, int64_t nrows, int k, float *const *out_cols,
                              int nbins, const float *lo, const float *hi, uint64_t *counts) {
    if (nbins > 0) memset(counts, 0, (size_t)k * nbins * sizeof(uint64_t));
    const int64_t nchunks = (nrows + ORACLE_CHUNK - 1) / ORACLE_CHUNK;
#pragma omp parallel
    {
        uint64_t *local = nbins > 0 ? (uint64_t *)calloc((size_t)k * nbins, sizeof(uint64_t)) : NULL;
#pragma omp for schedule(static, 1)
        for (int64_t c = 0; c < nchunks; ++c) {
            const int64_t r0 = c * ORACLE_CHUNK, r1 = r0 + ORACLE_CHUNK < nrows ? r0 + ORACLE_CHUNK : nrows;
            for (int j = 0; j < k; ++j) {
                const double *src = in_cols[j];
                float *dst = out_cols ? out_cols[j] : NULL;
                const float l = nbins > 0 ? lo[j] : 0.f, h = nbins > 0 ? hi[j] : 0.f;
                const float w = nbins > 0 ? oracle_bin_width(l, h, nbins) : 1.f;
                uint64_t *cnt = nbins > 0 ? local + (size_t)j * nbins : NULL;
                for (int64_t r = r0; r < r1; ++r) {
                    float f = cast_one(src[r]);
                    if (dst) dst[r] = f;
                    if (cnt) {
                        int b = bin_one(f, l, h, w, nbins);
                        if (b >= 0) cnt[b]++;
                    }
                }
            }
        }
        if (local) {
#pragma omp critical
            for (int64_t i = 0; i < (int64_t)k * nbins; ++i) counts[i] += local[i];
            free(local);
        }
    }
}

/* per-column value counts of byte columns; counts[k*256] overwritten */
void oracle_hist_u8_cols(const uint8_t *const *in_cols, int64_t nrows, int k, uint64_t *counts) {
    memset(counts, 0, (size_t)k * 256 * sizeof(uint64_t));
    const int64_t nchunks = (nrows + ORACLE_CHUNK - 1) / ORACLE_CHUNK;
#pragma omp parallel
    {
        uint64_t *local = (uint64_t *)calloc((size_t)k * 256, sizeof(uint64_t));
#pragma omp for schedule(static, 1)
        for (int64_t c = 0; c < nchunks; ++c) {
            const int64_t r0 = c * ORACLE_CHUNK, r1 = r0 + ORACLE_CHUNK < nrows ? r0 + ORACLE_CHUNK : nrows;
            for (int j = 0; j < k; ++j) {
                uint64_t *cnt = local + (size_t)j * 256;
                const uint8_t *src = in_cols[j];
                for (int64_t r = r0; r < r1; ++r) cnt[src[r]]++;
            }
        }
#pragma omp critical
        for (int64_t i = 0; i < (int64_t)k * 256; ++i) counts[i] += local[i];
        free(local);
    }
}

/* position-weighted checksum of bit patterns: sum bits(x[r]) * (2*(row_offset+r)+1) mod 2^64 */
uint64_t oracle_checksum(const void *col, int elem_size, int64_t nrows, int64_t row_offset) {
    uint64_t acc = 0;
    for (int64_t r = 0; r < nrows; ++r) {
        uint64_t bits = 0;
        memcpy(&bits, (const char *)col + r * elem_size, (size_t)elem_size); /* little endian */
        acc += bits * (2ull * (uint64_t)(row_offset + r) + 1ull);
    }
    return acc;
}

/* ---- counter-based synthetic tables (twin of lo_table_fill_synthetic_dev) ------------------- */
static inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline double bits_to_double(uint64_t b) {
    double d;
    memcpy(&d, &b, 8);
    return d;
}

static double special_value(int idx, double lo, double hi) {
    switch (idx) {
        case 0: return 0.0;
        case 1: return -0.0;
        case 2: return 1e-40;
        case 3: return 1e-46;
        case 4: return -1e-46;
        case 5: return 1e39;
        case 6: return -1e39;
        case 7: return bits_to_double(0x7ff8000000000000ull);
        case 8: return bits_to_double(0xfff4000000000001ull);
        case 9: return 1.0 + 5.9604644775390625e-08;
        case 10: return 1.0 + 1.7881393432617188e-07;
        case 11: return 16777217.0;
        case 12: return 3.4028235677973366e38;
        case 13: return hi;
        case 14: return lo;
        case 15: {
            uint64_t b;
            memcpy(&b, &hi, 8);
            b = hi > 0 ? b + 1 : b - 1;
            return bits_to_double(b);
        }
        case 16: return hi + (hi - lo) * 9.5367431640625e-07;
        case 17: return lo - (hi - lo) * 9.5367431640625e-07;
        case 18: return bits_to_double(0x7ff0000000000000ull);
        default: return bits_to_double(0xfff0000000000000ull);
    }
}

/* kind: 0 uniform, 1 + special values, 2 + constant column 0 */
static inline double synth_f64_one(int kind, uint64_t seed, int col, uint64_t g, double lo, double hi) {
    const uint64_t u = splitmix64(seed ^ ((uint64_t)col << 40) ^ g);
    double frac = (double)(u >> 11) * 1.1102230246251565e-16; /* 2^-53 */
    double span = hi - lo;
    double scaled = span * frac;
    double x = lo + scaled;
    if (kind >= 1 && (g % ORACLE_SPECIAL_PERIOD) == (uint64_t)(col % ORACLE_SPECIAL_PERIOD))
        x = special_value((int)((g / ORACLE_SPECIAL_PERIOD + (uint64_t)col) % ORACLE_NUM_SPECIALS), lo, hi);
    if (kind == 2 && col == 0) {
        double q = span * 0.75;
        x = lo + q;
    }
    return x;
}

void oracle_synth_f64(int kind, uint64_t seed, int col, int64_t row0, int64_t n, double lo, double hi, double *out) {
    for (int64_t r = 0; r < n; ++r) out[r] = synth_f64_one(kind, seed, col, (uint64_t)(row0 + r), lo, hi);
}

static inline uint8_t synth_u8_one(uint64_t seed, int col, uint64_t g) {
    const uint64_t u = splitmix64(seed ^ ((uint64_t)col << 40) ^ g);
    const int py = (col % 784) / 28, px = (col % 784) % 28;
    if (py >= 4 && py < 24 && px >= 4 && px < 24 && (u & 0xFFu) >= 0x99u) return (uint8_t)((u >> 8) & 0xFFu);
    return 0;
}

void oracle_synth_u8(uint64_t seed, int col, int64_t row0, int64_t n, uint8_t *out) {
    for (int64_t r = 0; r < n; ++r) out[r] = synth_u8_one(seed, col, (uint64_t)(row0 + r));
}

/* Generate k table columns (col_idx[j] of the synthetic table, global rows [row0, row0+n)) into cols[j], with the
 * SAME team and static chunk map as oracle_project_cast_hist: parallel first touch.  touch_out[j] (may be NULL) is
 * zero-filled the same way so the output pages are placed too. */
void oracle_synth_fill_f64_mt(int kind, uint64_t seed, const int32_t *col_idx, int k, int64_t row0, int64_t n,
                              double glo, double ghi, double *const *cols, float *const *touch_out) {
    const int64_t nchunks = (n + ORACLE_CHUNK - 1) / ORACLE_CHUNK;
#pragma omp parallel for schedule(static, 1)
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t r0 = c * ORACLE_CHUNK, r1 = r0 + ORACLE_CHUNK < n ? r0 + ORACLE_CHUNK : n;
        for (int j = 0; j < k; ++j) {
            for (int64_t r = r0; r < r1; ++r) cols[j][r] = synth_f64_one(kind, seed, col_idx[j], (uint64_t)(row0 + r), glo, ghi);
            if (touch_out && touch_out[j]) memset(touch_out[j] + r0, 0, (size_t)(r1 - r0) * sizeof(float));
        }
    }
}

void oracle_synth_fill_u8_mt(uint64_t seed, const int32_t *col_idx, int k, int64_t row0, int64_t n, uint8_t *const *cols) {
    const int64_t nchunks = (n + ORACLE_CHUNK - 1) / ORACLE_CHUNK;
#pragma omp parallel for schedule(static, 1)
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t r0 = c * ORACLE_CHUNK, r1 = r0 + ORACLE_CHUNK < n ? r0 + ORACLE_CHUNK : n;
        for (int j = 0; j < k; ++j)
            for (int64_t r = r0; r < r1; ++r) cols[j][r] = synth_u8_one(seed, col_idx[j], (uint64_t)(row0 + r));
    }
}

/* Streaming full-size check: regenerate rows [row0, row0+nrows) of the synthetic table, run
 * projection + cast + histogram on them without materialising anything, and return per projected
 * column the counts[k*nbins] and the checksum of the fp32 output slab (row offsets are GLOBAL). */
void oracle_synth_project_cast_hist(int kind, uint64_t seed, int64_t row0, int64_t nrows, double glo, double ghi,
                                    const int32_t *col_idx, int k, int nbins, const float *lo, const float *hi,
                                    uint64_t *counts, uint64_t *checksums) {
    memset(counts, 0, (size_t)k * nbins * sizeof(uint64_t));
    memset(checksums, 0, (size_t)k * sizeof(uint64_t));
    const int64_t chunk = 1 << 16;
    const int64_t nchunks = (nrows + chunk - 1) / chunk;
#pragma omp parallel
    {
        uint64_t *local = (uint64_t *)calloc((size_t)k * nbins + k, sizeof(uint64_t));
        uint64_t *lsum = local + (size_t)k * nbins;
#pragma omp for schedule(dynamic, 1) collapse(2)
        for (int j = 0; j < k; ++j) {
            for (int64_t c = 0; c < nchunks; ++c) {
                const int64_t r0 = c * chunk, r1 = r0 + chunk < nrows ? r0 + chunk : nrows;
                const float l = lo[j], h = hi[j], w = oracle_bin_width(l, h, nbins);
                uint64_t *cnt = local + (size_t)j * nbins;
                uint64_t acc = 0;
                for (int64_t r = r0; r < r1; ++r) {
                    const uint64_t g = (uint64_t)(row0 + r);
                    float f = cast_one(synth_f64_one(kind, seed, col_idx[j], g, glo, ghi));
                    uint32_t bits;
                    memcpy(&bits, &f, 4);
                    acc += (uint64_t)bits * (2ull * g + 1ull);
                    int b = bin_one(f, l, h, w, nbins);
                    if (b >= 0) cnt[b]++;
                }
                lsum[j] += acc;
            }
        }
#pragma omp critical
        {
            for (int64_t i = 0; i < (int64_t)k * nbins; ++i) counts[i] += local[i];
            for (int j = 0; j < k; ++j) checksums[j] += lsum[j];
        }
        free(local);
    }
}

/* same for the MNIST-shaped byte table: counts[k*256] */
void oracle_synth_hist_u8(uint64_t seed, int64_t row0, int64_t nrows, const int32_t *col_idx, int k,
                          uint64_t *counts) {
    memset(counts, 0, (size_t)k * 256 * sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < k; ++j) {
        uint64_t *cnt = counts + (size_t)j * 256;
        for (int64_t r = 0; r < nrows; ++r) cnt[synth_u8_one(seed, col_idx[j], (uint64_t)(row0 + r))]++;
    }
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU arm wants every core the process may run on */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
