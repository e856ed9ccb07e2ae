/* Plain-C consumer of include/loexec.h: proves the boundary needs nothing but a C compiler and the .so.
 * Exit codes: 0 = ran the hot path on the GPU and the invariants held; 3 = no usable GPU (LO_ERR_NO_DEVICE,
 * the documented behaviour on a CPU-only host); anything else = failure. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "loexec.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != LO_OK) {                                                      \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, lo_last_error());      \
            return rc_ == LO_ERR_NO_DEVICE ? 3 : 1;                              \
        }                                                                        \
    } while (0)

int main(void) {
    if (lo_abi_version() != LO_ABI_VERSION) return 2;
    lo_ctx *ctx = NULL;
    CHECK(lo_init(0, &ctx));
    const int64_t nrows = 1000003;
    const int32_t ncols = 6, k = 3, nbins = 256;
    const int32_t cols[3] = {4, 0, 2};
    lo_table *in = NULL, *out = NULL;
    CHECK(lo_table_alloc(ctx, LO_F64, nrows, ncols, &in));
    CHECK(lo_table_alloc(ctx, LO_F32, nrows, k, &out));
    CHECK(lo_table_fill_synthetic_dev(ctx, in, LO_SYNTH_UNIFORM, 20260921ull, 0, -1000.0, 1000.0, NULL));
    uint64_t *counts_dev = NULL;
    CHECK(lo_counts_alloc(ctx, (int64_t)k * nbins, &counts_dev));
    float lo[3] = {-1000.f, -1000.f, -1000.f}, hi[3] = {1000.f, 1000.f, 1000.f};
    lo_hist_spec spec = {nbins, 0, lo, hi};
    CHECK(lo_project_cast_hist_dev(ctx, in, cols, k, out, &spec, counts_dev, NULL));
    uint64_t *counts = (uint64_t *)malloc(sizeof(uint64_t) * k * nbins);
    CHECK(lo_counts_download(ctx, counts_dev, (int64_t)k * nbins, counts, NULL));
    for (int j = 0; j < k; ++j) {
        uint64_t total = 0;
        for (int b = 0; b < nbins; ++b) total += counts[j * nbins + b];
        if (total != (uint64_t)nrows) { fprintf(stderr, "column %d counted %llu rows\n", j, (unsigned long long)total); return 1; }
    }
    /* the same through the host-buffer entry point: download a column, push it back through the pipeline */
    double *col = (double *)malloc(sizeof(double) * nrows);
    float *res = (float *)malloc(sizeof(float) * nrows);
    CHECK(lo_table_download_col(ctx, in, cols[0], 0, col, nrows, NULL));
    const double *in_cols[1] = {col};
    float *out_cols[1] = {res};
    uint64_t hcounts[256];
    lo_hist_spec spec1 = {nbins, 0, lo, hi};
    lo_host_timing tm;
    CHECK(lo_project_cast_hist_host(ctx, in_cols, nrows, 1, out_cols, &spec1, hcounts, &tm));
    for (int b = 0; b < nbins; ++b)
        if (hcounts[b] != counts[b]) { fprintf(stderr, "host/dev counts differ at bin %d\n", b); return 1; }
    for (int64_t r = 0; r < nrows; r += 997)
        if (res[r] != (float)col[r]) { fprintf(stderr, "row %lld cast differs\n", (long long)r); return 1; }
    int64_t launches = 0;
    CHECK(lo_launch_count(ctx, &launches));
    printf("abi_smoke ok: %lld kernel launches, %.0f bytes h2d\n", (long long)launches, tm.h2d_bytes);
    free(counts); free(col); free(res);
    CHECK(lo_counts_free(ctx, counts_dev));
    CHECK(lo_table_free(ctx, in));
    CHECK(lo_table_free(ctx, out));
    CHECK(lo_shutdown(ctx));
    return 0;
}
