/*
 * oracle/record_harness.c -- TEST INFRASTRUCTURE.
 * A libspmv "platform" that records what a reference caller feeds the ABI:
 * on the first call it writes the six arguments to $SPMV_RECORD_PATH as raw
 * little-endian arrays (header: magic, fp32 flag, rows, nnz, ncols), then
 * computes the product with the reference semantics (libspmv/native-impl.c).
 * Used by tests/golden/make_golden.py to capture the exact CSR that
 * parboil/benchmarks/spmv/src/cpu/main.c:80-95 builds from a MatrixMarket file.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static int recorded;

static void record(const void *a, const void *iv, const int *rowstr, const int *colidx,
                   int rows, int es)
{
    const char *path = getenv("SPMV_RECORD_PATH");
    if (recorded || !path) return;
    recorded = 1;
    const int base = rowstr[0];
    const int64_t nnz = (int64_t)rowstr[rows] - base;
    int ncols = 0;
    for (int64_t k = 0; k < nnz; ++k)
        if (colidx[base - 1 + k] > ncols) ncols = colidx[base - 1 + k];
    FILE *f = fopen(path, "wb");
    if (!f) return;
    int64_t hdr[5] = {0x53504d56, es == 4, rows, nnz, ncols};
    fwrite(hdr, sizeof hdr, 1, f);
    fwrite(rowstr, sizeof(int), (size_t)rows + 1, f);
    fwrite(colidx + (base - 1), sizeof(int), (size_t)nnz, f);
    fwrite((const char *)a + (size_t)(base - 1) * es, (size_t)es, (size_t)nnz, f);
    fwrite(iv, (size_t)es, (size_t)ncols, f);
    fclose(f);
}

void *spmv_harness_(double *ov, double *a, double *iv, int *rowstr, int *colidx, int *rows)
{
    record(a, iv, rowstr, colidx, *rows, 8);
    for (int r = 0; r < *rows; ++r) {
        double acc = 0.0;
        for (int k = rowstr[r] - 1; k < rowstr[r + 1] - 1; ++k) acc = acc + a[k] * iv[colidx[k] - 1];
        ov[r] = acc;
    }
    return NULL;
}

void *f_spmv_harness_(float *ov, float *a, float *iv, int *rowstr, int *colidx, int *rows)
{
    record(a, iv, rowstr, colidx, *rows, 4);
    for (int r = 0; r < *rows; ++r) {
        float acc = 0.0f;
        for (int k = rowstr[r] - 1; k < rowstr[r + 1] - 1; ++k) acc = acc + a[k] * iv[colidx[k] - 1];
        ov[r] = acc;
    }
    return NULL;
}
