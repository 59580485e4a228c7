/*
 * oracle/spmv_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's sequential CSR SpMV, used only as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under lilac-benchmarks_b200/ may call,
 * link or load this file.
 *
 * What it restates (reference file:line):
 *   libspmv/native-impl.c:1-12   native_spmv    (fp64)
 *   libspmv/native-impl.c:14-25  f_native_spmv  (fp32)
 *   libspmv/native.c:3-11        the two exported Fortran-callable symbols
 *
 * Semantics kept to the letter: 1-based rowstr / colidx, the row sum starts at
 * +0.0 and is accumulated strictly left to right, the product and the sum are
 * rounded separately (the reference build has no FMA: -O3 -std=c99 without
 * -march; this file is built with -ffp-contract=off to guarantee the same),
 * y is overwritten, an empty row yields 0.0, rowstr[0] may be any base.
 *
 * Parity pinned: tests/test_oracle.py checks these functions bit-for-bit
 * against (a) the libspmv/test.cpp:44-49 known-answer vector, (b) the
 * reference's own native.so built into oracle/_ref/ (when /root/reference is
 * present), (c) the committed fixtures in tests/golden/.
 */
#include <stddef.h>
#include <stdint.h>
#include <math.h>

#define ORACLE_ROW_LOOP(T, y, val, x, rowstr, colidx, nrows)                 \
    do {                                                                     \
        for (int r = 0; r < (nrows); ++r) {                                  \
            T acc = (T)0.0;                                                  \
            const int lo = (rowstr)[r] - 1, hi = (rowstr)[r + 1] - 1;        \
            for (int k = lo; k < hi; ++k) {                                  \
                T prod = (val)[k] * (x)[(colidx)[k] - 1];                    \
                acc = acc + prod;                                            \
            }                                                                \
            (y)[r] = acc;                                                    \
        }                                                                    \
    } while (0)

/* fp64 oracle: follows libspmv/native-impl.c:1-12 */
void oracle_spmv_f64(double *y, const double *val, const double *x,
                     const int *rowstr, const int *colidx, const int *rows)
{
    ORACLE_ROW_LOOP(double, y, val, x, rowstr, colidx, *rows);
}

/* fp32 oracle: follows libspmv/native-impl.c:14-25 */
void oracle_spmv_f32(float *y, const float *val, const float *x,
                     const int *rowstr, const int *colidx, const int *rows)
{
    ORACLE_ROW_LOOP(float, y, val, x, rowstr, colidx, *rows);
}

/*
 * The same loop with rows spread over host threads: the multi-core CPU
 * baseline that stands in for the reference's MKL backend (libspmv/mkl.c),
 * which is not installed in this image.  Each row is still summed in the
 * reference order, so the result is bit-identical to oracle_spmv_f64.
 */
void oracle_spmv_f64_omp(double *y, const double *val, const double *x,
                         const int *rowstr, const int *colidx, const int *rows)
{
    const int n = *rows;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; ++r) {
        double acc = 0.0;
        const int lo = rowstr[r] - 1, hi = rowstr[r + 1] - 1;
        for (int k = lo; k < hi; ++k) {
            double prod = val[k] * x[colidx[k] - 1];
            acc = acc + prod;
        }
        y[r] = acc;
    }
}

void oracle_spmv_f32_omp(float *y, const float *val, const float *x,
                         const int *rowstr, const int *colidx, const int *rows)
{
    const int n = *rows;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; ++r) {
        float acc = 0.0f;
        const int lo = rowstr[r] - 1, hi = rowstr[r + 1] - 1;
        for (int k = lo; k < hi; ++k) {
            float prod = val[k] * x[colidx[k] - 1];
            acc = acc + prod;
        }
        y[r] = acc;
    }
}

/*
 * Error-metric helpers (not in the reference): per row, the long-double value
 * of the sum and the sum of |terms|, so tests can report both the strict
 * |dy|/|y_native| and the backward error |dy|/sum|a_ij x_j|.
 */
void oracle_spmv_f64_extended(double *y_ld, double *abs_terms, const double *val,
                              const double *x, const int *rowstr,
                              const int *colidx, const int *rows)
{
    const int n = *rows;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; ++r) {
        long double acc = 0.0L, mag = 0.0L;
        const int lo = rowstr[r] - 1, hi = rowstr[r + 1] - 1;
        for (int k = lo; k < hi; ++k) {
            long double prod = (long double)val[k] * (long double)x[colidx[k] - 1];
            acc += prod;
            mag += fabsl(prod);
        }
        y_ld[r] = (double)acc;
        abs_terms[r] = (double)mag;
    }
}

/* max(colidx) over the nnz entries: the column count every reference GPU/MKL
 * backend derives because the ABI does not pass it (libspmv/mkl.c:42-44,
 * libspmv/opencl.cpp:344-346; gpu.c:216-223 intends the same). */
int oracle_max_colidx(const int *rowstr, const int *colidx, const int *rows)
{
    int m = 0;
    for (long k = rowstr[0] - 1; k < (long)rowstr[*rows] - 1; ++k)
        if (colidx[k] > m) m = colidx[k];
    return m;
}

/*
 * The ABI names, so that liboracle.so can be dlopen'ed / linked wherever a
 * libX-spmv.so is expected (NPB driver run on the CPU, test.cpp).
 * Follows libspmv/native.c:3-11; returns NULL like libspmv/gpu.c:288 (the
 * reference's native version falls off the end of a void* function).
 */
void *spmv_harness_(double *ov, double *a, double *iv, int *rowstr, int *colidx, int *rows)
{
    oracle_spmv_f64(ov, a, iv, rowstr, colidx, rows);
    return NULL;
}

void *f_spmv_harness_(float *ov, float *a, float *iv, int *rowstr, int *colidx, int *rows)
{
    oracle_spmv_f32(ov, a, iv, rowstr, colidx, rows);
    return NULL;
}
