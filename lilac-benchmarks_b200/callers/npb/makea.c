/*
 * makea.c -- the NPB CG test-matrix generator, restated in C.
 *
 * Follows NPB3.3.1/CG/cg.f:
 *   makea   :650-735   outer loop over n random sparse vectors
 *   sparse  :740-905   A = sum_i size_i * v_i v_i^T (+ rcond - shift on the
 *                      diagonal), duplicates summed, columns kept sorted
 *   sprnvc  :911-965   random sparse vector with `nonzer` distinct positions
 *   icnvrt  :971-985   int(ipwr2 * x)
 *   vecset  :991-1019  force element i of the vector to 0.5
 *   randlc  : NPB3.3.1/common/randi8.f:1-30   x <- a*x mod 2^46
 *
 * The reference builds each row by sorted insertion, adding a duplicate's
 * value to the slot already there (cg.f:821-871).  The value of an entry is
 * therefore 0.0 + va_1 + va_2 + ... with the contributions taken in the order
 * of the generating vector i.  This file gets the same bits without the
 * O(row length) insertion: triples are bucketed by row in generation order,
 * each bucket is sorted by (column, arrival number) and runs of equal columns
 * are summed in arrival order.  The result is the reference's 1-based CSR.
 *
 * npb_makea_rows() is the row-block form used for sharded runs: every shard
 * replays the whole random stream and keeps only its own rows, so no shard
 * ever holds the full matrix.
 */
#include "npb_cg.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const npb_cg_class k_classes[] = {
    /* cls  na       nonzer niter shift   rcond  zeta_verify (cg.f:122-166) */
    {'S',   1400,    7,     15,   10.0,   0.1,   8.5971775078648},
    {'W',   7000,    8,     15,   12.0,   0.1,   10.362595087124},
    {'A',   14000,   11,    15,   20.0,   0.1,   17.130235054029},
    {'B',   75000,   13,    75,   60.0,   0.1,   22.712745482631},
    {'C',   150000,  15,    75,   110.0,  0.1,   28.973605592845},
    {'D',   1500000, 21,    100,  500.0,  0.1,   52.514532105794},
    {'E',   9000000, 26,    100,  1500.0, 0.1,   77.522164599383},
};

int npb_cg_class_lookup(char cls, npb_cg_class *out)
{
    if (cls >= 'a' && cls <= 'z') cls = (char)(cls - 'a' + 'A');
    for (size_t i = 0; i < sizeof k_classes / sizeof k_classes[0]; ++i)
        if (k_classes[i].cls == cls) { *out = k_classes[i]; return 0; }
    return -1;
}

/* randi8.f:1-30.  The Fortran multiplies two INTEGER*8 and masks to 46 bits;
 * only the low 46 bits of the product matter, so unsigned wrap-around is the
 * same arithmetic. */
double npb_randlc(double *x, double a)
{
    const uint64_t mask46 = ((uint64_t)1 << 46) - 1;
    uint64_t lx = (uint64_t)(int64_t)*x;
    uint64_t la = (uint64_t)(int64_t)a;
    lx = (lx * la) & mask46;
    *x = (double)lx;
    return ldexp((double)lx, -46);
}

typedef struct { double tran, amult; } urando_t;

/* cg.f:911-965 */
static void sprnvc(int n, int nz, int nn1, double *v, int *iv, urando_t *u)
{
    int nzv = 0;
    while (nzv < nz) {
        double vecelt = npb_randlc(&u->tran, u->amult);
        double vecloc = npb_randlc(&u->tran, u->amult);
        int i = (int)((double)nn1 * vecloc) + 1;          /* icnvrt, cg.f:971-985 */
        if (i > n) continue;
        int seen = 0;
        for (int ii = 0; ii < nzv; ++ii)
            if (iv[ii] == i) { seen = 1; break; }
        if (seen) continue;
        v[nzv] = vecelt;
        iv[nzv] = i;
        ++nzv;
    }
}

/* cg.f:991-1019 */
static void vecset(double *v, int *iv, int *nzv, int i, double val)
{
    int set = 0;
    for (int k = 0; k < *nzv; ++k)
        if (iv[k] == i) { v[k] = val; set = 1; }
    if (!set) {
        v[*nzv] = val;
        iv[*nzv] = i;
        ++*nzv;
    }
}

static int cmp_u64(const void *pa, const void *pb)
{
    uint64_t a = *(const uint64_t *)pa, b = *(const uint64_t *)pb;
    return (a > b) - (a < b);
}

void npb_csr_free(npb_csr *m)
{
    if (!m) return;
    free(m->rowstr); free(m->colidx); free(m->a);
    memset(m, 0, sizeof *m);
}

/* The n generating sparse vectors (cg.f:709-718).  They depend only on the
 * class, so a process that builds several row blocks keeps them. */
typedef struct { int n, ld; int *arow, *acol; double *aelt; } npb_vectors;
static npb_vectors g_vec;     /* cache of the last class generated */
static char g_vec_cls;

static int make_vectors(const npb_cg_class *c)
{
    if (g_vec.arow && g_vec_cls == c->cls && g_vec.n == c->na) return 0;
    free(g_vec.arow); free(g_vec.acol); free(g_vec.aelt);
    memset(&g_vec, 0, sizeof g_vec);
    const int n = c->na, nonzer = c->nonzer, ld = nonzer + 1;
    /* nn1: smallest power of two not less than n (cg.f:700-704) */
    int nn1 = 1;
    do { nn1 *= 2; } while (nn1 < n);
    /* random stream: cg.f:186-188 (one draw is consumed before makea) */
    urando_t u = {314159265.0, 1220703125.0};
    (void)npb_randlc(&u.tran, u.amult);
    int    *arow = (int *)malloc(sizeof(int) * (size_t)n);
    int    *acol = (int *)malloc(sizeof(int) * (size_t)n * ld);
    double *aelt = (double *)malloc(sizeof(double) * (size_t)n * ld);
    if (!arow || !acol || !aelt) { free(arow); free(acol); free(aelt); return -3; }
    for (int iouter = 1; iouter <= n; ++iouter) {
        int nzv = nonzer;
        int    *ivc = acol + (size_t)(iouter - 1) * ld;
        double *vc  = aelt + (size_t)(iouter - 1) * ld;
        sprnvc(n, nzv, nn1, vc, ivc, &u);
        vecset(vc, ivc, &nzv, iouter, 0.5);
        arow[iouter - 1] = nzv;
    }
    g_vec.n = n; g_vec.ld = ld; g_vec.arow = arow; g_vec.acol = acol; g_vec.aelt = aelt;
    g_vec_cls = c->cls;
    return 0;
}

/* the generating vectors and the size_i sequence (cg.f:876: size = size * ratio), for a
 * generator that assembles the rows elsewhere (include/b200_npb.h: on the device) */
int npb_vectors_get(const npb_cg_class *c, const int **arow, const int **acol, const double **aelt,
                    double **size_out)
{
    if (make_vectors(c)) return -3;
    *arow = g_vec.arow; *acol = g_vec.acol; *aelt = g_vec.aelt;
    if (size_out) {
        double *sz = (double *)malloc(sizeof(double) * (size_t)c->na);
        if (!sz) return -3;
        double size = 1.0;
        const double ratio = pow(c->rcond, 1.0 / (double)c->na);
        for (int i = 0; i < c->na; ++i) { sz[i] = size; size = size * ratio; }
        *size_out = sz;                         /* caller frees with npb_free */
    }
    return 0;
}

void npb_free(void *p) { free(p); }

void npb_makea_release_cache(void)
{
    free(g_vec.arow); free(g_vec.acol); free(g_vec.aelt);
    memset(&g_vec, 0, sizeof g_vec);
    g_vec_cls = 0;
}

int npb_makea_rows(const npb_cg_class *c, int row_lo, int row_hi, npb_csr *out)
{
    const int n = c->na, ld = c->nonzer + 1;
    const int nrows = row_hi - row_lo;
    memset(out, 0, sizeof *out);
    if (row_lo < 0 || row_hi > n || nrows < 0) return -2;
    if (make_vectors(c)) return -3;
    const int *arow = g_vec.arow, *acol = g_vec.acol;
    const double *aelt = g_vec.aelt;
    int64_t *start = (int64_t *)calloc((size_t)nrows + 1, sizeof(int64_t));
    if (!start) return -3;

    /* count the triples that land in each kept row (cg.f:778-790) */
    for (int i = 0; i < n; ++i)
        for (int nza = 0; nza < arow[i]; ++nza) {
            int j = acol[(size_t)i * ld + nza] - 1;        /* 0-based row */
            if (j >= row_lo && j < row_hi) start[j - row_lo + 1] += arow[i];
        }
    for (int j = 0; j < nrows; ++j) start[j + 1] += start[j];
    const int64_t ntrip = start[nrows];

    uint64_t *key  = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(ntrip ? ntrip : 1));
    double   *tval = (double *)malloc(sizeof(double) * (size_t)(ntrip ? ntrip : 1));
    int64_t  *fill = (int64_t *)malloc(sizeof(int64_t) * ((size_t)nrows + 1));
    if (!key || !tval || !fill) return -3;
    memcpy(fill, start, sizeof(int64_t) * ((size_t)nrows + 1));

    /* generate the values in the reference's traversal order (cg.f:809-876) */
    double size = 1.0;
    const double ratio = pow(c->rcond, 1.0 / (double)n);
    for (int i = 0; i < n; ++i) {
        const int    *ci = acol + (size_t)i * ld;
        const double *ei = aelt + (size_t)i * ld;
        for (int nza = 0; nza < arow[i]; ++nza) {
            const int j = ci[nza];                          /* 1-based row */
            if (j - 1 >= row_lo && j - 1 < row_hi) {
                const double scale = size * ei[nza];
                int64_t pos = fill[j - 1 - row_lo];
                const int64_t base = start[j - 1 - row_lo];
                for (int nzrow = 0; nzrow < arow[i]; ++nzrow) {
                    const int jcol = ci[nzrow];
                    double va = ei[nzrow] * scale;
                    if (jcol == j && j == i + 1)
                        va = va + c->rcond - c->shift;      /* cg.f:826-828 */
                    key[pos]  = ((uint64_t)(uint32_t)jcol << 32) | (uint64_t)(pos - base);
                    tval[pos] = va;
                    ++pos;
                }
                fill[j - 1 - row_lo] = pos;
            }
        }
        size = size * ratio;
    }
    free(fill);

    /* sort each row's triples by (column, arrival) and count distinct columns */
    int *rowstr = (int *)malloc(sizeof(int) * ((size_t)nrows + 1));
    int64_t *rownnz = (int64_t *)calloc((size_t)nrows + 1, sizeof(int64_t));
    if (!rowstr || !rownnz) return -3;
#pragma omp parallel for schedule(dynamic, 64)
    for (int j = 0; j < nrows; ++j) {
        const int64_t lo = start[j], hi = start[j + 1];
        qsort(key + lo, (size_t)(hi - lo), sizeof(uint64_t), cmp_u64);
        int64_t distinct = 0;
        uint32_t prev = 0;
        for (int64_t t = lo; t < hi; ++t) {
            uint32_t col = (uint32_t)(key[t] >> 32);
            if (t == lo || col != prev) ++distinct;
            prev = col;
        }
        rownnz[j + 1] = distinct;
    }
    for (int j = 0; j < nrows; ++j) rownnz[j + 1] += rownnz[j];
    const int64_t nnz = rownnz[nrows];
    if (nnz + 1 > (int64_t)INT32_MAX) {
        free(key); free(tval); free(start); free(rowstr); free(rownnz);
        return -1;                                         /* breaks the int32 ABI */
    }

    int    *colidx = (int *)malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
    double *a      = (double *)malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
    if (!colidx || !a) return -3;
#pragma omp parallel for schedule(dynamic, 64)
    for (int j = 0; j < nrows; ++j) {
        const int64_t lo = start[j], hi = start[j + 1];
        int64_t w = rownnz[j] - 1;
        uint32_t prev = 0;
        for (int64_t t = lo; t < hi; ++t) {
            const uint32_t col = (uint32_t)(key[t] >> 32);
            const double va = tval[lo + (int64_t)(uint32_t)key[t]];
            if (t == lo || col != prev) {
                ++w;
                colidx[w] = (int)col;
                a[w] = 0.0;                                /* cg.f:800-803, 846 */
            }
            a[w] = a[w] + va;                              /* cg.f:869 */
            prev = col;
        }
        rowstr[j] = (int)(rownnz[j] + 1);
    }
    rowstr[nrows] = (int)(nnz + 1);
    free(key); free(tval); free(start); free(rownnz);

    out->n = nrows;
    out->nnz = nnz;
    out->rowstr = rowstr;
    out->colidx = colidx;
    out->a = a;
    return 0;
}

int npb_makea(const npb_cg_class *c, npb_csr *out)
{
    return npb_makea_rows(c, 0, c->na, out);
}
