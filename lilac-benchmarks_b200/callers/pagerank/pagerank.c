/*
 * pagerank.c -- the reference's PageRank power iteration as a C caller of the
 * libspmv ABI.  Follows pagerank/main.cpp:
 *   :82-98    random start vector normalised to sum 1 (done by the caller here)
 *   :103-111  read .mtx, normalise, 1-based CSR, scale by d = 0.85
 *   :125-149  the timed loop: mean, harness(y, a, x, rowstr, colidx, rows),
 *             y += (1-d)*mean, x = y, error = l2 norm of the change
 * x and y keep their addresses for the whole run, as the two std::vectors do.
 */
#include "pagerank.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

double pr_power_iterations(int n, double *a, int *rowstr, int *colidx, double *x, double *y,
                           double d, int iters, pr_harness_fn harness, double *seconds)
{
    double *last = (double *)malloc(sizeof(double) * (size_t)(n ? n : 1));
    double error = 0.0;
    int rows = n;
    const double t0 = now_s();
    for (int it = 0; it < iters; ++it) {
        memcpy(last, x, sizeof(double) * (size_t)n);                /* last_vector = x */
        double sum = 0.0;
        for (int i = 0; i < n; ++i) sum += x[i];                    /* std::accumulate */
        const double add_term = (1.0 - d) * (sum / (double)n);
        harness(y, a, x, rowstr, colidx, &rows);
        for (int i = 0; i < n; ++i) y[i] += add_term;
        memcpy(x, y, sizeof(double) * (size_t)n);                   /* x = y */
        error = 0.0;
        for (int i = 0; i < n; ++i) {
            const double df = x[i] - last[i];
            error += df * df;
        }
        error = sqrt(error);
    }
    if (seconds) *seconds = now_s() - t0;
    free(last);
    return error;
}

typedef struct { int r, c; double v; } coo_t;
static int cmp_coo(const void *pa, const void *pb)
{
    const coo_t *a = (const coo_t *)pa, *b = (const coo_t *)pb;
    if (a->r != b->r) return a->r < b->r ? -1 : 1;
    return a->c < b->c ? -1 : (a->c > b->c);
}

int pr_load_mtx(const char *path, double d, int *n_out, int *nnz_out, int **rowstr_out,
                int **colidx_out, double **a_out)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[512];
    int pattern = 0, symmetric = 0;
    if (!fgets(line, sizeof line, f)) { fclose(f); return -2; }
    if (strstr(line, "pattern")) pattern = 1;
    if (strstr(line, "symmetric")) symmetric = 1;
    while (line[0] == '%')
        if (!fgets(line, sizeof line, f)) { fclose(f); return -2; }
    int rows, cols, entries;
    if (sscanf(line, "%d %d %d", &rows, &cols, &entries) != 3 || rows != cols) { fclose(f); return -2; }
    coo_t *e = (coo_t *)malloc(sizeof(coo_t) * (size_t)entries * (symmetric ? 2 : 1) + 1);
    int m = 0;
    for (int k = 0; k < entries; ++k) {
        int r, c;
        double v = 1.0;
        if (!fgets(line, sizeof line, f)) break;
        if (pattern ? sscanf(line, "%d %d", &r, &c) != 2 : sscanf(line, "%d %d %lf", &r, &c, &v) < 2) continue;
        e[m].r = r; e[m].c = c; e[m].v = fabs(v) > 0 ? fabs(v) : 1.0; ++m;
        if (symmetric && r != c) { e[m].r = c; e[m].c = r; e[m].v = e[m - 1].v; ++m; }
    }
    fclose(f);
    double *colsum = (double *)calloc((size_t)rows + 1, sizeof(double));
    for (int k = 0; k < m; ++k) colsum[e[k].c] += e[k].v;
    qsort(e, (size_t)m, sizeof(coo_t), cmp_coo);
    int *rowstr = (int *)calloc((size_t)rows + 2, sizeof(int));
    const size_t cap = m > 0 ? (size_t)m : 1;
    int *colidx = (int *)malloc(sizeof(int) * cap);
    double *a = (double *)malloc(sizeof(double) * cap);
    for (int k = 0; k < m; ++k) rowstr[e[k].r]++;
    int run = 1;
    for (int r = 1; r <= rows; ++r) { const int c = rowstr[r]; rowstr[r] = run; run += c; }
    rowstr[rows + 1] = run;
    for (int k = 0; k < m; ++k) {
        colidx[k] = e[k].c;
        a[k] = d * e[k].v / colsum[e[k].c];
    }
    /* shift to 0-based array of n+1 one-based offsets */
    memmove(rowstr, rowstr + 1, sizeof(int) * ((size_t)rows + 1));
    free(e); free(colsum);
    *n_out = rows; *nnz_out = m; *rowstr_out = rowstr; *colidx_out = colidx; *a_out = a;
    return 0;
}
