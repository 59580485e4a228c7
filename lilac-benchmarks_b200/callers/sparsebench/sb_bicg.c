/*
 * sb_bicg.c -- SparseBench BiCG on a CRS matrix, restated in C as a caller of
 * the libspmv ABI.  Follows SparseBench/SRC/reference:
 *   iter.f:18-104     bicg (prec = 0 branch: zz = r, zzl = rl)
 *   iter.f:281-306    matprod dispatcher -> random.f
 *   random.f:16-48    random_crs_matprod    -> call spmv_harness(y,val,x,ptr,idx,size)
 *   random.f:50-88    random_crs_matprod_t  -> zeroes y, then the SAME
 *                     non-transposed spmv_harness call (this fork replaced the
 *                     transposed loop by A*x; kept as is, not "fixed")
 *   vec.f:1,24,71,93,137   dotprod, vecnorm, x_is_x_plus_ay, x_is_ax_plus_y, veccopy
 *   main.f:341-346    x0 = 0, rhs = 1;  main.f:26,365  maxit = 100, rtol = 1e-6
 *   gen_crs.f:757-789 CRS file reader
 * All vector algebra is plain left-to-right host loops like the Fortran.
 */
#include "sb_bicg.h"

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

int sb_read_crs(const char *path, int *n_out, int *nnz_out, int **ptr_out, int **idx_out,
                double **val_out)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int n = 0, nnz = 0;
    if (fscanf(f, "%d %d", &n, &nnz) != 2 || n < 0 || nnz < 0) { fclose(f); return -2; }
    int *ptr = (int *)malloc(sizeof(int) * ((size_t)n + 1));
    int *idx = (int *)malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
    double *val = (double *)malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
    if (!ptr || !idx || !val) { fclose(f); return -3; }
    for (int i = 0; i <= n; ++i)
        if (fscanf(f, "%d", &ptr[i]) != 1) { fclose(f); return -2; }
    for (int k = 0; k < nnz; ++k)                     /* trailing extra lines are ignored */
        if (fscanf(f, "%d %lf", &idx[k], &val[k]) != 2) { fclose(f); return -2; }
    fclose(f);
    *n_out = n; *nnz_out = nnz; *ptr_out = ptr; *idx_out = idx; *val_out = val;
    return 0;
}

typedef struct { sb_harness_fn fn; double t; int calls; } prod_ctx;

/* random.f:16-48 */
static void matprod_n(prod_ctx *c, double *val, int *idx, int *ptr, double *x, double *y, int size)
{
    const double t = now_s();
    c->fn(y, val, x, ptr, idx, &size);
    c->t += now_s() - t;
    c->calls++;
}

/* random.f:50-88: y is zeroed, then the same non-transposed product */
static void matprod_t(prod_ctx *c, double *val, int *idx, int *ptr, double *x, double *y, int size)
{
    const double t = now_s();
    for (int row = 0; row < size; ++row) y[row] = 0.0;
    c->fn(y, val, x, ptr, idx, &size);
    c->t += now_s() - t;
    c->calls++;
}

static double dotprod(const double *x, const double *y, int n)     /* vec.f:1-22 */
{
    double d = 0.0;
    for (int i = 0; i < n; ++i) d = d + x[i] * y[i];
    return d;
}
static double vecnorm(const double *x, int n)                      /* vec.f:24-46 */
{
    double d = 0.0;
    for (int i = 0; i < n; ++i) d = d + x[i] * x[i];
    return sqrt(d);
}

int sb_bicg(int n, double *val, int *ptr, int *idx, sb_harness_fn harness,
            int maxit, double rtol, double *x, double *hist, sb_bicg_result *res)
{
    /* block(len, 9): p=1 ap=2 pl=3 apl=4 r=5 rl=6 z=7 zl=8 tmp=9 (iter.f:35) */
    enum { P = 0, AP, PL, APL, R, RL, Z, ZL, TMP, NBLK };
    const size_t len = (size_t)n;
    double *block = (double *)calloc(len * NBLK + 1, sizeof(double));
    double *rhs = (double *)malloc(sizeof(double) * (len + 1));
    if (!block || !rhs) return -3;
    double *b[NBLK];
    for (int k = 0; k < NBLK; ++k) b[k] = block + len * k;
    for (int i = 0; i < n; ++i) { x[i] = 0.0; rhs[i] = 1.0; }       /* main.f:343-346 */
    prod_ctx ctx = {harness, 0.0, 0};
    double rr = 0.0, rrp = 0.0, rn = 0.0, rn0 = 0.0, alpha, beta;
    int its = maxit, it;

    const double t0 = now_s();
    matprod_n(&ctx, val, idx, ptr, x, b[TMP], n);                   /* iter.f:46-47 */
    for (int i = 0; i < n; ++i) {
        b[R][i] = b[TMP][i] - rhs[i];
        b[RL][i] = b[R][i];
    }
    int converged = 0;
    for (it = 1; it <= maxit; ++it) {
        rn = vecnorm(b[R], n);                                      /* iter.f:55-61 */
        if (hist) hist[it - 1] = rn;
        if (it == 1) rn0 = rn;
        if (rn < rtol * rn0) { its = it; converged = 1; break; }
        if (it > 1) rrp = rr;                                       /* iter.f:70-71 (zz=r, zzl=rl) */
        rr = dotprod(b[R], b[RL], n);
        if (it == 1) {                                              /* iter.f:73-80 */
            memcpy(b[P], b[R], sizeof(double) * len);
            memcpy(b[PL], b[RL], sizeof(double) * len);
        } else {
            beta = rr / rrp;
            for (int i = 0; i < n; ++i) b[P][i] = beta * b[P][i] + b[R][i];
            for (int i = 0; i < n; ++i) b[PL][i] = beta * b[PL][i] + b[RL][i];
        }
        matprod_n(&ctx, val, idx, ptr, b[P], b[AP], n);             /* iter.f:82-85 */
        matprod_t(&ctx, val, idx, ptr, b[PL], b[APL], n);
        alpha = rr / dotprod(b[PL], b[AP], n);                      /* iter.f:87 */
        for (int i = 0; i < n; ++i) x[i] = x[i] + (-alpha) * b[P][i];       /* iter.f:90-92 */
        for (int i = 0; i < n; ++i) b[R][i] = b[R][i] + (-alpha) * b[AP][i];
        for (int i = 0; i < n; ++i) b[RL][i] = b[RL][i] + (-alpha) * b[APL][i];
    }
    if (!converged) its = -maxit;                                   /* iter.f:95 */
    res->its = its;
    res->rnorm0 = rn0;
    res->rnorm = rn;
    res->t_iter = now_s() - t0;
    res->t_matprod = ctx.t;
    res->matprod_calls = ctx.calls;
    free(block); free(rhs);
    return 0;
}
