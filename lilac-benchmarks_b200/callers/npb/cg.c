/*
 * cg.c -- NPB3.3.1 serial CG restated in C as a caller of the libspmv ABI.
 *
 * Follows NPB3.3.1/CG/cg.f:
 *   main program  :53-443   (class detection :122-166, untimed iteration
 *                            :233-272, timed loop :299-349, verification
 *                            :363-392, Mop/s :395-402)
 *   conj_grad     :447-644  (the two ABI call sites are :531-532 and :628)
 * and the result block of NPB3.3.1/common/print_results.f.
 *
 * All vector algebra stays on the host in plain left-to-right loops exactly
 * like the Fortran; the only thing delegated is `call spmv_harness(...)`.
 * The four work vectors are allocated once with na+2 elements (cg.f:75-80)
 * so the backend sees four stable host pointers, as it does under Fortran
 * COMMON storage.
 */
#include "npb_cg.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int g_spmv_calls;

/* cg.f:447-644 */
void npb_conj_grad(const npb_csr *m, spmv_harness_fn harness,
                   double *x, double *z, double *p, double *q, double *r,
                   double *rnorm)
{
    const int naa = m->n;
    const int cgitmax = 25;
    int rows = naa;                       /* lastrow - firstrow + 1 */
    double rho = 0.0, rho0, alpha, beta, d, sum;

    /* cg.f:484-489 (note the loop runs to naa+1) */
    for (int j = 0; j < naa + 1; ++j) {
        q[j] = 0.0;
        z[j] = 0.0;
        r[j] = x[j];
        p[j] = r[j];
    }
    /* cg.f:496-498 */
    for (int j = 0; j < naa; ++j) rho = rho + r[j] * r[j];

    for (int cgit = 1; cgit <= cgitmax; ++cgit) {
        /* q = A.p : cg.f:531-532 */
        rows = naa;
        harness(q, m->a, p, m->rowstr, m->colidx, &rows);
        ++g_spmv_calls;

        d = 0.0;                                            /* cg.f:573-576 */
        for (int j = 0; j < naa; ++j) d = d + p[j] * q[j];
        alpha = rho / d;                                    /* cg.f:581 */
        rho0 = rho;                                         /* cg.f:586 */
        rho = 0.0;
        for (int j = 0; j < naa; ++j) {                     /* cg.f:593-596 */
            z[j] = z[j] + alpha * p[j];
            r[j] = r[j] - alpha * q[j];
        }
        for (int j = 0; j < naa; ++j) rho = rho + r[j] * r[j];  /* cg.f:602-604 */
        beta = rho / rho0;                                  /* cg.f:609 */
        for (int j = 0; j < naa; ++j) p[j] = r[j] + beta * p[j]; /* cg.f:614-616 */
    }

    /* r = A.z : cg.f:628 */
    harness(r, m->a, z, m->rowstr, m->colidx, &rows);
    ++g_spmv_calls;

    sum = 0.0;                                              /* cg.f:633-637 */
    for (int j = 0; j < naa; ++j) {
        d = x[j] - r[j];
        sum = sum + d * d;
    }
    *rnorm = sqrt(sum);
}

int npb_cg_run(const npb_cg_class *c, const npb_csr *m, spmv_harness_fn harness,
               npb_cg_result *res, int verbose)
{
    const int na = c->na;
    if (m->n != na) return -1;
    double *x = (double *)calloc((size_t)na + 2, sizeof(double));
    double *z = (double *)calloc((size_t)na + 2, sizeof(double));
    double *p = (double *)calloc((size_t)na + 2, sizeof(double));
    double *q = (double *)calloc((size_t)na + 2, sizeof(double));
    double *r = (double *)calloc((size_t)na + 2, sizeof(double));
    if (!x || !z || !p || !q || !r) return -3;
    double zeta, rnorm = 0.0, norm_temp1, norm_temp2;
    g_spmv_calls = 0;

    if (verbose) {
        printf("\n\n NAS Parallel Benchmarks (NPB3.3-SER) - CG Benchmark\n\n");
        printf(" Size: %11d\n", na);
        printf(" Iterations: %5d\n\n", c->niter);
    }

    double t0 = wtime();
    for (int i = 0; i < na + 1; ++i) x[i] = 1.0;             /* cg.f:216-218 */
    zeta = 0.0;

    /* one untimed iteration: cg.f:233-272 */
    npb_conj_grad(m, harness, x, z, p, q, r, &rnorm);
    norm_temp1 = 0.0;
    norm_temp2 = 0.0;
    for (int j = 0; j < na; ++j) {
        norm_temp1 = norm_temp1 + x[j] * z[j];
        norm_temp2 = norm_temp2 + z[j] * z[j];
    }
    norm_temp2 = 1.0 / sqrt(norm_temp2);
    for (int j = 0; j < na; ++j) x[j] = norm_temp2 * z[j];

    for (int i = 0; i < na + 1; ++i) x[i] = 1.0;             /* cg.f:280-282 */
    zeta = 0.0;
    double t1 = wtime();
    if (verbose) printf(" Initialization time = %15.3f seconds\n", t1 - t0);

    /* timed section: cg.f:292-352 */
    for (int it = 1; it <= c->niter; ++it) {
        npb_conj_grad(m, harness, x, z, p, q, r, &rnorm);
        norm_temp1 = 0.0;
        norm_temp2 = 0.0;
        for (int j = 0; j < na; ++j) {
            norm_temp1 = norm_temp1 + x[j] * z[j];
            norm_temp2 = norm_temp2 + z[j] * z[j];
        }
        norm_temp2 = 1.0 / sqrt(norm_temp2);
        zeta = c->shift + 1.0 / norm_temp1;
        if (verbose) {
            if (it == 1) printf("\n   iteration           ||r||                 zeta\n");
            printf("    %5d       %20.14E%20.13f\n", it, rnorm, zeta);
        }
        if (res->zeta_hist)  res->zeta_hist[it - 1] = zeta;
        if (res->rnorm_hist) res->rnorm_hist[it - 1] = rnorm;
        for (int j = 0; j < na; ++j) x[j] = norm_temp2 * z[j];
    }
    double t2 = wtime();

    const double t = t2 - t1;
    const double err = fabs(zeta - c->zeta_verify) / c->zeta_verify;
    const int verified = err <= 1.0e-10;
    double mops = 0.0;
    if (t != 0.0) {                                          /* cg.f:395-402 */
        const double nz1 = (double)(c->nonzer * (c->nonzer + 1));
        mops = (double)(2.0 * c->niter * (double)na)
             * (3.0 + nz1 + 25.0 * (5.0 + nz1) + 3.0) / t / 1000000.0;
    }

    res->zeta = zeta;
    res->rnorm = rnorm;
    res->err = err;
    res->verified = verified;
    res->t_bench = t;
    res->t_init = t1 - t0;
    res->mops = mops;
    res->spmv_calls = g_spmv_calls;

    if (verbose) {
        printf(" Benchmark completed \n");
        if (verified) {
            printf(" VERIFICATION SUCCESSFUL \n");
            printf(" Zeta is    %20.13E\n", zeta);
            printf(" Error is   %20.13E\n", err);
        } else {
            printf(" VERIFICATION FAILED\n");
            printf(" Zeta                %20.13E\n", zeta);
            printf(" The correct zeta is %20.13E\n", c->zeta_verify);
        }
        /* NPB3.3.1/common/print_results.f */
        printf("\n\n CG Benchmark Completed.\n");
        printf(" Class           =             %12c\n", c->cls);
        printf(" Size            =             %12d\n", na);
        printf(" Iterations      =             %12d\n", c->niter);
        printf(" Time in seconds =             %12.2f\n", t);
        printf(" Mop/s total     =             %12.2f\n", mops);
        printf(" Operation type  =           floating point\n");
        printf(" Verification    =             %12s\n", verified ? "  SUCCESSFUL" : "UNSUCCESSFUL");
        printf(" Version         =             %12s\n", "3.3.1");
        printf(" SpMV calls      =             %12d\n", g_spmv_calls);
    }
    free(x); free(z); free(p); free(q); free(r);
    return 0;
}

double npb_time_spmv_calls(spmv_harness_fn harness, double *ov, double *a, double *const *xs, int nx,
                           int *rowstr, int *colidx, int rows, int calls)
{
    int n = rows;                       /* passed by reference, as Fortran does */
    const double t0 = wtime();
    for (int i = 0; i < calls; ++i) harness(ov, a, xs[i % nx], rowstr, colidx, &n);
    return wtime() - t0;
}

void npb_issue_exec_calls(spmv_exec_fn exec, void *matrix, void *const *d_xs, int nx, void *d_y,
                          void *stream, int calls)
{
    for (int i = 0; i < calls; ++i) exec(matrix, d_xs[i % nx], d_y, stream);
}
