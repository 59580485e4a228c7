/*
 * npb_cg.h -- C host-side mirror of the reference's primary caller of the
 * libspmv ABI, NPB3.3.1 serial CG (NPB3.3.1/CG/cg.f).  The image has no
 * Fortran compiler, so the driver is restated in C and calls the very same
 * six-pointer symbol (`spmv_harness_`) with the same 1-based CSR arrays.
 */
#ifndef B200_NPB_CG_H
#define B200_NPB_CG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the libspmv ABI (libspmv/native.c:3-6) */
typedef void *(*spmv_harness_fn)(double *ov, double *a, double *iv,
                                 int *rowstr, int *colidx, int *rows);

/* class table: NPB3.3.1/sys/setparams.c:551-590, CG/globals.h:22-78 */
typedef struct {
    char   cls;      /* 'S','W','A','B','C','D','E' */
    int    na;       /* matrix order */
    int    nonzer;   /* nonzeros per generating sparse vector */
    int    niter;    /* outer (inverse power) iterations */
    double shift;    /* diagonal shift */
    double rcond;    /* 0.1 for every class */
    double zeta_verify; /* cg.f:122-166 */
} npb_cg_class;

/* 0 on success, -1 for an unknown class letter */
int npb_cg_class_lookup(char cls, npb_cg_class *out);

/* 1-based CSR produced by makea (cg.f:650-905); arrays are malloc'ed */
typedef struct {
    int      n;
    int64_t  nnz;
    int     *rowstr;   /* n+1, 1-based offsets */
    int     *colidx;   /* nnz, 1-based, sorted within a row */
    double  *a;        /* nnz */
} npb_csr;

/* Generate the class matrix exactly as cg.f does (seed 314159265, one
 * warm-up randlc; cg.f:186-195).  Returns 0, or -1 if nnz overflows int32. */
int  npb_makea(const npb_cg_class *c, npb_csr *out);
void npb_csr_free(npb_csr *m);

/* Row-block variant for sharded runs: replays the whole generator stream and
 * keeps only rows [row_lo, row_hi) (0-based, half-open).  rowstr is local
 * (row_hi-row_lo+1 entries, starting at 1), colidx stays global. */
int  npb_makea_rows(const npb_cg_class *c, int row_lo, int row_hi, npb_csr *out);
/* the generating vectors of the last class are cached between calls (several
 * row blocks of one class replay the random stream once); this drops them */
void npb_makea_release_cache(void);
/* The n generating sparse vectors of the class (cached like above) and a malloc'ed copy of
 * the size_i sequence (free with npb_free): the sequential part of makea, for a generator
 * that assembles the rows elsewhere (include/b200_npb.h does it on the GPU).
 * arow[n], acol / aelt [n * (nonzer + 1)]. */
int  npb_vectors_get(const npb_cg_class *c, const int **arow, const int **acol, const double **aelt,
                     double **size_out);
void npb_free(void *p);

typedef struct {
    double zeta;            /* final zeta */
    double rnorm;           /* final ||r|| */
    double err;             /* |zeta - ref| / ref */
    int    verified;        /* err <= 1e-10 (cg.f:363-368) */
    double t_bench;         /* timed section seconds (cg.f:292-352) */
    double t_init;          /* makea + untimed iteration */
    double mops;            /* cg.f:395-402 */
    int    spmv_calls;      /* total calls through the ABI */
    /* per outer iteration history, niter entries each (caller provides or NULL) */
    double *zeta_hist;
    double *rnorm_hist;
} npb_cg_result;

/* Whole benchmark (cg.f:53-443) on an already generated matrix, through
 * `harness`.  verbose!=0 prints the NPB banner / iteration table / result
 * block in the reference's format. */
int npb_cg_run(const npb_cg_class *c, const npb_csr *m, spmv_harness_fn harness,
               npb_cg_result *res, int verbose);

/* One conj_grad call (cg.f:447-644) exposed for tests: 25 CG iterations +
 * the residual product, 26 ABI calls. */
void npb_conj_grad(const npb_csr *m, spmv_harness_fn harness,
                   double *x, double *z, double *p, double *q, double *r,
                   double *rnorm);

/* `calls` back-to-back products through `harness`, issued the way conj_grad's
 * hot loop issues them (cg.f:531-532): one matrix, x rotating over the nx
 * caller vectors xs[0..nx), y always into ov.  Returns the wall-clock
 * seconds of the loop (what a compiled caller pays per ABI call, without
 * any scripting-language marshalling). */
double npb_time_spmv_calls(spmv_harness_fn harness, double *ov, double *a, double *const *xs, int nx,
                           int *rowstr, int *colidx, int rows, int calls);

/* The same for a device-resident caller: `calls` back-to-back launches through the
 * resident-matrix entry point (b200_spmv_exec: matrix handle, device x, device y, stream),
 * x rotating over nx device vectors.  Launches only -- the caller brackets the loop with
 * events on `stream` and synchronises.  (A scripting-language loop cannot issue launches
 * faster than one per ~10 us, which hides kernels of a few microseconds.) */
typedef int (*spmv_exec_fn)(void *matrix, const void *d_x, void *d_y, void *stream);
void npb_issue_exec_calls(spmv_exec_fn exec, void *matrix, void *const *d_xs, int nx, void *d_y,
                          void *stream, int calls);

double npb_randlc(double *x, double a);

#ifdef __cplusplus
}
#endif
#endif
