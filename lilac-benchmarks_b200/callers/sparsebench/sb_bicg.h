/*
 * sb_bicg.h -- C host-side mirror of the reference's second caller of the
 * libspmv ABI: SparseBench's unpreconditioned BiCG on an irregular CRS matrix
 * (SparseBench/run_all:42 feeds "size,2,0,1": CRS storage, no preconditioner,
 * method 1 = BiCG).  The image has no Fortran compiler, so
 * SRC/reference/{main.f,iter.f,vec.f,random.f,gen_crs.f} are restated in C for
 * exactly that configuration.
 */
#ifndef B200_SB_BICG_H
#define B200_SB_BICG_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void *(*sb_harness_fn)(double *ov, double *a, double *iv,
                               int *rowstr, int *colidx, int *rows);

typedef struct {
    int     its;          /* iterations done; negative if maxit was reached (iter.f:95) */
    double  rnorm0;       /* hist(1) */
    double  rnorm;        /* last residual norm */
    double  t_iter;       /* seconds in the solver */
    double  t_matprod;    /* seconds inside the products ("Matrix multiply: Total time") */
    int     matprod_calls;
} sb_bicg_result;

/* CRS file of SparseBench/big_gen.py:52-57 / SRC/reference/gen_crs.f:757-789:
 * "n nnz" (2 x i12), n+1 row pointers (i12), then lines "col value"; the
 * reader consumes the first nnz entry lines.  Arrays are malloc'ed. */
int sb_read_crs(const char *path, int *n, int *nnz, int **ptr, int **idx, double **val);

/* BiCG (SRC/reference/iter.f:18-104) with x0 = 0, rhs = 1 (main.f:341-346),
 * at most maxit iterations, stop when ||r|| < rtol * ||r0||.  hist gets the
 * residual norm of every iteration (caller provides >= maxit doubles or NULL);
 * x (n doubles) receives the iterate. */
int sb_bicg(int n, double *val, int *ptr, int *idx, sb_harness_fn harness,
            int maxit, double rtol, double *x, double *hist, sb_bicg_result *res);

#ifdef __cplusplus
}
#endif
#endif
