/*
 * pagerank.h -- C mirror of the reference's third caller of the libspmv ABI,
 * pagerank/main.cpp (power iteration, backend chosen by dlopen).  The C++
 * program cannot be built here (it needs the un-vendored `mm` library,
 * pagerank/main.cpp:1), so its loop is restated in C.
 */
#ifndef B200_PAGERANK_H
#define B200_PAGERANK_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void *(*pr_harness_fn)(double *ov, double *a, double *iv,
                               int *rowstr, int *colidx, int *rows);

/* One run of `iters` power iterations (pagerank/main.cpp:125-149):
 *   mean = sum(x)/n; y = (d M) x; y += (1-d) mean; x = y; error = ||x - x_prev||_2
 * `a` already holds d*M (column-normalised, scaled by d = 0.85, main.cpp:103-111).
 * x is updated in place (n doubles); y is scratch (n doubles).  Returns the
 * last error; *seconds gets the wall time of the loop. */
double pr_power_iterations(int n, double *a, int *rowstr, int *colidx, double *x, double *y,
                           double d, int iters, pr_harness_fn harness, double *seconds);

/* MatrixMarket coordinate file -> 1-based CSR of d * (column-normalised matrix),
 * rows = destinations (main.cpp:103-111).  Arrays are malloc'ed. */
int pr_load_mtx(const char *path, double d, int *n, int *nnz, int **rowstr, int **colidx, double **a);

#ifdef __cplusplus
}
#endif
#endif
