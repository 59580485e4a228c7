/*
 * sb_main.c -- front end: `sparsebench crsmatNNNu [libX-spmv.so]`.
 * Reads the CRS file big_gen.py writes (SparseBench/run_all:35-38 generates it
 * before the run) and prints the lines SparseBench/run_all:42 scrapes.
 */
#include "sb_bicg.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

extern void *spmv_harness_(double *, double *, double *, int *, int *, int *) __attribute__((weak));

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s crsmat-file [libX-spmv.so]\n", argv[0]); return 1; }
    sb_harness_fn harness = NULL;
    if (argc >= 3) {
        void *lib = dlopen(argv[2], RTLD_NOW);
        if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
        harness = (sb_harness_fn)dlsym(lib, "spmv_harness_");
    } else if (spmv_harness_) {
        harness = spmv_harness_;
    }
    if (!harness) { fprintf(stderr, "no backend\n"); return 2; }
    int n, nnz, *ptr, *idx;
    double *val;
    if (sb_read_crs(argv[1], &n, &nnz, &ptr, &idx, &val)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 3; }
    double *x = (double *)malloc(sizeof(double) * ((size_t)n + 1));
    double hist[100];
    sb_bicg_result r;
    sb_bicg(n, val, ptr, idx, harness, 100, 1e-6, x, hist, &r);
    printf(" Iterative method BiCG chosen:\n");
    printf(" Size: %d  nnz: %d\n", n, nnz);
    printf(" Iterations: %d   residual %.6e -> %.6e\n", r.its, r.rnorm0, r.rnorm);
    printf(" Matrix multiply\n   Total time: %.6f   calls: %d\n", r.t_matprod, r.matprod_calls);
    printf(" Solver time: %.6f\n", r.t_iter);
    return 0;
}
