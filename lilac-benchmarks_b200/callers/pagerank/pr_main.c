/*
 * pr_main.c -- `pagerank libX-spmv.so platform-label graph.mtx`, the command
 * line of pagerank/main.cpp:171-188; prints `platform,PageRank,impl,matrix,t1..t5`.
 */
#include "pagerank.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv)
{
    if (argc != 4) { fprintf(stderr, "usage: %s libX-spmv.so label graph.mtx\n", argv[0]); return 1; }
    void *lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    pr_harness_fn harness = (pr_harness_fn)dlsym(lib, "spmv_harness_");
    if (!harness) { fprintf(stderr, "%s\n", dlerror()); return 3; }
    int n, nnz, *rowstr, *colidx;
    double *a;
    if (pr_load_mtx(argv[3], 0.85, &n, &nnz, &rowstr, &colidx, &a)) { fprintf(stderr, "cannot read %s\n", argv[3]); return 4; }
    double *x = (double *)malloc(sizeof(double) * (size_t)n), *y = (double *)calloc((size_t)n, sizeof(double));
    double s = 0.0;
    srand(12345);
    for (int i = 0; i < n; ++i) { x[i] = (double)rand() / RAND_MAX; s += x[i]; }
    for (int i = 0; i < n; ++i) x[i] /= s;
    const char *base = strrchr(argv[1], '/');
    base = base ? base + 1 : argv[1];
    printf("%s,PageRank,%s,%s", argv[2], base, argv[3]);
    for (int run = 0; run < 5; ++run) {
        double sec;
        pr_power_iterations(n, a, rowstr, colidx, x, y, 0.85, 1024, harness, &sec);
        printf(",%f", sec);
    }
    printf("\n");
    return 0;
}
