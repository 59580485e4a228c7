/*
 * cg_main.c -- command-line front end of the C NPB CG caller.
 *
 *   cg CLASS                 backend chosen at LINK time (-l<platform>-spmv,
 *                            as NPB3.3.1/config/make.def:91-109 does)
 *   cg CLASS path/to/lib.so  backend chosen at RUN time by dlopen + dlsym of
 *                            "spmv_harness_" (as pagerank/main.cpp:17-43 does)
 *
 * Prints the reference's result block, so NPB3.3.1/run_all:21's
 * "Time in seconds" scrape works unchanged.
 */
#include "npb_cg.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

extern void *spmv_harness_(double *, double *, double *, int *, int *, int *)
    __attribute__((weak));

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s CLASS [libX-spmv.so]\n", argv[0]);
        return 1;
    }
    npb_cg_class c;
    if (npb_cg_class_lookup(argv[1][0], &c)) {
        fprintf(stderr, "unknown class %s\n", argv[1]);
        return 1;
    }
    spmv_harness_fn harness = NULL;
    if (argc >= 3) {
        void *lib = dlopen(argv[2], RTLD_NOW);
        if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
        harness = (spmv_harness_fn)dlsym(lib, "spmv_harness_");
        if (!harness) { fprintf(stderr, "%s\n", dlerror()); return 3; }
    } else if (spmv_harness_) {
        harness = spmv_harness_;
    } else {
        fprintf(stderr, "no backend linked and no library path given\n");
        return 2;
    }
    npb_csr m;
    int rc = npb_makea(&c, &m);
    if (rc) { fprintf(stderr, "makea failed (%d)\n", rc); return 4; }
    npb_cg_result res = {0};
    rc = npb_cg_run(&c, &m, harness, &res, 1);
    npb_csr_free(&m);
    return rc ? 5 : (res.verified ? 0 : 6);
}
