/*
 * b200_npb.h -- the NPB CG test matrix assembled on the device (SURVEY.md section 8f,
 * "next" row 2: on-device / sharded makea for classes D and E).
 *
 * The reference generates the matrix on the host, sequentially, inside the benchmark
 * (NPB3.3.1/CG/cg.f:650-905 makea / sparse); for class D that is minutes of host time
 * to set up a 1.5 ms product, and class E cannot be built on one host at all
 * (CG/globals.h:80-82).  Only the random stream that draws the n generating sparse
 * vectors is sequential; the caller draws them on the host (callers/npb/makea.c,
 * npb_vectors_get) and this entry point assembles rows [row_lo, row_hi) of
 * A = sum_i size_i v_i v_i^T + (rcond - shift) I on the current CUDA device, bit for
 * bit what cg.f produces (same additions in the same order), as a 1-based CSR in
 * DEVICE memory ready for b200_spmv_upload_device().  No PCIe traffic beyond the
 * vectors (12 bytes per vector entry).
 */
#ifndef B200_NPB_H
#define B200_NPB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int      rows;       /* row_hi - row_lo */
    int64_t  nnz;
    int     *d_rowstr;   /* rows + 1, 1-based local offsets (device) */
    int     *d_colidx;   /* nnz, 1-based global columns, sorted within a row (device) */
    double  *d_a;        /* nnz (device) */
} b200_npb_csr;

/* na          matrix order; ld = nonzer + 1 (leading dimension of acol / aelt)
 * arow[na]    entries of generating vector i (nonzer or nonzer + 1, cg.f:709-718)
 * acol, aelt  [na * ld] their 1-based positions and values
 * size[na]    size_i = ratio^i accumulated by repeated multiplication as cg.f:876 does
 * All four are HOST arrays.  Returns 0; -1 when the block breaks the int32 ABI (use more
 * row blocks); -2 on bad arguments; -4 on a CUDA error. */
int b200_npb_makea_device(int na, int ld, const int *arow, const int *acol, const double *aelt,
                          const double *size, double rcond, double shift, int row_lo, int row_hi,
                          b200_npb_csr *out);
void b200_npb_csr_free(b200_npb_csr *m);
/* copy the block to HOST arrays of rows + 1, nnz and nnz elements (for the CPU checker) */
int  b200_npb_csr_to_host(const b200_npb_csr *m, int *rowstr, int *colidx, double *a);

#ifdef __cplusplus
}
#endif
#endif
