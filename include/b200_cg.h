/*
 * b200_cg.h -- device-resident NPB CG on top of the resident-matrix API
 * ("next" row 1).
 *
 * The drop-in symbols keep NPB's vector algebra on the host, so every product
 * pays two PCIe transfers (NPB3.3.1/CG/cg.f:531-532 through libspmv/gpu.c:264,
 * 285).  Here x, z, p, q, r never leave HBM: conj_grad (cg.f:447-644) runs as
 * the SpMV kernel plus three fused vector kernels per CG iteration, the
 * 25-iteration sweep is captured once in a CUDA graph and replayed for every
 * outer iteration, and only (zeta, ||r||) come back per outer iteration.
 * Precedent inside the reference: SNU_NPB/NPB3.3-OCL/CG/cg_gpu.cl:139-371 keeps
 * the whole CG on the device, too.
 *
 * Dot products are block-parallel with a fixed reduction order (deterministic
 * run to run), not the host's left-to-right order: zeta is verified against
 * cg.f:122-166 (1e-10), it is not bit-identical to the host-algebra run.
 */
#ifndef B200_CG_H
#define B200_CG_H

#include "b200_spmv.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    double zeta;            /* final zeta (cg.f:334) */
    double rnorm;           /* final ||x - A z|| (cg.f:633-639) */
    double seconds;         /* timed section, cg.f:292-352 */
    double mops;            /* cg.f:395-402 */
    int    spmv_launches;   /* products launched in the timed section */
    int    vector_launches; /* fused vector kernels launched in the timed section */
} b200_cg_result;

/* Whole NPB CG benchmark (cg.f:53-443) on a resident fp64 matrix of order na:
 * one untimed outer iteration, then `niter` timed ones.  zeta_hist / rnorm_hist
 * (niter doubles each, may be NULL) receive the per-iteration values cg.f
 * prints.  use_graph != 0 replays a captured CUDA graph per conj_grad call.
 * Returns 0, or -1 if the matrix is not square fp64. */
int b200_cg_npb_run(b200_matrix *m, int nonzer, int niter, double shift, int use_graph,
                    double *zeta_hist, double *rnorm_hist, b200_cg_result *res);

/* The fused vector kernels on DEVICE pointers and a stream (used by the
 * multi-GPU driver, where the reductions are completed by an allreduce).
 * `partial` arrays hold b200_cg_partials() doubles. */
int  b200_cg_partials(void);
void b200_cg_dot(const double *x, const double *y, int n, double *partial, void *stream);
/* z += alpha p; r -= alpha q; partial <- blockwise sum of r*r, with
 * alpha = rho / d, both read from device scalars */
void b200_cg_update_zr(double *z, double *r, const double *p, const double *q, int n,
                       const double *rho, const double *d, double *partial, void *stream);
/* p = r + beta p with beta = rho_new / rho_old (device scalars) */
void b200_cg_update_p(double *p, const double *r, int n, const double *rho_new,
                      const double *rho_old, void *stream);
/* out[0] = sum of the partials in fixed order */
void b200_cg_finish(const double *partial, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif
