/*
 * cg_device.cu -- NPB CG with every vector resident in HBM (include/b200_cg.h).
 *
 * Follows NPB3.3.1/CG/cg.f: main loop :233-352, conj_grad :447-644.  Per CG
 * iteration the reference does (host loops over na elements):
 *     q = A p                      :531-532  -> the resident SpMV kernel
 *     d = p.q                      :573-576  -> cg_dot_kernel
 *     alpha = rho/d; z += alpha p; r -= alpha q; rho' = r.r     :581-604
 *                                            -> cg_update_zr_kernel (fused)
 *     beta = rho'/rho; p = r + beta p        :609-616 -> cg_update_p_kernel
 * Scalars stay on the device: each consumer kernel re-reduces the producer's
 * per-block partial sums in a fixed order, so no host round trip and no
 * atomics (bitwise reproducible run to run).
 */
#include "../../include/b200_cg.h"

#include <cuda_runtime.h>

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

namespace {

constexpr int kBlocks = 296;      /* 2 per SM on a 148-SM B200 */
constexpr int kThreads = 256;

#define CG_OK(call)                                                            \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            fprintf(stderr, "libb200-spmv: fatal: %s failed at %s:%d: %s\n", #call, __FILE__,     \
                    __LINE__, cudaGetErrorString(e_));                         \
            abort();                                                           \
        }                                                                      \
    } while (0)

__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double red[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += red[w];
    }
    __syncthreads();
    return s;                      /* valid in thread 0 */
}

/* every thread gets the sum of the partial array, reduced by the whole block in a fixed
 * order (pairs kThreads apart, xor-shuffle tree, warp sums left to right): the same bits on
 * every launch and in every block, and ~6x shorter than one thread adding kBlocks values */
__device__ __forceinline__ double sum_partials(const double *__restrict__ partial)
{
    static_assert(kBlocks <= 2 * kThreads, "two partials per thread at most");
    __shared__ double wsum[kThreads / 32];
    __shared__ double total;
    const int t = threadIdx.x;
    double v = t < kBlocks ? partial[t] : 0.0;
    if (t + kThreads < kBlocks) v += partial[t + kThreads];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((t & 31) == 0) wsum[t >> 5] = v;
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += wsum[w];
        total = s;
    }
    __syncthreads();
    const double r = total;
    __syncthreads();
    return r;
}

/* cg.f:484-498: q = 0, z = 0, r = x, p = r; partial <- r.r */
__global__ void __launch_bounds__(kThreads)
cg_init_kernel(const double *__restrict__ x, double *z, double *p, double *q, double *r, int n,
               double *__restrict__ partial)
{
    double acc = 0.0;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += kBlocks * kThreads) {
        const double xi = x[i];
        q[i] = 0.0; z[i] = 0.0; r[i] = xi; p[i] = xi;
        acc += xi * xi;
    }
    const double s = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(kThreads)
cg_dot_kernel(const double *__restrict__ x, const double *__restrict__ y, int n,
              double *__restrict__ partial)
{
    double acc = 0.0;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += kBlocks * kThreads) acc += x[i] * y[i];
    const double s = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

/* out[0] = sum(partial) */
__global__ void cg_finish_kernel(const double *__restrict__ partial, double *out)
{
    const double t = sum_partials(partial);
    if (threadIdx.x == 0) out[0] = t;
}

/* rho and d given either as finished device scalars or as partial arrays */
__global__ void __launch_bounds__(kThreads)
cg_update_zr_kernel(double *z, double *r, const double *__restrict__ p, const double *__restrict__ q,
                    int n, const double *rho_scalar, const double *d_partial, const double *d_scalar,
                    double *__restrict__ partial_out)
{
    const double d = d_scalar ? d_scalar[0] : sum_partials(d_partial);
    const double alpha = rho_scalar[0] / d;                       /* cg.f:581 */
    double acc = 0.0;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += kBlocks * kThreads) {
        z[i] = z[i] + alpha * p[i];                               /* cg.f:593-596 */
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        acc += ri * ri;                                           /* cg.f:602-604 */
    }
    const double s = block_sum(acc);
    if (threadIdx.x == 0) partial_out[blockIdx.x] = s;
}

__global__ void __launch_bounds__(kThreads)
cg_update_p_kernel(double *p, const double *__restrict__ r, int n, const double *rho_partial,
                   const double *rho_new_scalar, const double *rho_old, double *rho_next)
{
    const double rho_new = rho_new_scalar ? rho_new_scalar[0] : sum_partials(rho_partial);
    const double beta = rho_new / rho_old[0];                     /* cg.f:609 */
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += kBlocks * kThreads)
        p[i] = r[i] + beta * p[i];                                /* cg.f:614-616 */
    if (rho_next && blockIdx.x == 0 && threadIdx.x == 0) rho_next[0] = rho_new;
}

/* cg.f:633-639: partial <- sum (x - r)^2 */
__global__ void __launch_bounds__(kThreads)
cg_resid_kernel(const double *__restrict__ x, const double *__restrict__ r, int n,
                double *__restrict__ partial)
{
    double acc = 0.0;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += kBlocks * kThreads) {
        const double d = x[i] - r[i];
        acc += d * d;
    }
    const double s = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

/* cg.f:315-321: partials of x.z and z.z */
__global__ void __launch_bounds__(kThreads)
cg_norms_kernel(const double *__restrict__ x, const double *__restrict__ z, int n,
                double *__restrict__ partial_xz, double *__restrict__ partial_zz)
{
    double a = 0.0, b = 0.0;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += kBlocks * kThreads) {
        const double zi = z[i];
        a += x[i] * zi;
        b += zi * zi;
    }
    const double sa = block_sum(a);
    const double sb = block_sum(b);
    if (threadIdx.x == 0) { partial_xz[blockIdx.x] = sa; partial_zz[blockIdx.x] = sb; }
}

/* cg.f:324-346: out = {x.z, z.z, sum of the residual partials}; x = z / sqrt(z.z) */
__global__ void __launch_bounds__(kThreads)
cg_scale_x_kernel(double *x, const double *__restrict__ z, int n, const double *partial_xz,
                  const double *partial_zz, const double *partial_res, double *out)
{
    const double xz = sum_partials(partial_xz);
    const double zz = sum_partials(partial_zz);
    const double rs = sum_partials(partial_res);
    const double s = 1.0 / sqrt(zz);
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += kBlocks * kThreads) x[i] = s * z[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = xz; out[1] = zz; out[2] = rs; }
}

__global__ void cg_fill_kernel(double *x, int n, double v)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = v;
}

double wall(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

struct CgBuffers {
    double *x, *z, *p, *q, *r;
    double *part_a, *part_b, *part_res, *part_xz, *part_zz;   /* kBlocks each */
    double *rho;                                              /* [2] ping-pong */
    double *out;                                              /* [4] */
};

/* one conj_grad call (cg.f:447-644) + the zeta / normalisation step (:315-346) */
void enqueue_outer_iteration(b200_matrix *m, const CgBuffers &b, int n, cudaStream_t s,
                             int *spmv_count, int *vec_count)
{
    const int ndot = b200_spmv_dot_partials(m);
    const char *fd = getenv("B200_CG_FUSED_DOT");                 /* 0: keep the separate dot kernel */
    const bool fused_dot = ndot > 0 && ndot <= kBlocks && !(fd && atoi(fd) == 0);
    cg_init_kernel<<<kBlocks, kThreads, 0, s>>>(b.x, b.z, b.p, b.q, b.r, n, b.part_a);
    cg_finish_kernel<<<1, kThreads, 0, s>>>(b.part_a, b.rho + 0);
    if (fused_dot)      /* entries [ndot, kBlocks) of part_a must read zero for the fused dot */
        cudaMemsetAsync(b.part_a, 0, kBlocks * sizeof(double), s);
    *vec_count += 2;
    for (int cgit = 0; cgit < 25; ++cgit) {
        const double *rho_old = b.rho + (cgit & 1);
        double *rho_next = b.rho + ((cgit + 1) & 1);
        /* q = A p and d = p.q: in one launch when the product kernel has the fused epilogue
         * (its CTAs fill the first entries of part_a, the rest stay zero) */
        if (fused_dot) {
            b200_spmv_exec_dot(m, b.p, b.q, b.p, b.part_a, (void *)s);
        } else {
            b200_spmv_exec(m, b.p, b.q, (void *)s);                                  /* q = A p */
            cg_dot_kernel<<<kBlocks, kThreads, 0, s>>>(b.p, b.q, n, b.part_a);        /* d = p.q */
            *vec_count += 1;
        }
        cg_update_zr_kernel<<<kBlocks, kThreads, 0, s>>>(b.z, b.r, b.p, b.q, n, rho_old, b.part_a,
                                                         nullptr, b.part_b);
        cg_update_p_kernel<<<kBlocks, kThreads, 0, s>>>(b.p, b.r, n, b.part_b, nullptr, rho_old,
                                                        rho_next);
        *spmv_count += 1;
        *vec_count += 2;
    }
    b200_spmv_exec(m, b.z, b.r, (void *)s);                                          /* r = A z */
    cg_resid_kernel<<<kBlocks, kThreads, 0, s>>>(b.x, b.r, n, b.part_res);
    cg_norms_kernel<<<kBlocks, kThreads, 0, s>>>(b.x, b.z, n, b.part_xz, b.part_zz);
    cg_scale_x_kernel<<<kBlocks, kThreads, 0, s>>>(b.x, b.z, n, b.part_xz, b.part_zz, b.part_res, b.out);
    *spmv_count += 1;
    *vec_count += 3;
}

}  // namespace

extern "C" int b200_cg_partials(void) { return kBlocks; }

extern "C" void b200_cg_dot(const double *x, const double *y, int n, double *partial, void *stream)
{
    cg_dot_kernel<<<kBlocks, kThreads, 0, (cudaStream_t)stream>>>(x, y, n, partial);
}

extern "C" void b200_cg_update_zr(double *z, double *r, const double *p, const double *q, int n,
                                  const double *rho, const double *d, double *partial, void *stream)
{
    cg_update_zr_kernel<<<kBlocks, kThreads, 0, (cudaStream_t)stream>>>(z, r, p, q, n, rho, nullptr, d,
                                                                        partial);
}

extern "C" void b200_cg_update_p(double *p, const double *r, int n, const double *rho_new,
                                 const double *rho_old, void *stream)
{
    cg_update_p_kernel<<<kBlocks, kThreads, 0, (cudaStream_t)stream>>>(p, r, n, nullptr, rho_new,
                                                                       rho_old, nullptr);
}

extern "C" void b200_cg_finish(const double *partial, double *out, void *stream)
{
    cg_finish_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(partial, out);
}

extern "C" int b200_cg_npb_run(b200_matrix *m, int nonzer, int niter, double shift, int use_graph,
                               double *zeta_hist, double *rnorm_hist, b200_cg_result *res)
{
    const int n = b200_spmv_rows(m);
    if (b200_spmv_ncols(m) > n || b200_spmv_algorithmic_bytes(m) <= 0) return -1;
    if (b200_spmv_algorithmic_bytes(m) != 12 * b200_spmv_nnz(m) + 4 * ((int64_t)n + 1) +
                                              8 * (int64_t)b200_spmv_ncols(m) + 8 * (int64_t)n)
        return -1;                                   /* not an fp64 matrix */
    cudaStream_t s;
    CG_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CgBuffers b;
    double *pool = nullptr;
    const size_t vec = (size_t)n + 2;                /* cg.f:75-80: na+2 elements each */
    const size_t total = 5 * vec + 5 * (size_t)kBlocks + 2 + 4;
    CG_OK(cudaMalloc((void **)&pool, total * sizeof(double)));
    CG_OK(cudaMemsetAsync(pool, 0, total * sizeof(double), s));
    b.x = pool; b.z = b.x + vec; b.p = b.z + vec; b.q = b.p + vec; b.r = b.q + vec;
    b.part_a = b.r + vec; b.part_b = b.part_a + kBlocks; b.part_res = b.part_b + kBlocks;
    b.part_xz = b.part_res + kBlocks; b.part_zz = b.part_xz + kBlocks;
    b.rho = b.part_zz + kBlocks; b.out = b.rho + 2;

    int spmv_count = 0, vec_count = 0;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;

    /* untimed iteration (cg.f:233-272); it also warms the kernels up before any capture */
    cg_fill_kernel<<<kBlocks, kThreads, 0, s>>>(b.x, n + 1, 1.0);               /* cg.f:216-218 */
    enqueue_outer_iteration(m, b, n, s, &spmv_count, &vec_count);
    CG_OK(cudaStreamSynchronize(s));
    int per_iter_spmv = spmv_count, per_iter_vec = vec_count;
    if (use_graph) {
        int dummy_a = 0, dummy_b = 0;
        CG_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        enqueue_outer_iteration(m, b, n, s, &dummy_a, &dummy_b);
        CG_OK(cudaStreamEndCapture(s, &graph));
        CG_OK(cudaGraphInstantiate(&exec, graph, 0));
    }
    cg_fill_kernel<<<kBlocks, kThreads, 0, s>>>(b.x, n + 1, 1.0);               /* cg.f:280-282 */
    CG_OK(cudaStreamSynchronize(s));

    double out[4] = {0, 0, 0, 0};
    double zeta = 0.0, rnorm = 0.0;
    const double t0 = wall();
    for (int it = 0; it < niter; ++it) {                                        /* cg.f:299-349 */
        if (exec) {
            CG_OK(cudaGraphLaunch(exec, s));
        } else {
            int a_ = 0, b_ = 0;
            enqueue_outer_iteration(m, b, n, s, &a_, &b_);
        }
        CG_OK(cudaMemcpyAsync(out, b.out, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
        CG_OK(cudaStreamSynchronize(s));
        zeta = shift + 1.0 / out[0];                                            /* cg.f:334 */
        rnorm = sqrt(out[2]);
        if (zeta_hist) zeta_hist[it] = zeta;
        if (rnorm_hist) rnorm_hist[it] = rnorm;
    }
    const double t = wall() - t0;

    res->zeta = zeta;
    res->rnorm = rnorm;
    res->seconds = t;
    const double nz1 = (double)nonzer * (double)(nonzer + 1);
    res->mops = t > 0 ? 2.0 * niter * (double)n * (3.0 + nz1 + 25.0 * (5.0 + nz1) + 3.0) / t / 1e6 : 0.0;
    res->spmv_launches = per_iter_spmv * niter;
    res->vector_launches = per_iter_vec * niter;
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    cudaFree(pool);
    cudaStreamDestroy(s);
    return 0;
}
