/*
 * l2_to_smem_probe.cu -- how fast can every SM pull the same L2-resident
 * vector into shared memory with TMA bulk copies, with and without cluster
 * multicast?  (The PANEL kernels re-read x once per CTA; this bounds that
 * traffic.)   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2probe l2_to_smem_probe.cu
 * usage: l2probe [x_megabytes=12] [slice_kb=96] [iters=20]
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n.reg .pred p;\nW:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(bar)), "r"(parity) : "memory");
}

template <int CS>
__global__ void __launch_bounds__(128, 1)
probe(const char *__restrict__ x, size_t xbytes, uint32_t slice, int passes, unsigned long long *sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    unsigned char *buf = smem + 128;
    const int tid = threadIdx.x;
    uint32_t rank = 0;
    if (CS > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (CS > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    const int P = (int)(xbytes / slice);
    unsigned long long acc = 0;
    int it = 0;
    auto issue = [&](int k) {
        const int p = k % P;
        const int b = k & 1;
        unsigned char *dst = buf + (size_t)b * slice;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                     ::"r"(s32(&bars[b])), "r"(slice) : "memory");
        if (CS == 1) {
            for (uint32_t o = 0; o < slice; o += 32768) {
                const uint32_t n = slice - o < 32768 ? slice - o : 32768;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(dst + o)), "l"(x + (size_t)p * slice + o), "r"(n), "r"(s32(&bars[b])) : "memory");
            }
        } else {
            const uint32_t share = slice / CS;             /* multiple of 16 by construction */
            const uint32_t o0 = rank * share;
            for (uint32_t o = 0; o < share; o += 32768) {
                const uint32_t n = share - o < 32768 ? share - o : 32768;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                             ::"r"(s32(dst + o0 + o)), "l"(x + (size_t)p * slice + o0 + o), "r"(n),
                               "r"(s32(&bars[b])), "h"((unsigned short)((1u << CS) - 1)) : "memory");
            }
        }
    };
    const int total = P * passes;
    if (tid == 0) issue(0);
    for (it = 0; it < total; ++it) {
        if (CS > 1) {
            /* every CTA of the cluster is done with the buffer about to be refilled */
            asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        } else {
            __syncthreads();
        }
        if (tid == 0 && it + 1 < total) issue(it + 1);
        mbar_wait(&bars[it & 1], (uint32_t)((it >> 1) & 1));
        acc += *reinterpret_cast<unsigned long long *>(buf + (size_t)(it & 1) * slice + (size_t)tid * 8);
    }
    if (CS > 1) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (acc == 0x1234567ull) sink[0] = acc;
}

template <int CS>
static void run(const char *x, size_t xbytes, uint32_t slice, int iters, int grid, unsigned long long *sink)
{
    const size_t smem = 128 + 2 * (size_t)slice;
    CK(cudaFuncSetAttribute(probe<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) CK(cudaFuncSetAttribute(probe<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid / CS * CS);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, probe<CS>, &cfg);
    const int passes = 4;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaLaunchKernelEx(&cfg, probe<CS>, x, xbytes, slice, passes, sink));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) CK(cudaLaunchKernelEx(&cfg, probe<CS>, x, xbytes, slice, passes, sink));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = (double)(xbytes / slice) * slice * passes * cfg.gridDim.x;
    const double tbs = bytes * iters / (ms * 1e-3) / 1e12;
    printf("cluster=%2d grid=%3d max_active_clusters=%3d (%s) slice=%u KB  delivered %.2f TB/s  (%.1f GB/s per SM, %.1f us per pass)\n",
           CS, cfg.gridDim.x, nclusters, e == cudaSuccess ? "ok" : cudaGetErrorString(e), slice / 1024,
           tbs, tbs * 1e3 / cfg.gridDim.x, ms * 1e3 / iters / passes);
}

int main(int argc, char **argv)
{
    const size_t mb = argc > 1 ? atoi(argv[1]) : 12;
    const uint32_t slice = (argc > 2 ? atoi(argv[2]) : 96) * 1024u;
    const int iters = argc > 3 ? atoi(argv[3]) : 20;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t xbytes = mb << 20;
    char *x; unsigned long long *sink;
    CK(cudaMalloc(&x, xbytes)); CK(cudaMemset(x, 1, xbytes));
    CK(cudaMalloc(&sink, 8));
    printf("SMs=%d x=%zu MB\n", sms, mb);
    run<1>(x, xbytes, slice, iters, sms, sink);
    run<2>(x, xbytes, slice, iters, sms, sink);
    run<4>(x, xbytes, slice, iters, sms, sink);
    run<8>(x, xbytes, slice, iters, sms, sink);
    run<16>(x, xbytes, slice, iters, sms, sink);
    return 0;
}
