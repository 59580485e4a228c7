/*
 * panel_common.cuh -- device primitives shared by the PANEL kernels
 * (spmv_panel.cu: two rows per lane, paired; spmv_panelg.cu: G rows per lane,
 * flagged streams): separately rounded multiply / add, mbarrier + TMA bulk
 * copies, the prefetch chunk and the read cursor over a lane stream.
 */
#pragma once
#include "spmv_kernels.cuh"

namespace b200 {

__device__ __forceinline__ double pmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  pmul(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double padd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  padd(float a, float b)   { return __fadd_rn(a, b); }

template <typename T> struct PairT;
template <> struct PairT<double> { using type = double2; };
template <> struct PairT<float>  { using type = float2; };

/* ---- mbarrier / TMA bulk-copy primitives (sm_90+ PTX) --------------------- */
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
/* global -> shared bulk copy; bytes multiple of 16, both addresses 16-byte aligned */
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

/* after a __threadfence_system(): fence + relaxed store is a release without a fence per store */
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

/* x that arrives slice by slice (XFlags: from other GPUs over NVLink, or from the host through
 * the copy engine while the product runs): before a panel's x slice is requested, the slices
 * holding its columns [cbase, cbase + cw) must have landed.  Columns are walked left to right,
 * so `ready` (slices [0, ready) have arrived) only grows.  One thread calls this.
 * The spin is out of line (it is off the common path and the paired kernel has no register to
 * spare) and, when the launcher asks for it, guarded by a watchdog: a product that waits for
 * copies issued on ANOTHER stream never ends if something serialises the streams (a profiler
 * replaying kernels, CUDA_LAUNCH_BLOCKING, one hardware queue) -- after timeout_ns it reports
 * *timed_out = 1 and goes on with whatever x holds; the host redoes the call the plain way. */
static __device__ __noinline__ int wait_x_slices_spin(const unsigned long long *flags, unsigned long long epoch,
                                                      int ready, int r1, int *timed_out,
                                                      unsigned long long timeout_ns)
{
    unsigned long long t0 = 0;
    for (; ready <= r1; ++ready) {
        while (ld_acquire_sys_u64(flags + ready) < epoch) {
            if (timed_out) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > timeout_ns) {
                    *reinterpret_cast<volatile int *>(timed_out) = 1;
                    return 1 << 30;                  /* this CTA waits for nothing any more */
                }
            }
        }
    }
    return ready;
}

__device__ __forceinline__ void wait_x_slices(const XFlags &xf, int &ready, int cbase, int cw)
{
    const int r1 = min((cbase + cw - 1) / xf.cols_per_rank, xf.nranks - 1);
    if (r1 < ready) return;
    ready = wait_x_slices_spin(xf.flags, xf.epoch, ready, r1, xf.timed_out, xf.timeout_ns);
    /* the slice was written through the generic proxy (another GPU) or by a copy engine, the
     * bulk copy reads it through the async proxy */
    asm volatile("fence.proxy.async.global;" ::: "memory");
}

/* host side: true when the calling device has been seen before (bit per device ordinal);
 * used to set function attributes once per device (callers hold the library lock, or
 * race benignly: setting an attribute twice is harmless) */
inline bool attr_done(unsigned *mask)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned bit = 1u << (dev & 31);
    if (*mask & bit) return true;
    *mask |= bit;
    return false;
}

/* loads of the matrix stream: read once.  B200_STREAM_LD selects the cache operator:
 * 1 = ld.global.nc.L1::no_allocate (SASS LDG.E.NA; class C 74.4 us), 0 = ld.global.cs
 * (evict first; 75.7 us), 2 = plain non-coherent (77.9 us) -- profiles/r01_run42_sweep_ld_operator.txt */
#ifndef B200_STREAM_LD
#define B200_STREAM_LD 1
#endif
__device__ __forceinline__ double2 ld_matrix_stream(const double2 *p)
{
#if B200_STREAM_LD == 1
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
#elif B200_STREAM_LD == 2
    return __ldg(p);
#else
    return __ldcs(p);
#endif
}
__device__ __forceinline__ float2 ld_matrix_stream(const float2 *p) { return __ldcs(p); }
__device__ __forceinline__ uint32_t ld_matrix_stream(const uint32_t *p)
{
#if B200_STREAM_LD == 1
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
#elif B200_STREAM_LD == 2
    return __ldg(p);
#else
    return __ldcs(p);
#endif
}

template <typename T, int U>
struct Chunk {
    typename PairT<T>::type v[U];
    uint32_t c[U];
};

template <typename T, int U>
__device__ __forceinline__ void load_chunk(Chunk<T, U> &ch, const typename PairT<T>::type *vp,
                                           const uint32_t *cp, int kp, int npair)
{
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (kp + u < npair) {
            ch.v[u] = ld_matrix_stream(vp + (size_t)(kp + u) * 32);
            ch.c[u] = ld_matrix_stream(cp + (size_t)(kp + u) * 32);
        }
    }
}

/* pairs walked for a slice of npair pairs: whole double chunks, and one double chunk
 * even for an empty slice -- the consumer loop runs once for it too, so cursor and
 * consumer always spend the same number of chunks per panel */
template <int U>
__device__ __forceinline__ int cursor_rounds(int npair)
{
    return (max(npair, 1) + 2 * U - 1) / (2 * U) * (2 * U);
}

/* Read cursor over a lane stream: walks the pairs of panel 0, 1, ... of this
 * warp's slices in chunks of U pairs, so that the matrix stream stays
 * requested ahead of its use across panel boundaries.  Every panel is walked
 * in an even number of chunks (consumption alternates two register sets). */
struct StreamCursor {
    int p;          /* panel being requested */
    int kp;         /* next pair inside that panel's slice */
    int npair;      /* pairs of that slice */
    int nround;     /* pairs walked for that slice: npair rounded up to 2U */
    size_t base;    /* pair offset of the slice + lane */
};

template <typename T, int U>
__device__ __forceinline__ void cursor_load(Chunk<T, U> &ch, StreamCursor &cur,
                                            const typename PairT<T>::type *val2,
                                            const uint32_t *col2, const int2 *s_slice, int spb,
                                            int warp, int lane, int P)
{
    if (cur.p < P) {
        load_chunk<T, U>(ch, val2 + cur.base, col2 + cur.base, cur.kp, cur.npair);
        cur.kp += U;
        if (cur.kp >= cur.nround) {
            ++cur.p;
            if (cur.p < P) {
                const int2 so = s_slice[cur.p * spb + warp];
                cur.kp = 0;
                cur.npair = so.y;
                cur.nround = cursor_rounds<U>(so.y);
                cur.base = (size_t)(so.x >> 1) + lane;
            }
        }
    }
}

/* ---- flagged streams (spmv_panelg.cu, spmv_panelr.cu) ---------------------- */
/* the G row ids of a lane, 16 bits each, consumed front to back */
template <int G> struct RowIds { uint32_t w[G / 2]; };

template <int G>
__device__ __forceinline__ RowIds<G> load_ids(const uint16_t *p)
{
    RowIds<G> r;
    if constexpr (G == 2) {
        r.w[0] = __ldg(reinterpret_cast<const uint32_t *>(p));
    } else if constexpr (G == 4) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        r.w[0] = v.x; r.w[1] = v.y;
    } else {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
        r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
    }
    return r;
}

template <int G>
__device__ __forceinline__ int pop_id(RowIds<G> &ids)
{
    const int r = (int)(ids.w[0] & 0xFFFFu);
#pragma unroll
    for (int i = 0; i + 1 < G / 2; ++i) ids.w[i] = __funnelshift_r(ids.w[i], ids.w[i + 1], 16);
    ids.w[G / 2 - 1] >>= 16;
    return r;
}

}  // namespace b200
