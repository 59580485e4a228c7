/*
 * spmv_small.cu -- SMALL: the whole x in shared memory, one row block per SM
 *
 * For matrices that sit in L2 and whose x fits one SM's shared memory beside a tile of
 * products (NPB classes S, W, A; parboil's and bfs's inputs -- BASELINE config 1): the
 * launch-bound regime, where a kernel is as fast as its longest chain of dependent memory
 * round trips.  Arithmetic as libspmv/native-impl.c:1-12 (fp64) / :14-25 (fp32): every
 * product rounded, every row added left to right from 0 -- for ANY column order.
 *
 * One wave, one CTA per SM, an nnz-balanced row block each (build_small_locked).  The CTA
 *  (1) requests the first two batches of its slice of the matrix -- the values as uploaded,
 *      in pairs (one 128-bit load per two entries), the columns from a 16-bit 0-based copy
 *      made at upload (one 32-bit load per two entries; 10 instead of 12 bytes per entry) --
 *      BEFORE griddepcontrol.wait: the matrix is immutable, so under programmatic dependent
 *      launch these loads fly while the kernel in front drains.  A block that starts or ends
 *      on an odd entry walks the neighbour's entry of that pair too and never adds it;
 *  (2) brings x in with ONE TMA bulk copy per 64 KB (cp.async.bulk + mbarrier; SASS UBLKCP):
 *      no registers, no per-thread round trips; x vectors that are not 16-byte aligned take a
 *      cooperative loop instead;
 *  (3) forms the products with every thread, x gathered from shared memory, into a
 *      shared-memory tile, the next batch always requested before the current one is used;
 *  (4) adds each row's products left to right, one thread per row.
 * Only (4) is serial, and its chain is the longest row -- against one lane walking loads,
 * gathers and additions of whole rows in the panel kernels.
 *
 * Measured (launches from a C loop, profiles/r02_run40 / run42 / run44): class A 15.2 us on the
 * panel layout -> 8.4 us (scalar loads, cooperative x) -> 8.3 us (TMA x: the x round trips
 * were not the limit) -> 7.7 us (pairs + 16-bit columns: the ncu capture r02_run43 showed 26 %
 * of the stall samples in the load/store queue -- 24 scalar loads per thread); classes W / S
 * 4.3 / 3.0-3.8 us.  What is left for class A: 147 CTAs x (112 KB of x + 126 KB of matrix)
 * = 35 MB through the L2 per launch, and the serial additions of the longest row.
 */
#include "panel_common.cuh"

namespace b200 {

/* dynamic shared memory: [0, 16) the x barrier, [16, 272) 32 warp sums, then x, then the products */
constexpr int kSmallHeader = 16 + 32 * 8;

/* column pairs: int2 of the 1-based columns as uploaded, or ushort2 of the 0-based 16-bit
 * private copy (x fits shared memory, so there are fewer than 65 536 columns) */
template <typename CT> struct ColPair;
template <> struct ColPair<int>      { using type = int2;    static constexpr int base = 1; };
template <> struct ColPair<uint16_t> { using type = ushort2; static constexpr int base = 0; };

template <typename T, typename CT, int U>
struct SmallBatch {
    typename PairT<T>::type v[U];
    typename ColPair<CT>::type c[U];
};

/* pair p of the block = entries 2p, 2p + 1 counted from the even entry at or below the block's
 * first one: 128-bit value loads and 64- / 32-bit column loads, coalesced, read once */
template <typename T, typename CT, int THREADS, int U>
__device__ __forceinline__ void small_load(SmallBatch<T, CT, U> &b, const typename PairT<T>::type *__restrict__ v2,
                                           const typename ColPair<CT>::type *__restrict__ c2, int p0, int np)
{
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int p = p0 + u * THREADS;
        if (p < np) {
            b.v[u] = __ldcs(v2 + p);
            b.c[u] = __ldcs(c2 + p);
        }
    }
}

template <typename T, typename CT, int THREADS, int U>
__device__ __forceinline__ void small_products(const SmallBatch<T, CT, U> &b, const T *xs,
                                               typename PairT<T>::type *prod2, int p0, int np)
{
    constexpr int cb = ColPair<CT>::base;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int p = p0 + u * THREADS;
        if (p < np) {
            typename PairT<T>::type r;
            r.x = pmul(b.v[u].x, xs[(int)b.c[u].x - cb]);
            r.y = pmul(b.v[u].y, xs[(int)b.c[u].y - cb]);
            prod2[p] = r;
        }
    }
}

template <typename T, typename CT, int THREADS, int U>
__global__ void __launch_bounds__(THREADS, 1)
spmv_small_kernel(const T *__restrict__ val, const CT *__restrict__ col,
                  const int *__restrict__ rowptr, const int *__restrict__ rowblk,
                  const T *__restrict__ x, T *__restrict__ y, int ncols, int xpad, int use_tma,
                  const T *__restrict__ dotv, T *__restrict__ dot_partial)
{
    using P2 = typename PairT<T>::type;
    using C2 = typename ColPair<CT>::type;
    extern __shared__ __align__(16) unsigned char small_smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(small_smem);          /* header: barrier, warp sums */
    T *red = reinterpret_cast<T *>(small_smem + 16);                   /* [32] */
    T *xs = reinterpret_cast<T *>(small_smem + kSmallHeader);
    T *prod = xs + xpad;                                               /* 16-byte aligned: xpad % 4 == 0 */
    constexpr int STEP = THREADS * U;
    /* a kernel behind this one in the stream may start its CTAs as ours finish */
    asm volatile("griddepcontrol.launch_dependents;");
    const int tid = threadIdx.x;
    const int r0 = __ldg(rowblk + blockIdx.x), r1 = __ldg(rowblk + blockIdx.x + 1);
    const int lo = __ldg(rowptr + r0), n = __ldg(rowptr + r1) - lo;
    /* walk whole pairs: the entry in front of an odd start and the one behind an odd end are
     * multiplied too (entries of the neighbouring blocks, or the zero padding behind the
     * arrays) and never added */
    const int base = lo & ~1, head = lo - base;
    const int np = (n + head + 1) >> 1;
    const P2 *v2 = reinterpret_cast<const P2 *>(val + base);
    const C2 *c2 = reinterpret_cast<const C2 *>(col + base);
    P2 *prod2 = reinterpret_cast<P2 *>(prod);

    SmallBatch<T, CT, U> a, b;
    small_load<T, CT, THREADS, U>(a, v2, c2, tid, np);
    small_load<T, CT, THREADS, U>(b, v2, c2, tid + STEP, np);
    if (use_tma && tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    /* everything above read only the resident matrix; x and y belong to the kernel in front */
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (use_tma) {
        if (tid == 0) {
            constexpr int VE = 16 / (int)sizeof(T);
            const int al = ncols & ~(VE - 1);
            for (int i = al; i < ncols; ++i) xs[i] = __ldcg(x + i);   /* ragged tail (other bytes than the copy's) */
            uint32_t left = (uint32_t)al * (uint32_t)sizeof(T);
            mbar_expect_tx(bar, left);
            const char *src = reinterpret_cast<const char *>(x);
            char *d = reinterpret_cast<char *>(xs);
            while (left) {
                const uint32_t m = left > 65536u ? 65536u : left;
                tma_bulk_g2s(d, src, m, bar);
                d += m; src += m; left -= m;
            }
        }
    } else {
        for (int i = tid; i < ncols; i += THREADS) xs[i] = __ldcg(x + i);
    }
    __syncthreads();                      /* barrier initialised, tail / cooperative x visible */
    if (use_tma) mbar_wait(bar, 0);

    for (int p0 = tid; p0 < np; p0 += 2 * STEP) {
        small_products<T, CT, THREADS, U>(a, xs, prod2, p0, np);
        small_load<T, CT, THREADS, U>(a, v2, c2, p0 + 2 * STEP, np);
        small_products<T, CT, THREADS, U>(b, xs, prod2, p0 + STEP, np);
        small_load<T, CT, THREADS, U>(b, v2, c2, p0 + 3 * STEP, np);
    }
    __syncthreads();
    T dacc = (T)0;
    for (int r = r0 + tid; r < r1; r += THREADS) {
        const int s = __ldg(rowptr + r) - base, e = __ldg(rowptr + r + 1) - base;
        T acc = (T)0;
#pragma unroll 8
        for (int k = s; k < e; ++k) acc = padd(acc, prod[k]);
        y[r] = acc;
        if (dotv) dacc += acc * dotv[r];
    }
    /* optionally the block's share of dotv . y on the way out (NPB conj_grad's d = p.q,
     * cg.f:573-576, as the PANEL epilogue): fixed order -- per-thread stride, xor-shuffle tree,
     * warp sums left to right --, the same result on every launch */
    if (dotv) {                                           /* block-uniform */
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
        if ((tid & 31) == 0) red[tid >> 5] = dacc;
        __syncthreads();
        if (tid == 0) {
            T tot = (T)0;
            for (int w = 0; w < THREADS / 32; ++w) tot += red[w];
            dot_partial[blockIdx.x] = tot;
        }
    }
}

template <typename T, typename CT, int THREADS, int U>
static void launch_small_cfg(const DevSmall &sm, const DevCsr &m, const CT *col, const T *x, T *y, cudaStream_t s,
                             const T *dotv, T *dot_partial)
{
    static unsigned attr_mask = 0;
    if (!attr_done(&attr_mask))
        cudaFuncSetAttribute(spmv_small_kernel<T, CT, THREADS, U>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024);
    /* + 4: the two entries of the neighbouring blocks that whole pairs may bring along */
    const size_t smem = kSmallHeader + ((size_t)sm.xpad + (size_t)sm.tile + 4) * sizeof(T);
    const int use_tma = sm.use_tma && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)sm.nblk);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    /* dependent launches outside a graph capture only (as the PANEL kernel) */
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusActive; }
    cfg.numAttrs = (cap == cudaStreamCaptureStatusNone && sm.pdl) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, spmv_small_kernel<T, CT, THREADS, U>, static_cast<const T *>(m.val), col,
                       (const int *)m.rowptr, (const int *)sm.rowblk, x, y, sm.ncols, sm.xpad, use_tma,
                       dotv, dot_partial);
}

template <typename T, typename CT>
static void launch_small_cols(const DevSmall &sm, const DevCsr &m, const CT *col, const T *x, T *y, cudaStream_t s,
                              const T *dotv, T *dot_partial)
{
    if (sm.cfg & 1) launch_small_cfg<T, CT, 512, 6>(sm, m, col, x, y, s, dotv, dot_partial);
    else            launch_small_cfg<T, CT, 1024, 3>(sm, m, col, x, y, s, dotv, dot_partial);
}

/* 16-bit 0-based copy of the columns, padding included (upload) */
__global__ void small_col16_kernel(const int *__restrict__ col, uint16_t *__restrict__ col16, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) col16[i] = (uint16_t)(col[i] - 1);
}

void launch_small_col16(const int *col, uint16_t *col16, size_t n, cudaStream_t s)
{
    if (n == 0) return;
    small_col16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(col, col16, n);
}

template <typename T>
void launch_small(const DevSmall &sm, const DevCsr &m, const T *x, T *y, cudaStream_t s, const T *dotv,
                  T *dot_partial)
{
    if (sm.nblk <= 0) return;
    if (sm.col16) launch_small_cols<T, uint16_t>(sm, m, sm.col16, x, y, s, dotv, dot_partial);
    else          launch_small_cols<T, int>(sm, m, (const int *)m.col, x, y, s, dotv, dot_partial);
}
template void launch_small<double>(const DevSmall &, const DevCsr &, const double *, double *, cudaStream_t,
                                   const double *, double *);
template void launch_small<float>(const DevSmall &, const DevCsr &, const float *, float *, cudaStream_t,
                                  const float *, float *);

}  // namespace b200
