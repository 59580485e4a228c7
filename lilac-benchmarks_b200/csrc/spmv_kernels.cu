/*
 * spmv_kernels.cu -- hand-written sm_100a kernels of the b200 libspmv platform.
 *
 * The arithmetic being replaced is libspmv/native-impl.c:1-25 (CPU loop) and
 * the closed-source cusparse{D,S}csrmv_mp called at libspmv/gpu.c:270,350.
 *
 * ORDERED (spmv_stream_ordered)
 *   nnz-split: the matrix is cut at upload into row blocks of at most one
 *   shared-memory tile of nonzeros.  Phase 1: the CTA streams its contiguous
 *   slice of val / col with 128-bit / 64-bit coalesced, L1-bypassing,
 *   evict-first loads, gathers x through the read-only path (x lives in L2),
 *   forms the products with a separately rounded multiply and parks them in
 *   shared memory.  Phase 2: one thread per row adds that row's products
 *   strictly left to right with a separately rounded add.  Same operations
 *   in the same order as the reference loop => bit-identical y.
 *   A row longer than a tile is reduced by the whole CTA (tree order).
 *
 * VECTOR (spmv_vector)
 *   2..32 lanes per row, lane-strided partial sums, xor-shuffle reduction.
 *   Re-orders the sum; kept for rows/shapes where a per-row chain is too long.
 *
 * Tensor cores are deliberately unused: SpMV is a bandwidth-bound gather.
 */
#include "spmv_kernels.cuh"

#include <algorithm>

namespace b200 {

/* ---- rounding-exact scalar ops (never contracted into FMA) -------------- */
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  mul_rn(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  add_rn(float a, float b)   { return __fadd_rn(a, b); }

template <typename T> struct Pair;
template <> struct Pair<double> { using type = double2; };
template <> struct Pair<float>  { using type = float2; };

/* streaming (evict-first, no L1 allocation) loads for the matrix stream */
__device__ __forceinline__ double2 ld_stream(const double2 *p) { return __ldcs(p); }
__device__ __forceinline__ float2  ld_stream(const float2 *p)  { return __ldcs(p); }
__device__ __forceinline__ int2    ld_stream(const int2 *p)    { return __ldcs(p); }
__device__ __forceinline__ double  ld_stream(const double *p)  { return __ldcs(p); }
__device__ __forceinline__ float   ld_stream(const float *p)   { return __ldcs(p); }
__device__ __forceinline__ int     ld_stream(const int *p)     { return __ldcs(p); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = add_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

/* ------------------------------------------------------------------------
 * ORDERED: nnz-split row blocks, shared-memory product tile, in-order rows
 * ---------------------------------------------------------------------- */
template <typename T, int THREADS, int TILE>
__global__ void __launch_bounds__(THREADS)
spmv_stream_ordered(const T *__restrict__ val, const int *__restrict__ col,
                    const int *__restrict__ rowptr, const int *__restrict__ rowblk,
                    const T *__restrict__ xm1,   /* x - 1: indexed by 1-based col */
                    T *__restrict__ y)
{
    using P = typename Pair<T>::type;
    constexpr int U = 4;                         /* independent load batches per thread */
    __shared__ __align__(16) T prod[TILE + 2];
    __shared__ T red[THREADS / 32];

    const int tid = threadIdx.x;
    const int r0 = rowblk[blockIdx.x], r1 = rowblk[blockIdx.x + 1];
    const int lo = rowptr[r0], hi = rowptr[r1];

    if (hi - lo > TILE) {
        /* a single row longer than the tile: CTA-wide reduction (re-ordered) */
        T acc = (T)0;
        for (int i = lo + tid; i < hi; i += THREADS)
            acc = add_rn(acc, mul_rn(ld_stream(val + i), __ldg(xm1 + ld_stream(col + i))));
        acc = warp_sum(acc);
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            T s = (T)0;
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) s = add_rn(s, red[w]);
            y[r0] = s;
        }
        return;
    }

    /* phase 1: products of the slice [lo_al, hi) into shared memory */
    const int lo_al = lo & ~1;                   /* 16-byte aligned pair start */
    const int npair = (hi - lo_al + 1) >> 1;
    const P    *val2 = reinterpret_cast<const P *>(val + lo_al);
    const int2 *col2 = reinterpret_cast<const int2 *>(col + lo_al);
    P *prod2 = reinterpret_cast<P *>(prod);

    for (int p0 = tid; p0 < npair; p0 += THREADS * U) {
        int2 c[U];
        P    v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = min(p0 + u * THREADS, npair - 1);
            c[u] = ld_stream(col2 + p);
            v[u] = ld_stream(val2 + p);
        }
        T xa[U], xb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            xa[u] = __ldg(xm1 + c[u].x);
            xb[u] = __ldg(xm1 + c[u].y);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + u * THREADS;
            if (p < npair) {
                P pr;
                pr.x = mul_rn(v[u].x, xa[u]);
                pr.y = mul_rn(v[u].y, xb[u]);
                prod2[p] = pr;
            }
        }
    }
    __syncthreads();

    /* phase 2: one thread per row, strict left-to-right sum */
    for (int r = r0 + tid; r < r1; r += THREADS) {
        const int s = rowptr[r] - lo_al, e = rowptr[r + 1] - lo_al;
        T acc = (T)0;
#pragma unroll 8
        for (int k = s; k < e; ++k) acc = add_rn(acc, prod[k]);
        y[r] = acc;
    }
}

template <typename T>
void launch_ordered(const DevCsr &m, const T *x, T *y, cudaStream_t s)
{
    if (m.nblk <= 0) return;
    constexpr int TILE = sizeof(T) == 8 ? kTileF64 : kTileF32;
    spmv_stream_ordered<T, kStreamThreads, TILE><<<m.nblk, kStreamThreads, 0, s>>>(
        static_cast<const T *>(m.val), m.col, m.rowptr, m.rowblk, x - 1, y);
}
template void launch_ordered<double>(const DevCsr &, const double *, double *, cudaStream_t);
template void launch_ordered<float>(const DevCsr &, const float *, float *, cudaStream_t);

int tile_elems(bool f32) { return f32 ? kTileF32 : kTileF64; }

/* ------------------------------------------------------------------------
 * VECTOR: V lanes per row
 * ---------------------------------------------------------------------- */
template <typename T, int V>
__global__ void __launch_bounds__(256)
spmv_vector(const T *__restrict__ val, const int *__restrict__ col,
            const int *__restrict__ rowptr, const T *__restrict__ xm1,
            T *__restrict__ y, int rows)
{
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = gt / V, lane = gt % V;
    T acc = (T)0;
    if (row < rows) {
        const int e = rowptr[row + 1];
        for (int i = rowptr[row] + lane; i < e; i += V)
            acc = add_rn(acc, mul_rn(ld_stream(val + i), __ldg(xm1 + ld_stream(col + i))));
    }
#pragma unroll
    for (int o = V / 2; o > 0; o >>= 1) acc = add_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if (row < rows && lane == 0) y[row] = acc;
}

template <typename T, int V>
static void launch_vector_v(const DevCsr &m, const T *x, T *y, cudaStream_t s)
{
    const long long threads = (long long)m.rows * V;
    const int grid = (int)((threads + 255) / 256);
    spmv_vector<T, V><<<grid, 256, 0, s>>>(static_cast<const T *>(m.val), m.col, m.rowptr,
                                           x - 1, y, m.rows);
}

template <typename T>
void launch_vector(const DevCsr &m, int lanes, const T *x, T *y, cudaStream_t s)
{
    if (m.rows <= 0) return;
    switch (lanes) {
    case 2:  launch_vector_v<T, 2>(m, x, y, s); break;
    case 4:  launch_vector_v<T, 4>(m, x, y, s); break;
    case 8:  launch_vector_v<T, 8>(m, x, y, s); break;
    case 16: launch_vector_v<T, 16>(m, x, y, s); break;
    default: launch_vector_v<T, 32>(m, x, y, s); break;
    }
}
template void launch_vector<double>(const DevCsr &, int, const double *, double *, cudaStream_t);
template void launch_vector<float>(const DevCsr &, int, const float *, float *, cudaStream_t);

/* ------------------------------------------------------------------------
 * upload-time passes
 * ---------------------------------------------------------------------- */
__global__ void rebase_rowptr_kernel(int *rowptr, int n, int base)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rowptr[i] -= base;
}

void launch_rebase_rowptr(int *rowptr, int rows_plus_1, int base, cudaStream_t s)
{
    if (rows_plus_1 <= 0) return;
    rebase_rowptr_kernel<<<(rows_plus_1 + 255) / 256, 256, 0, s>>>(rowptr, rows_plus_1, base);
}

/* One pass over rowptr and col: column range (the ABI does not pass the
 * column count -- mkl.c:42-44, gpu.c:216-223), row-length histogram, and
 * whether every row has non-decreasing columns (needed by the panel layout). */
__global__ void upload_scan_kernel(const int *__restrict__ rowptr, const int *__restrict__ col,
                                   int rows, int nnz, UploadScan *out)
{
    __shared__ unsigned int sh_hist[32];
    __shared__ int sh_max_col, sh_min_col, sh_max_len, sh_min_len, sh_unsorted;
    if (threadIdx.x < 32) sh_hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        sh_max_col = 0; sh_min_col = 0x7fffffff;
        sh_max_len = 0; sh_min_len = 0x7fffffff; sh_unsorted = 0;
    }
    __syncthreads();

    const int stride = gridDim.x * blockDim.x;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    int mx = 0, mn = 0x7fffffff;
    for (int i = gt; i < nnz; i += stride) {
        const int c = col[i];
        mx = max(mx, c);
        mn = min(mn, c);
    }
    int mxl = 0, mnl = 0x7fffffff, uns = 0;
    for (int r = gt; r < rows; r += stride) {
        const int b = rowptr[r], e = rowptr[r + 1];
        const int len = e - b;
        mxl = max(mxl, len);
        mnl = min(mnl, len);
        const int bin = len <= 1 ? 0 : 32 - __clz(len - 1);
        atomicAdd(&sh_hist[bin], 1u);
        /* sortedness is only probed for moderate rows; long rows never take
         * the panel layout */
        if (len <= 4096) {
            int prev = 0, bad = 0;
            for (int i = b; i < e; ++i) {
                const int c = col[i];
                bad |= (c < prev);
                prev = c;
            }
            uns += bad;
        } else {
            uns += 1;
        }
    }
    atomicMax(&sh_max_col, mx);
    atomicMin(&sh_min_col, mn);
    atomicMax(&sh_max_len, mxl);
    atomicMin(&sh_min_len, mnl);
    atomicAdd(&sh_unsorted, uns);
    __syncthreads();
    if (threadIdx.x < 32 && sh_hist[threadIdx.x])
        atomicAdd(&out->hist[threadIdx.x], (unsigned long long)sh_hist[threadIdx.x]);
    if (threadIdx.x == 0) {
        atomicMax(&out->max_col, sh_max_col);
        atomicMin(&out->min_col, sh_min_col);
        atomicMax(&out->max_len, sh_max_len);
        atomicMin(&out->min_len, sh_min_len);
        atomicAdd(&out->rows_unsorted, sh_unsorted);
    }
}

void launch_upload_scan(const int *rowptr, const int *col, int rows, int nnz,
                        UploadScan *out, cudaStream_t s)
{
    UploadScan init;
    init.max_col = 0; init.min_col = 0x7fffffff;
    init.max_len = 0; init.min_len = 0x7fffffff;
    init.rows_unsorted = 0; init.pad = 0;
    for (int i = 0; i < 32; ++i) init.hist[i] = 0;
    cudaMemcpyAsync(out, &init, sizeof init, cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);           /* `init` is a stack object */
    const int grid = 148 * 8;
    upload_scan_kernel<<<grid, 256, 0, s>>>(rowptr, col, rows, nnz, out);
}

/* ------------------------------------------------------------------------
 * x staging: pinned host memory -> device vector, read over PCIe by the SMs
 * ---------------------------------------------------------------------- */
__global__ void copy_in_kernel(const int4 *__restrict__ src, int4 *__restrict__ dst, size_t n16,
                               const unsigned char *__restrict__ src_tail,
                               unsigned char *__restrict__ dst_tail, int tail)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        dst[i] = src[i];
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}

__global__ void copy_in_bytes_kernel(const unsigned char *__restrict__ src,
                                     unsigned char *__restrict__ dst, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

struct CopyDst { unsigned char *p[8]; int n; };

__global__ void copy_in_multi_kernel(const unsigned char *__restrict__ src, CopyDst dst, size_t bytes, int vec)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t done = 0;
    if (vec) {
        const size_t n16 = bytes / 16;
        for (size_t i = t0; i < n16; i += stride) {
            const int4 v = reinterpret_cast<const int4 *>(src)[i];
            for (int j = 0; j < dst.n; ++j) reinterpret_cast<int4 *>(dst.p[j])[i] = v;
        }
        done = n16 * 16;
    }
    for (size_t i = done + t0; i < bytes; i += stride) {
        const unsigned char v = src[i];
        for (int j = 0; j < dst.n; ++j) dst.p[j][i] = v;
    }
}

struct CopyFlags { unsigned long long *p[8]; unsigned long long epoch; unsigned int *counter; };

__global__ void copy_in_multi_flagged_kernel(const unsigned char *__restrict__ src, CopyDst dst, size_t bytes,
                                             int vec, CopyFlags fl)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t done = 0;
    if (vec) {
        const size_t n16 = bytes / 16;
        for (size_t i = t0; i < n16; i += stride) {
            const int4 v = reinterpret_cast<const int4 *>(src)[i];
            for (int j = 0; j < dst.n; ++j) reinterpret_cast<int4 *>(dst.p[j])[i] = v;
        }
        done = n16 * 16;
    }
    for (size_t i = done + t0; i < bytes; i += stride) {
        const unsigned char v = src[i];
        for (int j = 0; j < dst.n; ++j) dst.p[j][i] = v;
    }
    /* the block that arrives last publishes: one system-scope fence per block after the
     * barrier (cumulative), one more in the last block, then relaxed flag stores */
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int prev = atomicAdd(fl.counter, 1u);
        if (prev == gridDim.x - 1) {
            *fl.counter = 0;
            __threadfence_system();
            for (int j = 0; j < dst.n; ++j)
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(fl.p[j]), "l"(fl.epoch) : "memory");
        }
    }
}

void launch_copy_in_multi_flagged(const void *src, void *const *dst, int ndst, size_t bytes,
                                  unsigned long long *const *flag, unsigned long long epoch,
                                  unsigned int *counter, cudaStream_t s)
{
    if (ndst <= 0) return;
    CopyDst d;
    CopyFlags f;
    d.n = ndst > 8 ? 8 : ndst;
    uintptr_t al = reinterpret_cast<uintptr_t>(src);
    for (int j = 0; j < d.n; ++j) {
        d.p[j] = static_cast<unsigned char *>(dst[j]);
        f.p[j] = flag[j];
        al |= reinterpret_cast<uintptr_t>(dst[j]);
    }
    f.epoch = epoch;
    f.counter = counter;
    const int grid = (int)std::min<size_t>((bytes / 16 + 255) / 256 + 1, 148 * 4);
    copy_in_multi_flagged_kernel<<<grid, 256, 0, s>>>(static_cast<const unsigned char *>(src), d, bytes,
                                                      (al & 15) == 0 ? 1 : 0, f);
}

struct ProbeArgs { size_t off[4]; unsigned long long val[4]; };

__device__ __forceinline__ unsigned long long probe_word(const unsigned char *p, int es)
{
    return es == 8 ? *reinterpret_cast<const volatile unsigned long long *>(p)
                   : (unsigned long long)*reinterpret_cast<const volatile unsigned int *>(p);
}

__global__ void probe_x_kernel(const unsigned char *alias, ProbeArgs a, int es, int *bad)
{
    if (threadIdx.x < 4 && probe_word(alias + a.off[threadIdx.x], es) != a.val[threadIdx.x]) *bad = 1;
}

__global__ void probe_y_kernel(const unsigned char *alias, ProbeArgs a, int es, size_t limit,
                               unsigned long long *out)
{
    if (threadIdx.x < 4 && a.off[threadIdx.x] < limit) out[threadIdx.x] = probe_word(alias + a.off[threadIdx.x], es);
}

void launch_probe_x(const void *alias, const size_t off[4], const unsigned long long val[4], int es,
                    int *bad, cudaStream_t s)
{
    ProbeArgs a;
    for (int k = 0; k < 4; ++k) { a.off[k] = off[k]; a.val[k] = es == 8 ? val[k] : (val[k] & 0xffffffffull); }
    probe_x_kernel<<<1, 32, 0, s>>>(static_cast<const unsigned char *>(alias), a, es, bad);
}

void launch_probe_y(const void *alias, const size_t off[4], int es, size_t limit, unsigned long long *out,
                    cudaStream_t s)
{
    ProbeArgs a;
    for (int k = 0; k < 4; ++k) { a.off[k] = off[k]; a.val[k] = 0; }
    probe_y_kernel<<<1, 32, 0, s>>>(static_cast<const unsigned char *>(alias), a, es, limit, out);
}

__global__ void wait_flags_kernel(const unsigned long long *flags, int n, unsigned long long epoch)
{
    if ((int)threadIdx.x < n) {
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
        } while (v < epoch);
    }
}

void launch_wait_flags(const unsigned long long *flags, int n, unsigned long long epoch, cudaStream_t s)
{
    if (n > 0) wait_flags_kernel<<<1, 32, 0, s>>>(flags, n, epoch);
}

void launch_copy_in_multi(const void *src, void *const *dst, int ndst, size_t bytes, cudaStream_t s)
{
    if (bytes == 0 || ndst <= 0) return;
    CopyDst d;
    d.n = ndst > 8 ? 8 : ndst;
    uintptr_t al = reinterpret_cast<uintptr_t>(src);
    for (int j = 0; j < d.n; ++j) {
        d.p[j] = static_cast<unsigned char *>(dst[j]);
        al |= reinterpret_cast<uintptr_t>(dst[j]);
    }
    const int grid = (int)std::min<size_t>((bytes / 16 + 255) / 256 + 1, 148 * 4);
    copy_in_multi_kernel<<<grid, 256, 0, s>>>(static_cast<const unsigned char *>(src), d, bytes,
                                              (al & 15) == 0 ? 1 : 0);
}

void launch_copy_in(const void *src, void *dst, size_t bytes, cudaStream_t s)
{
    if (bytes == 0) return;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        const size_t n16 = bytes / 16;
        const int tail = (int)(bytes - n16 * 16);
        const int grid = (int)std::min<size_t>((n16 + 255) / 256 + 1, 148 * 8);
        copy_in_kernel<<<grid, 256, 0, s>>>(
            static_cast<const int4 *>(src), static_cast<int4 *>(dst), n16,
            static_cast<const unsigned char *>(src) + n16 * 16,
            static_cast<unsigned char *>(dst) + n16 * 16, tail);
    } else {
        const int grid = (int)std::min<size_t>((bytes + 255) / 256, 148 * 8);
        copy_in_bytes_kernel<<<grid, 256, 0, s>>>(static_cast<const unsigned char *>(src),
                                                  static_cast<unsigned char *>(dst), bytes);
    }
}

}  // namespace b200
