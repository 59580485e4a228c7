/*
 * spmv_sell.cu -- the SELL kernel family: order-preserving CSR SpMV for
 * matrices whose column range is too wide for shared-memory x slices
 * (NPB class D row blocks, SparseBench crsmat, graphs).  x is gathered
 * through the read-only path and lives in B200's 126 MB L2.
 *
 * Layout (built on the device at upload, same lane-stream idea as the PANEL
 * family, spmv_panel.cu): rows are cut into tiles of R rows; inside a tile
 * they are sorted by length and dealt to the T = R/G threads longest with
 * shortest, so all lanes of a warp -- and all warps -- carry about the same
 * number of entries.  A lane stream is row A's entries, padded to a pair,
 * followed by row B's; warp slices are stored SELL-style in pairs
 * (pair kp of lane l at slice_off + kp*64 + l*2) with 32-bit global columns.
 * Every lane walks its rows left to right with a separately rounded multiply
 * and a separately rounded add: bit-identical to libspmv/native-impl.c:1-12
 * whatever the column order (unsorted and repeated columns included).
 *
 * Rows longer than the cap (65534 entries, or far above the mean) would
 * serialise one lane; they are excluded from the tiles and reduced by a whole
 * CTA each (tree order) from the CSR copy.
 *
 * Roofline: HBM stream 12 B per entry; the binding resource on B200 is the
 * L1TEX wavefront rate of the divergent x gather (one wavefront per entry per
 * SM clock), i.e. about half the HBM roofline for fp64.
 */
#include "spmv_kernels.cuh"

namespace b200 {

__device__ __forceinline__ double smul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  smul(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double sadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  sadd(float a, float b)   { return __fadd_rn(a, b); }

template <typename T> struct SPair;
template <> struct SPair<double> { using type = double2; };
template <> struct SPair<float>  { using type = float2; };

/* ---- build: row lengths per tile (rows above the cap are left out) -------- */
__global__ void sell_rowlen_kernel(const int *__restrict__ rowptr, int rows, int cap,
                                   uint16_t *__restrict__ seglen)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int len = rowptr[r + 1] - rowptr[r];
    seglen[r] = (uint16_t)(len <= cap ? len : 0);
}

void launch_sell_rowlen(const int *rowptr, int rows, int R, int cap, uint16_t *seglen, cudaStream_t s)
{
    (void)R;   /* tiles are consecutive runs of R rows: seglen[tile * R + rr] == seglen[row] */
    if (rows <= 0) return;
    sell_rowlen_kernel<<<(rows + 255) / 256, 256, 0, s>>>(rowptr, rows, cap, seglen);
}

/* ---- build: scatter CSR entries into the lane streams ---------------------- */
template <typename T>
__global__ void sell_fill_kernel(const T *__restrict__ val, const int *__restrict__ col,
                                 const int *__restrict__ rowptr, int rows, int R, int G,
                                 const ushort4 *__restrict__ meta, const int *__restrict__ slice_off,
                                 int nslices, T *__restrict__ val_out, int *__restrict__ col_out)
{
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= nslices) return;
    const int Tn = R / G;
    const int spb = Tn >> 5;
    const int tile = gw / spb, w = gw - tile * spb;
    const ushort4 mt = meta[(size_t)tile * Tn + w * 32 + lane];
    const int off = slice_off[gw];
    const int npair = (slice_off[gw + 1] - off) >> 6;
    int kp = 0;
    for (int g = 0; g < G; ++g) {
        const int r = tile * R + (g == 0 ? mt.x : mt.y);
        const int len = g == 0 ? mt.z : mt.w;
        const int src = r < rows ? rowptr[r] : 0;
        for (int k = 0; k < len; k += 2, ++kp) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                T v = (T)0;
                int c = 1;
                if (k + e < len) {
                    v = val[src + k + e];
                    c = col[src + k + e];
                }
                const size_t idx = (size_t)off + (size_t)kp * 64 + lane * 2 + e;
                val_out[idx] = v;
                col_out[idx] = c;
            }
        }
    }
    for (; kp < npair; ++kp) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const size_t idx = (size_t)off + (size_t)kp * 64 + lane * 2 + e;
            val_out[idx] = (T)0;
            col_out[idx] = 1;
        }
    }
}

template <typename T>
void launch_sell_fill(const T *val, const int *col, const int *rowptr, int rows, const DevSell &sm,
                      const uint16_t *seglen, T *val_out, int *col_out, cudaStream_t s)
{
    (void)seglen;
    const int nslices = sm.nblk * (sm.R / sm.G / 32);
    if (nslices <= 0) return;
    const long long threads = (long long)nslices * 32;
    sell_fill_kernel<T><<<(int)((threads + 255) / 256), 256, 0, s>>>(
        val, col, rowptr, rows, sm.R, sm.G, sm.meta, sm.slice_off, nslices, val_out, col_out);
}
template void launch_sell_fill<double>(const double *, const int *, const int *, int, const DevSell &,
                                       const uint16_t *, double *, int *, cudaStream_t);
template void launch_sell_fill<float>(const float *, const int *, const int *, int, const DevSell &,
                                      const uint16_t *, float *, int *, cudaStream_t);

/* ------------------------------------------------------------------------
 * the product
 * ---------------------------------------------------------------------- */
template <typename T, int U>
struct SChunk {
    typename SPair<T>::type v[U];
    int2 c[U];
};

template <typename T, int U>
__device__ __forceinline__ void sell_load(SChunk<T, U> &ch, const typename SPair<T>::type *vp,
                                          const int2 *cp, int kp, int npair)
{
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (kp + u < npair) {
            ch.v[u] = __ldcs(vp + (size_t)(kp + u) * 32);
            ch.c[u] = __ldcs(cp + (size_t)(kp + u) * 32);
        }
    }
}

/* entries [0, lenA) belong to row A, [swE, endE) to row B (swE = lenA rounded
 * up to a pair); everything else in the slice is padding and is skipped */
template <typename T, int U>
__device__ __forceinline__ T sell_consume(const SChunk<T, U> &ch, const T *__restrict__ xm1,
                                          T *y, T acc, int kp, int npair,
                                          int lenA, int swE, int endE, int rowA, int rowB, int rows)
{
    T xa[U], xb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (kp + u < npair) {
            xa[u] = __ldg(xm1 + ch.c[u].x);
            xb[u] = __ldg(xm1 + ch.c[u].y);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (kp + u < npair) {
            const int e0 = 2 * (kp + u);
            if (e0 == swE && swE < endE) {         /* row A complete, row B starts */
                if (rowA < rows) y[rowA] = acc;
                acc = (T)0;
            }
            if (e0 < lenA || (e0 >= swE && e0 < endE)) acc = sadd(acc, smul(ch.v[u].x, xa[u]));
            if (e0 + 1 < lenA || (e0 + 1 >= swE && e0 + 1 < endE)) acc = sadd(acc, smul(ch.v[u].y, xb[u]));
        }
    }
    (void)rowB;
    return acc;
}

template <typename T, int U, int MAXT>
__global__ void __launch_bounds__(MAXT, 2)
spmv_sell_kernel(const T *__restrict__ val, const int *__restrict__ col,
                 const ushort4 *__restrict__ meta, const int *__restrict__ slice_off,
                 const T *__restrict__ xm1, T *__restrict__ y, int rows, int R, int G)
{
    using P2 = typename SPair<T>::type;
    const int Tn = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int spb = Tn >> 5;
    const int tile = blockIdx.x;
    const ushort4 mt = meta[(size_t)tile * Tn + tid];
    const int off = slice_off[tile * spb + warp];
    const int npair = (slice_off[tile * spb + warp + 1] - off) >> 6;
    const P2 *vp = reinterpret_cast<const P2 *>(val) + (size_t)(off >> 1) + lane;
    const int2 *cp = reinterpret_cast<const int2 *>(col) + (size_t)(off >> 1) + lane;

    /* the tile's y is assembled in shared memory and stored coalesced */
    extern __shared__ __align__(16) unsigned char sell_smem[];
    T *ytile = reinterpret_cast<T *>(sell_smem);
    T *yout = y;
    y = ytile;
    const int rows_g = rows;
    rows = R;                                   /* bounds for the staged stores: inside the tile */
    const int rowA = mt.x, rowB = mt.y;
    const int lenA = mt.z;
    const int swE = G == 2 ? (lenA + 1) & ~1 : 0x7fffffff;
    const int endE = G == 2 ? swE + mt.w : lenA;

    SChunk<T, U> a, b;
    sell_load<T, U>(a, vp, cp, 0, npair);
    sell_load<T, U>(b, vp, cp, U, npair);
    T acc = (T)0;
    for (int kp = 0; kp < npair; kp += 2 * U) {
        acc = sell_consume<T, U>(a, xm1, y, acc, kp, npair, lenA, swE, endE, rowA, rowB, rows);
        sell_load<T, U>(a, vp, cp, kp + 2 * U, npair);
        acc = sell_consume<T, U>(b, xm1, y, acc, kp + U, npair, lenA, swE, endE, rowA, rowB, rows);
        sell_load<T, U>(b, vp, cp, kp + 3 * U, npair);
    }
    /* which row does the running sum belong to?  B if it was entered */
    if (G == 2 && swE < endE) {             /* A was stored at the switch */
        if (rowB < rows) y[rowB] = acc;
    } else {
        if (rowA < rows) y[rowA] = acc;
        if (G == 2 && rowB < rows && rowB != rowA) y[rowB] = (T)0;   /* B has no entries */
    }
    __syncthreads();
    for (int i = tid; i < R; i += Tn) {
        const int r = tile * R + i;
        if (r < rows_g) yout[r] = ytile[i];
    }
}

/* Long rows (above the cap): nnz-split.  Every long row is cut into chunks of
 * kChunk entries; one warp reduces one chunk (4 independent loads per lane in
 * flight, xor-shuffle tree).  A row that fits one chunk is finished there;
 * otherwise the chunk sums are carried to a fix-up kernel that adds them in
 * chunk order -- deterministic, no atomics.  This is the load-balanced path
 * for heavy-tailed row lengths (power-law graphs); it re-orders the sum. */
constexpr int kChunk = 512;

template <typename T, int V>
__global__ void __launch_bounds__(256)
spmv_long_chunks_kernel(const T *__restrict__ val, const int *__restrict__ col,
                        const int4 *__restrict__ chunks, int n_chunks,
                        const T *__restrict__ xm1, T *__restrict__ y, T *__restrict__ carry)
{
    /* V lanes per chunk (adaptive vector-per-row: 8 lanes for chunks of a few
     * dozen entries, a whole warp otherwise) */
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) / V;
    const int lane = threadIdx.x % V;
    const bool live = g < n_chunks;
    const int4 c = live ? chunks[g] : make_int4(0, 0, 0, -1);   /* {row, lo, hi, carry slot or -1} */
    T a0 = (T)0, a1 = (T)0, a2 = (T)0, a3 = (T)0;
    int i = c.y + lane;
    for (; i + 3 * V < c.z; i += 4 * V) {
        const T v0 = __ldcs(val + i), v1 = __ldcs(val + i + V), v2 = __ldcs(val + i + 2 * V),
                v3 = __ldcs(val + i + 3 * V);
        const int c0 = __ldcs(col + i), c1 = __ldcs(col + i + V), c2 = __ldcs(col + i + 2 * V),
                  c3 = __ldcs(col + i + 3 * V);
        a0 = sadd(a0, smul(v0, __ldg(xm1 + c0)));
        a1 = sadd(a1, smul(v1, __ldg(xm1 + c1)));
        a2 = sadd(a2, smul(v2, __ldg(xm1 + c2)));
        a3 = sadd(a3, smul(v3, __ldg(xm1 + c3)));
    }
    for (; i < c.z; i += V) a0 = sadd(a0, smul(__ldcs(val + i), __ldg(xm1 + __ldcs(col + i))));
    T acc = sadd(sadd(a0, a1), sadd(a2, a3));
#pragma unroll
    for (int o = V / 2; o > 0; o >>= 1) acc = sadd(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if (live && lane == 0) {
        if (c.w < 0) y[c.x] = acc;
        else carry[c.w] = acc;
    }
}

/* one warp per multi-chunk row: y[row] = carry[first] + carry[first+1] + ... in chunk order.
 * The lanes fetch 32 carries at a time (one coalesced load instead of a chain of dependent
 * ones: a 65536-entry row has 128 chunks) and the sum walks them in order through shuffles,
 * so the result does not depend on the lane count: same bits as a single thread adding. */
template <typename T>
__global__ void __launch_bounds__(256)
spmv_long_fixup_kernel(const int2 *__restrict__ multi, int n_multi,
                       const int *__restrict__ rows, const T *__restrict__ carry,
                       T *__restrict__ y)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= n_multi) return;                  /* whole warps leave together */
    const int2 m = multi[t];                   /* {first carry slot, count} */
    T acc = (T)0;
    for (int k0 = 0; k0 < m.y; k0 += 32) {
        const int n = min(32, m.y - k0);
        const T v = lane < n ? carry[m.x + k0 + lane] : (T)0;
        for (int k = 0; k < n; ++k) acc = sadd(acc, __shfl_sync(0xffffffffu, v, k));
    }
    if (lane == 0) y[rows[t]] = acc;
}

int sell_chunk_entries() { return kChunk; }
int sell_short_chunk_entries() { return 96; }

template <typename T, int U>
static void launch_sell_u(const DevSell &sm, const T *x, T *y, cudaStream_t s)
{
    spmv_sell_kernel<T, U, 256><<<sm.nblk, sm.R / sm.G, (size_t)sm.R * sizeof(T), s>>>(
        static_cast<const T *>(sm.val), sm.col, sm.meta, sm.slice_off, x - 1, y, sm.rows, sm.R, sm.G);
}

template <typename T>
void launch_sell(const DevSell &sm, const DevCsr &csr, const T *x, T *y, cudaStream_t s)
{
    if (sm.nblk > 0 && sm.fmt == 1) {
        launch_sellu<T>(sm, x, y, s);
    } else if (sm.nblk > 0) {
        if (sm.U >= 6) launch_sell_u<T, 6>(sm, x, y, s);
        else if (sm.U <= 2) launch_sell_u<T, 2>(sm, x, y, s);
        else launch_sell_u<T, 4>(sm, x, y, s);
    }
    /* chunks = [n_chunks_short chunks of <= kShortChunk entries][the rest] */
    if (sm.n_chunks_short > 0)
        spmv_long_chunks_kernel<T, 8><<<(sm.n_chunks_short + 31) / 32, 256, 0, s>>>(
            static_cast<const T *>(csr.val), csr.col, sm.chunks, sm.n_chunks_short, x - 1, y,
            static_cast<T *>(sm.carry));
    if (sm.n_chunks > sm.n_chunks_short)
        spmv_long_chunks_kernel<T, 32><<<(sm.n_chunks - sm.n_chunks_short + 7) / 8, 256, 0, s>>>(
            static_cast<const T *>(csr.val), csr.col, sm.chunks + sm.n_chunks_short,
            sm.n_chunks - sm.n_chunks_short, x - 1, y, static_cast<T *>(sm.carry));
    if (sm.n_multi > 0)
        spmv_long_fixup_kernel<T><<<(sm.n_multi + 7) / 8, 256, 0, s>>>(
            sm.multi, sm.n_multi, sm.multi_rows, static_cast<const T *>(sm.carry), y);
}
template void launch_sell<double>(const DevSell &, const DevCsr &, const double *, double *, cudaStream_t);
template void launch_sell<float>(const DevSell &, const DevCsr &, const float *, float *, cudaStream_t);

}  // namespace b200
