/*
 * spmv_panelg.cu -- PANEL for wide matrices: G = 2, 4 or 8 rows per lane
 * stream ("flagged streams").
 *
 * Every CTA of a PANEL kernel reads the whole x once per row block, so the
 * x traffic from L2 is  ceil(rows / R) * ncols * sizeof(T)  and only a tall
 * row block (large R) keeps it below the matrix stream: NPB class D has 1.5 M
 * columns (12 MB of x) and ~463 entries per row, so R = 1024 reloads 9.4 KB
 * of x per row for 4.9 KB of matrix.  This variant raises R to G * T rows
 * (T <= 512 threads, R <= 4096) without growing the CTA:
 *
 *   - a lane stream is the entries of its G rows back to back with no
 *     per-row padding; bit 15 of the 16-bit panel-local column marks the
 *     first entry of a row, so the consumer needs no switch table;
 *   - the per-panel metadata is just the G tile-local row ids of the lane
 *     (2 bytes per (row, panel)), fetched one panel ahead and consumed like a
 *     shift register;
 *   - rows are dealt to the lanes boustrophedon over the per-tile sorted
 *     order (ranks t, 2T-1-t, 2T+t, 4T-1-t, ...), so the streams of a slice
 *     have nearly equal length and rows with no entry in the panel come last
 *     in every lane (they are never switched to).
 *
 * Everything else is the PANEL design (spmv_panel.cu): x slices by TMA bulk
 * copies into shared memory, the matrix stream prefetched by a cursor that
 * runs across panel boundaries, running sums carried in shared memory,
 * separately rounded multiply and add in the reference's order
 * (libspmv/native-impl.c:1-12), hence bit-identical results for sorted rows.
 * Padding entries are (+0.0, slot W) with slot W holding +0.0.
 */
#include "panel_common.cuh"

namespace b200 {

/* ---- build: sort the rows of a tile, deal them to the lanes ---------------- */
__global__ void __launch_bounds__(1024)
panelg_sort_kernel(const uint16_t *__restrict__ seglen, int R, int G, int NK,
                   uint16_t *__restrict__ rowids, int *__restrict__ slice_elems)
{
    extern __shared__ uint32_t key[];
    const int tile = blockIdx.x, t = threadIdx.x;
    for (int i = t; i < NK; i += 1024)
        key[i] = i < R ? ((uint32_t)seglen[(size_t)tile * R + i] << 16) | (uint32_t)(0xFFFF - i) : 0u;
    __syncthreads();
    for (int k = 2; k <= NK; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < NK; i += 1024) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t a = key[i], b = key[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { key[i] = b; key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    const int T = R / G;                      /* <= 1024, a multiple of 32 */
    /* the g-th rows of the 32 lanes of a slice are neighbours in the sorted order, so their
     * counts are (nearly) equal: every such row slot is padded to the slice's longest, which
     * makes the row switches of a slice fall on the same stream position in every lane */
    int len = 0;
    for (int g = 0; g < G; ++g) {
        int lg = 0;
        if (t < T) {
            const int rank = (g & 1) ? (g + 1) * T - 1 - t : g * T + t;
            const uint32_t kk = key[rank];
            lg = (int)(kk >> 16);
            rowids[((size_t)tile * T + t) * G + g] = (uint16_t)(0xFFFF - (kk & 0xFFFFu));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lg = max(lg, __shfl_xor_sync(0xffffffffu, lg, o));
        len += lg;
    }
    if (t < T && (t & 31) == 0) slice_elems[tile * (T >> 5) + (t >> 5)] = ((len + 1) >> 1) * 64;
}

void launch_panelg_sort(const uint16_t *seglen, int ntiles, int R, int G, uint16_t *rowids,
                        int *slice_elems, cudaStream_t s)
{
    if (ntiles <= 0) return;
    int NK = 1024;
    while (NK < R) NK <<= 1;
    panelg_sort_kernel<<<ntiles, 1024, NK * sizeof(uint32_t), s>>>(seglen, R, G, NK, rowids, slice_elems);
}

/* ---- build: scatter CSR entries into the flagged lane streams -------------- */
template <typename T>
__global__ void panelg_fill_kernel(const T *__restrict__ val, const int *__restrict__ col,
                                   const int *__restrict__ rowptr, int rows, int R, int G, int P, int W,
                                   const uint16_t *__restrict__ seglen,
                                   const uint16_t *__restrict__ rowids,
                                   const int *__restrict__ slice_off, int nslices, int wmajor,
                                   T *__restrict__ val_out, uint16_t *__restrict__ col_out)
{
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   /* global slice id */
    const int lane = threadIdx.x & 31;
    if (gw >= nslices) return;
    const int Tn = R / G;
    const int spb = Tn >> 5;
    const int tile = gw / spb, w = gw - tile * spb;
    const int rb = tile / P, p = tile - rb * P;
    /* slice order in memory: (row block, panel, warp), or (row block, warp, panel) for the
     * ring kernel, whose warps stream their slices of all panels back to back */
    const int si = wmajor ? (rb * spb + w) * P + p : gw;
    const int off = slice_off[si];
    const int nent = (slice_off[si + 1] - off) >> 5;               /* entries per lane */
    const uint16_t *ids = rowids + ((size_t)tile * Tn + w * 32 + lane) * G;
    int k = 0;
    for (int g = 0; g < G; ++g) {
        const int rr = ids[g];
        const int r = rb * R + rr;
        const int len = r < rows ? (int)seglen[(size_t)tile * R + rr] : 0;
        int slot = len;                                            /* row slot g: the slice's longest */
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) slot = max(slot, __shfl_xor_sync(0xffffffffu, slot, o));
        const int k_end = k + slot;
        if (len > 0) {
            /* first entry of the row with column >= p * W (columns are sorted) */
            int lo = rowptr[r], hi = rowptr[r + 1];
            const int c0 = p * W + 1;                              /* 1-based */
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (col[mid] < c0) lo = mid + 1; else hi = mid;
            }
            for (int e = 0; e < len; ++e, ++k) {
                const size_t idx = (size_t)off + (size_t)(k >> 1) * 64 + lane * 2 + (k & 1);
                val_out[idx] = val[lo + e];
                col_out[idx] = (uint16_t)((col[lo + e] - c0) | (e == 0 ? 0x8000 : 0));
            }
        }
        for (; k < k_end; ++k) {                                   /* +0.0 * x[W] = +0.0 */
            const size_t idx = (size_t)off + (size_t)(k >> 1) * 64 + lane * 2 + (k & 1);
            val_out[idx] = (T)0;
            col_out[idx] = (uint16_t)W;
        }
    }
    for (; k < nent; ++k) {
        const size_t idx = (size_t)off + (size_t)(k >> 1) * 64 + lane * 2 + (k & 1);
        val_out[idx] = (T)0;
        col_out[idx] = (uint16_t)W;
    }
}

template <typename T>
void launch_panelg_fill(const T *val, const int *col, const int *rowptr, int rows,
                        const DevPanel &pm, const uint16_t *seglen, T *val_out, uint16_t *col_out,
                        cudaStream_t s)
{
    const int nslices = pm.nblk * pm.P * (pm.R / pm.G / 32);
    if (nslices <= 0) return;
    const long long threads = (long long)nslices * 32;
    panelg_fill_kernel<T><<<(int)((threads + 255) / 256), 256, 0, s>>>(
        val, col, rowptr, rows, pm.R, pm.G, pm.P, pm.W, seglen, pm.rowids, pm.slice_off, nslices,
        pm.fmt == 2, val_out, col_out);
}
template void launch_panelg_fill<double>(const double *, const int *, const int *, int, const DevPanel &,
                                         const uint16_t *, double *, uint16_t *, cudaStream_t);
template void launch_panelg_fill<float>(const float *, const int *, const int *, int, const DevPanel &,
                                        const uint16_t *, float *, uint16_t *, cudaStream_t);

/* ------------------------------------------------------------------------
 * the product
 * ---------------------------------------------------------------------- */
/* consume U pairs in order; a flagged entry parks the running sum of the
 * current row in shared memory and picks up the next row of the lane */
template <typename T, int U, int G>
__device__ __forceinline__ T consume_flagged(const Chunk<T, U> &ch, const T *xs, T *sums, T acc,
                                             int kp, int npair, int &cur, RowIds<G> &ids)
{
    T xa[U], xb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (kp + u < npair) {
            xa[u] = xs[ch.c[u] & 0x7FFFu];
            xb[u] = xs[(ch.c[u] >> 16) & 0x7FFFu];
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (kp + u < npair) {
            const uint32_t c = ch.c[u];
            if (c & 0x8000u) {
                sums[cur] = acc;
                cur = pop_id<G>(ids);
                acc = sums[cur];
            }
            acc = padd(acc, pmul(ch.v[u].x, xa[u]));
            if (c & 0x80000000u) {
                sums[cur] = acc;
                cur = pop_id<G>(ids);
                acc = sums[cur];
            }
            acc = padd(acc, pmul(ch.v[u].y, xb[u]));
        }
    }
    return acc;
}

/* the read cursor (panel_common.cuh) over a one-int-per-slice offset table */
template <int U>
__device__ __forceinline__ void cursor_open(StreamCursor &cur, const int *s_off, int slice, int lane)
{
    const int o = s_off[slice];
    cur.kp = 0;
    cur.npair = (s_off[slice + 1] - o) >> 6;
    cur.nround = cursor_rounds<U>(cur.npair);
    cur.base = (size_t)(o >> 1) + lane;
}

template <typename T, int U>
__device__ __forceinline__ void cursor_next(Chunk<T, U> &ch, StreamCursor &cur,
                                            const typename PairT<T>::type *val2,
                                            const uint32_t *col2, const int *s_off, int spb,
                                            int warp, int lane, int P)
{
    if (cur.p < P) {
        load_chunk<T, U>(ch, val2 + cur.base, col2 + cur.base, cur.kp, cur.npair);
        cur.kp += U;
        if (cur.kp >= cur.nround) {
            ++cur.p;
            if (cur.p < P) cursor_open<U>(cur, s_off, cur.p * spb + warp, lane);
        }
    }
}

template <typename T, int U, int MAXT, int G>
__global__ void __launch_bounds__(MAXT, 1)
spmv_panelg_kernel(const T *__restrict__ val, const uint16_t *__restrict__ col,
                   const uint16_t *__restrict__ rowids, const int *__restrict__ slice_off,
                   const T *__restrict__ x, T *__restrict__ y,
                   int rows, int ncols, int P, int W, int R, int use_tma, int nbuf)
{
    using P2 = typename PairT<T>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    /* layout: [mbarriers 16 B][slice offsets P*spb + 1 int][sums R + 1 (dummy)][xbuf0 W+pad][xbuf1 W+pad];
     * the slices of a row block are contiguous in (panel, warp) order, so one
     * offset per slice gives its size too */
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    const int Tn = blockDim.x;
    const int spb = Tn >> 5;
    int *s_off = reinterpret_cast<int *>(smem_raw + 16);
    const size_t soff = (16 + ((size_t)P * spb + 1) * sizeof(int) + 7) & ~(size_t)7;
    T *sums = reinterpret_cast<T *>(smem_raw + soff);
    const size_t xoff = (soff + (size_t)(R + 1) * sizeof(T) + 15) & ~(size_t)15;
    const int WS = W + (16 / (int)sizeof(T));
    T *xbuf = reinterpret_cast<T *>(smem_raw + xoff);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rb = blockIdx.x;
    const P2 *val2 = reinterpret_cast<const P2 *>(val);
    const uint32_t *col2 = reinterpret_cast<const uint32_t *>(col);

    for (int i = tid; i <= R; i += Tn) sums[i] = (T)0;
    for (int i = tid; i <= P * spb; i += Tn) s_off[i] = slice_off[(size_t)rb * P * spb + i];
    if (tid == 0) {
        xbuf[W] = (T)0;                                   /* padding slot, never overwritten */
        if (nbuf == 2) xbuf[WS + W] = (T)0;
        if (use_tma) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    const uint16_t *my_ids = rowids + ((size_t)rb * P * Tn + tid) * G;
    RowIds<G> ids_next = load_ids<G>(my_ids);
    __syncthreads();

    auto issue_panel = [&](int p) {                       /* thread 0 only (TMA path) */
        const int cbase = p * W;
        const int cw = min(W, ncols - cbase);
        T *dst = xbuf + (size_t)(p & (nbuf - 1)) * WS;
        constexpr int VE = 16 / sizeof(T);
        const int cw_al = cw & ~(VE - 1);
        for (int i = cw_al; i < cw; ++i) dst[i] = __ldg(x + cbase + i);   /* ragged tail */
        fence_proxy_async();
        uint64_t *bar = &bars[p & (nbuf - 1)];
        if (cw_al > 0) {
            mbar_expect_tx(bar, (uint32_t)(cw_al * sizeof(T)));
            uint32_t left = (uint32_t)(cw_al * sizeof(T));
            const char *src = reinterpret_cast<const char *>(x + cbase);
            char *d = reinterpret_cast<char *>(dst);
            while (left) {
                const uint32_t n = left > 32768u ? 32768u : left;
                tma_bulk_g2s(d, src, n, bar);
                d += n; src += n; left -= n;
            }
        } else {
            mbar_expect_tx(bar, 0);
        }
    };
    auto coop_panel = [&](int p) {                        /* all threads (fallback path) */
        const int cbase = p * W;
        const int cw = min(W, ncols - cbase);
        T *dst = xbuf + (size_t)(p & (nbuf - 1)) * WS;
        for (int i = tid; i < cw; i += Tn) dst[i] = __ldg(x + cbase + i);
    };

    if (use_tma) {
        if (tid == 0) issue_panel(0);
    } else {
        coop_panel(0);
    }

    StreamCursor cur;
    cur.p = 0;
    cursor_open<U>(cur, s_off, warp, lane);
    Chunk<T, U> a, b;
    cursor_next<T, U>(a, cur, val2, col2, s_off, spb, warp, lane, P);
    cursor_next<T, U>(b, cur, val2, col2, s_off, spb, warp, lane, P);

    for (int p = 0; p < P; ++p) {
        RowIds<G> ids = ids_next;
        if (p + 1 < P) ids_next = load_ids<G>(my_ids + (size_t)(p + 1) * Tn * G);
        const int npair = (s_off[p * spb + warp + 1] - s_off[p * spb + warp]) >> 6;

        if (use_tma) {
            if (tid == 0) {
                if (nbuf == 2 && p + 1 < P) issue_panel(p + 1);
                if (nbuf == 1 && p > 0) issue_panel(p);
            }
            mbar_wait(&bars[p & (nbuf - 1)], (uint32_t)((p >> (nbuf - 1)) & 1));
        } else {
            if (nbuf == 2 && p + 1 < P) coop_panel(p + 1);
            if (nbuf == 1 && p > 0) coop_panel(p);
            __syncthreads();
        }
        const T *xs = xbuf + (size_t)(p & (nbuf - 1)) * WS;

        int row_cur = R;                                  /* dummy slot until the first flag */
        T acc = (T)0;
        /* an empty slice still takes one (fully predicated) round: see cursor_rounds() */
        for (int kp = 0; kp < max(npair, 1); kp += 2 * U) {
            acc = consume_flagged<T, U, G>(a, xs, sums, acc, kp, npair, row_cur, ids);
            cursor_next<T, U>(a, cur, val2, col2, s_off, spb, warp, lane, P);
            acc = consume_flagged<T, U, G>(b, xs, sums, acc, kp + U, npair, row_cur, ids);
            cursor_next<T, U>(b, cur, val2, col2, s_off, spb, warp, lane, P);
        }
        sums[row_cur] = acc;
        __syncthreads();            /* panel p consumed: its x buffer and the sums are free */
    }
    for (int i = tid; i < R; i += Tn) {
        const int row = rb * R + i;
        if (row < rows) y[row] = sums[i];
    }
}

size_t panelg_smem_bytes(const DevPanel &pm, bool f32)
{
    const size_t es = f32 ? 4 : 8;
    const size_t soff = (16 + ((size_t)pm.P * (pm.R / pm.G / 32) + 1) * 4 + 7) & ~(size_t)7;
    const size_t xoff = (soff + (size_t)(pm.R + 1) * es + 15) & ~(size_t)15;
    const size_t ws = (size_t)pm.W + 16 / es;
    return xoff + (size_t)pm.nbuf * ws * es;
}

template <typename T, int U, int MAXT, int G>
static void launch_panelg_cfg(const DevPanel &pm, const T *x, T *y, cudaStream_t s)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(spmv_panelg_kernel<T, U, MAXT, G>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    const size_t smem = panelg_smem_bytes(pm, sizeof(T) == 4);
    const int use_tma = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && pm.use_tma;
    spmv_panelg_kernel<T, U, MAXT, G><<<pm.nblk, pm.R / pm.G, smem, s>>>(
        static_cast<const T *>(pm.val), pm.col, pm.rowids, pm.slice_off, x, y,
        pm.rows, pm.ncols, pm.P, pm.W, pm.R, use_tma, pm.nbuf);
}

template <typename T, int G>
static void launch_panelg_g(const DevPanel &pm, const T *x, T *y, cudaStream_t s)
{
    const int threads = pm.R / pm.G;
    /* fewer warps leave more registers per lane: keep more of the stream in flight */
    if (threads <= 256)      launch_panelg_cfg<T, 10, 256, G>(pm, x, y, s);
    else if (threads <= 384) launch_panelg_cfg<T, 7, 384, G>(pm, x, y, s);
    else                     launch_panelg_cfg<T, 5, 512, G>(pm, x, y, s);
}

template <typename T>
void launch_panelg(const DevPanel &pm, const T *x, T *y, cudaStream_t s)
{
    if (pm.nblk <= 0) return;
    if (pm.G == 2)      launch_panelg_g<T, 2>(pm, x, y, s);
    else if (pm.G == 4) launch_panelg_g<T, 4>(pm, x, y, s);
    else                launch_panelg_g<T, 8>(pm, x, y, s);
}
template void launch_panelg<double>(const DevPanel &, const double *, double *, cudaStream_t);
template void launch_panelg<float>(const DevPanel &, const float *, float *, cudaStream_t);

}  // namespace b200
