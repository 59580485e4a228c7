/*
 * spmv_panel.cu -- the PANEL kernel family: order-preserving CSR SpMV with
 * the x slice of the current column panel staged in shared memory.
 *
 * Why: on B200 a warp-wide gather of 8-byte x entries from L2 costs one
 * L1TEX wavefront per lane (32 per instruction) and drags a 32-byte sector
 * per entry across the L2 fabric; for NPB class C that alone is ~1 cycle per
 * nonzero per SM, more than the HBM roofline allows, and it clogs the
 * load/store pipe for every other access (profiles/r01_run1_*).  A gather
 * from shared memory costs ~7 wavefronts per 32 lanes.  So the matrix is
 * re-laid out once, at upload, into (row block x column panel) tiles whose x
 * slice fits in shared memory; the CTA walks the panels left to right and
 * every row's running sum is carried from panel to panel.  With sorted
 * columns (NPB's makea keeps them sorted, cg.f:838-850) this visits every
 * row's entries in their original order, so the result is bit-identical to
 * the reference loop (libspmv/native-impl.c:1-12): separately rounded
 * multiply, separately rounded add, left to right.
 *
 * Tile layout (built on the device at upload):
 *   inside tile (rb, p) the R rows are sorted by their entry count in panel
 *   p (descending).  A CTA has T = R/G threads and every thread ("lane
 *   stream") owns G rows of the tile: with G = 2, thread t takes sorted ranks
 *   t and R-1-t (longest with shortest), so all lane streams -- and with them
 *   all warps -- carry nearly the same number of entries and the per-panel
 *   barrier costs little.  A lane stream is its rows' entries back to back,
 *   each row padded to a pair; warp slice w (threads 32w..32w+31) is stored
 *   SELL-style in pairs:  pair kp of lane l at  slice_off + kp*64 + l*2,
 *   padded to the slice's longest stream.  Padding entries hold value +0.0
 *   and the panel-local column W, a shared-memory slot that always contains
 *   +0.0, so they add +0.0 and change no bit of the sum.  meta[tile][t] =
 *   {row A, row B, pair index where B starts}; the running sums live in
 *   shared memory between panels.
 *
 * Data movement: the matrix stream (8 B value + 2 B panel-local column per
 * entry) is read with 128-bit / 32-bit coalesced evict-first loads, software
 * prefetched one chunk ahead and across the panel switch; x slices arrive by
 * TMA bulk copies (cp.async.bulk + mbarrier) into a double buffer, one panel
 * ahead of the compute.
 *
 * Algorithmic bytes are unchanged; the private layout streams
 * ~10 B per nonzero plus padding.
 *
 * Two things around the kernel proper.  (1) Programmatic dependent launch: the CTAs release
 * their dependents at once and wait (griddepcontrol.wait) only after everything that reads
 * nothing but the immutable matrix -- so back-to-back products overlap the tail of one
 * with the head of the next (class C 74.0 -> 70.3 us).  (2) A flagged instance (template
 * parameter XF) for x that is still on its way from the host or from other GPUs: thread 0
 * waits for the chunks a panel needs just before it requests the panel's slice.
 */
#include "panel_common.cuh"

#include <stdlib.h>

namespace b200 {

/* ---- build: entries of every (row, panel) -------------------------------- */
__global__ void panel_count_kernel(const int *__restrict__ rowptr, const int *__restrict__ col,
                                   int rows, int P, int W, int R,
                                   uint16_t *__restrict__ seglen, int *overflow)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int rb = r / R, rr = r - rb * R;
    const int b = rowptr[r], e = rowptr[r + 1];
    int p_cur = 0, cnt = 0;
    for (int i = b; i < e; ++i) {
        const int p = (col[i] - 1) / W;
        while (p_cur < p) {
            if (cnt > 65534) atomicExch(overflow, 1);
            seglen[((size_t)rb * P + p_cur) * R + rr] = (uint16_t)cnt;
            cnt = 0;
            ++p_cur;
        }
        ++cnt;
    }
    while (p_cur < P) {
        if (cnt > 65534) atomicExch(overflow, 1);
        seglen[((size_t)rb * P + p_cur) * R + rr] = (uint16_t)cnt;
        cnt = 0;
        ++p_cur;
    }
}

void launch_panel_count(const int *rowptr, const int *col, int rows, int P, int W, int R,
                        uint16_t *seglen, int *overflow, cudaStream_t s)
{
    if (rows <= 0) return;
    panel_count_kernel<<<(rows + 127) / 128, 128, 0, s>>>(rowptr, col, rows, P, W, R, seglen, overflow);
}

/* ---- build: sort the rows of every tile by entry count ------------------- */
/* one CTA of 1024 threads per tile; bitonic sort of up to 1024 keys
 * (count << 16 | 0xFFFF - row) in descending order => longest first, ties by
 * ascending row; then the lane streams are formed (G rows per thread). */
__global__ void __launch_bounds__(1024)
panel_sort_kernel(const uint16_t *__restrict__ seglen, int R, int G, int mode,
                  ushort4 *__restrict__ meta, int *__restrict__ slice_elems)
{
    __shared__ uint32_t key[1024];
    const int tile = blockIdx.x, t = threadIdx.x;
    key[t] = t < R ? ((uint32_t)seglen[(size_t)tile * R + t] << 16) | (uint32_t)(0xFFFF - t) : 0u;
    __syncthreads();
    for (int k = 2; k <= 1024; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int ixj = t ^ j;
            if (ixj > t) {
                const uint32_t a = key[t], b = key[ixj];
                const bool desc = (t & k) == 0;
                if (desc ? (a < b) : (a > b)) { key[t] = b; key[ixj] = a; }
            }
            __syncthreads();
        }
    }
    const int T = R / G;
    int pairs = 0;
    if (t < T) {
        const uint32_t ka = key[t];
        const int la = (int)(ka >> 16);
        ushort4 mt;
        mt.x = (unsigned short)(0xFFFF - (ka & 0xFFFFu));
        mt.y = mt.x;
        mt.z = 0xFFFF;                       /* never switches */
        mt.w = 0;
        pairs = (la + 1) >> 1;
        if (G == 2) {
            const uint32_t kb = key[R - 1 - t];
            const int lb = (int)(kb >> 16);
            mt.y = (unsigned short)(0xFFFF - (kb & 0xFFFFu));
            mt.z = (unsigned short)pairs;    /* row B starts at this pair */
            pairs += (lb + 1) >> 1;
            if (mode == 1) mt.w = (unsigned short)lb;
        }
        if (mode == 1) mt.z = (unsigned short)la;   /* SELL family: entry counts of A and B */
        meta[(size_t)tile * T + t] = mt;
    }
    /* all 32 warps take part in the reduction; threads >= T contribute 0 */
    int mx = pairs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (t < T && (t & 31) == 0) slice_elems[tile * (T >> 5) + (t >> 5)] = mx * 64;
}

void launch_panel_sort(const uint16_t *seglen, int ntiles, int R, int G, int mode, ushort4 *meta,
                       int *slice_elems, cudaStream_t s)
{
    if (ntiles <= 0) return;
    panel_sort_kernel<<<ntiles, 1024, 0, s>>>(seglen, R, G, mode, meta, slice_elems);
}

/* ---- build: scatter CSR entries into the padded tile order ---------------- */
template <typename T>
__global__ void panel_fill_kernel(const T *__restrict__ val, const int *__restrict__ col,
                                  const int *__restrict__ rowptr, int rows, int R, int G, int P, int W,
                                  const uint16_t *__restrict__ seglen,
                                  const ushort4 *__restrict__ meta,
                                  const int *__restrict__ slice_off, int nslices,
                                  T *__restrict__ val_out, uint16_t *__restrict__ col_out)
{
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   /* global slice id */
    const int lane = threadIdx.x & 31;
    if (gw >= nslices) return;
    const int Tn = R / G;
    const int spb = Tn >> 5;
    const int tile = gw / spb, w = gw - tile * spb;
    const int rb = tile / P, p = tile - rb * P;
    const ushort4 mt = meta[(size_t)tile * Tn + w * 32 + lane];
    const int off = slice_off[gw];
    const int npair = (slice_off[gw + 1] - off) >> 6;
    int kp = 0;
    for (int g = 0; g < G; ++g) {
        const int rr = g == 0 ? mt.x : mt.y;
        const int r = rb * R + rr;
        int len = 0, src = 0;
        if (r < rows) {
            len = seglen[(size_t)tile * R + rr];
            src = rowptr[r];
            for (int q = 0; q < p; ++q) src += seglen[((size_t)rb * P + q) * R + rr];
        }
        for (int k = 0; k < len; k += 2, ++kp) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                T v = (T)0;
                int c = W;                                 /* the +0.0 slot */
                if (k + e < len) {
                    v = val[src + k + e];
                    c = col[src + k + e] - 1 - p * W;
                }
                const size_t idx = (size_t)off + (size_t)kp * 64 + lane * 2 + e;
                val_out[idx] = v;
                col_out[idx] = (uint16_t)c;
            }
        }
    }
    for (; kp < npair; ++kp) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const size_t idx = (size_t)off + (size_t)kp * 64 + lane * 2 + e;
            val_out[idx] = (T)0;
            col_out[idx] = (uint16_t)W;
        }
    }
}

template <typename T>
void launch_panel_fill(const T *val, const int *col, const int *rowptr, int rows,
                       const DevPanel &pm, const uint16_t *seglen, T *val_out, uint16_t *col_out,
                       cudaStream_t s)
{
    const int nslices = pm.nblk * pm.P * (pm.R / pm.G / 32);
    if (nslices <= 0) return;
    const long long threads = (long long)nslices * 32;
    panel_fill_kernel<T><<<(int)((threads + 255) / 256), 256, 0, s>>>(
        val, col, rowptr, rows, pm.R, pm.G, pm.P, pm.W, seglen, pm.meta, pm.slice_off, nslices,
        val_out, col_out);
}
template void launch_panel_fill<double>(const double *, const int *, const int *, int, const DevPanel &,
                                        const uint16_t *, double *, uint16_t *, cudaStream_t);
template void launch_panel_fill<float>(const float *, const int *, const int *, int, const DevPanel &,
                                       const uint16_t *, float *, uint16_t *, cudaStream_t);

/* ------------------------------------------------------------------------
 * the product
 * ---------------------------------------------------------------------- */
/* consume U pairs in order; `sw` is the pair index at which the lane stream
 * moves on to its second row (running sums parked in shared memory) */
template <typename T, int U>
__device__ __forceinline__ T consume_chunk(const Chunk<T, U> &ch, const T *xs, T *sums, T acc,
                                           int kp, int npair, int sw, int &cur, int nxt)
{
    T xa[U], xb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (kp + u < npair) {
            xa[u] = xs[ch.c[u] & 0xFFFFu];
            xb[u] = xs[ch.c[u] >> 16];
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (kp + u < npair) {
            if (kp + u == sw) {
                sums[cur] = acc;
                cur = nxt;
                acc = sums[cur];
            }
            acc = padd(acc, pmul(ch.v[u].x, xa[u]));
            acc = padd(acc, pmul(ch.v[u].y, xb[u]));
        }
    }
    return acc;
}

template <typename T, int U, int MAXT, bool XF>
__global__ void __launch_bounds__(MAXT, 1)
spmv_panel_kernel(const T *__restrict__ val, const uint16_t *__restrict__ col,
                  const ushort4 *__restrict__ meta, const int *__restrict__ slice_off,
                  const T *__restrict__ x, T *__restrict__ y,
                  int rows, int ncols, int P, int W, int R, int use_tma, int nbuf,
                  const T *__restrict__ dotv, T *__restrict__ dot_partial, XFlags xf)
{
    using P2 = typename PairT<T>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    /* layout: [mbarriers 16 B][slice table P*spb int2][sums R][xbuf0 W+pad][xbuf1 W+pad] */
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    const int Tn = blockDim.x;
    const int spb = Tn >> 5;
    int2 *s_slice = reinterpret_cast<int2 *>(smem_raw + 16);
    const size_t soff = 16 + (size_t)P * spb * sizeof(int2);
    T *sums = reinterpret_cast<T *>(smem_raw + soff);
    const size_t xoff = (soff + (size_t)R * sizeof(T) + 15) & ~(size_t)15;
    const int WS = W + (16 / (int)sizeof(T));            /* buffer stride keeps 16-byte alignment */
    T *xbuf = reinterpret_cast<T *>(smem_raw + xoff);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rb = blockIdx.x;
    /* a kernel behind this one in the stream may start its CTAs as ours finish (see below) */
    asm volatile("griddepcontrol.launch_dependents;");
    const P2 *val2 = reinterpret_cast<const P2 *>(val);
    const uint32_t *col2 = reinterpret_cast<const uint32_t *>(col);

    for (int i = tid; i < R; i += Tn) sums[i] = (T)0;
    for (int i = tid; i < P * spb; i += Tn) {
        const int o = slice_off[(size_t)rb * P * spb + i];
        const int e = slice_off[(size_t)rb * P * spb + i + 1];
        s_slice[i] = make_int2(o, (e - o) >> 6);
    }
    if (tid == 0) {
        xbuf[W] = (T)0;                                   /* padding slot, never overwritten */
        if (nbuf == 2) xbuf[WS + W] = (T)0;
        if (use_tma) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    ushort4 mt_next = meta[(size_t)rb * P * Tn + tid];
    __syncthreads();

    /* x still on its way from the host (drop-in path: the copy engine delivers it chunk by
     * chunk while the product runs) or from other GPUs: wait for the chunks a panel needs
     * just before the panel is requested.  XF is a template parameter: the kernel a caller
     * with x resident in HBM gets (XF = false) carries none of this state -- the class C
     * instance sits at its 128-register limit */
    int x_ready = XF ? xf.ready0 : 0;
    auto issue_panel = [&](int p) {                       /* called by thread 0 only (TMA path) */
        const int cbase = p * W;
        const int cw = min(W, ncols - cbase);
        T *dst = xbuf + (size_t)(p & (nbuf - 1)) * WS;
        constexpr int VE = 16 / sizeof(T);
        const int cw_al = cw & ~(VE - 1);
        if (XF) wait_x_slices(xf, x_ready, cbase, cw);
        if (cw_al < cw) {                                 /* ragged tail: generic stores, then the fence */
            for (int i = cw_al; i < cw; ++i) dst[i] = XF ? __ldcg(x + cbase + i) : __ldg(x + cbase + i);
            fence_proxy_async();
        }
        uint64_t *bar = &bars[p & (nbuf - 1)];
        if (cw_al > 0) {
            mbar_expect_tx(bar, (uint32_t)(cw_al * sizeof(T)));
            /* bulk copies of at most 64 KB each */
            uint32_t left = (uint32_t)(cw_al * sizeof(T));
            const char *src = reinterpret_cast<const char *>(x + cbase);
            char *d = reinterpret_cast<char *>(dst);
            while (left) {
                const uint32_t n = left > 65536u ? 65536u : left;
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
        if constexpr (XF) {                               /* block-uniform */
            if (tid == 0) wait_x_slices(xf, x_ready, cbase, cw);
            __syncthreads();
            for (int i = tid; i < cw; i += Tn) dst[i] = __ldcg(x + cbase + i);
        } else {
            for (int i = tid; i < cw; i += Tn) dst[i] = __ldg(x + cbase + i);
        }
    };

    /* request the first two chunks of the matrix stream */
    StreamCursor cur;
    {
        const int2 so = s_slice[warp];
        cur.p = 0; cur.kp = 0; cur.npair = so.y;
        cur.nround = cursor_rounds<U>(so.y);
        cur.base = (size_t)(so.x >> 1) + lane;
    }
    Chunk<T, U> a, b;
    cursor_load<T, U>(a, cur, val2, col2, s_slice, spb, warp, lane, P);
    cursor_load<T, U>(b, cur, val2, col2, s_slice, spb, warp, lane, P);

    /* Everything above read only the resident matrix (immutable) and wrote only shared memory.
     * Launched with programmatic stream serialisation (launch_panel_xf), this grid's CTAs
     * take their SMs as the CTAs of the kernel in front of it in the stream finish one by
     * one, do all of that -- table, sums, the first matrix chunks on their way -- and wait
     * HERE for the rest of that kernel before they touch x or y.  (No-op otherwise.) */
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (use_tma) {
        if (tid == 0) issue_panel(0);
    } else {
        coop_panel(0);
    }

    for (int p = 0; p < P; ++p) {
        /* double buffer: slot (p+1)&1 was released by the barrier that ended
         * panel p-1, so panel p+1 is requested while panel p is computed.
         * single buffer (wide panels): panel p is requested here, after the
         * barrier that ended panel p-1.  The request comes first: thread 0's
         * warp is on everybody's critical path at the next barrier, and the
         * proxy fence of a ragged tail would wait for any load queued before it. */
        if (use_tma && tid == 0) {
            if (nbuf == 2 && p + 1 < P) issue_panel(p + 1);
            if (nbuf == 1 && p > 0) issue_panel(p);
        }
        const ushort4 mt = mt_next;
        if (p + 1 < P) mt_next = meta[((size_t)rb * P + p + 1) * Tn + tid];
        const int npair = s_slice[p * spb + warp].y;

        if (use_tma) {
            mbar_wait(&bars[p & (nbuf - 1)], (uint32_t)((p >> (nbuf - 1)) & 1));
        } else {
            if (nbuf == 2 && p + 1 < P) coop_panel(p + 1);
            if (nbuf == 1 && p > 0) coop_panel(p);
            __syncthreads();
        }
        const T *xs = xbuf + (size_t)(p & (nbuf - 1)) * WS;

        int row_cur = mt.x;
        const int row_nxt = mt.y, sw = mt.z;
        T acc = sums[row_cur];
        /* an empty slice still takes one (fully predicated) round: see cursor_rounds() */
        for (int kp = 0; kp < max(npair, 1); kp += 2 * U) {
            acc = consume_chunk<T, U>(a, xs, sums, acc, kp, npair, sw, row_cur, row_nxt);
            cursor_load<T, U>(a, cur, val2, col2, s_slice, spb, warp, lane, P);
            acc = consume_chunk<T, U>(b, xs, sums, acc, kp + U, npair, sw, row_cur, row_nxt);
            cursor_load<T, U>(b, cur, val2, col2, s_slice, spb, warp, lane, P);
        }
        sums[row_cur] = acc;
        __syncthreads();            /* panel p consumed: its x buffer and the sums are free */
    }
    /* epilogue: the rows of the block, coalesced; optionally the block's share of
     * dotv . y on the way out (NPB conj_grad's d = p.q, cg.f:573-576: q is in shared memory
     * right here, so the dot product costs one read of p instead of a kernel of its own).
     * Fixed order -- per-thread stride, xor-shuffle tree, warp sums left to right --, so the
     * result is the same on every launch. */
    T dacc = (T)0;
    for (int i = tid; i < R; i += Tn) {
        const int row = rb * R + i;
        if (row < rows) {
            const T v = sums[i];
            y[row] = v;
            if (dotv) dacc += v * dotv[row];
        }
    }
    if (dotv) {                                           /* block-uniform */
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
        __syncthreads();                                  /* everybody has read its sums */
        if (lane == 0) sums[warp] = dacc;
        __syncthreads();
        if (tid == 0) {
            T tot = (T)0;
            for (int w = 0; w < spb; ++w) tot += sums[w];
            dot_partial[rb] = tot;
        }
    }
}

size_t panel_smem_bytes(const DevPanel &pm, bool f32)
{
    const size_t es = f32 ? 4 : 8;
    const size_t soff = 16 + (size_t)pm.P * (pm.R / pm.G / 32) * 8;
    const size_t xoff = (soff + (size_t)pm.R * es + 15) & ~(size_t)15;
    const size_t ws = (size_t)pm.W + 16 / es;
    return xoff + (size_t)pm.nbuf * ws * es;
}

template <typename T, int U, int MAXT, bool XF>
static void launch_panel_xf(const DevPanel &pm, const T *x, T *y, const T *dotv, T *dot_partial,
                            const XFlags &xf, cudaStream_t s)
{
    static unsigned attr_set = 0;                  /* function attributes are per device */
    if (!attr_done(&attr_set))
        cudaFuncSetAttribute(spmv_panel_kernel<T, U, MAXT, XF>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const size_t smem = panel_smem_bytes(pm, sizeof(T) == 4);
    const int use_tma = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && pm.use_tma;
    /* B200_SPMV_PDL (default 1): programmatic dependent launch -- back-to-back products (a
     * CG loop, bench.py's timed steps) overlap one kernel's prologue with the tail of the one
     * before.  Not while the stream is being captured into a graph. */
    static int pdl = -1;
    if (pdl < 0) { const char *v = getenv("B200_SPMV_PDL"); pdl = (v && *v) ? atoi(v) : 1; }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (pdl && (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)) {
        cudaGetLastError();
        cap = cudaStreamCaptureStatusActive;
    }
    if (pdl && cap == cudaStreamCaptureStatusNone) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)pm.nblk);
        cfg.blockDim = dim3((unsigned)(pm.R / pm.G));
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, spmv_panel_kernel<T, U, MAXT, XF>,
                           static_cast<const T *>(pm.val), (const uint16_t *)pm.col, (const ushort4 *)pm.meta,
                           (const int *)pm.slice_off, x, y, pm.rows, pm.ncols, pm.P, pm.W, pm.R, use_tma, pm.nbuf,
                           dotv, dot_partial, xf);
        return;
    }
    spmv_panel_kernel<T, U, MAXT, XF><<<pm.nblk, pm.R / pm.G, smem, s>>>(
        static_cast<const T *>(pm.val), pm.col, pm.meta, pm.slice_off, x, y,
        pm.rows, pm.ncols, pm.P, pm.W, pm.R, use_tma, pm.nbuf, dotv, dot_partial, xf);
}

template <typename T, int U, int MAXT>
static void launch_panel_cfg(const DevPanel &pm, const T *x, T *y, const T *dotv, T *dot_partial,
                             const XFlags &xf, cudaStream_t s)
{
    if (xf.flags) launch_panel_xf<T, U, MAXT, true>(pm, x, y, dotv, dot_partial, xf, s);
    else          launch_panel_xf<T, U, MAXT, false>(pm, x, y, dotv, dot_partial, xf, s);
}

template <typename T>
void launch_panel(const DevPanel &pm, const T *x, T *y, cudaStream_t s, const T *dotv, T *dot_partial,
                  const XFlags *flags)
{
    if (pm.nblk <= 0) return;
    const XFlags none = {nullptr, 0ull, 1, 0, nullptr, 0ull, 0};
    const XFlags &xf = flags ? *flags : none;
    const int threads = pm.R / pm.G;
    if (threads <= 256 && pm.U >= 5) {
        /* few warps per SM (NPB class A / B sized row blocks): the register file
         * is free, so each lane keeps twice as many pairs of the stream in flight (12 pairs
         * per chunk spill and are no faster: profiles/r02_run6_sweep.txt) */
        if (pm.U >= 10) launch_panel_cfg<T, 10, 256>(pm, x, y, dotv, dot_partial, xf, s);
        else launch_panel_cfg<T, 8, 256>(pm, x, y, dotv, dot_partial, xf, s);     /* best on class B (profiles/r01_run23) */
    } else if (pm.U >= 5) {
        launch_panel_cfg<T, 5, 512>(pm, x, y, dotv, dot_partial, xf, s);
    } else if (pm.U == 3) {
        launch_panel_cfg<T, 3, 512>(pm, x, y, dotv, dot_partial, xf, s);
    } else {
        launch_panel_cfg<T, 4, 512>(pm, x, y, dotv, dot_partial, xf, s);           /* <= 128 registers per thread */
    }
}
template void launch_panel<double>(const DevPanel &, const double *, double *, cudaStream_t, const double *, double *,
                                   const XFlags *);
template void launch_panel<float>(const DevPanel &, const float *, float *, cudaStream_t, const float *, float *,
                                  const XFlags *);

}  // namespace b200
