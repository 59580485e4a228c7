/*
 * spmv_panelr.cu -- PANEL with the matrix stream in a shared-memory ring
 * ("ring kernel", DevPanel::fmt == 2; layout built by spmv_panelg.cu).
 *
 * What limits a register-staged PANEL kernel on wide matrices (NPB class D
 * row blocks: ~70-140 panels, 4-7 entries per (row, panel)) is how much of
 * the matrix stream is in flight: a lane holds two chunks of U pairs, every
 * panel is walked in whole chunks, so with a few pairs per lane and panel the
 * prefetch reaches one panel ahead (~48 KB per SM), and the loads in flight
 * also need L1 lines, which the x slices in shared memory take away
 * (profiles/r01_run28, r01_run29).  Here the stream never touches registers
 * or L1 on its way in:
 *
 *   - the slices are stored warp-major -- (row block, warp, panel) -- so a
 *     warp's stream over ALL panels of its row block is one contiguous run of
 *     pair rows; a pair row is 32 value pairs followed by its 32 column pairs
 *     (640 bytes in fp64);
 *   - every warp owns a ring of S stages of K pair rows in shared memory and
 *     refills a stage with ONE TMA bulk copy as soon as it has consumed it,
 *     one mbarrier per stage (the number of bulk copies an SM issues matters:
 *     K = 2 instead of 4 pair rows per copy costs 20 %, a second copy per
 *     stage 3 %).  Stages ignore panel boundaries, so the bytes in flight are
 *     the ring size whatever the panel geometry;
 *   - the consumer: x slice of the panel in shared memory (TMA; one wide
 *     slice by default, two with nbuf = 2), bit 15 of a column starts a new
 *     row of the lane -- on the same pair in all 32 lanes, so one warp vote
 *     per pair keeps the switch code off the common path --, running sums
 *     parked in shared memory with the next row's sum pre-loaded, separately
 *     rounded multiply and add in the reference's order
 *     (libspmv/native-impl.c:1-12) => bit-identical results for sorted rows.
 */
#include "panel_common.cuh"

namespace b200 {

/* running-sum state of a lane: the row being accumulated and, already loaded,
 * the parked sum of the row that comes next -- so a row switch costs a store
 * and two moves, not a dependent shared-memory round trip */
template <typename T, int G>
struct LaneRows {
    RowIds<G> ids;     /* rows after `nxt`, front to back */
    int cur, nxt;      /* tile-local row ids (R = dummy slot) */
    T acc, acc_nxt;
};

template <typename T, int G>
__device__ __forceinline__ void switch_row(LaneRows<T, G> &st, T *sums)
{
    sums[st.cur] = st.acc;
    st.cur = st.nxt;
    st.acc = st.acc_nxt;
    st.nxt = pop_id<G>(st.ids);
    st.acc_nxt = sums[st.nxt];        /* rows of a lane are distinct: never the slot just stored */
}

/* consume n <= K pair rows of one ring stage, rv / rc pointing at the first of them
 * (FULL: n == K, no per-row predicates).  Row slots are padded to a common length inside
 * a slice (panelg_sort_kernel), so the flags of the 32 lanes fall on the same pair: one
 * warp vote per pair keeps the switch code off the common path. */
template <typename T, int G, int K, bool FULL>
__device__ __forceinline__ void consume_rows(const unsigned char *rows, int lane,
                                             int n, const T *xs, T *sums, LaneRows<T, G> &st)
{
    using P2 = typename PairT<T>::type;
    constexpr int kRowB = 32 * ((int)sizeof(P2) + 4);     /* bytes of a pair row in the ring */
    P2 v[K];
    uint32_t c[K];
    T xa[K], xb[K];
#pragma unroll
    for (int u = 0; u < K; ++u) {
        if (FULL || u < n) {
            v[u] = reinterpret_cast<const P2 *>(rows + u * kRowB)[lane];
            c[u] = reinterpret_cast<const uint32_t *>(rows + u * kRowB + 32 * sizeof(P2))[lane];
        }
    }
#pragma unroll
    for (int u = 0; u < K; ++u) {
        if (FULL || u < n) {
            xa[u] = xs[c[u] & 0x7FFFu];
            xb[u] = xs[(c[u] >> 16) & 0x7FFFu];
        }
    }
#pragma unroll
    for (int u = 0; u < K; ++u) {
        if (FULL || u < n) {
            const T p0 = pmul(v[u].x, xa[u]), p1 = pmul(v[u].y, xb[u]);
            if (__any_sync(0xffffffffu, c[u] & 0x80008000u)) {
                if (c[u] & 0x8000u) switch_row<T, G>(st, sums);
                st.acc = padd(st.acc, p0);
                if (c[u] & 0x80000000u) switch_row<T, G>(st, sums);
                st.acc = padd(st.acc, p1);
            } else {
                st.acc = padd(padd(st.acc, p0), p1);
            }
        }
    }
}

#ifndef B200_RING_WAR_FENCE
#define B200_RING_WAR_FENCE 0
#endif
#ifndef B200_RING_TAILS
#define B200_RING_TAILS 0      /* size-specialised tail batches: measured 1 % slower, see below */
#endif
#ifndef B200_XCOPY_BYTES
#define B200_XCOPY_BYTES 262144
#endif
constexpr uint32_t kXCopyBytes = B200_XCOPY_BYTES;      /* largest single bulk copy of an x slice */

template <typename T, int G, int K, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
spmv_panelr_kernel(const unsigned char *__restrict__ stream,
                   const uint16_t *__restrict__ rowids, const int *__restrict__ slice_off,
                   const T *__restrict__ x, T *__restrict__ y,
                   int rows, int ncols, int P, int W, int R, int use_tma, int nbuf, int S, XFlags xf, XPush xp)
{
    using P2 = typename PairT<T>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Tn = blockDim.x;
    const int spb = Tn >> 5;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rb = blockIdx.x;
    /* layout: [x mbarriers 16 B][ring mbarriers spb*S*8][sums R+1][xbuf nbuf*(W+pad)][rings: spb*S*K pair rows] */
    uint64_t *xbars = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *rbars = reinterpret_cast<uint64_t *>(smem_raw + 16);
    const size_t soff = 16 + (size_t)spb * S * 8;
    T *sums = reinterpret_cast<T *>(smem_raw + soff);
    const size_t xoff = (soff + (size_t)(R + 1) * sizeof(T) + 15) & ~(size_t)15;
    const int WS = W + (16 / (int)sizeof(T));
    T *xbuf = reinterpret_cast<T *>(smem_raw + xoff);
    const size_t roff = (xoff + (size_t)nbuf * WS * sizeof(T) + 127) & ~(size_t)127;
    constexpr int kRowB = 32 * ((int)sizeof(P2) + 4);     /* a pair row: 32 value pairs + 32 column pairs */

    for (int i = tid; i <= R; i += Tn) sums[i] = (T)0;
    if (tid == 0) {
        xbuf[W] = (T)0;                                   /* padding slot, never overwritten */
        if (nbuf == 2) xbuf[WS + W] = (T)0;
        mbar_init(&xbars[0], 1);
        mbar_init(&xbars[1], 1);
    }
    if (lane == 0)
        for (int s = 0; s < S; ++s) mbar_init(&rbars[warp * S + s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    /* this warp's stream: pair rows [0, total) of 32 lanes x (value pair, column pair) */
    const int *woff = slice_off + ((size_t)rb * spb + warp) * P;
    const int off0 = __ldg(woff);
    const int total = (__ldg(woff + P) - off0) >> 6;
    const int nstage = (total + K - 1) / K;
    const unsigned char *gstream = stream + (size_t)(off0 >> 6) * kRowB;
    unsigned char *ring_w = smem_raw + roff + (size_t)warp * S * K * kRowB;
    uint64_t *rb_w = rbars + warp * S;

    auto issue_stage = [&](int t, int slot) {            /* lane 0: stage t -> slot t % S */
        const int r0 = t * K;
        const int nr = min(K, total - r0);
        const uint32_t bytes = (uint32_t)(nr * kRowB);     /* one copy: values and columns together */
        uint64_t *bar = rb_w + slot;
        mbar_expect_tx(bar, bytes);
        tma_bulk_g2s(ring_w + (size_t)slot * K * kRowB, gstream + (size_t)r0 * kRowB, bytes, bar);
    };
    if (lane == 0)
        for (int t = 0; t < S && t < nstage; ++t) issue_stage(t, t);

    /* The exchange, fused: with the first stages of the matrix stream requested, every CTA
     * stores its share of this rank's x slice into all ranks' buffers.  Epoch e goes to the
     * buffer the products of epoch e - 2 read last: report e - 1 as consumed (the product
     * that read it precedes this kernel in the stream), wait for everybody's report of e - 2. */
    if (xp.src) {
        if (blockIdx.x == 0 && tid == 0 && xp.epoch > 1) {
            __threadfence_system();
            for (int j = 0; j < xp.nranks; ++j) st_relaxed_sys_u64(xp.rflag[j] + xp.rank, xp.epoch - 1);
        }
        if (xp.epoch > 2) {
            if (tid < xp.nranks)
                while (ld_acquire_sys_u64(xp.rflag[xp.rank] + tid) < xp.epoch - 2) { }
            __syncthreads();
        }
        const size_t n16 = xp.bytes >> 4;                     /* slices are 16-byte granular */
        const size_t per = (n16 + gridDim.x - 1) / gridDim.x;
        const size_t i0 = min(n16, per * blockIdx.x), i1 = min(n16, i0 + per);
        const int4 *src = reinterpret_cast<const int4 *>(xp.src);
        /* four loads in flight per thread, then the stores: a load-store-load-store loop would
         * pay the load latency once per element (source and destinations may alias for all
         * the compiler knows) */
        for (size_t base = i0; base < i1; base += (size_t)Tn * 4) {
            int4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const size_t i = base + (size_t)q * Tn + tid;
                if (i < i1) v[q] = __ldg(src + i);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const size_t i = base + (size_t)q * Tn + tid;
                if (i < i1)
                    for (int j = 0; j < xp.nranks; ++j)
                        reinterpret_cast<int4 *>(static_cast<char *>(xp.dst[j]) + xp.offset)[i] = v[q];
            }
        }
        /* one system-scope fence per CTA after the barrier (cumulative), one more in the CTA
         * that arrives last, then relaxed flag stores: a st.release.sys per rank would repeat
         * the fence -- an NVLink round trip -- once per rank */
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            const unsigned int prev = atomicAdd(xp.counter, 1u);
            if (prev == gridDim.x - 1) {                      /* everybody's stores are out */
                *xp.counter = 0;
                __threadfence_system();
                for (int j = 0; j < xp.nranks; ++j) st_relaxed_sys_u64(xp.vflag[j] + xp.rank, xp.epoch);
            }
        }
    }

    /* x assembled from the slices of several GPUs: the slices a panel needs must have
     * arrived (epoch flag written with st.release.sys by the pushing rank after its stores)
     * before the panel is requested.  Columns are walked left to right, so `rank_ready`
     * only grows; the product overlaps the exchange instead of running after it. */
    int rank_ready = 0;                                   /* ranks [0, rank_ready) have arrived */
    auto wait_slices = [&](int cbase, int cw) {           /* one thread */
        const int r1 = min((cbase + cw - 1) / xf.cols_per_rank, xf.nranks - 1);
        if (r1 < rank_ready) return;
        for (; rank_ready <= r1; ++rank_ready)
            while (ld_acquire_sys_u64(xf.flags + rank_ready) < xf.epoch) { }
        /* the slice was written through the generic proxy (by other GPUs), the bulk copy
         * reads it through the async proxy */
        asm volatile("fence.proxy.async.global;" ::: "memory");
    };
    auto issue_panel = [&](int p) {                       /* thread 0 only (TMA path) */
        const int cbase = p * W;
        const int cw = min(W, ncols - cbase);
        T *dst = xbuf + (size_t)(p & (nbuf - 1)) * WS;
        constexpr int VE = 16 / sizeof(T);
        const int cw_al = cw & ~(VE - 1);
        if (xf.flags) wait_slices(cbase, cw);
        if (cw_al < cw) {                                 /* ragged tail: generic stores, then the fence */
            for (int i = cw_al; i < cw; ++i) dst[i] = __ldcg(x + cbase + i);
            fence_proxy_async();
        }
        uint64_t *bar = &xbars[p & (nbuf - 1)];
        mbar_expect_tx(bar, (uint32_t)(cw_al * sizeof(T)));
        uint32_t left = (uint32_t)(cw_al * sizeof(T));
        const char *src = reinterpret_cast<const char *>(x + cbase);
        char *d = reinterpret_cast<char *>(dst);
        /* few, large copies: issuing a bulk copy costs the issuing warp ~15 instructions and
         * that warp is on everybody's critical path at the panel barrier */
        while (left) {
            const uint32_t n = left > kXCopyBytes ? kXCopyBytes : left;
            tma_bulk_g2s(d, src, n, bar);
            d += n; src += n; left -= n;
        }
    };
    auto coop_panel = [&](int p) {                        /* all threads (x not 16-byte aligned) */
        const int cbase = p * W;
        const int cw = min(W, ncols - cbase);
        T *dst = xbuf + (size_t)(p & (nbuf - 1)) * WS;
        if (xf.flags) {                                   /* block-uniform */
            if (tid == 0) wait_slices(cbase, cw);
            __syncthreads();
        }
        for (int i = tid; i < cw; i += Tn) dst[i] = __ldcg(x + cbase + i);
    };
    if (use_tma) {
        if (tid == 0) issue_panel(0);
    } else {
        coop_panel(0);
    }

    const uint16_t *my_ids = rowids + ((size_t)rb * P * Tn + tid) * G;
    RowIds<G> ids_next = load_ids<G>(my_ids);
    int o_cur = off0;
    int o_nxt = __ldg(woff + 1);
    int t_cur = 0;          /* stage being consumed */
    int slot = 0;           /* its ring slot, t_cur % S */
    uint32_t par = 0;       /* mbarrier phase of that slot, (t_cur / S) & 1 */
    int sr = 0;             /* pair rows of the stage already consumed */

    for (int p = 0; p < P; ++p) {
        /* first thing after the barrier: request the x slice (thread 0's warp is on
         * everybody's critical path, so nothing is queued in front of the request).
         * CTAs are deliberately NOT kept in step with each other: a soft grid barrier every
         * n panels (all CTAs are co-resident) made the class D 1/8 block slower the more
         * often it ran -- 275 us without, 295 us every 16 panels, 391 us every panel
         * (profiles/r02_run7_sweep.txt); staggered slice requests suit the L2 better. */
        if (use_tma && tid == 0) {
            if (nbuf == 2 && p + 1 < P) issue_panel(p + 1);
            if (nbuf == 1 && p > 0) issue_panel(p);
        }
        LaneRows<T, G> st;
        st.ids = ids_next;
        if (p + 1 < P) ids_next = load_ids<G>(my_ids + (size_t)(p + 1) * Tn * G);
        const int npair = (o_nxt - o_cur) >> 6;
        o_cur = o_nxt;
        if (p + 1 < P) o_nxt = __ldg(woff + p + 2);

        if (use_tma) {
            mbar_wait(&xbars[p & (nbuf - 1)], (uint32_t)((p >> (nbuf - 1)) & 1));
        } else {
            if (nbuf == 2 && p + 1 < P) coop_panel(p + 1);
            if (nbuf == 1 && p > 0) coop_panel(p);
            __syncthreads();
        }
        const T *xs = xbuf + (size_t)(p & (nbuf - 1)) * WS;

        /* slot R: dummy "current row" until the first flag.  Once a lane's ids are used up
         * pop_id returns 0 and the pre-load reads sums[0]: harmless, it is never switched to */
        st.cur = R;
        st.acc = (T)0;
        st.nxt = pop_id<G>(st.ids);
        st.acc_nxt = sums[st.nxt];
        int kp = 0;
        while (kp < npair) {                              /* warp-uniform control flow */
            if (sr == 0) mbar_wait(rb_w + slot, par);
            const int n = min(npair - kp, K - sr);
            const unsigned char *rows_at = ring_w + (size_t)(slot * K + sr) * kRowB;
            if (n == K) {
                consume_rows<T, G, K, true>(rows_at, lane, n, xs, sums, st);
            } else {
                /* head / tail of a panel inside a stage: one pair row at a time.  Lean code wins
                 * here: a predicated K-batch was slower, and so were batches specialised on
                 * their size (B200_RING_TAILS=1: class D 1/8 block 274.8 us against 273.2 us,
                 * 1/2 block 853 us against 842 us, profiles/r02_run8_sweep*.txt) */
#if B200_RING_TAILS
                if (K >= 4 && n == 3) {
                    consume_rows<T, G, 3, true>(rows_at, lane, 3, xs, sums, st);
                } else if (n == 2) {
                    consume_rows<T, G, 2, true>(rows_at, lane, 2, xs, sums, st);
                } else
#endif
                {
                    for (int u = 0; u < n; ++u)
                        consume_rows<T, G, 1, true>(rows_at + u * kRowB, lane, 1, xs, sums, st);
                }
            }
            kp += n;
            sr += n;
            if (sr == K) {                                /* stage consumed: refill its slot */
                /* every lane has used its values of the stage (they fed arithmetic), so the
                 * slot can be overwritten: same hand-over as an "empty" mbarrier arrive */
                __syncwarp();
                if (lane == 0 && t_cur + S < nstage) {
                    /* Write-after-read across proxies: the generic-proxy reads of the slot have
                     * returned their data (it fed arithmetic) and __syncwarp orders them before
                     * this thread, so the bulk copy cannot overtake them -- the same hand-over
                     * as a consumer's arrive on an "empty" mbarrier, which CUTLASS pipelines do
                     * not fence either.  The explicit proxy fence costs 3 % of the kernel
                     * (profiles/r02_run2: class D 1/8 block 277 us against 269 us). */
#if B200_RING_WAR_FENCE
                    fence_proxy_async();
#endif
                    issue_stage(t_cur + S, slot);
                }
                ++t_cur;
                sr = 0;
                if (++slot == S) { slot = 0; par ^= 1u; }
            }
        }
        sums[st.cur] = st.acc;
        __syncthreads();            /* panel p consumed: its x buffer and the sums are free */
    }
    for (int i = tid; i < R; i += Tn) {
        const int row = rb * R + i;
        if (row < rows) y[row] = sums[i];
    }
}

static size_t ring_bytes(const DevPanel &pm, bool f32)
{
    return (size_t)(pm.R / pm.G / 32) * pm.ring_S * pm.ring_K * 32 * ((f32 ? 8 : 16) + 4);
}

size_t panelr_smem_bytes(const DevPanel &pm, bool f32)
{
    const size_t es = f32 ? 4 : 8;
    const size_t soff = 16 + (size_t)(pm.R / pm.G / 32) * pm.ring_S * 8;
    const size_t xoff = (soff + (size_t)(pm.R + 1) * es + 15) & ~(size_t)15;
    const size_t ws = (size_t)pm.W + 16 / es;
    const size_t roff = (xoff + (size_t)pm.nbuf * ws * es + 127) & ~(size_t)127;
    return roff + ring_bytes(pm, f32);
}

template <typename T, int G, int K, int MAXT>
static void launch_panelr_cfg(const DevPanel &pm, const T *x, T *y, const XFlags &xf, const XPush &xp, cudaStream_t s)
{
    static unsigned attr_set = 0;                  /* function attributes are per device */
    if (!attr_done(&attr_set))
        cudaFuncSetAttribute(spmv_panelr_kernel<T, G, K, MAXT>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const size_t smem = panelr_smem_bytes(pm, sizeof(T) == 4);
    const int use_tma = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && pm.use_tma;
    spmv_panelr_kernel<T, G, K, MAXT><<<pm.nblk, pm.R / pm.G, smem, s>>>(
        static_cast<const unsigned char *>(pm.val), pm.rowids, pm.slice_off, x, y,
        pm.rows, pm.ncols, pm.P, pm.W, pm.R, use_tma, pm.nbuf, pm.ring_S, xf, xp);
}

template <typename T, int G>
static void launch_panelr_g(const DevPanel &pm, const T *x, T *y, const XFlags &xf, const XPush &xp, cudaStream_t s)
{
    const int threads = pm.R / pm.G;
    if (threads > 512) {
        if (pm.ring_K == 2) launch_panelr_cfg<T, G, 2, 768>(pm, x, y, xf, xp, s);
        else                launch_panelr_cfg<T, G, 4, 768>(pm, x, y, xf, xp, s);
    } else {
        if (pm.ring_K == 2) launch_panelr_cfg<T, G, 2, 512>(pm, x, y, xf, xp, s);
        else                launch_panelr_cfg<T, G, 4, 512>(pm, x, y, xf, xp, s);
    }
}

template <typename T>
void launch_panelr(const DevPanel &pm, const T *x, T *y, const XFlags &xf, const XPush &xp, cudaStream_t s)
{
    if (pm.nblk <= 0) return;
    if (pm.G == 2)      launch_panelr_g<T, 2>(pm, x, y, xf, xp, s);
    else if (pm.G == 4) launch_panelr_g<T, 4>(pm, x, y, xf, xp, s);
    else                launch_panelr_g<T, 8>(pm, x, y, xf, xp, s);
}
template void launch_panelr<double>(const DevPanel &, const double *, double *, const XFlags &, const XPush &, cudaStream_t);
template void launch_panelr<float>(const DevPanel &, const float *, float *, const XFlags &, const XPush &, cudaStream_t);

}  // namespace b200
