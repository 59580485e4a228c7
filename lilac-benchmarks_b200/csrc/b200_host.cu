/*
 * b200_host.cu -- host layer of libb200-spmv.so, part 1: per-device contexts,
 * upload (layout selection and build) and launch of a resident CSR row block,
 * i.e. the resident-matrix API of include/b200_spmv.h part 2.  Thin C-style
 * code over the CUDA runtime only (no cuSPARSE, no cuBLAS, no CPU fallback).
 * The libspmv ABI itself (cache, write guard, x / y movement, several devices)
 * lives in b200_dropin.cu.
 *
 * Reference behaviour mirrored (and where it deliberately differs):
 *   libspmv/gpu.c:36-85    setup(): lazy one-time init, device buffers sized
 *                          from (rows, cols, nnz).  Here buffers belong to a
 *                          resident matrix and are freed with it (gpu.c leaks
 *                          the old ones on every shape change).
 *   libspmv/gpu.c:213-223  column count = max(colidx); gpu.c scans the host
 *                          array with an off-by-one loop, here a device
 *                          reduction over the uploaded colidx does it.
 *   libspmv/gpu.c:42-80    errors: assert -> here message on stderr + abort().
 *
 * State is per device (DevCtx): a process may hold resident matrices on several
 * GPUs; every entry point makes the device of the object it works on current
 * and restores the caller's device before it returns.
 */
#include "host_internal.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <cmath>
#include <vector>

using namespace b200;

#define B200_VERSION "b200-spmv 0.2 (sm_100a)"

namespace b200 {

void die(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "libb200-spmv: fatal: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    abort();
}

double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
int g_verbose = 0;

static bool g_ready = false;
static int g_default_device = -1;
static int g_device_count = 0;
static DevCtx *g_ctx[kMaxDevices * 8];      /* indexed by device ordinal */

void ensure_init_locked(int device)
{
    if (!g_ready) {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            die("no CUDA device available (%s); this platform has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        g_device_count = std::min(count, (int)(sizeof g_ctx / sizeof g_ctx[0]));
        g_verbose = env_int("B200_SPMV_VERBOSE", 0);
        g_ready = true;
    }
    /* the device of the drop-in path: an explicit b200_spmv_init(d) wins, then
     * B200_SPMV_DEVICE, then whatever device is current at the first call */
    if (device >= 0) {
        if (device >= g_device_count) die("device %d does not exist (%d visible)", device, g_device_count);
        g_default_device = device;
    } else if (g_default_device < 0) {
        int d = env_int("B200_SPMV_DEVICE", -1);
        if (d < 0) CUDA_OK(cudaGetDevice(&d));
        if (d >= g_device_count) die("device %d does not exist (%d visible)", d, g_device_count);
        g_default_device = d;
    }
}

int default_device_locked(void) { return g_default_device; }

DevCtx *ctx_for_device_locked(int device)
{
    if (device < 0 || device >= g_device_count) die("device %d out of range", device);
    if (g_ctx[device]) return g_ctx[device];
    DeviceScope scope(device);
    DevCtx *c = (DevCtx *)calloc(1, sizeof *c);
    c->device = device;
    CUDA_OK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreate(&c->ev0));
    CUDA_OK(cudaEventCreate(&c->ev1));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_x, cudaEventDisableTiming));
    g_ctx[device] = c;
    return c;
}

}  // namespace b200

/* ------------------------------------------------------------------------
 * upload
 * ---------------------------------------------------------------------- */
/* ------------------------------------------------------------------------
 * upload
 * ---------------------------------------------------------------------- */
static int kernel_from_env(int requested)
{
    if (requested != B200_KERNEL_AUTO) return requested;
    const char *v = getenv("B200_SPMV_KERNEL");
    if (!v || !*v) return B200_KERNEL_AUTO;
    if (!strcmp(v, "ordered")) return B200_KERNEL_ORDERED;
    if (!strcmp(v, "vector")) return B200_KERNEL_VECTOR;
    if (!strcmp(v, "panel")) return B200_KERNEL_PANEL;
    if (!strcmp(v, "merge")) return B200_KERNEL_MERGE;
    if (!strcmp(v, "sell")) return B200_KERNEL_SELL;
    if (!strcmp(v, "small")) return B200_KERNEL_SMALL;
    if (!strcmp(v, "auto")) return B200_KERNEL_AUTO;
    die("B200_SPMV_KERNEL=%s is not one of auto|ordered|vector|panel|sell|merge|small", v);
    return 0;
}

/* greedy nnz-split: consecutive rows are packed into a block until the next
 * row would overflow the tile (or the row cap); a row longer than the tile
 * forms a block of its own. */
static void build_row_blocks(const int *rowstr, int rows, int tile, std::vector<int> &blk)
{
    blk.clear();
    blk.push_back(0);
    const int cap = tile - 2;          /* slack for the 16-byte aligned start */
    int r = 0;
    while (r < rows) {
        const int start = r;
        const long long base = rowstr[r];
        long long used = 0;
        while (r < rows && r - start < kRowsPerBlock) {
            const long long len = (long long)rowstr[r + 1] - rowstr[r];
            if (used + len > cap) break;
            used += len;
            ++r;
        }
        if (r == start) ++r;           /* single long row */
        (void)base;
        blk.push_back(r);
    }
}

/* Column-panel layout: worth it when x does not fit L1 but the matrix is dense
 * enough per (row, panel) that staging x slices in shared memory pays:
 * sorted rows (order-preserving), a handful of panels, >= 4 entries per
 * (row, panel) on average. */
static const size_t kSmemMax = 227 * 1024;

struct PanelPlan { int P, W, R, G, nbuf, fmt, ring_K, ring_S; };

/* RING layout (spmv_panelg.cu + spmv_panelr.cu): tall row blocks of G * T rows for
 * wide matrices.  One CTA per SM and as few passes over x as the shared-memory
 * budget (R + 1 running sums) allows; the rest of the 227 KB holds the per-warp
 * rings of the matrix stream and two x slices. */
static bool panel_plan_ring(const b200_matrix *m, PanelPlan *pl)
{
    const int fmt = 2;
    if (m->rows <= 0 || m->nnz <= 0 || m->ncols <= 0) return false;
    if (m->scan.rows_unsorted != 0) return false;
    const size_t es = elem_size(m->dtype);
    /* up to 256 panels: NPB class D blocks need 71; class E blocks (9 M columns, 426 panels,
     * 1.6 entries per (row, panel)) are left on the SELL path they were measured on */
    const int kRmax = 4096, kMaxPanels = 256;
    const int kTmax = std::max(64, std::min(768, env_int("B200_SPMV_PANEL_TMAX", 512)));
    int R = env_int("B200_SPMV_PANEL_ROWS", 0);
    if (R <= 0) {
        const int passes = (int)((m->rows + (long long)m->ctx->sm_count * kRmax - 1) / ((long long)m->ctx->sm_count * kRmax));
        R = (int)((m->rows + (long long)m->ctx->sm_count * passes - 1) / ((long long)m->ctx->sm_count * passes));
    }
    R = std::max(64, std::min(kRmax, R));
    int G = env_int("B200_SPMV_PANEL_G", 0);
    if (G != 2 && G != 4 && G != 8) G = R <= 2 * kTmax ? 2 : R <= 4 * kTmax ? 4 : 8;
    int Tn = (R + G - 1) / G;
    Tn = std::max(32, std::min(kTmax, (Tn + 31) & ~31));
    R = Tn * G;
    const int spb = Tn / 32;
    /* one wide x slice beats two narrower ones here: the per-panel cost (barrier, hand-over)
     * outweighs the exposed slice load, 169 KB at ~216 GB/s per SM (profiles/r01_run47_sweep.txt) */
    int nbuf = env_int("B200_SPMV_PANEL_NBUF", 1) == 2 ? 2 : 1;
    /* the stream lands in a shared-memory ring (S stages of K pair rows per warp), nothing
     * is left for the L1 to do, so the whole 227 KB is used */
    int ring_K = 0, ring_S = 0;
    size_t ring = 0;
    const size_t budget = std::min<size_t>(kSmemMax, (size_t)env_int("B200_SPMV_PANEL_SMEM_KB", 227) * 1024);
    {
        ring_K = env_int("B200_SPMV_PANEL_RING_K", 4) == 2 ? 2 : 4;
        const size_t stage = (size_t)spb * ring_K * 32 * (2 * es + 4);
        /* two stages per warp: a deeper ring takes the room from the x slices, and more,
         * narrower panels cost more than the extra bytes in flight gain (profiles/r01_run34) */
        ring_S = std::max(2, std::min(32, env_int("B200_SPMV_PANEL_RING_S", 2)));
        ring = stage * ring_S + (size_t)spb * ring_S * 8 + 256;
    }
    auto width_that_fits = [&](int P) {
        (void)P;
        const size_t fixed = (((size_t)16 + (size_t)(R + 1) * es + 15) & ~(size_t)15) + 64 + ring;
        if (fixed >= budget) return 0;
        return (int)(std::min<long long>(32736, (long long)((budget - fixed) / (nbuf * es)) - 4) & ~31);
    };
    int wmax = env_int("B200_SPMV_PANEL_COLS", 0);
    int P = 1, W = 0;
    for (int it = 0; it < 8; ++it) {                 /* P and the slice table size depend on each other */
        int w = width_that_fits(P);
        if (wmax > 0) w = std::min(w, wmax & ~31);
        if (w < 32) return false;
        const int Pn = (m->ncols + w - 1) / w;
        W = w;
        if (Pn <= P) break;
        P = Pn;
    }
    if (P > kMaxPanels || (long long)P * W < m->ncols) return false;
    W = std::min(W, ((((m->ncols + P - 1) / P) + 31) & ~31));
    P = (m->ncols + W - 1) / W;                      /* rounding W up may have emptied the last panels */
    pl->P = P; pl->W = W; pl->R = R; pl->G = G; pl->nbuf = nbuf; pl->fmt = fmt;
    pl->ring_K = ring_K; pl->ring_S = ring_S;
    return true;
}

/* modelled launch time of a ring plan against the SELL family, both fitted on NPB
 * class D row blocks (profiles/r01_run47_sweep.txt, r01_run48_sweep_ring_single_buffer.txt):
 * the ring kernel streams ~5.5 TB/s of stored entries and pays ~1.4 us per panel
 * (barrier, x slice hand-over and load) in every wave of CTAs; SELL is bound by the
 * L1TEX wavefront rate of its gather, ~1.29 clk per entry and SM. */
static bool panel_plan_beats_sell(const b200_matrix *m, const PanelPlan &pl)
{
    const double es = (double)elem_size(m->dtype);
    const double nblk = (double)((m->rows + pl.R - 1) / pl.R);
    const double stream = (double)m->nnz * (es + 2) * 1.06 + (double)m->rows * pl.P * 2;
    const double waves = std::ceil(nblk / m->ctx->sm_count);
    const double per_panel = pl.nbuf == 1 ? 1.4e-6 : 0.65e-6;
    const double t_panel = stream / 5.5e12 * std::max(1.0, waves * m->ctx->sm_count / nblk) + waves * pl.P * per_panel;
    const double t_sell = (double)m->nnz * 1.29 / (m->ctx->sm_count * 1.965e9);
    return t_panel < 0.92 * t_sell;
}

static bool panel_applicable(const b200_matrix *m, int g_default, int *P_out, int *W_out, int *R_out,
                             int *G_out, int *nbuf_out)
{
    if (m->rows <= 0 || m->nnz <= 0 || m->ncols <= 0) return false;
    if (m->scan.rows_unsorted != 0) return false;
    const size_t es = elem_size(m->dtype);
    /* g_default: two rows per lane (longest with shortest), or one row per lane when there are
     * so few rows that every lane counts (NPB class A: 95 rows per SM; 16.5 us against 18.5 us,
     * profiles/r02_run1_sweep.txt) */
    int G = env_int("B200_SPMV_PANEL_G", g_default);
    G = G == 1 ? 1 : 2;
    int R = env_int("B200_SPMV_PANEL_ROWS", 0);
    if (R <= 0) R = (m->rows + m->ctx->sm_count - 1) / m->ctx->sm_count;
    const int gran = 32 * G;
    R = std::max(gran, std::min(512 * G, (R + gran - 1) / gran * gran));   /* CTA <= 512 threads */
    const int spb = R / G / 32;
    const int kMaxPanels = 128;
    /* shared memory: 16 B barriers + slice table + R sums + nbuf * (W + pad) x entries */
    auto width_that_fits = [&](int nbuf, int P) {
        const size_t fixed = ((16 + (size_t)P * spb * 8 + (size_t)R * es + 15) & ~(size_t)15) + 64;
        if (fixed >= kSmemMax) return 0;
        return (int)(std::min<long long>(65504, (long long)((kSmemMax - fixed) / (nbuf * es)) - 4) & ~31);
    };
    /* measured on NPB class C (profiles/r01_run5_sweep.txt): 96 KB double-buffered
     * x slices are the sweet spot; wider single-buffered slices only pay when
     * the column range is so wide that the double buffer would leave fewer
     * than ~8 entries per (row, panel). */
    const int w_one = width_that_fits(1, 1);
    const int w_pref = std::min<int>(width_that_fits(2, 16), (int)(96 * 1024 / es));
    int P, W, nbuf;
    const int force_nbuf = env_int("B200_SPMV_PANEL_NBUF", 0);
    if (m->ncols <= w_one && !getenv("B200_SPMV_PANEL_COLS")) {
        P = 1; nbuf = 1;
        W = (m->ncols + 31) & ~31;
    } else {
        int wmax = env_int("B200_SPMV_PANEL_COLS", w_pref);
        nbuf = 2;
        int Pd = (m->ncols + std::max(32, wmax & ~31) - 1) / std::max(32, wmax & ~31);
        const double seg_d = (double)m->nnz / ((double)m->rows * Pd);
        if (force_nbuf == 1 || (force_nbuf == 0 && !getenv("B200_SPMV_PANEL_COLS") && seg_d < 8.0)) {
            nbuf = 1;
            const int w1 = width_that_fits(1, std::min(kMaxPanels, (m->ncols + 16383) / 16384 + 1));
            wmax = getenv("B200_SPMV_PANEL_COLS") ? std::min(wmax, w1) : w1;
        }
        wmax = std::max(32, std::min(wmax, width_that_fits(nbuf, std::min(kMaxPanels, Pd + 1)))) & ~31;
        P = (m->ncols + wmax - 1) / wmax;
        W = (((m->ncols + P - 1) / P) + 31) & ~31;
    }
    const double seg = (double)m->nnz / ((double)m->rows * P);
    if (P > kMaxPanels || (seg < 4.0 && P > 1)) return false;
    *P_out = P; *W_out = W; *R_out = R; *G_out = G; *nbuf_out = nbuf;
    return true;
}

static bool build_panel_locked(b200_matrix *m, bool forced)
{
    int P, W, R, G, nbuf, fmt = 0, ring_K = 0, ring_S = 0;
    const int want_fmt = env_int("B200_SPMV_PANEL_FMT", -1);
    bool ok = false;
    /* one row per lane first when rows are scarce (see panel_applicable), else two */
    const int g_first = m->rows <= 128 * m->ctx->sm_count ? 1 : 2;
    for (int g_try = g_first; g_try <= 2 && !ok && want_fmt <= 0; ++g_try) {
        if (!panel_applicable(m, g_try, &P, &W, &R, &G, &nbuf)) continue;
        /* every CTA loads the whole x once: only worth it while that stays below the
         * matrix stream itself (short, wide row blocks fail this: NPB class D shards) */
        const double x_bytes = (double)((m->rows + R - 1) / R) * (double)P * W * elem_size(m->dtype);
        const double a_bytes = (double)m->nnz * (elem_size(m->dtype) + 2);
        /* ... or while everything sits in L2 anyway (NPB class W: 11 us against 21 us on SELL) */
        ok = forced || x_bytes <= a_bytes || x_bytes + a_bytes < 48e6;
    }
    if (!ok && want_fmt != 0) {
        /* wide matrices: tall row blocks, G rows per lane (spmv_panelg.cu) */
        PanelPlan pl;
        if (panel_plan_ring(m, &pl) && (forced || panel_plan_beats_sell(m, pl))) {
            P = pl.P; W = pl.W; R = pl.R; G = pl.G; nbuf = pl.nbuf; fmt = pl.fmt;
            ring_K = pl.ring_K; ring_S = pl.ring_S;
            ok = true;
        }
    }
    if (!ok) return false;
    const size_t es = elem_size(m->dtype);
    const int nblk = (m->rows + R - 1) / R;
    const int Tn = R / G;
    const int spb = Tn / 32;
    const int ntiles = nblk * P;
    const size_t nseg = (size_t)ntiles * R;
    const int nslices = ntiles * spb;
    int *d_overflow = nullptr, *d_cnt = nullptr;
    uint16_t *d_seglen = nullptr;
    CUDA_OK(cudaMalloc((void **)&d_seglen, nseg * sizeof(uint16_t)));
    CUDA_OK(cudaMemsetAsync(d_seglen, 0, nseg * sizeof(uint16_t), m->ctx->stream));
    CUDA_OK(cudaMalloc((void **)&d_overflow, sizeof(int)));
    CUDA_OK(cudaMemsetAsync(d_overflow, 0, sizeof(int), m->ctx->stream));
    launch_panel_count(m->d_rowptr, m->d_col, m->rows, P, W, R, d_seglen, d_overflow, m->ctx->stream);
    /* fmt 0: one ushort4 per lane and tile; fmt 2: G row ids per lane and tile */
    const size_t meta_bytes = fmt == 2 ? (size_t)ntiles * R * sizeof(uint16_t) : (size_t)ntiles * Tn * sizeof(ushort4);
    CUDA_OK(cudaMalloc((void **)&m->d_meta, meta_bytes));
    CUDA_OK(cudaMalloc((void **)&d_cnt, ((size_t)nslices + 1) * sizeof(int)));
    if (fmt == 2)
        launch_panelg_sort(d_seglen, ntiles, R, G, reinterpret_cast<uint16_t *>(m->d_meta), d_cnt, m->ctx->stream);
    else
        launch_panel_sort(d_seglen, ntiles, R, G, 0, m->d_meta, d_cnt, m->ctx->stream);
    CUDA_OK(cudaGetLastError());
    int overflow = 0;
    std::vector<int> cnt((size_t)nslices + 1);
    CUDA_OK(cudaMemcpyAsync(&overflow, d_overflow, sizeof(int), cudaMemcpyDeviceToHost, m->ctx->stream));
    CUDA_OK(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)nslices * sizeof(int), cudaMemcpyDeviceToHost, m->ctx->stream));
    CUDA_OK(cudaStreamSynchronize(m->ctx->stream));
    CUDA_OK(cudaFree(d_overflow));
    /* exclusive scan of the padded slice sizes (host; a few 10^4 entries) */
    long long run = 0;
    if (fmt == 2) {
        /* ring kernel: a warp's slices of all panels back to back -- (row block, warp, panel) */
        std::vector<int> off((size_t)nslices + 1);
        for (int rb = 0; rb < nblk; ++rb)
            for (int w = 0; w < spb; ++w)
                for (int pp = 0; pp < P; ++pp) {
                    off[((size_t)rb * spb + w) * P + pp] = (int)std::min<long long>(run, 0x7fffffffLL);
                    run += cnt[((size_t)rb * P + pp) * spb + w];
                }
        cnt.swap(off);
    } else {
        for (int i = 0; i < nslices; ++i) { const int c = cnt[i]; cnt[i] = (int)run; run += c; }
    }
    if (overflow || run > 0x7fffff00LL) {
        CUDA_OK(cudaFree(d_cnt));
        CUDA_OK(cudaFree(d_seglen));
        CUDA_OK(cudaFree(m->d_meta));
        m->d_meta = nullptr;
        return false;
    }
    cnt[nslices] = (int)run;
    m->d_slice_off = d_cnt;
    CUDA_OK(cudaMemcpyAsync(m->d_slice_off, cnt.data(), ((size_t)nslices + 1) * sizeof(int),
                            cudaMemcpyHostToDevice, m->ctx->stream));
    const size_t nval = (size_t)run + 64;
    if (fmt == 2) {
        /* one stream: per pair row 32 value pairs followed by 32 column pairs */
        CUDA_OK(cudaMalloc(&m->d_pval, nval * (es + 2)));
        CUDA_OK(cudaMemsetAsync((char *)m->d_pval + (size_t)run * (es + 2), 0, 64 * (es + 2), m->ctx->stream));
        m->d_pcol = nullptr;
    } else {
        CUDA_OK(cudaMalloc(&m->d_pval, nval * es));
        CUDA_OK(cudaMalloc((void **)&m->d_pcol, nval * sizeof(uint16_t)));
        CUDA_OK(cudaMemsetAsync((char *)m->d_pval + (size_t)run * es, 0, 64 * es, m->ctx->stream));
        CUDA_OK(cudaMemsetAsync(m->d_pcol + run, 0, 64 * sizeof(uint16_t), m->ctx->stream));
    }
    DevPanel &pm = m->panel;
    pm.val = m->d_pval; pm.col = m->d_pcol; pm.meta = m->d_meta; pm.slice_off = m->d_slice_off;
    pm.rowids = reinterpret_cast<const uint16_t *>(m->d_meta);
    pm.fmt = fmt; pm.ring_K = ring_K; pm.ring_S = ring_S;
    pm.rows = m->rows; pm.ncols = m->ncols; pm.R = R; pm.G = G; pm.P = P; pm.W = W; pm.nblk = nblk;
    pm.U = env_int("B200_SPMV_PANEL_U", 5);
    pm.use_tma = env_int("B200_SPMV_PANEL_TMA", 1);
    pm.nbuf = nbuf;
    pm.padded = run;
    if (fmt == 2) {
        if (m->dtype == B200_F64)
            launch_panelg_fill<double>((const double *)m->d_val, m->d_col, m->d_rowptr, m->rows, pm,
                                       d_seglen, (unsigned char *)m->d_pval, m->ctx->stream);
        else
            launch_panelg_fill<float>((const float *)m->d_val, m->d_col, m->d_rowptr, m->rows, pm,
                                      d_seglen, (unsigned char *)m->d_pval, m->ctx->stream);
    } else if (m->dtype == B200_F64)
        launch_panel_fill<double>((const double *)m->d_val, m->d_col, m->d_rowptr, m->rows, pm,
                                  d_seglen, (double *)m->d_pval, m->d_pcol, m->ctx->stream);
    else
        launch_panel_fill<float>((const float *)m->d_val, m->d_col, m->d_rowptr, m->rows, pm,
                                 d_seglen, (float *)m->d_pval, m->d_pcol, m->ctx->stream);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(m->ctx->stream));
    CUDA_OK(cudaFree(d_seglen));
    if ((fmt == 2 ? panelr_smem_bytes(pm, m->dtype == B200_F32) : panel_smem_bytes(pm, m->dtype == B200_F32)) > kSmemMax)
        die("panel shared-memory budget exceeded (W=%d R=%d)", W, R);
    /* the CSR copy of val / col is no longer needed */
    CUDA_OK(cudaFree(m->d_val)); m->d_val = nullptr;
    CUDA_OK(cudaFree(m->d_col)); m->d_col = nullptr;
    m->dev.val = nullptr; m->dev.col = nullptr;
    m->resident_bytes = (int64_t)(nval * es + nval * 2 + meta_bytes + ((size_t)nslices + 1) * 4 +
                                  ((size_t)m->rows + 1) * 4);
    return true;
}
/* SELL layout: lane streams with global columns, x gathered through L2.
 * Applicable to every matrix; rows above the cap go to the long-row kernel. */
static bool build_sell_locked(b200_matrix *m, const int *rowstr, bool all_rows_split)
{
    if (m->rows <= 0 || m->nnz <= 0) return false;
    const size_t es = elem_size(m->dtype);
    /* tile geometry measured on NPB class D row blocks, crsmat170 and the
     * power-law graph (profiles/r01_run11_sweep_sell.txt): 256-row tiles, two
     * rows per lane (longest with shortest), 2-pair chunks (72 registers, so
     * 7 CTAs of 128 threads per SM keep the L1TEX gather pipe full). */
    const double mean_len = (double)m->nnz / m->rows;
    /* short rows: uniform row slots (spmv_sellu.cu); long rows (NPB class D / E blocks): paired
     * rows with 128-bit loads (spmv_sell.cu).  Measured, kernel only (profiles/r02_run6_sweep.txt):
     * crsmat170 127 -> 120 us, power-law 2^22 370 -> 324 us, class D 1/8 block 383 us paired
     * against 412-558 us with slots. */
    const int fmt = env_int("B200_SPMV_SELL_FMT", mean_len <= 64.0 ? 1 : 0) == 0 ? 0 : 1;
    int G, R;
    if (fmt == 1) {
        /* 128 threads, G rows per lane; small tiles keep the grid many waves deep */
        G = env_int("B200_SPMV_SELL_G", 0);
        if (G != 2 && G != 4 && G != 8 && G != 16) G = mean_len <= 8 ? 4 : mean_len <= 24 ? 8 : 4;
        R = sellu_threads() * G;
    } else {
        G = env_int("B200_SPMV_SELL_G", 2);
        if (G != 1 && G != 2) G = 2;
        R = env_int("B200_SPMV_SELL_ROWS", 128 * G);
        const int gran = 32 * G;
        R = std::max(gran, std::min(256 * G, (R + gran - 1) / gran * gran));
    }
    const double mean = (double)m->nnz / m->rows;
    /* cap: rows above it leave the tiles for the nnz-split path.  A narrow
     * length distribution (NPB, crsmat: max <= 4 x mean) keeps every row in the
     * tiles; a heavy tail is cut at ~2.5 x mean, because one long row pads its
     * whole 32-lane slice (profiles/r01_run19_sweep_merge.txt). */
    int cap = env_int("B200_SPMV_SELL_CAP", 0);
    if (cap <= 0) {
        if ((double)m->scan.max_len <= 4.0 * std::max(mean, 8.0)) cap = m->scan.max_len;
        else cap = (int)std::max(32.0, 2.5 * mean);
    }
    cap = std::min(cap, 65534);
    if (all_rows_split) cap = 0;            /* MERGE: every row goes through the nnz-split path */
    const int nblk = (m->rows + R - 1) / R;
    const int Tn = R / G, spb = Tn / 32;
    const int nslices = nblk * spb;
    /* long rows (host pass over the caller's rowstr) */
    /* long rows -> nnz-split chunks (host pass over the caller's rowstr) */
    std::vector<int4> chunks;
    std::vector<int2> multi;
    std::vector<int> multi_rows;
    int n_long = 0, n_carry = 0;
    if (m->scan.max_len > cap) {
        const int CH = sell_chunk_entries();
        const int base = rowstr[0];
        for (int r = 0; r < m->rows; ++r) {
            const int len = rowstr[r + 1] - rowstr[r];
            if (len <= cap) continue;
            ++n_long;
            const int lo = rowstr[r] - base, hi = rowstr[r + 1] - base;
            const int nch = (len + CH - 1) / CH;
            if (nch == 1) {
                chunks.push_back(make_int4(r, lo, hi, -1));
            } else {
                multi.push_back(make_int2(n_carry, nch));
                multi_rows.push_back(r);
                for (int k = 0; k < nch; ++k)
                    chunks.push_back(make_int4(r, lo + k * CH, std::min(hi, lo + (k + 1) * CH), n_carry++));
            }
        }
    }
    uint16_t *d_seglen = nullptr;
    int *d_cnt = nullptr;
    const size_t nseg = (size_t)nblk * R;
    CUDA_OK(cudaMalloc((void **)&d_seglen, nseg * sizeof(uint16_t)));
    CUDA_OK(cudaMemsetAsync(d_seglen, 0, nseg * sizeof(uint16_t), m->ctx->stream));
    launch_sell_rowlen(m->d_rowptr, m->rows, R, cap, d_seglen, m->ctx->stream);
    size_t meta_bytes;
    uint16_t *d_slotlen = nullptr;
    if (fmt == 1) {
        /* rowids [nblk][G][T] followed by slotlen [nblk][4][G] */
        const size_t n_ids = (size_t)nblk * R, n_sl = (size_t)nblk * spb * G;
        meta_bytes = (n_ids + n_sl) * sizeof(uint16_t);
        CUDA_OK(cudaMalloc((void **)&m->d_meta, meta_bytes));
        d_slotlen = reinterpret_cast<uint16_t *>(m->d_meta) + n_ids;
        CUDA_OK(cudaMalloc((void **)&d_cnt, ((size_t)nslices + 1) * sizeof(int)));
        launch_sellu_sort(d_seglen, nblk, R, G, reinterpret_cast<uint16_t *>(m->d_meta), d_slotlen, d_cnt,
                          m->ctx->stream);
    } else {
        meta_bytes = (size_t)nblk * Tn * sizeof(ushort4);
        CUDA_OK(cudaMalloc((void **)&m->d_meta, meta_bytes));
        CUDA_OK(cudaMalloc((void **)&d_cnt, ((size_t)nslices + 1) * sizeof(int)));
        launch_panel_sort(d_seglen, nblk, R, G, 1, m->d_meta, d_cnt, m->ctx->stream);
    }
    CUDA_OK(cudaGetLastError());
    std::vector<int> cnt((size_t)nslices + 1);
    CUDA_OK(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)nslices * sizeof(int), cudaMemcpyDeviceToHost, m->ctx->stream));
    CUDA_OK(cudaStreamSynchronize(m->ctx->stream));
    long long run = 0;
    for (int i = 0; i < nslices; ++i) { const int c = cnt[i]; cnt[i] = (int)run; run += c; }
    if (run > 0x7fffff00LL) {
        CUDA_OK(cudaFree(d_cnt)); CUDA_OK(cudaFree(d_seglen)); CUDA_OK(cudaFree(m->d_meta));
        m->d_meta = nullptr;
        return false;
    }
    cnt[nslices] = (int)run;
    m->d_slice_off = d_cnt;
    CUDA_OK(cudaMemcpyAsync(m->d_slice_off, cnt.data(), ((size_t)nslices + 1) * sizeof(int),
                            cudaMemcpyHostToDevice, m->ctx->stream));
    const size_t nval = (size_t)run + 64;
    CUDA_OK(cudaMalloc(&m->d_pval, nval * es));
    CUDA_OK(cudaMalloc((void **)&m->d_scol, nval * sizeof(int)));
    DevSell &sm = m->sell;
    sm.val = m->d_pval; sm.col = m->d_scol; sm.meta = m->d_meta; sm.slice_off = m->d_slice_off;
    sm.rows = m->rows; sm.R = R; sm.G = G; sm.nblk = nblk; sm.padded = run;
    sm.fmt = fmt;
    sm.rowids = reinterpret_cast<const uint16_t *>(m->d_meta);
    sm.slotlen = d_slotlen;
    sm.U = env_int("B200_SPMV_SELL_U", fmt == 1 ? 4 : 2);
    sm.n_long = n_long;
    /* short single-chunk rows first: they run with 8 lanes per row */
    const int short_len = sell_short_chunk_entries();
    auto mid = std::stable_partition(chunks.begin(), chunks.end(), [&](const int4 &c) {
        return c.w < 0 && c.z - c.y <= short_len;
    });
    sm.n_chunks_short = (int)(mid - chunks.begin());
    sm.n_chunks = (int)chunks.size();
    sm.n_multi = (int)multi.size();
    sm.chunks = nullptr; sm.multi = nullptr; sm.multi_rows = nullptr; sm.carry = nullptr;
    if (sm.n_chunks > 0) {
        CUDA_OK(cudaMalloc((void **)&m->d_chunks, chunks.size() * sizeof(int4)));
        CUDA_OK(cudaMemcpy(m->d_chunks, chunks.data(), chunks.size() * sizeof(int4), cudaMemcpyHostToDevice));
        sm.chunks = m->d_chunks;
    }
    if (sm.n_multi > 0) {
        CUDA_OK(cudaMalloc((void **)&m->d_multi, multi.size() * sizeof(int2)));
        CUDA_OK(cudaMemcpy(m->d_multi, multi.data(), multi.size() * sizeof(int2), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc((void **)&m->d_multi_rows, multi_rows.size() * sizeof(int)));
        CUDA_OK(cudaMemcpy(m->d_multi_rows, multi_rows.data(), multi_rows.size() * sizeof(int),
                           cudaMemcpyHostToDevice));
        CUDA_OK(cudaMalloc(&m->d_carry, (size_t)std::max(n_carry, 1) * es));
        sm.multi = m->d_multi; sm.multi_rows = m->d_multi_rows; sm.carry = m->d_carry;
    }
    if (fmt == 1) {
        if (m->dtype == B200_F64)
            launch_sellu_fill<double>((const double *)m->d_val, m->d_col, m->d_rowptr, m->rows, sm, cap,
                                      (double *)m->d_pval, m->d_scol, m->ctx->stream);
        else
            launch_sellu_fill<float>((const float *)m->d_val, m->d_col, m->d_rowptr, m->rows, sm, cap,
                                     (float *)m->d_pval, m->d_scol, m->ctx->stream);
    } else if (m->dtype == B200_F64)
        launch_sell_fill<double>((const double *)m->d_val, m->d_col, m->d_rowptr, m->rows, sm,
                                 d_seglen, (double *)m->d_pval, m->d_scol, m->ctx->stream);
    else
        launch_sell_fill<float>((const float *)m->d_val, m->d_col, m->d_rowptr, m->rows, sm,
                                d_seglen, (float *)m->d_pval, m->d_scol, m->ctx->stream);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(m->ctx->stream));
    CUDA_OK(cudaFree(d_seglen));
    m->resident_bytes = (int64_t)(nval * (es + 4) + meta_bytes + ((size_t)nslices + 1) * 4 +
                                  ((size_t)m->rows + 1) * 4);
    if (sm.n_chunks == 0) {
        CUDA_OK(cudaFree(m->d_val)); m->d_val = nullptr;
        CUDA_OK(cudaFree(m->d_col)); m->d_col = nullptr;
        m->dev.val = nullptr; m->dev.col = nullptr;
    } else {
        m->resident_bytes += (int64_t)(((size_t)m->nnz + kPadElems) * (es + 4));
    }
    return true;
}

/* SMALL: x (ncols elements) and a tile of products must share one SM's shared memory, and
 * the nnz-balanced row blocks must form ONE wave (one CTA per SM).  The tile grows from
 * nnz / SMs until the greedy blocks fit the SM count; a row longer than the tile rules the
 * family out (the panel kernels take such matrices). */
static bool build_small_locked(b200_matrix *m, const int *rowstr)
{
    if (m->rows <= 0 || m->nnz <= 0 || m->ncols <= 0) return false;
    const size_t es = elem_size(m->dtype);
    const int xpad = (m->ncols + 3) & ~3;
    const size_t x_bytes = (size_t)xpad * es;
    if (x_bytes + 4096 > kSmemMax) return false;
    const long long cap = (long long)((kSmemMax - x_bytes - 512) / es);   /* header, pair slack */
    const int sms = m->ctx->sm_count;
    long long tile = std::max<long long>((m->nnz + sms - 1) / sms, m->scan.max_len);
    tile = std::max<long long>(tile, 64);
    std::vector<int> blk;
    for (;;) {
        if (tile > cap) return false;
        blk.clear();
        blk.push_back(0);
        int r = 0;
        bool ok = true;
        while (r < m->rows) {
            long long used = 0;
            const int start = r;
            while (r < m->rows) {
                const long long len = (long long)rowstr[r + 1] - rowstr[r];
                if (used + len > tile) break;
                used += len;
                ++r;
            }
            if (r == start) { ok = false; break; }         /* a row longer than the tile */
            blk.push_back(r);
        }
        if (ok && (int)blk.size() - 1 <= sms) break;
        if (tile == cap) return false;
        tile = std::min<long long>(cap, tile + std::max<long long>(tile / 64, 8));
    }
    CUDA_OK(cudaMalloc((void **)&m->d_small_blk, blk.size() * sizeof(int)));
    CUDA_OK(cudaMemcpy(m->d_small_blk, blk.data(), blk.size() * sizeof(int), cudaMemcpyHostToDevice));
    m->small_.rowblk = m->d_small_blk;
    m->small_.nblk = (int)blk.size() - 1;
    m->small_.tile = (int)tile;
    m->small_.xpad = xpad;
    m->small_.ncols = m->ncols;
    m->small_.use_tma = env_int("B200_SPMV_SMALL_TMA", 1);
    m->small_.pdl = env_int("B200_SPMV_PDL", 1);
    m->small_.cfg = env_int("B200_SPMV_SMALL_CFG", 0);
    m->small_.col16 = nullptr;
    if (env_int("B200_SPMV_SMALL_COL16", 1)) {
        /* fewer than 65 536 columns (x fits shared memory): 10 instead of 12 bytes per entry */
        const size_t n16 = (size_t)m->nnz + kPadElems;
        CUDA_OK(cudaMalloc((void **)&m->d_small_col16, n16 * sizeof(uint16_t)));
        launch_small_col16(m->d_col, m->d_small_col16, n16, m->ctx->stream);
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaStreamSynchronize(m->ctx->stream));
        m->small_.col16 = m->d_small_col16;
        m->resident_bytes += (int64_t)(n16 * sizeof(uint16_t));
    }
    m->resident_bytes += (int64_t)(blk.size() * sizeof(int));
    return true;
}

namespace b200 {

/* `on_device`: a / rowstr / colidx are DEVICE arrays of m's device (same 1-based
 * contents); the row pointers are read back for the host passes below. */
static b200_matrix *upload_any_locked(DevCtx *ctx, const void *a, const int *rowstr_in,
                                      const int *colidx, int rows, int dtype, int kernel,
                                      bool on_device)
{
    if (rows < 0) die("negative row count %d", rows);
    if (dtype != B200_F64 && dtype != B200_F32) die("unknown dtype %d", dtype);
    DeviceScope scope(ctx->device);
    const size_t es = elem_size(dtype);
    std::vector<int> rowstr_copy;
    const int *rowstr = rowstr_in;
    if (on_device && rows > 0) {
        rowstr_copy.resize((size_t)rows + 1);
        CUDA_OK(cudaMemcpy(rowstr_copy.data(), rowstr_in, ((size_t)rows + 1) * sizeof(int),
                           cudaMemcpyDeviceToHost));
        rowstr = rowstr_copy.data();
    }
    const int base1 = rows > 0 ? rowstr[0] : 1;               /* 1-based offset */
    /* native-impl.c:4-10 reads a[rowstr[i]-1 ...]: an offset below 1 would read in front of
     * the arrays; offsets must not decrease (the row-block builders rely on it) */
    if (base1 < 1) die("rowstr[0] = %d; row offsets are 1-based", base1);
    for (int r = 0; r < rows; ++r)
        if (rowstr[r + 1] < rowstr[r])
            die("rowstr is not non-decreasing at row %d (%d > %d)", r, rowstr[r], rowstr[r + 1]);
    const int64_t nnz = rows > 0 ? (int64_t)rowstr[rows] - base1 : 0;

    b200_matrix *m = (b200_matrix *)calloc(1, sizeof *m);
    m->dtype = dtype;
    m->rows = rows;
    m->nnz = nnz;
    m->device = ctx->device;
    m->ctx = ctx;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;

    const size_t nval = (size_t)nnz + kPadElems;
    CUDA_OK(cudaMalloc(&m->d_val, nval * es));
    CUDA_OK(cudaMalloc((void **)&m->d_col, nval * sizeof(int)));
    CUDA_OK(cudaMalloc((void **)&m->d_rowptr, ((size_t)rows + 1) * sizeof(int)));
    m->resident_bytes = (int64_t)(nval * es + nval * 4 + ((size_t)rows + 1) * 4);

    /* padding: value 0, column 1 (a valid 1-based index) */
    CUDA_OK(cudaMemsetAsync((char *)m->d_val + (size_t)nnz * es, 0, kPadElems * es, m->ctx->stream));
    {
        int ones[kPadElems];
        for (int i = 0; i < kPadElems; ++i) ones[i] = 1;
        CUDA_OK(cudaMemcpyAsync(m->d_col + nnz, ones, sizeof ones, cudaMemcpyHostToDevice, m->ctx->stream));
        CUDA_OK(cudaStreamSynchronize(m->ctx->stream));
    }
    if (nnz > 0) {
        CUDA_OK(cudaMemcpyAsync(m->d_val, (const char *)a + (size_t)(base1 - 1) * es,
                                (size_t)nnz * es, kind, m->ctx->stream));
        CUDA_OK(cudaMemcpyAsync(m->d_col, colidx + (base1 - 1), (size_t)nnz * sizeof(int),
                                kind, m->ctx->stream));
    }
    if (rows > 0) {
        CUDA_OK(cudaMemcpyAsync(m->d_rowptr, rowstr_in, ((size_t)rows + 1) * sizeof(int),
                                kind, m->ctx->stream));
        launch_rebase_rowptr(m->d_rowptr, rows + 1, base1, m->ctx->stream);
    } else {
        CUDA_OK(cudaMemsetAsync(m->d_rowptr, 0, sizeof(int), m->ctx->stream));
    }

    /* device-side scan: column count, histogram, sortedness */
    UploadScan *d_scan = nullptr;
    CUDA_OK(cudaMalloc((void **)&d_scan, sizeof(UploadScan)));
    launch_upload_scan(m->d_rowptr, m->d_col, rows, (int)nnz, d_scan, m->ctx->stream);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(&m->scan, d_scan, sizeof(UploadScan), cudaMemcpyDeviceToHost, m->ctx->stream));
    CUDA_OK(cudaStreamSynchronize(m->ctx->stream));
    CUDA_OK(cudaFree(d_scan));
    if (nnz == 0) { m->scan.max_col = 0; m->scan.min_col = 1; }
    if (rows == 0) { m->scan.max_len = 0; m->scan.min_len = 0; }
    if (m->scan.min_col < 1) die("colidx holds %d; indices are 1-based", m->scan.min_col);
    m->ncols = m->scan.max_col;

    /* row blocks of the nnz-split kernel (host greedy over the caller's rowstr) */
    std::vector<int> blk;
    build_row_blocks(rowstr, rows, tile_elems(dtype == B200_F32), blk);
    const int nblk = (int)blk.size() - 1;
    CUDA_OK(cudaMalloc((void **)&m->d_rowblk, blk.size() * sizeof(int)));
    CUDA_OK(cudaMemcpy(m->d_rowblk, blk.data(), blk.size() * sizeof(int), cudaMemcpyHostToDevice));
    m->resident_bytes += (int64_t)(blk.size() * sizeof(int));

    m->dev.val = m->d_val;
    m->dev.col = m->d_col;
    m->dev.rowptr = m->d_rowptr;
    m->dev.rowblk = m->d_rowblk;
    m->dev.rows = rows;
    m->dev.nblk = rows > 0 ? nblk : 0;
    m->dev.nnz = (int)nnz;

    /* kernel choice from the histogram */
    kernel = kernel_from_env(kernel);
    /* the launch-bound regime first: x and one SM's share of the products in shared memory */
    if (kernel == B200_KERNEL_SMALL || (kernel == B200_KERNEL_AUTO && env_int("B200_SPMV_SMALL", 1))) {
        if (build_small_locked(m, rowstr)) kernel = B200_KERNEL_SMALL;
        else if (kernel == B200_KERNEL_SMALL) kernel = B200_KERNEL_AUTO;
    }
    if (kernel == B200_KERNEL_AUTO || kernel == B200_KERNEL_PANEL) {
        if (build_panel_locked(m, kernel == B200_KERNEL_PANEL)) kernel = B200_KERNEL_PANEL;
        else kernel = B200_KERNEL_SELL;
    }
    if ((kernel == B200_KERNEL_SELL || kernel == B200_KERNEL_MERGE) &&
        !build_sell_locked(m, rowstr, kernel == B200_KERNEL_MERGE))
        kernel = B200_KERNEL_ORDERED;
    m->kernel = kernel;
    {
        const double mean = rows > 0 ? (double)nnz / rows : 0.0;
        int lanes = 2;
        while (lanes < 32 && lanes < mean / 2) lanes *= 2;
        m->lanes = env_int("B200_SPMV_LANES", lanes);
    }
    if (g_verbose)
        fprintf(stderr,
                "libb200-spmv: dev %d: uploaded %s matrix rows=%d cols=%d nnz=%lld len[min=%d max=%d] "
                "unsorted_rows=%d blocks=%d kernel=%s panel[fmt=%d R=%d G=%d P=%d W=%d nbuf=%d ring=%dx%d padded=%lld] sell[R=%d G=%d padded=%lld long=%d/%d]\n",
                m->device, dtype == B200_F32 ? "f32" : "f64", rows, m->ncols, (long long)nnz,
                m->scan.min_len, m->scan.max_len, m->scan.rows_unsorted, nblk,
                b200_spmv_kernel_name(m), m->panel.fmt, m->panel.R, m->panel.G, m->panel.P, m->panel.W, m->panel.nbuf, m->panel.ring_S, m->panel.ring_K, m->panel.padded,
                m->sell.R, m->sell.G, m->sell.padded, m->sell.n_long, m->sell.n_chunks);
    return m;
}

b200_matrix *upload_locked(DevCtx *ctx, const void *a, const int *rowstr, const int *colidx,
                           int rows, int dtype, int kernel)
{
    return upload_any_locked(ctx, a, rowstr, colidx, rows, dtype, kernel, false);
}

void release_locked(b200_matrix *m)
{
    if (!m) return;
    DeviceScope scope(m->device);
    cudaFree(m->d_val); cudaFree(m->d_col); cudaFree(m->d_rowptr); cudaFree(m->d_rowblk);
    cudaFree(m->d_pval); cudaFree(m->d_pcol); cudaFree(m->d_meta); cudaFree(m->d_slice_off);
    cudaFree(m->d_scol); cudaFree(m->d_chunks); cudaFree(m->d_multi); cudaFree(m->d_multi_rows);
    cudaFree(m->d_carry);
    cudaFree(m->d_small_blk); cudaFree(m->d_small_col16);
    free(m);
}

bool exec_waits_in_kernel(const b200_matrix *m)
{
    return m->kernel == B200_KERNEL_PANEL && m->panel.fmt == 2;
}

bool exec_takes_flags(const b200_matrix *m)
{
    return m->kernel == B200_KERNEL_PANEL && (m->panel.fmt == 2 || m->panel.fmt == 0);
}

/* ... with a watchdog on the wait (the drop-in path's overlapped upload): the paired kernel's
 * flagged instance only; the ring kernel's spin is left exactly as the multi-GPU runs measured it */
bool exec_takes_guarded_flags(const b200_matrix *m)
{
    return m->kernel == B200_KERNEL_PANEL && m->panel.fmt == 0;
}

int exec_locked(b200_matrix *m, const void *d_x, void *d_y, cudaStream_t s, const SliceFlags *sf,
                const XPush *xp)
{
    if (m->rows == 0) return 0;
    int launched_kernels = 1;
    if (m->kernel == B200_KERNEL_PANEL) {
        if (m->panel.fmt == 2) {
            XFlags xf = {nullptr, 0ull, 1, 0, nullptr, 0ull, 0};
            if (sf) { xf.flags = sf->flags; xf.epoch = sf->epoch; xf.cols_per_rank = sf->cols_per_rank; xf.nranks = sf->nranks;
                      xf.timed_out = sf->timed_out; xf.timeout_ns = sf->timeout_ns; xf.ready0 = sf->ready0; }
            XPush none;
            none.src = nullptr;
            const XPush &push = xp ? *xp : none;
            if (m->dtype == B200_F64)
                launch_panelr<double>(m->panel, (const double *)d_x, (double *)d_y, xf, push, s);
            else
                launch_panelr<float>(m->panel, (const float *)d_x, (float *)d_y, xf, push, s);
        } else {
            XFlags xf = {nullptr, 0ull, 1, 0, nullptr, 0ull, 0};
            if (sf) { xf.flags = sf->flags; xf.epoch = sf->epoch; xf.cols_per_rank = sf->cols_per_rank; xf.nranks = sf->nranks;
                      xf.timed_out = sf->timed_out; xf.timeout_ns = sf->timeout_ns; xf.ready0 = sf->ready0; }
            if (m->dtype == B200_F64)
                launch_panel<double>(m->panel, (const double *)d_x, (double *)d_y, s, nullptr, nullptr,
                                     xf.flags ? &xf : nullptr);
            else
                launch_panel<float>(m->panel, (const float *)d_x, (float *)d_y, s, nullptr, nullptr,
                                    xf.flags ? &xf : nullptr);
        }
    } else if (m->kernel == B200_KERNEL_SMALL) {
        if (m->dtype == B200_F64)
            launch_small<double>(m->small_, m->dev, (const double *)d_x, (double *)d_y, s);
        else
            launch_small<float>(m->small_, m->dev, (const float *)d_x, (float *)d_y, s);
    } else if (m->kernel == B200_KERNEL_SELL || m->kernel == B200_KERNEL_MERGE) {
        if (m->dtype == B200_F64)
            launch_sell<double>(m->sell, m->dev, (const double *)d_x, (double *)d_y, s);
        else
            launch_sell<float>(m->sell, m->dev, (const float *)d_x, (float *)d_y, s);
        launched_kernels = 1 + (m->sell.n_chunks_short > 0) +
                           (m->sell.n_chunks > m->sell.n_chunks_short) + (m->sell.n_multi > 0);
    } else if (m->dtype == B200_F64) {
        if (m->kernel == B200_KERNEL_VECTOR)
            launch_vector<double>(m->dev, m->lanes, (const double *)d_x, (double *)d_y, s);
        else
            launch_ordered<double>(m->dev, (const double *)d_x, (double *)d_y, s);
    } else {
        if (m->kernel == B200_KERNEL_VECTOR)
            launch_vector<float>(m->dev, m->lanes, (const float *)d_x, (float *)d_y, s);
        else
            launch_ordered<float>(m->dev, (const float *)d_x, (float *)d_y, s);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) die("kernel launch failed: %s", cudaGetErrorString(e));
    return launched_kernels;
}

}  // namespace b200

/* ------------------------------------------------------------------------
 * public resident-matrix API
 * ---------------------------------------------------------------------- */
extern "C" int b200_spmv_init(int device)
{
    pthread_mutex_lock(&g_lock);
    ensure_init_locked(device);
    const int d = default_device_locked();
    ctx_for_device_locked(d);
    /* as before: the chosen device is also made current for the caller */
    CUDA_OK(cudaSetDevice(d));
    pthread_mutex_unlock(&g_lock);
    return d;
}

static DevCtx *current_ctx_locked(void)
{
    ensure_init_locked(-1);
    int d = 0;
    CUDA_OK(cudaGetDevice(&d));
    return ctx_for_device_locked(d);
}

extern "C" b200_matrix *b200_spmv_upload(const void *a, const int *rowstr, const int *colidx,
                                         int rows, int dtype, int kernel)
{
    pthread_mutex_lock(&g_lock);
    b200_matrix *m = upload_any_locked(current_ctx_locked(), a, rowstr, colidx, rows, dtype, kernel, false);
    pthread_mutex_unlock(&g_lock);
    return m;
}

extern "C" b200_matrix *b200_spmv_upload_device(const void *d_a, const int *d_rowstr, const int *d_colidx,
                                                int rows, int dtype, int kernel)
{
    pthread_mutex_lock(&g_lock);
    b200_matrix *m = upload_any_locked(current_ctx_locked(), d_a, d_rowstr, d_colidx, rows, dtype, kernel, true);
    pthread_mutex_unlock(&g_lock);
    return m;
}

extern "C" void b200_spmv_release(b200_matrix *m)
{
    pthread_mutex_lock(&g_lock);
    release_locked(m);
    pthread_mutex_unlock(&g_lock);
}

extern "C" int b200_spmv_exec(b200_matrix *m, const void *d_x, void *d_y, void *stream)
{
    if (!m) die("b200_spmv_exec: null matrix");
    DeviceScope scope(m->device);
    return exec_locked(m, d_x, d_y, (cudaStream_t)stream, nullptr);
}

extern "C" int b200_spmv_exec_sliced(b200_matrix *m, const void *d_x, void *d_y, void *stream,
                                     const unsigned long long *flags, unsigned long long epoch,
                                     int cols_per_rank, int nranks)
{
    if (!m) die("b200_spmv_exec_sliced: null matrix");
    if (!exec_takes_flags(m)) return -1;
    DeviceScope scope(m->device);
    SliceFlags sf = {flags, epoch, cols_per_rank, nranks, nullptr, 0ull, 0};
    return exec_locked(m, d_x, d_y, (cudaStream_t)stream, &sf);
}

/* defined in cg_peer.cu: the pointers and flags of a peer group for one push */
struct b200_peer_group;
extern "C" int b200_peer_describe_push(b200_peer_group *g, const double *v, int n_local, int64_t lo,
                                       uint64_t e, b200::XPush *out, const double **xbuf,
                                       const unsigned long long **vflags, int *cols_per_rank);

extern "C" int b200_spmv_exec_pushed(b200_matrix *m, void *d_y, void *stream, void *peer_group,
                                     const double *v_local, int n_local, int64_t lo, uint64_t epoch,
                                     int cols_per_rank)
{
    if (!m || !peer_group) die("b200_spmv_exec_pushed: null argument");
    /* the publishing CTA waits for all others: they must be co-resident (one per SM) */
    if (!exec_waits_in_kernel(m) || m->dtype != B200_F64 || m->panel.nblk > m->ctx->sm_count) return -1;
    XPush xp;
    const double *xbuf = nullptr;
    const unsigned long long *vflags = nullptr;
    int cpr = 0;
    if (b200_peer_describe_push((b200_peer_group *)peer_group, v_local, n_local, lo, epoch, &xp, &xbuf,
                                &vflags, &cpr) != 0)
        return -1;
    DeviceScope scope(m->device);
    SliceFlags sf = {vflags, epoch, cols_per_rank, xp.nranks, nullptr, 0ull, 0};
    return exec_locked(m, xbuf, d_y, (cudaStream_t)stream, &sf, &xp);
}

/* y = A x and, in the same launch, partial[b] = (share of CTA b of) dotv . y for the first
 * b200_spmv_dot_partials(m) entries of `partial`.  The paired PANEL kernel and the SMALL
 * kernel have the fused epilogue; returns -1 without launching otherwise. */
extern "C" int b200_spmv_dot_partials(const b200_matrix *m)
{
    if (m->dtype != B200_F64) return 0;
    if (m->kernel == B200_KERNEL_SMALL) return m->small_.nblk;
    return (m->kernel == B200_KERNEL_PANEL && m->panel.fmt == 0) ? m->panel.nblk : 0;
}

extern "C" int b200_spmv_exec_dot(b200_matrix *m, const void *d_x, void *d_y, const void *d_dotv,
                                  void *d_partial, void *stream)
{
    if (!m) die("b200_spmv_exec_dot: null matrix");
    if (b200_spmv_dot_partials(m) <= 0) return -1;
    DeviceScope scope(m->device);
    if (m->kernel == B200_KERNEL_SMALL)
        launch_small<double>(m->small_, m->dev, (const double *)d_x, (double *)d_y, (cudaStream_t)stream,
                             (const double *)d_dotv, (double *)d_partial);
    else
        launch_panel<double>(m->panel, (const double *)d_x, (double *)d_y, (cudaStream_t)stream,
                             (const double *)d_dotv, (double *)d_partial);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) die("kernel launch failed: %s", cudaGetErrorString(e));
    return 1;
}

extern "C" int b200_spmv_can_push(const b200_matrix *m)
{
    return exec_waits_in_kernel(m) && m->dtype == B200_F64 && m->panel.nblk <= m->ctx->sm_count ? 1 : 0;
}

extern "C" int b200_spmv_device(const b200_matrix *m) { return m->device; }
extern "C" int b200_spmv_waits_in_kernel(const b200_matrix *m) { return exec_waits_in_kernel(m) ? 1 : 0; }
extern "C" int b200_spmv_rows(const b200_matrix *m) { return m->rows; }
extern "C" int b200_spmv_ncols(const b200_matrix *m) { return m->ncols; }
extern "C" int64_t b200_spmv_nnz(const b200_matrix *m) { return m->nnz; }
extern "C" int b200_spmv_kernel(const b200_matrix *m) { return m->kernel; }
extern "C" const char *b200_spmv_kernel_name(const b200_matrix *m)
{
    switch (m->kernel) {
    case B200_KERNEL_ORDERED: return "ordered";
    case B200_KERNEL_VECTOR:  return "vector";
    case B200_KERNEL_PANEL:   return "panel";
    case B200_KERNEL_MERGE:   return "merge";
    case B200_KERNEL_SELL:    return "sell";
    case B200_KERNEL_SMALL:   return "small";
    default: return "auto";
    }
}
extern "C" int b200_spmv_launches_per_exec(const b200_matrix *m)
{
    if (m->rows <= 0) return 0;
    return (m->kernel == B200_KERNEL_SELL || m->kernel == B200_KERNEL_MERGE)
               ? 1 + (m->sell.n_chunks_short > 0) + (m->sell.n_chunks > m->sell.n_chunks_short) +
                     (m->sell.n_multi > 0)
               : 1;
}
extern "C" int64_t b200_spmv_algorithmic_bytes(const b200_matrix *m)
{
    const int64_t es = (int64_t)elem_size(m->dtype);
    return (es + 4) * m->nnz + 4 * ((int64_t)m->rows + 1) + es * m->ncols + es * m->rows;
}
extern "C" int64_t b200_spmv_resident_bytes(const b200_matrix *m) { return m->resident_bytes; }

extern "C" void b200_spmv_row_histogram(const b200_matrix *m, int64_t bins[32],
                                        int *min_len, int *max_len)
{
    for (int i = 0; i < 32; ++i) bins[i] = (int64_t)m->scan.hist[i];
    if (min_len) *min_len = m->scan.min_len;
    if (max_len) *max_len = m->scan.max_len;
}

extern "C" void b200_spmv_partition_rows(const int *rowstr, int rows, int parts, int *bounds)
{
    const int64_t base = rows > 0 ? rowstr[0] : 0;
    const int64_t nnz = rows > 0 ? (int64_t)rowstr[rows] - base : 0;
    bounds[0] = 0;
    for (int p = 1; p < parts; ++p) {
        const int64_t target = base + (nnz * p) / parts;
        /* first row whose start offset is >= target */
        const int *it = std::lower_bound(rowstr, rowstr + rows + 1, (int)target);
        int r = (int)(it - rowstr);
        if (r > rows) r = rows;
        if (r < bounds[p - 1]) r = bounds[p - 1];
        bounds[p] = r;
    }
    bounds[parts] = rows;
}
extern "C" const char *b200_spmv_version(void) { return B200_VERSION; }
