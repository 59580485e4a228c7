/*
 * spmv_kernels.cuh -- device-side view of a resident CSR row block and the
 * launchers of the sm_100a kernel families.
 * Internal to libb200-spmv; the public surface is include/b200_spmv.h.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

/* tile geometry of the nnz-split (ORDERED) kernel */
constexpr int kStreamThreads = 256;     /* threads per CTA */
constexpr int kTileF64       = 4096;    /* products staged per CTA, fp64 (32 KB) */
constexpr int kTileF32       = 8192;    /* fp32 (32 KB) */
constexpr int kRowsPerBlock  = 1024;    /* row cap of one row block */
constexpr int kPadElems      = 16;      /* zero padding behind val / col */

struct DevCsr {
    const void *val;      /* T[nnz + kPadElems], padding = 0 */
    const int  *col;      /* int[nnz + kPadElems], 1-based, padding = 1 */
    const int  *rowptr;   /* int[rows + 1], 0-based offsets into val / col */
    const int  *rowblk;   /* int[nblk + 1], first row of every row block */
    int rows;
    int nblk;
    int nnz;
};

/* statistics produced on the device at upload */
struct UploadScan {
    int max_col;          /* max(colidx) = number of columns (gpu.c:216-223 intent) */
    int min_col;          /* must be >= 1 */
    int max_len;
    int min_len;
    int rows_unsorted;    /* rows whose columns are not non-decreasing */
    int pad;
    unsigned long long hist[32];
};

/* y = A x, every row summed left to right (bit-identical to the CPU loop) */
template <typename T>
void launch_ordered(const DevCsr &m, const T *x, T *y, cudaStream_t s);

/* y = A x, `lanes` (2..32, power of two) threads per row, shuffle reduction */
template <typename T>
void launch_vector(const DevCsr &m, int lanes, const T *x, T *y, cudaStream_t s);

/* upload-time passes */
void launch_rebase_rowptr(int *rowptr, int rows_plus_1, int base, cudaStream_t s);
void launch_upload_scan(const int *rowptr, const int *col, int rows, int nnz,
                        UploadScan *out, cudaStream_t s);

int tile_elems(bool f32);

/* SMALL: whole x in shared memory, one nnz-balanced row block per SM (spmv_small.cu) */
struct DevSmall {
    const int *rowblk;    /* int[nblk + 1], first row of every row block */
    int nblk;
    int tile;             /* products a CTA can hold (entries) */
    int xpad;             /* elements reserved for x in shared memory */
    int ncols;
    int use_tma;          /* x by TMA bulk copy when it is 16-byte aligned (B200_SPMV_SMALL_TMA) */
    int pdl;              /* programmatic dependent launch (B200_SPMV_PDL) */
    int cfg;              /* 0 = 1024 threads x 3 pairs per batch, 1 = 512 x 6 */
    const uint16_t *col16;/* 0-based 16-bit copy of the columns (nullptr: the int32 columns as uploaded) */
};
void launch_small_col16(const int *col, uint16_t *col16, size_t n, cudaStream_t s);
/* dotv != nullptr: dot_partial[b] = row block b's share of dotv . y, b < sm.nblk (fixed order) */
template <typename T>
void launch_small(const DevSmall &sm, const DevCsr &m, const T *x, T *y, cudaStream_t s,
                  const T *dotv = nullptr, T *dot_partial = nullptr);

/* dst[0..bytes) = src[0..bytes): src is a device alias of pinned host memory,
 * so the loads travel over PCIe; runs on the SMs, in stream order */
void launch_copy_in(const void *src, void *dst, size_t bytes, cudaStream_t s);
/* the same with up to 8 destinations (the local buffer and the peers' over NVLink):
 * the source is read once.  All pointers 16-byte aligned relative to each other. */
void launch_copy_in_multi(const void *src, void *const *dst, int ndst, size_t bytes, cudaStream_t s);
/* ... and, once every block's stores are out, flag[j][0] = epoch for every destination j
 * (st.release.sys; `counter` is a zeroed device word used to find the last block): consumers
 * on the destination devices wait on their flag instead of on a CUDA event */
void launch_copy_in_multi_flagged(const void *src, void *const *dst, int ndst, size_t bytes,
                                  unsigned long long *const *flag, unsigned long long epoch,
                                  unsigned int *counter, cudaStream_t s);
/* probes of a host range registered by the library on the caller's behalf (b200_dropin.cu):
 * bad <- 1 if any of four `es`-byte words read through the device alias differs from what the
 * host saw; out[k] <- the word at off[k] as read through the alias (words at or beyond `limit`
 * are skipped) */
void launch_probe_x(const void *alias, const size_t off[4], const unsigned long long val[4], int es,
                    int *bad, cudaStream_t s);
void launch_probe_y(const void *alias, const size_t off[4], int es, size_t limit, unsigned long long *out,
                    cudaStream_t s);
/* spin (one warp) until flags[i] >= epoch for all i < n: for kernels that cannot wait themselves */
void launch_wait_flags(const unsigned long long *flags, int n, unsigned long long epoch, cudaStream_t s);

/* ------------------------------------------------------------------------
 * PANEL: private column-panel layout (built once at upload, on the device);
 * see spmv_panel.cu for the format.
 * ---------------------------------------------------------------------- */
struct DevPanel {
    const void     *val;        /* fmt 0: T[padded], tile-major SELL-pair order; fmt 2: the unified stream,
                                 * per pair row 32 value pairs then 32 column pairs (col unused) */
    const uint16_t *col;        /* u16[padded], 0-based column inside its panel; W = +0.0 slot */
    const ushort4  *meta;       /* fmt 0: [nblk * P * R/G]: {row A, row B, pair where B starts, 0} */
    const uint16_t *rowids;     /* fmt 2: [nblk * P * R]: the G tile-local row ids of every lane */
    const int      *slice_off;  /* int[nblk * P * R/32 + 1], element offsets (multiples of 64) */
    int rows, ncols;
    int R;                      /* rows per row block (multiple of 32 G) */
    int G;                      /* rows per thread (fmt 0: 1 or 2; fmt 2: 2, 4 or 8); the CTA has R/G threads */
    int fmt;                    /* 0: paired rows (spmv_panel.cu); 2: flagged streams, warp-major,
                                 * through a shared-memory ring (spmv_panelg.cu, spmv_panelr.cu) */
    int ring_K, ring_S;         /* fmt 2: pair rows per ring stage, stages per warp */
    int U;                      /* pairs per prefetch chunk (tuning) */
    int P;                      /* number of column panels */
    int W;                      /* columns per panel (even) */
    int nblk;                   /* row blocks = CTAs */
    int use_tma;                /* x slices by cp.async.bulk (else cooperative loads) */
    int nbuf;                   /* x slice buffers in shared memory: 2 (prefetch) or 1 (wide) */
    long long padded;           /* stored entries including padding */
};

/* build passes (device side) */
void launch_panel_count(const int *rowptr, const int *col, int rows, int P, int W, int R,
                        uint16_t *seglen, int *overflow, cudaStream_t s);
void launch_panel_sort(const uint16_t *seglen, int ntiles, int R, int G, int mode, ushort4 *meta,
                       int *slice_elems, cudaStream_t s);
template <typename T>
void launch_panel_fill(const T *val, const int *col, const int *rowptr, int rows,
                       const DevPanel &pm, const uint16_t *seglen, T *val_out, uint16_t *col_out,
                       cudaStream_t s);
/* y = A x on the panel layout; every row summed left to right.  dotv != NULL: CTA b also
 * writes its share of dotv . y to dot_partial[b] (nblk values, fixed reduction order) */
struct XFlags;
/* flags != NULL: x arrives chunk by chunk while the kernel runs, see XFlags below */
template <typename T>
void launch_panel(const DevPanel &pm, const T *x, T *y, cudaStream_t s, const T *dotv = nullptr,
                  T *dot_partial = nullptr, const XFlags *flags = nullptr);
size_t panel_smem_bytes(const DevPanel &pm, bool f32);

/* flagged-stream layout for wide matrices: build passes (spmv_panelg.cu) */
void launch_panelg_sort(const uint16_t *seglen, int ntiles, int R, int G, uint16_t *rowids,
                        int *slice_elems, cudaStream_t s);
template <typename T>
void launch_panelg_fill(const T *val, const int *col, const int *rowptr, int rows,
                        const DevPanel &pm, const uint16_t *seglen, unsigned char *stream_out,
                        cudaStream_t s);
/* x that arrives slice by slice from other GPUs (include/b200_peer.h): flags[r] >= epoch
 * <=> columns [r * cols_per_rank, (r + 1) * cols_per_rank) are in the buffer.  flags == NULL:
 * x is complete when the kernel starts. */
struct XFlags {
    const unsigned long long *flags;
    unsigned long long epoch;
    int cols_per_rank;
    int nranks;
    int *timed_out;                     /* watchdog (NULL: wait for ever): set to 1 when a slice did not */
    unsigned long long timeout_ns;      /* ... arrive within this time; the kernel then goes on regardless */
    int ready0;                         /* slices [0, ready0) are in the buffer by stream order: never waited for */
};
/* The exchange fused into the product: the kernel's CTAs first store this rank's slice into
 * every rank's buffer of the epoch (NVLink peer stores), the last one publishes the epoch;
 * then they walk the panels, waiting per slice (XFlags).  src == NULL: no push.
 * Only for grids whose CTAs are all co-resident (nblk <= SM count): the publishing CTA
 * waits for the others. */
struct XPush {
    const void *src;                    /* local slice */
    size_t bytes;                       /* its size, and */
    size_t offset;                      /* its byte offset inside the full vector */
    int rank, nranks;
    unsigned long long epoch;
    void *dst[8];                       /* every rank's full-length buffer of this epoch */
    unsigned long long *vflag[8];       /* every rank's row of vector flags */
    unsigned long long *rflag[8];       /* every rank's row of consumed flags */
    unsigned int *counter;              /* local arrival counter (zero between launches) */
};
/* ... and its kernel: the matrix stream through per-warp shared-memory rings (spmv_panelr.cu);
 * waits per x slice on `xf` before it requests the panels that need the slice */
template <typename T>
void launch_panelr(const DevPanel &pm, const T *x, T *y, const XFlags &xf, const XPush &xp, cudaStream_t s);
size_t panelr_smem_bytes(const DevPanel &pm, bool f32);

/* ------------------------------------------------------------------------
 * SELL: the same lane-stream layout without column panels -- x is gathered
 * through L2 with full 32-bit column indices.  Order-preserving for every
 * row up to `cap` entries, any column order; see spmv_sell.cu.
 * ---------------------------------------------------------------------- */
struct DevSell {
    const void    *val;        /* T[padded], SELL-pair order */
    const int     *col;        /* int[padded], 1-based global column (padding: 1) */
    const ushort4 *meta;       /* [nblk * R/G]: {row A, row B, entries of A, entries of B} */
    const int     *slice_off;  /* int[nblk * R/G/32 + 1] element offsets (multiples of 64) */
    int rows, R, G, nblk, U;
    long long padded;
    /* fmt 1 ("SELLU", spmv_sellu.cu): uniform row slots, entry-granular streams */
    int fmt;                   /* 0: paired rows (spmv_sell.cu); 1: uniform slots (spmv_sellu.cu) */
    const uint16_t *rowids;    /* fmt 1: [nblk][G][128] tile-local row of slot g of thread t */
    const uint16_t *slotlen;   /* fmt 1: [nblk][4][G] entries of slot g in every lane of warp w */
    /* rows above the cap: nnz-split chunks + ordered carry fix-up (re-ordering) */
    const int4 *chunks;        /* {row, lo, hi, carry slot or -1} per chunk; short chunks first */
    int n_chunks;
    int n_chunks_short;        /* chunks of <= sell_short_chunk_entries(): 8 lanes each */
    const int2 *multi;         /* {first carry slot, count} per multi-chunk row */
    const int  *multi_rows;    /* its row id */
    int n_multi;
    void *carry;               /* T[number of carried chunks] */
    int n_long;                /* rows above the cap */
};
void launch_sell_rowlen(const int *rowptr, int rows, int R, int cap, uint16_t *seglen, cudaStream_t s);
template <typename T>
void launch_sell_fill(const T *val, const int *col, const int *rowptr, int rows, const DevSell &sm,
                      const uint16_t *seglen, T *val_out, int *col_out, cudaStream_t s);
/* SELLU build passes and product (spmv_sellu.cu); the long-row path stays launch_sell's */
int sellu_threads();
void launch_sellu_sort(const uint16_t *seglen, int ntiles, int R, int G, uint16_t *rowids,
                       uint16_t *slotlen, int *slice_elems, cudaStream_t s);
template <typename T>
void launch_sellu_fill(const T *val, const int *col, const int *rowptr, int rows, const DevSell &sm, int cap,
                       T *val_out, int *col_out, cudaStream_t s);
template <typename T>
void launch_sellu(const DevSell &sm, const T *x, T *y, cudaStream_t s);
int sell_chunk_entries();
int sell_short_chunk_entries();
template <typename T>
void launch_sell(const DevSell &sm, const DevCsr &csr, const T *x, T *y, cudaStream_t s);

}  // namespace b200
