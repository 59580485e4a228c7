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

/* ------------------------------------------------------------------------
 * PANEL: private column-panel layout (built once at upload, on the device)
 *
 *   rows are cut into row blocks of R rows (one CTA each, R threads, one row
 *   per thread); columns into P panels of W columns whose x slice fits in
 *   shared memory.  Tile (rb, p) stores, warp slice by warp slice, the
 *   entries of each row that fall into panel p, k-major and compacted
 *   ("ragged": entry k of lane l sits at slice_off + sum_{k'<k} active(k')
 *   + rank of l among the lanes active at k), so a warp reads contiguous
 *   memory while every lane walks its own row left to right.
 * ---------------------------------------------------------------------- */
struct DevPanel {
    const void     *val;        /* T[nnz + pad], tile-major ragged order */
    const uint16_t *col;        /* u16[nnz + pad], 0-based column inside its panel */
    const uint16_t *seglen;     /* u16[nblk * P * R], entries of (row, panel) */
    const int      *slice_off;  /* int[nblk * P * R/32 + 1] */
    int rows, ncols;
    int R;                      /* rows per row block = threads per CTA (multiple of 32) */
    int P;                      /* number of column panels */
    int W;                      /* columns per panel */
    int nblk;                   /* row blocks = CTAs */
};

/* build passes (device side) */
void launch_panel_count(const int *rowptr, const int *col, int rows, int P, int W, int R,
                        uint16_t *seglen, int *overflow, cudaStream_t s);
void launch_panel_slice_sizes(const uint16_t *seglen, int nslices, int *slice_cnt, cudaStream_t s);
template <typename T>
void launch_panel_fill(const T *val, const int *col, const int *rowptr, int rows,
                       const DevPanel &pm, T *val_out, uint16_t *col_out, cudaStream_t s);
/* y = A x on the panel layout; every row summed left to right */
template <typename T>
void launch_panel(const DevPanel &pm, const T *x, T *y, cudaStream_t s);
size_t panel_smem_bytes(const DevPanel &pm, bool f32);

}  // namespace b200
