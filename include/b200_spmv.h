/*
 * b200_spmv.h -- C ABI of the `b200` libspmv platform (libb200-spmv.so).
 *
 * Part 1 is the drop-in boundary: exactly the two symbols every libspmv
 * backend of mob-group/lilac-benchmarks exports, so the reference's callers
 * relink (-lb200-spmv) or dlopen it unchanged.  Part 2 is the resident-matrix
 * API the drop-in symbols are built on; it takes DEVICE pointers and a CUDA
 * stream and is what bench.py, the tests and the multi-GPU row-block path
 * drive.  Plain C types only; no torch types anywhere.
 *
 * There is no CPU fallback: every entry point needs a CUDA device and aborts
 * (stderr message + abort(), like the asserts of libspmv/gpu.c:42-80) when
 * there is none.
 */
#ifndef B200_SPMV_H
#define B200_SPMV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * Part 1 -- the libspmv ABI
 *
 * Replaces: libspmv/native.c:3-6   (spmv_harness_, CPU loop)
 *           libspmv/gpu.c:211-289  (spmv_harness_, cuSPARSE csrmv_mp)
 *           libspmv/mkl.c:28-75    (spmv_harness_, MKL)
 * and the fp32 twins native.c:8-11 / gpu.c:291-369.
 *
 *   ov      out  y[*rows]                      (overwritten, alpha=1 beta=0)
 *   a       in   values[nnz]
 *   iv      in   x[>= max(colidx)]
 *   rowstr  in   [*rows+1] 1-based offsets, nnz = rowstr[*rows]-rowstr[0]
 *   colidx  in   [nnz] 1-based column indices, any order, repeats allowed
 *   rows    in   pointer to the row count (Fortran by-reference)
 *
 * The matrix is uploaded on first sight and stays resident in HBM, keyed by
 * (a, rowstr, colidx, *rows, nnz) like gpu.c:227-262; each call moves only x
 * (host->device) and y (device->host).  Returns NULL (gpu.c:288); no caller
 * reads the result.
 * ---------------------------------------------------------------------- */
void *spmv_harness_(double *ov, double *a, double *iv,
                    int *rowstr, int *colidx, int *rows);
void *f_spmv_harness_(float *ov, float *a, float *iv,
                      int *rowstr, int *colidx, int *rows);

/* ------------------------------------------------------------------------
 * Part 2 -- resident-matrix API (new; the reference has no equivalent: its
 * GPU state is file-static in gpu.c:12-34)
 * ---------------------------------------------------------------------- */

typedef struct b200_matrix b200_matrix;     /* opaque, one resident CSR (row block) */

/* kernel families */
enum {
    B200_KERNEL_AUTO    = 0,  /* chosen from the row-length histogram at upload */
    B200_KERNEL_ORDERED = 1,  /* nnz-split row blocks, products staged in shared
                                 memory, each row summed strictly left to right:
                                 bit-identical to native-impl.c:1-12 */
    B200_KERNEL_VECTOR  = 2,  /* 2..32 lanes per row + warp-shuffle reduction */
    B200_KERNEL_PANEL   = 3,  /* ORDERED on the column-panel private layout with
                                 x slices staged in shared memory (sorted rows) */
    B200_KERNEL_MERGE   = 4,  /* fixed-nnz split with carry-out fix-up for
                                 heavily skewed row lengths */
    B200_KERNEL_SELL    = 5,  /* lane streams over the sorted-SELL private layout,
                                 x gathered through L2; left-to-right rows for any
                                 column order (wide matrices, unsorted rows) */
    B200_KERNEL_SMALL   = 6   /* small matrices: the whole x in shared memory, one
                                 nnz-balanced row block per SM, products staged, one
                                 thread per row adds left to right (any column order) */
};

enum { B200_F64 = 0, B200_F32 = 1 };

/* Select / initialise the device of the drop-in path (-1 = keep cudaGetDevice()).
 * Optional: every entry point initialises lazily, as dlopen callers need
 * (pagerank/main.cpp:19).  b200_spmv_upload / _upload_device build on the device that
 * is current when they are called; exec / release switch to the matrix's device and
 * restore the caller's, so one process can hold matrices on several GPUs. */
int b200_spmv_init(int device);

/* Upload a 1-based CSR row block from HOST arrays and keep it resident.
 * `rowstr` points at the block's first row pointer (rows+1 entries); a/colidx
 * are the bases the offsets refer to, exactly as in the ABI, so a row block
 * of a larger matrix is (a, colidx, rowstr + row_lo, row_hi - row_lo).
 * dtype: B200_F64 / B200_F32.  kernel: B200_KERNEL_*.  Aborts on error. */
b200_matrix *b200_spmv_upload(const void *a, const int *rowstr, const int *colidx,
                              int rows, int dtype, int kernel);
void b200_spmv_release(b200_matrix *m);

/* y[0..rows) = A x on `stream` (a cudaStream_t passed as void*; NULL = the
 * legacy default stream).  d_x, d_y are DEVICE pointers of the matrix dtype;
 * d_x must hold at least b200_spmv_ncols(m) elements.  Asynchronous. */
int b200_spmv_exec(b200_matrix *m, const void *d_x, void *d_y, void *stream);

/* The same on x that is still arriving slice by slice -- from `nranks` GPUs
 * (include/b200_peer.h), or from the host through a copy engine (what the drop-in symbols
 * do per call): flags[r] >= epoch <=> columns [r * cols_per_rank, (r + 1) * cols_per_rank)
 * are in d_x.  The kernel waits per slice, just before the panels that need it, so the
 * product overlaps the transfer.  Both PANEL kernels can (column panels walked left to
 * right); returns -1 without launching for the other families (wait for the vector first,
 * then b200_spmv_exec). */
int b200_spmv_exec_sliced(b200_matrix *m, const void *d_x, void *d_y, void *stream,
                          const unsigned long long *flags, unsigned long long epoch,
                          int cols_per_rank, int nranks);

/* The exchange fused into the product (include/b200_peer.h): the kernel first stores this
 * rank's slice v_local[0..n_local) -- element offset `lo` of the full vector -- into every
 * rank's buffer of `epoch` over NVLink peer memory and publishes the epoch, then runs the
 * product on the local buffer, waiting per slice.  One launch per sharded step.  Returns -1
 * without launching when the kernel family or the grid cannot do that (more row blocks
 * than SMs, odd slice boundaries): use b200_peer_post + b200_spmv_exec_sliced then. */
int b200_spmv_exec_pushed(b200_matrix *m, void *d_y, void *stream, void *peer_group,
                          const double *v_local, int n_local, int64_t lo, uint64_t epoch,
                          int cols_per_rank);

/* y = A x with a dot product fused into the epilogue: CTA b adds its rows' share of
 * dotv . y into partial[b], b < b200_spmv_dot_partials(m) (0: this matrix's kernel has no
 * fused epilogue, b200_spmv_exec_dot returns -1 without launching).  Fixed reduction order.
 * NPB conj_grad's d = p.q (cg.f:573-576) right where q is produced. */
int b200_spmv_dot_partials(const b200_matrix *m);
int b200_spmv_exec_dot(b200_matrix *m, const void *d_x, void *d_y, const void *d_dotv,
                       void *d_partial, void *stream);
int b200_spmv_can_push(const b200_matrix *m);    /* 1: b200_spmv_exec_pushed works for this matrix */

/* Upload from DEVICE arrays of the current device (same 1-based contents as the ABI;
 * the on-device NPB generator produces them): no PCIe traffic. */
b200_matrix *b200_spmv_upload_device(const void *d_a, const int *d_rowstr, const int *d_colidx,
                                     int rows, int dtype, int kernel);

int         b200_spmv_device(const b200_matrix *m);  /* device the matrix lives on */
int         b200_spmv_waits_in_kernel(const b200_matrix *m);  /* 1: the wide-matrix (ring) PANEL kernel: sliced and pushed products */
int         b200_spmv_rows(const b200_matrix *m);
int         b200_spmv_ncols(const b200_matrix *m);   /* max(colidx) */
int64_t     b200_spmv_nnz(const b200_matrix *m);
int         b200_spmv_kernel(const b200_matrix *m);  /* family actually chosen */
const char *b200_spmv_kernel_name(const b200_matrix *m);
int         b200_spmv_launches_per_exec(const b200_matrix *m);
/* algorithmic bytes of one product: 12 nnz + 4 (rows+1) + 8 ncols + 8 rows for
 * fp64, 8/4/4/4 for fp32 */
int64_t     b200_spmv_algorithmic_bytes(const b200_matrix *m);
/* bytes the resident private layout occupies in HBM */
int64_t     b200_spmv_resident_bytes(const b200_matrix *m);

/* row-length histogram gathered at upload: bin k counts rows with
 * 2^(k-1) < len <= 2^k (bin 0: empty rows and len 1), 32 bins */
void b200_spmv_row_histogram(const b200_matrix *m, int64_t bins[32],
                             int *min_len, int *max_len);

/* nnz-balanced contiguous row partition: writes
 * parts+1 boundaries into `bounds` so that every part holds ~nnz/parts. */
void b200_spmv_partition_rows(const int *rowstr, int rows, int parts, int *bounds);

/* Resident cache of the drop-in symbols.
 *
 * Staleness: like libspmv/gpu.c:140-262 the host pages of a resident matrix are
 * write-protected and a write to them drops the device copy (B200_SPMV_GUARD=0 turns
 * that off; then only the sampled per-call fingerprint notices a change).  What
 * neither can see -- arrays freed and mapped again at the same addresses with other
 * contents of the same shape -- MUST be announced with b200_spmv_invalidate().
 *
 * Several devices: B200_SPMV_DEVICES="0,1,2,3" (or "all") makes the drop-in symbols
 * spread every matrix of at least B200_SPMV_MULTI_MIN_NNZ (default 4 Mi) nonzeros over
 * those GPUs in nnz-balanced row blocks; callers are unchanged. */
void b200_spmv_invalidate(void);   /* forget every cached matrix (host arrays changed) */
int  b200_spmv_devices_in_use(void);   /* most devices any cached matrix is spread over */

typedef struct {
    uint64_t calls;            /* ABI calls served */
    uint64_t uploads;          /* matrix uploads (cache misses) */
    uint64_t kernel_launches;  /* SpMV kernels launched by the ABI calls */
    double   kernel_ms;        /* CUDA-event time of those kernels */
    double   e2e_ms;           /* wall time inside the ABI calls, uploads excluded */
    double   upload_ms;        /* wall time of the uploads */
    uint64_t h2d_bytes;        /* x traffic */
    uint64_t d2h_bytes;        /* y traffic */
    uint64_t auto_pinned_calls;   /* calls that moved x or y through a vector this library registered */
    uint64_t auto_pin_revoked;    /* registrations dropped because the owner had remapped the memory */
    uint64_t x_overlapped_calls;  /* calls whose x went up chunk by chunk (copy engine) while the product ran */
    uint64_t x_overlap_timeouts;  /* ... whose product gave up waiting for a chunk (streams serialised by a
                                   * profiler, CUDA_LAUNCH_BLOCKING, ...): redone the plain way, overlap off */
} b200_spmv_stats;
void b200_spmv_get_stats(b200_spmv_stats *out);
void b200_spmv_reset_stats(void);
/* CUDA-event timing of the product inside every drop-in call (feeds kernel_ms above).  Off by
 * default -- two event records and a query per call are ~3 us of a 130 us call --;
 * B200_SPMV_TIME_KERNELS=1 or this switch turn it on. */
void b200_spmv_set_time_kernels(int on);
/* The copy the drop-in path uses from a pageable caller vector into its pinned bounce buffer
 * (cache-bypassing stores + store fence; B200_SPMV_NT_COPY=0: memcpy).  Host only. */
void b200_spmv_host_copy(void *dst, const void *src, size_t bytes);

/* Pin a caller-owned host vector so x / y move in place over PCIe instead of through the
 * library's pinned bounce buffer (class C: 0.12 ms per call instead of 0.23 ms).  Either the
 * owner does it (b200_spmv_pin_host, or its own cudaHostAlloc / cudaHostRegister), or the
 * library does: with B200_SPMV_PIN_HOST=N / b200_spmv_set_auto_pin(N), N > 0, the drop-in
 * symbols register a vector once they have seen it N times at the same address (NPB's COMMON
 * vectors, pagerank's two std::vectors), and verify on every call -- sample words of x and y
 * as seen through the GPU mapping against the host's view -- that its owner has not remapped
 * the memory; a stale registration is dropped and the call redone through the bounce buffer.
 * Off by default: see b200_dropin.cu.  set_auto_pin(0) unregisters everything again. */
void b200_spmv_set_auto_pin(int sightings);
int b200_spmv_pin_host(void *ptr, size_t bytes);
int b200_spmv_unpin_host(void *ptr);

const char *b200_spmv_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_SPMV_H */
