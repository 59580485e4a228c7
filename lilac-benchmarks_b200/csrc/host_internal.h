/*
 * host_internal.h -- what the two host translation units of libb200-spmv share:
 *   b200_host.cu    per-device contexts, upload (layout selection and build),
 *                   launch, the resident-matrix API of include/b200_spmv.h part 2
 *   b200_dropin.cu  the libspmv ABI (spmv_harness_ / f_spmv_harness_): resident
 *                   cache keyed by host pointers, write guard, x / y movement,
 *                   and the multi-device ("ABI mode") path
 * Not installed; the public surface is include/b200_spmv.h.
 */
#pragma once
#include "../../include/b200_spmv.h"
#include "spmv_kernels.cuh"

#include <cuda_runtime.h>
#include <pthread.h>
#include <stdint.h>

namespace b200 {

[[noreturn]] void die(const char *fmt, ...);

#define CUDA_OK(call)                                                          \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess)                                                 \
            ::b200::die("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,   \
                        cudaGetErrorString(e_));                               \
    } while (0)

double now_ms(void);
int env_int(const char *name, int dflt);
inline size_t elem_size(int dtype) { return dtype == B200_F32 ? 4 : 8; }

constexpr int kMaxDevices = 8;

/* One per device this process has touched: the library's stream and timing events
 * there.  Library state is per device, never "the" device: every entry point sets
 * the device of the object it works on and restores the caller's afterwards. */
struct DevCtx {
    int device;
    int sm_count;
    cudaStream_t stream;          /* non-blocking; uploads and the drop-in path run here */
    cudaStream_t copy_stream;     /* drop-in path: x chunks host -> device (copy engine) while the product runs */
    cudaEvent_t ev0, ev1;         /* kernel timing of the drop-in path */
    cudaEvent_t ev_x;             /* multi-device: "my x slice is in every device's buffer" */
};

/* RAII: make `device` current, restore the previous one on scope exit */
struct DeviceScope {
    int prev;
    explicit DeviceScope(int device)
    {
        prev = -1;
        cudaGetDevice(&prev);
        if (prev != device) CUDA_OK(cudaSetDevice(device));
        else prev = -1;
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

extern pthread_mutex_t g_lock;    /* one lock for all host state (the ABI is not re-entrant anyway) */
extern int g_verbose;

/* all of these expect g_lock to be held */
void ensure_init_locked(int device);
int default_device_locked(void);
DevCtx *ctx_for_device_locked(int device);

}  // namespace b200

/* resident matrix (opaque in the public header) */
struct b200_matrix {
    int dtype;                 /* B200_F64 / B200_F32 */
    int kernel;                /* family in use */
    int lanes;                 /* VECTOR: lanes per row */
    int device;
    b200::DevCtx *ctx;
    int rows, ncols;
    int64_t nnz;
    b200::DevCsr dev;          /* device pointers */
    void *d_val; int *d_col; int *d_rowptr; int *d_rowblk;
    int64_t resident_bytes;
    b200::UploadScan scan;
    /* PANEL layout (when kernel == B200_KERNEL_PANEL) */
    b200::DevPanel panel;
    void *d_pval; uint16_t *d_pcol; ushort4 *d_meta; int *d_slice_off;
    /* SELL layout (when kernel == B200_KERNEL_SELL) */
    b200::DevSell sell;
    /* SMALL (when kernel == B200_KERNEL_SMALL): the CSR arrays above + its own row blocks */
    b200::DevSmall small_;
    int *d_small_blk; uint16_t *d_small_col16;
    int *d_scol; int4 *d_chunks; int2 *d_multi; int *d_multi_rows; void *d_carry;
};

namespace b200 {

/* x arrives slice by slice (include/b200_peer.h): flag[r] >= epoch means that the slice
 * of rank r -- columns [r * cols_per_rank, (r + 1) * cols_per_rank) -- is in the buffer */
struct SliceFlags {
    const unsigned long long *flags;
    unsigned long long epoch;
    int cols_per_rank;
    int nranks;
    int *timed_out;                     /* watchdog, see XFlags; NULL: none */
    unsigned long long timeout_ns;
    int ready0;                         /* slices [0, ready0) are complete by stream order */
};

b200_matrix *upload_locked(DevCtx *ctx, const void *a, const int *rowstr, const int *colidx,
                           int rows, int dtype, int kernel);
void release_locked(b200_matrix *m);
/* launches on `s` (a stream of m->device, which must be current); returns the number of
 * kernels launched.  `sf` != NULL: the kernel itself waits for the x slices (only the
 * PANEL kernels can -- exec_takes_flags; for the others the caller must have waited) */
int exec_locked(b200_matrix *m, const void *d_x, void *d_y, cudaStream_t s, const SliceFlags *sf,
                const XPush *xp = nullptr);
bool exec_takes_flags(const b200_matrix *m);        /* paired PANEL and RING kernels */
bool exec_takes_guarded_flags(const b200_matrix *m);    /* paired PANEL kernel: flags with a watchdog on the wait */
bool exec_waits_in_kernel(const b200_matrix *m);    /* RING kernel: flags and the fused push (multi-GPU paths) */

}  // namespace b200
