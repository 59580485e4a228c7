/*
 * spmv_panel.cu -- the PANEL kernel family: order-preserving CSR SpMV with
 * the x slice of the current column panel staged in shared memory.
 *
 * Why: on B200 a warp-wide gather of 8-byte x entries from L2 costs one
 * L1TEX wavefront per lane (32 per instruction) and drags a 32-byte sector
 * per entry across the L2 fabric; for NPB class C that alone is ~1 cycle per
 * nonzero per SM, more than the HBM roofline allows, and it clogs the
 * load/store pipe for every other access (profiles/r01_run1_*).  A gather
 * from shared memory costs ~6 wavefronts per 32 lanes.  So the matrix is
 * re-laid out once, at upload, into (row block x column panel) tiles whose x
 * slice fits in shared memory; the CTA walks the panels left to right, each
 * thread carrying its row's running sum in a register.  With sorted columns
 * (NPB's makea keeps them sorted, cg.f:838-850) this visits every row's
 * entries in their original order, so the result is bit-identical to the
 * reference loop (libspmv/native-impl.c:1-12): separately rounded multiply,
 * separately rounded add, left to right.
 *
 * Algorithmic bytes are unchanged; the private layout stores
 * 16-bit panel-local column indices, so the HBM stream is 10 B per nonzero.
 */
#include "spmv_kernels.cuh"

namespace b200 {

__device__ __forceinline__ double pmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  pmul(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double padd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  padd(float a, float b)   { return __fadd_rn(a, b); }

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

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
            if (cnt > 65535) atomicExch(overflow, 1);
            seglen[((size_t)rb * P + p_cur) * R + rr] = (uint16_t)cnt;
            cnt = 0;
            ++p_cur;
        }
        ++cnt;
    }
    while (p_cur < P) {
        if (cnt > 65535) atomicExch(overflow, 1);
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

/* ---- build: nonzeros of every warp slice --------------------------------- */
__global__ void panel_slice_sizes_kernel(const uint16_t *__restrict__ seglen, int nslices,
                                         int *__restrict__ slice_cnt)
{
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= nslices) return;
    int v = seglen[(size_t)gw * 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) slice_cnt[gw] = v;
}

void launch_panel_slice_sizes(const uint16_t *seglen, int nslices, int *slice_cnt, cudaStream_t s)
{
    if (nslices <= 0) return;
    const long long threads = (long long)nslices * 32;
    panel_slice_sizes_kernel<<<(int)((threads + 255) / 256), 256, 0, s>>>(seglen, nslices, slice_cnt);
}

/* ---- build: scatter CSR entries into the ragged tile order --------------- */
template <typename T>
__global__ void panel_fill_kernel(const T *__restrict__ val, const int *__restrict__ col,
                                  const int *__restrict__ rowptr, int rows, int R, int P, int W,
                                  const uint16_t *__restrict__ seglen,
                                  const int *__restrict__ slice_off, int nslices,
                                  T *__restrict__ val_out, uint16_t *__restrict__ col_out)
{
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   /* global slice id */
    const int lane = threadIdx.x & 31;
    if (gw >= nslices) return;
    const int spb = R / 32;                   /* slices per (row block, panel) */
    const int tile = gw / spb, w = gw - tile * spb;
    const int rb = tile / P, p = tile - rb * P;
    const int rr = w * 32 + lane;
    const int r = rb * R + rr;
    int len = 0, src = 0;
    if (r < rows) {
        len = seglen[((size_t)rb * P + p) * R + rr];
        src = rowptr[r];
        for (int q = 0; q < p; ++q) src += seglen[((size_t)rb * P + q) * R + rr];
    }
    int off = slice_off[gw];
    const int maxlen = __reduce_max_sync(0xffffffffu, len);
    const unsigned lt = lanemask_lt();
    for (int k = 0; k < maxlen; ++k) {
        const bool act = k < len;
        const unsigned m = __ballot_sync(0xffffffffu, act);
        if (act) {
            const int idx = off + __popc(m & lt);
            val_out[idx] = val[src + k];
            col_out[idx] = (uint16_t)(col[src + k] - 1 - p * W);
        }
        off += __popc(m);
    }
}

template <typename T>
void launch_panel_fill(const T *val, const int *col, const int *rowptr, int rows,
                       const DevPanel &pm, T *val_out, uint16_t *col_out, cudaStream_t s)
{
    const int nslices = pm.nblk * pm.P * (pm.R / 32);
    if (nslices <= 0) return;
    const long long threads = (long long)nslices * 32;
    panel_fill_kernel<T><<<(int)((threads + 255) / 256), 256, 0, s>>>(
        val, col, rowptr, rows, pm.R, pm.P, pm.W, pm.seglen, pm.slice_off, nslices, val_out, col_out);
}
template void launch_panel_fill<double>(const double *, const int *, const int *, int,
                                        const DevPanel &, double *, uint16_t *, cudaStream_t);
template void launch_panel_fill<float>(const float *, const int *, const int *, int,
                                       const DevPanel &, float *, uint16_t *, cudaStream_t);

/* ------------------------------------------------------------------------
 * the product
 * ---------------------------------------------------------------------- */
template <typename T, int U>
__global__ void __launch_bounds__(1024, 1)
spmv_panel_kernel(const T *__restrict__ val, const uint16_t *__restrict__ col,
                  const uint16_t *__restrict__ seglen, const int *__restrict__ slice_off,
                  const T *__restrict__ x, T *__restrict__ y,
                  int rows, int ncols, int P, int W)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *xs = reinterpret_cast<T *>(smem_raw);

    const int R = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int spb = R >> 5;
    const int rb = blockIdx.x;
    const int row = rb * R + tid;
    const unsigned lt = lanemask_lt();
    const bool x_vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((W * sizeof(T)) % 16 == 0);

    T acc = (T)0;
    for (int p = 0; p < P; ++p) {
        const int cbase = p * W;
        const int cw = min(W, ncols - cbase);
        __syncthreads();                       /* previous panel fully consumed */
        if (x_vec_ok) {
            constexpr int VE = 16 / sizeof(T);
            const int nv = cw / VE;
            const int4 *src = reinterpret_cast<const int4 *>(x + cbase);
            int4 *dst = reinterpret_cast<int4 *>(xs);
            for (int i = tid; i < nv; i += R) dst[i] = __ldg(src + i);
            for (int i = nv * VE + tid; i < cw; i += R) xs[i] = __ldg(x + cbase + i);
        } else {
            for (int i = tid; i < cw; i += R) xs[i] = __ldg(x + cbase + i);
        }
        const size_t tile = (size_t)rb * P + p;
        const int len = seglen[tile * R + tid];
        int off = slice_off[tile * spb + warp];
        __syncthreads();                       /* x slice visible */

        const int maxlen = __reduce_max_sync(0xffffffffu, len);
        for (int k0 = 0; k0 < maxlen; k0 += U) {
            int  idx[U];
            bool act[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                act[u] = (k0 + u) < len;
                const unsigned m = __ballot_sync(0xffffffffu, act[u]);
                idx[u] = off + __popc(m & lt);
                off += __popc(m);
            }
            T v[U];
            int c[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v[u] = (T)0;
                c[u] = 0;
                if (act[u]) {
                    v[u] = __ldcs(val + idx[u]);
                    c[u] = __ldcs(col + idx[u]);
                }
            }
            T xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = xs[c[u]];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (act[u]) acc = padd(acc, pmul(v[u], xv[u]));
        }
    }
    if (row < rows) y[row] = acc;
}

size_t panel_smem_bytes(const DevPanel &pm, bool f32)
{
    return (size_t)pm.W * (f32 ? 4 : 8);
}

template <typename T>
void launch_panel(const DevPanel &pm, const T *x, T *y, cudaStream_t s)
{
    if (pm.nblk <= 0) return;
    constexpr int U = 8;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(spmv_panel_kernel<T, U>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024);
        attr_set = true;
    }
    const size_t smem = panel_smem_bytes(pm, sizeof(T) == 4);
    spmv_panel_kernel<T, U><<<pm.nblk, pm.R, smem, s>>>(
        static_cast<const T *>(pm.val), pm.col, pm.seglen, pm.slice_off, x, y,
        pm.rows, pm.ncols, pm.P, pm.W);
}
template void launch_panel<double>(const DevPanel &, const double *, double *, cudaStream_t);
template void launch_panel<float>(const DevPanel &, const float *, float *, cudaStream_t);

}  // namespace b200
