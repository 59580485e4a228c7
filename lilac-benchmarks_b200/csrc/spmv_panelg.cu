/*
 * spmv_panelg.cu -- build passes of the flagged-stream layout for wide
 * matrices (consumed by the ring kernel, spmv_panelr.cu).
 *
 * Every CTA of a PANEL kernel reads the whole x once per row block, so the
 * x traffic from L2 is  ceil(rows / R) * ncols * sizeof(T)  and only a tall
 * row block (large R) keeps it below the matrix stream: NPB class D has 1.5 M
 * columns (12 MB of x) and ~463 entries per row, so R = 1024 reloads 9.4 KB
 * of x per row for 4.9 KB of matrix.  This layout raises R to G * T rows
 * (G = 2, 4 or 8 rows per lane, T <= 512 threads, R <= 4096):
 *
 *   - a lane stream is the entries of its G rows back to back; bit 15 of the
 *     16-bit panel-local column marks the first entry of a row, so the
 *     consumer needs no switch table;
 *   - the per-panel metadata is just the G tile-local row ids of the lane
 *     (2 bytes per (row, panel));
 *   - rows are dealt to the lanes boustrophedon over the per-tile sorted
 *     order (ranks t, 2T-1-t, 2T+t, 4T-1-t, ...), so rows with no entry in
 *     the panel come last in every lane (they are never switched to), and
 *     the g-th rows of the 32 lanes of a slice are neighbours in the sorted
 *     order: every such row slot is padded to the slice's longest (< 1 %
 *     extra entries on NPB class D), which puts the row switches of all
 *     lanes of a warp on the same pair;
 *   - slices are stored (row block, warp, panel)-major, so a warp's stream
 *     over all panels is one contiguous run of pair rows; a pair row is 32
 *     value pairs followed by its 32 column pairs (640 bytes in fp64), so any
 *     run of pair rows is one contiguous block of bytes.
 *
 * Padding entries are (+0.0, slot W) with slot W of the x slice holding +0.0:
 * they add +0.0 to a running sum that is never -0.0, i.e. change no bit.
 * A register-staged kernel on this layout (two chunks of U pairs per lane, as
 * in spmv_panel.cu) was measured and dropped: with ~4 entries per (row,
 * panel) its prefetch reaches one panel ahead and it stays below the ring
 * kernel everywhere (profiles/r01_run28, r01_run36).
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
                                   unsigned char *__restrict__ stream_out)
{
    /* one pair row of a slice = 32 value pairs followed by 32 column pairs (kRowB bytes):
     * a ring stage of K pair rows is then ONE contiguous block for ONE TMA bulk copy */
    constexpr int kRowB = 32 * (2 * (int)sizeof(T) + 4);
    auto put = [&](size_t pair_row, int lane_, int half, T v, uint16_t c) {
        unsigned char *row = stream_out + pair_row * kRowB;
        reinterpret_cast<T *>(row)[lane_ * 2 + half] = v;
        reinterpret_cast<uint16_t *>(row + 64 * sizeof(T))[lane_ * 2 + half] = c;
    };
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   /* global slice id */
    const int lane = threadIdx.x & 31;
    if (gw >= nslices) return;
    const int Tn = R / G;
    const int spb = Tn >> 5;
    const int tile = gw / spb, w = gw - tile * spb;
    const int rb = tile / P, p = tile - rb * P;
    /* slice order in memory: (row block, warp, panel) -- the warps of the ring kernel
     * stream their slices of all panels back to back */
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
            for (int e = 0; e < len; ++e, ++k)
                put((size_t)(off >> 6) + (k >> 1), lane, k & 1, val[lo + e],
                    (uint16_t)((col[lo + e] - c0) | (e == 0 ? 0x8000 : 0)));
        }
        for (; k < k_end; ++k)                                     /* +0.0 * x[W] = +0.0 */
            put((size_t)(off >> 6) + (k >> 1), lane, k & 1, (T)0, (uint16_t)W);
    }
    for (; k < nent; ++k) put((size_t)(off >> 6) + (k >> 1), lane, k & 1, (T)0, (uint16_t)W);
}

template <typename T>
void launch_panelg_fill(const T *val, const int *col, const int *rowptr, int rows,
                        const DevPanel &pm, const uint16_t *seglen, unsigned char *stream_out,
                        cudaStream_t s)
{
    const int nslices = pm.nblk * pm.P * (pm.R / pm.G / 32);
    if (nslices <= 0) return;
    const long long threads = (long long)nslices * 32;
    panelg_fill_kernel<T><<<(int)((threads + 255) / 256), 256, 0, s>>>(
        val, col, rowptr, rows, pm.R, pm.G, pm.P, pm.W, seglen, pm.rowids, pm.slice_off, nslices,
        pm.fmt == 2, stream_out);
}
template void launch_panelg_fill<double>(const double *, const int *, const int *, int, const DevPanel &,
                                         const uint16_t *, unsigned char *, cudaStream_t);
template void launch_panelg_fill<float>(const float *, const int *, const int *, int, const DevPanel &,
                                        const uint16_t *, unsigned char *, cudaStream_t);

}  // namespace b200
