/*
 * spmv_sellu.cu -- SELL with uniform row slots ("SELLU", DevSell::fmt == 1): the
 * order-preserving gather kernel for matrices with MANY SHORT rows (SparseBench crsmat:
 * 4.9 entries per row; graphs) and the fallback for everything the panel kernels refuse.
 *
 * What the first SELL layout (spmv_sell.cu) loses on such matrices
 * (profiles/r02_run1_sell_crsmat170_*): two rows per lane dealt longest-with-shortest puts
 * the tile's longest rows into warp 0 and pads every 32-lane slice to its longest lane --
 * 32 % padding on crsmat170 -- and the CTA ends in a barrier that waits for that warp
 * (22 % of the stall samples); 19 000 CTAs of 256 rows each pay their start-up latency
 * chain (meta -> offsets -> stream -> gather) for 15 KB of stream.
 *
 * Layout (built on the device at upload):
 *   - a tile is R = T * G rows (T = 128 threads, G = 2..16 rows per lane, so a lane stream
 *     is ~100 entries whatever the mean row length); its rows are sorted by length and
 *     dealt rank g * T + t  ->  slot g of thread t: the 32 rows that share a slot of a warp
 *     are neighbours in the sorted order, so padding every slot to its warp's longest row
 *     costs a few per cent, and all four warps get the same mix of lengths;
 *   - a slot ends at the same stream position in all 32 lanes: the row switch is
 *     warp-uniform (no flags, no divergence), its positions are G 16-bit counts per warp;
 *   - streams are stored entry by entry (value 8 B, 32-bit global column 4 B), position s
 *     of lane l at  off + 32 s + l: every load is one fully coalesced line or two;
 *     padding entries carry column 0 (columns are 1-based) and are skipped, not added.
 * The product: each lane walks its stream left to right with separately rounded multiply
 * and add (libspmv/native-impl.c:1-12 order => bit-identical, any column order), two
 * chunks of U entries in flight, x gathered with an L2 evict-last policy (the matrix
 * stream is evict-first), y staged in shared memory and stored coalesced.
 * Rows above the cap are left to the nnz-split long-row path of spmv_sell.cu.
 */
#include "spmv_kernels.cuh"

namespace b200 {

namespace {

__device__ __forceinline__ double umul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  umul(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double uadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  uadd(float a, float b)   { return __fadd_rn(a, b); }

constexpr int kT = 128;            /* threads per tile */
constexpr int kW = kT / 32;

/* x gathers: keep x in L2 against the matrix stream that flows through it */
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double ld_x(const double *p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_x(const float *p, uint64_t pol)
{
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}

/* ---- build: sort the rows of a tile, deal them to the slots ------------------ */
__global__ void __launch_bounds__(1024)
sellu_sort_kernel(const uint16_t *__restrict__ seglen, int R, int G, int NK,
                  uint16_t *__restrict__ rowids, uint16_t *__restrict__ slotlen,
                  int *__restrict__ slice_elems)
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
    if (t >= kT) return;                       /* whole warps leave together */
    const int w = t >> 5;
    int total = 0;
    for (int g = 0; g < G; ++g) {
        const uint32_t kk = key[g * kT + t];
        int len = (int)(kk >> 16);
        rowids[((size_t)tile * G + g) * kT + t] = (uint16_t)(0xFFFF - (kk & 0xFFFFu));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
        if ((t & 31) == 0) slotlen[((size_t)tile * kW + w) * G + g] = (uint16_t)len;
        total += len;
    }
    if ((t & 31) == 0) slice_elems[tile * kW + w] = total * 32;
}

/* ---- build: scatter CSR entries into the slot streams ------------------------- */
template <typename T>
__global__ void sellu_fill_kernel(const T *__restrict__ val, const int *__restrict__ col,
                                  const int *__restrict__ rowptr, int rows, int R, int G, int cap,
                                  const uint16_t *__restrict__ rowids, const uint16_t *__restrict__ slotlen,
                                  const int *__restrict__ slice_off, int nslices,
                                  T *__restrict__ val_out, int *__restrict__ col_out)
{
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;        /* (tile, warp) */
    const int lane = threadIdx.x & 31;
    if (gw >= nslices) return;
    const int tile = gw / kW, w = gw - tile * kW;
    const size_t off = (size_t)slice_off[gw];
    size_t pos = 0;
    for (int g = 0; g < G; ++g) {
        const int r = tile * R + rowids[((size_t)tile * G + g) * kT + w * 32 + lane];
        int len = 0, src = 0;
        if (r < rows) {
            src = rowptr[r];
            len = rowptr[r + 1] - src;
            if (len > cap) len = 0;                                     /* long-row path */
        }
        const int sl = slotlen[((size_t)tile * kW + w) * G + g];
        for (int k = 0; k < sl; ++k) {
            const size_t idx = off + (pos + k) * 32 + lane;
            val_out[idx] = k < len ? val[src + k] : (T)0;
            col_out[idx] = k < len ? col[src + k] : 0;                  /* 0: padding, skipped */
        }
        pos += sl;
    }
}

/* ---- the product --------------------------------------------------------------- */
template <typename T, int U>
struct UChunk { T v[U]; int c[U]; };

template <typename T, int U>
__device__ __forceinline__ void sellu_load(UChunk<T, U> &ch, const T *vp, const int *cp, int s, int L)
{
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (s + u < L) {
            ch.v[u] = __ldcs(vp + (size_t)(s + u) * 32);
            ch.c[u] = __ldcs(cp + (size_t)(s + u) * 32);
        }
    }
}

template <typename T, int U>
__global__ void __launch_bounds__(kT)
spmv_sellu_kernel(const T *__restrict__ val, const int *__restrict__ col,
                  const uint16_t *__restrict__ rowids, const uint16_t *__restrict__ slotlen,
                  const int *__restrict__ slice_off, const T *__restrict__ xm1, T *__restrict__ y,
                  int rows, int R, int G)
{
    extern __shared__ __align__(16) unsigned char sellu_smem[];
    T *ystage = reinterpret_cast<T *>(sellu_smem);                                  /* [R] */
    uint16_t *rid = reinterpret_cast<uint16_t *>(sellu_smem + (size_t)R * sizeof(T));   /* [G][kT] */
    uint16_t *sl = rid + (size_t)G * kT;                                            /* [kW][G] */
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int tile = blockIdx.x;
    for (int g = 0; g < G; ++g) rid[g * kT + tid] = rowids[((size_t)tile * G + g) * kT + tid];
    if (lane < G) sl[w * G + lane] = slotlen[((size_t)tile * kW + w) * G + lane];
    const int off = slice_off[tile * kW + w];
    const int L = (slice_off[tile * kW + w + 1] - off) >> 5;
    const T *vp = val + (size_t)off + lane;
    const int *cp = col + (size_t)off + lane;
    const uint64_t pol = policy_evict_last();
    UChunk<T, U> a, b;
    sellu_load<T, U>(a, vp, cp, 0, L);
    sellu_load<T, U>(b, vp, cp, U, L);
    __syncwarp();                              /* rid / sl written by this warp's lanes */

    int g = 0;
    int slot_end = sl[w * G];
    T acc = (T)0;
    auto consume = [&](const UChunk<T, U> &ch, int s) {
        T xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (s + u < L && ch.c[u] != 0) xv[u] = ld_x(xm1 + ch.c[u], pol);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (s + u < L) {
                while (s + u == slot_end) {                     /* warp-uniform: slot g is complete */
                    ystage[rid[g * kT + tid]] = acc;
                    acc = (T)0;
                    ++g;
                    slot_end += sl[w * G + g];                  /* g < G here: entries remain */
                }
                if (ch.c[u] != 0) acc = uadd(acc, umul(ch.v[u], xv[u]));
            }
        }
    };
    for (int s = 0; s < L; s += 2 * U) {
        consume(a, s);
        sellu_load<T, U>(a, vp, cp, s + 2 * U, L);
        consume(b, s + U);
        sellu_load<T, U>(b, vp, cp, s + 3 * U, L);
    }
    for (; g < G; ++g) {                                        /* the last slot, and empty ones after it */
        ystage[rid[g * kT + tid]] = acc;
        acc = (T)0;
    }
    __syncthreads();
    for (int i = tid; i < R; i += kT) {
        const int r = tile * R + i;
        if (r < rows) y[r] = ystage[i];
    }
}

}  // namespace

int sellu_threads() { return kT; }

void launch_sellu_sort(const uint16_t *seglen, int ntiles, int R, int G, uint16_t *rowids,
                       uint16_t *slotlen, int *slice_elems, cudaStream_t s)
{
    if (ntiles <= 0) return;
    int NK = 1024;
    while (NK < R) NK <<= 1;
    sellu_sort_kernel<<<ntiles, 1024, NK * sizeof(uint32_t), s>>>(seglen, R, G, NK, rowids, slotlen, slice_elems);
}

template <typename T>
void launch_sellu_fill(const T *val, const int *col, const int *rowptr, int rows, const DevSell &sm, int cap,
                       T *val_out, int *col_out, cudaStream_t s)
{
    const int nslices = sm.nblk * kW;
    if (nslices <= 0) return;
    const long long threads = (long long)nslices * 32;
    sellu_fill_kernel<T><<<(int)((threads + 255) / 256), 256, 0, s>>>(
        val, col, rowptr, rows, sm.R, sm.G, cap, sm.rowids, sm.slotlen, sm.slice_off, nslices, val_out, col_out);
}
template void launch_sellu_fill<double>(const double *, const int *, const int *, int, const DevSell &, int,
                                        double *, int *, cudaStream_t);
template void launch_sellu_fill<float>(const float *, const int *, const int *, int, const DevSell &, int,
                                       float *, int *, cudaStream_t);

template <typename T>
void launch_sellu(const DevSell &sm, const T *x, T *y, cudaStream_t s)
{
    if (sm.nblk <= 0) return;
    const size_t smem = (size_t)sm.R * sizeof(T) + ((size_t)sm.G * kT + (size_t)kW * sm.G) * sizeof(uint16_t);
    if (sm.U >= 8)
        spmv_sellu_kernel<T, 8><<<sm.nblk, kT, smem, s>>>(static_cast<const T *>(sm.val), sm.col, sm.rowids,
                                                          sm.slotlen, sm.slice_off, x - 1, y, sm.rows, sm.R, sm.G);
    else if (sm.U >= 6)
        spmv_sellu_kernel<T, 6><<<sm.nblk, kT, smem, s>>>(static_cast<const T *>(sm.val), sm.col, sm.rowids,
                                                          sm.slotlen, sm.slice_off, x - 1, y, sm.rows, sm.R, sm.G);
    else
        spmv_sellu_kernel<T, 4><<<sm.nblk, kT, smem, s>>>(static_cast<const T *>(sm.val), sm.col, sm.rowids,
                                                          sm.slotlen, sm.slice_off, x - 1, y, sm.rows, sm.R, sm.G);
}
template void launch_sellu<double>(const DevSell &, const double *, double *, cudaStream_t);
template void launch_sellu<float>(const DevSell &, const float *, float *, cudaStream_t);

}  // namespace b200
