/*
 * npb_gen.cu -- the NPB CG test matrix assembled ON THE DEVICE (include/b200_npb.h;
 * SURVEY.md section 8f row 2).
 *
 * NPB3.3.1/CG/cg.f builds A = sum_i size_i v_i v_i^T (+ rcond - shift on the diagonal)
 * from n random sparse vectors v_i:
 *   makea :650-735   draws the vectors (sprnvc :911-965, vecset :991-1019) from one
 *                    sequential random stream and calls
 *   sparse :740-905  which inserts every triple (row, column, value) in generation order
 *                    into sorted rows and ADDS a duplicate's value to the slot that is
 *                    already there (:821-871).
 * The random stream is inherently sequential and cheap (the host draws the vectors:
 * callers/npb/makea.c, ~0.5 s for class D); the expensive part -- 726 M triples for
 * class D, 6.6 G for class E -- is embarrassingly parallel BY ROW once one knows which
 * vectors hit a row.  Here:
 *   1. count, per row of the block, the vectors that contain it (atomics), prefix sums;
 *   2. fill and sort every row's hit list by vector number (= generation order);
 *   3. one CTA per row: re-create the row's triples in shared memory in generation
 *      order, bitonic-sort the keys (column, arrival number), and add runs of equal
 *      columns in arrival order starting from 0.0 -- the very additions sparse performs,
 *      in the same order => the same bits as the Fortran / the host generator.
 *      Run twice: once to count the distinct columns of every row, once to write.
 * Triples are never stored: a class D row block costs its CSR plus ~0.5 GB of vectors.
 * Products are rounded separately (__dmul_rn / __dadd_rn: the reference build has no FMA).
 */
#include "../../include/b200_npb.h"

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

namespace {

#define GEN_OK(call)                                                           \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            fprintf(stderr, "libb200-spmv: npb generator: %s failed at %s:%d: %s\n", #call,       \
                    __FILE__, __LINE__, cudaGetErrorString(e_));               \
            return -4;                                                         \
        }                                                                      \
    } while (0)

constexpr int kGenThreads = 128;

/* pass 1: how many vectors hit each row of the block, and how many triples that makes */
__global__ void gen_count_kernel(const int *__restrict__ arow, const int *__restrict__ acol, int n, int ld,
                                 int row_lo, int row_hi, int *__restrict__ hits, int *__restrict__ trips)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * ld) return;
    const int i = (int)(e / ld), nza = (int)(e - (long long)i * ld);
    const int cnt = arow[i];
    if (nza >= cnt) return;
    const int j = acol[e] - 1;
    if (j < row_lo || j >= row_hi) return;
    atomicAdd(hits + (j - row_lo), 1);
    atomicAdd(trips + (j - row_lo), cnt);
}

__global__ void gen_fill_kernel(const int *__restrict__ arow, const int *__restrict__ acol, int n, int ld,
                                int row_lo, int row_hi, const int *__restrict__ hstart,
                                int *__restrict__ fill, int *__restrict__ list)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n * ld) return;
    const int i = (int)(e / ld), nza = (int)(e - (long long)i * ld);
    if (nza >= arow[i]) return;
    const int j = acol[e] - 1;
    if (j < row_lo || j >= row_hi) return;
    const int slot = atomicAdd(fill + (j - row_lo), 1);
    list[hstart[j - row_lo] + slot] = (int)e;            /* entry id = i * ld + nza */
}

/* generation order = ascending vector number = ascending entry id (a vector holds a row once) */
__global__ void gen_sort_hits_kernel(const int *__restrict__ hstart, int nrows, int *__restrict__ list)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    int *l = list + hstart[r];
    const int m = hstart[r + 1] - hstart[r];
    for (int a = 1; a < m; ++a) {
        const int v = l[a];
        int b = a - 1;
        while (b >= 0 && l[b] > v) { l[b + 1] = l[b]; --b; }
        l[b + 1] = v;
    }
}

/* one CTA per row.  WRITE == 0: rownnz[row] = distinct columns.  WRITE == 1: the row's
 * (column, value) pairs go to colidx / a at rowoff[row]. */
template <int CAP, int WRITE>
__global__ void __launch_bounds__(kGenThreads)
gen_rows_kernel(const int *__restrict__ arow, const int *__restrict__ acol, const double *__restrict__ aelt,
                const double *__restrict__ size, int ld, double rcond, double shift, int row_lo, int nrows,
                const int *__restrict__ hstart, const int *__restrict__ list,
                int *__restrict__ rownnz, const long long *__restrict__ rowoff,
                int *__restrict__ colidx, double *__restrict__ a, int *__restrict__ overflow)
{
    __shared__ unsigned long long key[CAP];       /* column << 16 | arrival number */
    __shared__ double val[CAP];                   /* indexed by arrival number */
    __shared__ int hbase[130];                    /* arrival number of each hit's first triple */
    __shared__ int warp_cnt[kGenThreads / 32 + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kGenThreads / 32;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int h0 = hstart[row], nh = hstart[row + 1] - h0;
        if (nh > 128) { if (tid == 0) atomicExch(overflow, 1); continue; }      /* block-uniform */
        if (tid == 0) {
            int run = 0;
            for (int k = 0; k < nh; ++k) {
                hbase[k] = run;
                run += arow[list[h0 + k] / ld];
            }
            hbase[nh] = run;
        }
        __syncthreads();
        const int T = hbase[nh];
        if (T > CAP) { if (tid == 0) atomicExch(overflow, 1); __syncthreads(); continue; }
        const int j1 = row_lo + row + 1;                   /* 1-based row */
        /* a warp per hit: lane nzrow re-creates one triple (cg.f:809-876) */
        for (int k = warp; k < nh; k += NW) {
            const int e = list[h0 + k];
            const int i = e / ld;
            const int cnt = arow[i];
            const double scale = __dmul_rn(size[i], aelt[e]);
            if (lane < cnt) {
                const long long src = (long long)i * ld + lane;
                const int jcol = acol[src];
                double va = __dmul_rn(aelt[src], scale);
                if (jcol == j1 && j1 == i + 1) va = __dadd_rn(__dadd_rn(va, rcond), -shift);   /* :826-828 */
                const int arr = hbase[k] + lane;
                key[arr] = ((unsigned long long)(unsigned)jcol << 16) | (unsigned long long)arr;
                val[arr] = va;
            }
        }
        for (int f = T + tid; f < CAP; f += kGenThreads) key[f] = ~0ull;
        __syncthreads();
        /* bitonic sort, ascending, over the smallest power of two >= T */
        int np2 = 32;
        while (np2 < T) np2 <<= 1;
        for (int kk = 2; kk <= np2; kk <<= 1) {
            for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                for (int f = tid; f < np2; f += kGenThreads) {
                    const int g = f ^ jj;
                    if (g > f) {
                        const unsigned long long x = key[f], y = key[g];
                        const bool up = (f & kk) == 0;
                        if (up ? (x > y) : (x < y)) { key[f] = y; key[g] = x; }
                    }
                }
                __syncthreads();
            }
        }
        /* runs of equal columns: every thread owns a contiguous chunk of the sorted keys */
        const int per = (T + kGenThreads - 1) / kGenThreads;
        const int f0 = min(T, tid * per), f1 = min(T, f0 + per);
        int heads = 0;
        for (int f = f0; f < f1; ++f)
            heads += (f == 0 || (key[f] >> 16) != (key[f - 1] >> 16)) ? 1 : 0;
        /* exclusive scan of the per-thread head counts */
        int incl = heads;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_cnt[warp] = incl;
        __syncthreads();
        int before = incl - heads;
        for (int w = 0; w < warp; ++w) before += warp_cnt[w];
        int total = 0;
        for (int w = 0; w < NW; ++w) total += warp_cnt[w];
        if (WRITE == 0) {
            if (tid == 0) rownnz[row] = total;
        } else {
            const long long out0 = rowoff[row];
            int rank = before;
            for (int f = f0; f < f1; ++f) {
                const unsigned col = (unsigned)(key[f] >> 16);
                if (f == 0 || (unsigned)(key[f - 1] >> 16) != col) {
                    double acc = 0.0;                                        /* cg.f:800-803, 846 */
                    for (int g = f; g < T && (unsigned)(key[g] >> 16) == col; ++g)
                        acc = __dadd_rn(acc, val[(int)(key[g] & 0xFFFFull)]);  /* cg.f:869 */
                    colidx[out0 + rank] = (int)col;
                    a[out0 + rank] = acc;
                    ++rank;
                }
            }
        }
        __syncthreads();               /* key / val / hbase / warp_cnt are reused by the next row */
    }
}

__global__ void gen_rowstr_kernel(const long long *__restrict__ rowoff, int nrows, int *__restrict__ rowstr)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= nrows) rowstr[r] = (int)(rowoff[r] + 1);                       /* 1-based */
}

template <int CAP>
int run_rows(int write, const int *arow, const int *acol, const double *aelt, const double *size, int ld,
             double rcond, double shift, int row_lo, int nrows, const int *hstart, const int *list,
             int *rownnz, const long long *rowoff, int *colidx, double *a, int *overflow)
{
    const int grid = nrows < 148 * 64 ? nrows : 148 * 64;
    if (grid <= 0) return 0;
    if (write)
        gen_rows_kernel<CAP, 1><<<grid, kGenThreads>>>(arow, acol, aelt, size, ld, rcond, shift, row_lo, nrows,
                                                       hstart, list, rownnz, rowoff, colidx, a, overflow);
    else
        gen_rows_kernel<CAP, 0><<<grid, kGenThreads>>>(arow, acol, aelt, size, ld, rcond, shift, row_lo, nrows,
                                                       hstart, list, rownnz, rowoff, colidx, a, overflow);
    return 0;
}

}  // namespace

extern "C" void b200_npb_csr_free(b200_npb_csr *m)
{
    if (!m) return;
    cudaFree(m->d_rowstr); cudaFree(m->d_colidx); cudaFree(m->d_a);
    m->d_rowstr = nullptr; m->d_colidx = nullptr; m->d_a = nullptr;
    m->rows = 0; m->nnz = 0;
}

extern "C" int b200_npb_csr_to_host(const b200_npb_csr *m, int *rowstr, int *colidx, double *a)
{
    if (!m || !m->d_rowstr) return -2;
    GEN_OK(cudaMemcpy(rowstr, m->d_rowstr, sizeof(int) * ((size_t)m->rows + 1), cudaMemcpyDeviceToHost));
    if (m->nnz > 0) {
        GEN_OK(cudaMemcpy(colidx, m->d_colidx, sizeof(int) * (size_t)m->nnz, cudaMemcpyDeviceToHost));
        GEN_OK(cudaMemcpy(a, m->d_a, sizeof(double) * (size_t)m->nnz, cudaMemcpyDeviceToHost));
    }
    return 0;
}

extern "C" int b200_npb_makea_device(int na, int ld, const int *arow, const int *acol, const double *aelt,
                                     const double *size, double rcond, double shift, int row_lo, int row_hi,
                                     b200_npb_csr *out)
{
    if (!out || row_lo < 0 || row_hi > na || row_hi < row_lo || ld < 1 || ld > 32) return -2;
    out->rows = row_hi - row_lo; out->nnz = 0;
    out->d_rowstr = nullptr; out->d_colidx = nullptr; out->d_a = nullptr;
    const int nrows = row_hi - row_lo;
    const long long nent = (long long)na * ld;
    if (nent > 0x7fffffffLL) return -2;
    int *d_arow = nullptr, *d_acol = nullptr, *d_hits = nullptr, *d_trips = nullptr, *d_hstart = nullptr,
        *d_list = nullptr, *d_rownnz = nullptr, *d_overflow = nullptr;
    double *d_aelt = nullptr, *d_size = nullptr;
    long long *d_rowoff = nullptr;
    int rc = 0;
    /* everything on the default stream of the current device; a one-off setup step */
    GEN_OK(cudaMalloc((void **)&d_arow, sizeof(int) * (size_t)na));
    GEN_OK(cudaMalloc((void **)&d_acol, sizeof(int) * (size_t)nent));
    GEN_OK(cudaMalloc((void **)&d_aelt, sizeof(double) * (size_t)nent));
    GEN_OK(cudaMalloc((void **)&d_size, sizeof(double) * (size_t)na));
    GEN_OK(cudaMemcpy(d_arow, arow, sizeof(int) * (size_t)na, cudaMemcpyHostToDevice));
    GEN_OK(cudaMemcpy(d_acol, acol, sizeof(int) * (size_t)nent, cudaMemcpyHostToDevice));
    GEN_OK(cudaMemcpy(d_aelt, aelt, sizeof(double) * (size_t)nent, cudaMemcpyHostToDevice));
    GEN_OK(cudaMemcpy(d_size, size, sizeof(double) * (size_t)na, cudaMemcpyHostToDevice));
    GEN_OK(cudaMalloc((void **)&d_hits, sizeof(int) * ((size_t)nrows + 1)));
    GEN_OK(cudaMalloc((void **)&d_trips, sizeof(int) * ((size_t)nrows + 1)));
    GEN_OK(cudaMemset(d_hits, 0, sizeof(int) * ((size_t)nrows + 1)));
    GEN_OK(cudaMemset(d_trips, 0, sizeof(int) * ((size_t)nrows + 1)));
    GEN_OK(cudaMalloc((void **)&d_overflow, sizeof(int)));
    GEN_OK(cudaMemset(d_overflow, 0, sizeof(int)));
    const int eblocks = (int)((nent + 255) / 256);
    gen_count_kernel<<<eblocks, 256>>>(d_arow, d_acol, na, ld, row_lo, row_hi, d_hits, d_trips);
    GEN_OK(cudaGetLastError());
    /* prefix sums on the host: a few MB */
    std::vector<int> hits((size_t)nrows + 1), trips((size_t)nrows + 1);
    GEN_OK(cudaMemcpy(hits.data(), d_hits, sizeof(int) * (size_t)nrows, cudaMemcpyDeviceToHost));
    GEN_OK(cudaMemcpy(trips.data(), d_trips, sizeof(int) * (size_t)nrows, cudaMemcpyDeviceToHost));
    long long run = 0;
    int max_trips = 0, max_hits = 0;
    for (int r = 0; r < nrows; ++r) {
        const int h = hits[r];
        if (h > max_hits) max_hits = h;
        if (trips[r] > max_trips) max_trips = trips[r];
        hits[r] = (int)run;
        run += h;
    }
    hits[nrows] = (int)run;
    if (run > 0x7fffffffLL || max_hits > 128 || max_trips > 2048) rc = -1;
    if (rc == 0) {
        GEN_OK(cudaMalloc((void **)&d_hstart, sizeof(int) * ((size_t)nrows + 1)));
        GEN_OK(cudaMemcpy(d_hstart, hits.data(), sizeof(int) * ((size_t)nrows + 1), cudaMemcpyHostToDevice));
        GEN_OK(cudaMalloc((void **)&d_list, sizeof(int) * (size_t)(run > 0 ? run : 1)));
        GEN_OK(cudaMemset(d_hits, 0, sizeof(int) * ((size_t)nrows + 1)));           /* reused as fill cursors */
        gen_fill_kernel<<<eblocks, 256>>>(d_arow, d_acol, na, ld, row_lo, row_hi, d_hstart, d_hits, d_list);
        if (nrows > 0) gen_sort_hits_kernel<<<(nrows + 127) / 128, 128>>>(d_hstart, nrows, d_list);
        GEN_OK(cudaGetLastError());
        GEN_OK(cudaMalloc((void **)&d_rownnz, sizeof(int) * ((size_t)nrows + 1)));
        GEN_OK(cudaMalloc((void **)&d_rowoff, sizeof(long long) * ((size_t)nrows + 1)));
        auto rows_pass = [&](int write, int *colidx, double *a) {
            if (max_trips <= 1024)
                return run_rows<1024>(write, d_arow, d_acol, d_aelt, d_size, ld, rcond, shift, row_lo, nrows,
                                      d_hstart, d_list, d_rownnz, d_rowoff, colidx, a, d_overflow);
            return run_rows<2048>(write, d_arow, d_acol, d_aelt, d_size, ld, rcond, shift, row_lo, nrows,
                                  d_hstart, d_list, d_rownnz, d_rowoff, colidx, a, d_overflow);
        };
        rows_pass(0, nullptr, nullptr);
        GEN_OK(cudaGetLastError());
        std::vector<int> rownnz((size_t)nrows + 1);
        GEN_OK(cudaMemcpy(rownnz.data(), d_rownnz, sizeof(int) * (size_t)nrows, cudaMemcpyDeviceToHost));
        std::vector<long long> rowoff((size_t)nrows + 1);
        long long nnz = 0;
        for (int r = 0; r < nrows; ++r) { rowoff[r] = nnz; nnz += rownnz[r]; }
        rowoff[nrows] = nnz;
        int overflow = 0;
        GEN_OK(cudaMemcpy(&overflow, d_overflow, sizeof(int), cudaMemcpyDeviceToHost));
        if (overflow || nnz + 1 > 0x7fffffffLL) rc = -1;            /* breaks the int32 ABI: use more row blocks */
        if (rc == 0) {
            GEN_OK(cudaMemcpy(d_rowoff, rowoff.data(), sizeof(long long) * ((size_t)nrows + 1), cudaMemcpyHostToDevice));
            GEN_OK(cudaMalloc((void **)&out->d_rowstr, sizeof(int) * ((size_t)nrows + 1)));
            GEN_OK(cudaMalloc((void **)&out->d_colidx, sizeof(int) * (size_t)(nnz > 0 ? nnz : 1)));
            GEN_OK(cudaMalloc((void **)&out->d_a, sizeof(double) * (size_t)(nnz > 0 ? nnz : 1)));
            rows_pass(1, out->d_colidx, out->d_a);
            gen_rowstr_kernel<<<(nrows + 256) / 256, 256>>>(d_rowoff, nrows, out->d_rowstr);
            GEN_OK(cudaGetLastError());
            GEN_OK(cudaDeviceSynchronize());
            out->nnz = nnz;
        }
    }
    cudaFree(d_arow); cudaFree(d_acol); cudaFree(d_aelt); cudaFree(d_size); cudaFree(d_hits); cudaFree(d_trips);
    cudaFree(d_hstart); cudaFree(d_list); cudaFree(d_rownnz); cudaFree(d_rowoff); cudaFree(d_overflow);
    if (rc != 0) b200_npb_csr_free(out);
    return rc;
}
