/*
 * cg_peer.cu -- fused compute + exchange kernels over peer memory
 * (include/b200_peer.h).  The algebra is conj_grad of NPB3.3.1/CG/cg.f:447-644;
 * the exchange replaces what an MPI / NCCL version would do with an allgather
 * of p and two allreduces per CG iteration.
 *
 * Memory model: data stores to peers are plain st.global; after the CTA barrier
 * that follows a block's stores ONE thread executes __threadfence_system() (the
 * barrier makes the fence cumulative over the block's stores) and bumps a local
 * counter; the block that finishes last fences once more and publishes the epoch
 * to every rank with relaxed system-scope stores (fence + relaxed store = release;
 * a st.release.sys per rank would pay the fence -- an NVLink round trip -- once
 * per rank).  Consumers spin with ld.acquire.sys.  One rank per GPU, so a
 * spinning kernel never waits on work queued behind it on its own device.
 */
#include "../../include/b200_peer.h"
#include "spmv_kernels.cuh"

#include <cuda_runtime.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int kBlocks = 296;
constexpr int kThreads = 256;
static_assert(kBlocks <= 2 * kThreads, "publish_scalar reduces two partials per thread at most");
constexpr int NR = B200_PEER_MAX_RANKS;

#define PEER_OK(call)                                                          \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            fprintf(stderr, "libb200-spmv: fatal: %s failed at %s:%d: %s\n", #call, __FILE__,     \
                    __LINE__, cudaGetErrorString(e_));                         \
            abort();                                                           \
        }                                                                      \
    } while (0)

/* device view of the group: pointers into every rank's segment */
struct PeerDev {
    int rank, nranks;
    double *xfull[NR];
    double *xalt[NR];                   /* second vector buffer: odd epochs of b200_peer_post */
    double *scal[NR];                   /* [B200_PEER_SLOTS][NR] */
    unsigned long long *sflag[NR];      /* [B200_PEER_SLOTS][NR] */
    unsigned long long *vflag[NR];      /* [NR] vector published */
    unsigned long long *rflag[NR];      /* [NR] vector consumed (safe to overwrite) */
    unsigned int *counter;              /* local: last-block detection, one per kernel kind */
    double *partial;                    /* local: kBlocks doubles */
};

/* release = ONE system-scope fence, then relaxed stores: a st.release.sys per destination
 * repeats the fence -- a round trip over NVLink each -- once per rank (the exchange kernel took
 * 50 us on 4 GPUs against 11 us on one, profiles/r02_run20_step_breakdown_n4.txt) */
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double red[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += red[w];
    }
    __syncthreads();
    return s;
}

/* true in every thread of the block that arrives last.  One thread per block fences at system
 * scope after the CTA barrier (causality is cumulative: the barrier orders the other threads'
 * stores before it), not every thread; the last block fences once more (acquire side of the
 * counter, release side of whatever it publishes next with relaxed stores). */
__device__ __forceinline__ bool last_block(unsigned int *counter)
{
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int prev = atomicAdd(counter, 1u);
        last = prev == gridDim.x - 1;
        if (last) {
            *counter = 0;                        /* ready for the next launch */
            __threadfence_system();
        }
    }
    __syncthreads();
    return last;
}

/* spin until every rank published >= e in my flag row */
__device__ __forceinline__ void wait_flags(const unsigned long long *flags, int nranks,
                                           unsigned long long e)
{
    if (threadIdx.x < nranks)
        while (ld_acquire_sys(flags + threadIdx.x) < e) { }
    __syncthreads();
}

/* rank-ordered sum of a scalar slot in my segment (same bits on every rank) */
__device__ __forceinline__ double slot_sum(const PeerDev &g, int slot)
{
    const double *s = g.scal[g.rank] + slot * NR;
    double t = 0.0;
    for (int r = 0; r < g.nranks; ++r) t += s[r];
    return t;
}

/* the last block reduces the block partials in a fixed order -- pairs kThreads apart, an
 * xor-shuffle tree, the warp sums left to right; the whole block takes part, one thread
 * adding 296 values cost ~6 us per scalar -- and publishes the sum to every rank */
__device__ __forceinline__ void publish_scalar(const PeerDev &g, int slot, unsigned long long e,
                                               unsigned int *counter)
{
    __shared__ double wsum[kThreads / 32];
    if (last_block(counter)) {                           /* block-uniform */
        const int t = threadIdx.x, nb = (int)gridDim.x;  /* nb <= 2 * kThreads */
        double v = t < nb ? g.partial[t] : 0.0;
        if (t + kThreads < nb) v += g.partial[t + kThreads];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((t & 31) == 0) wsum[t >> 5] = v;
        __syncthreads();
        if (t == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) tot += wsum[w];
            for (int j = 0; j < g.nranks; ++j) g.scal[j][slot * NR + g.rank] = tot;
            __threadfence_system();
            for (int j = 0; j < g.nranks; ++j) st_relaxed_sys(g.sflag[j] + slot * NR + g.rank, e);
        }
    }
}

/* dst[j][lo + 2 i .. 2 i + 1] = v2[i] for every rank j, grid-stride: the loads of four
 * elements are issued before their stores, so a thread pays the load latency once per four
 * elements, not once per element (a load-store-load-store loop serialises: the compiler must
 * assume the peer buffers alias the source) */
__device__ __forceinline__ void push_slice_vec(const double2 *__restrict__ v2, int n2, double *const *dst,
                                               long long lo, int nranks)
{
    const int stride = gridDim.x * kThreads;
    for (int base = blockIdx.x * kThreads + threadIdx.x; base < n2; base += 4 * stride) {
        double2 val[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (base + q * stride < n2) val[q] = v2[base + q * stride];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (base + q * stride < n2)
                for (int j = 0; j < nranks; ++j)
                    reinterpret_cast<double2 *>(dst[j] + lo)[base + q * stride] = val[q];
    }
}

__global__ void __launch_bounds__(kThreads)
peer_push_kernel(PeerDev g, const double *__restrict__ v, int n, long long lo, unsigned long long e,
                 unsigned long long e_consumed)
{
    /* do not overwrite a peer's vector before it has finished reading the old one */
    if (e_consumed) wait_flags(g.rflag[g.rank], g.nranks, e_consumed);
    /* 16-byte stores over NVLink when the slice is 16-byte aligned on both sides */
    const bool vec_ok = ((lo & 1) == 0) && ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
    const int n2 = vec_ok ? n >> 1 : 0;
    const double2 *v2 = reinterpret_cast<const double2 *>(v);
    push_slice_vec(v2, n2, g.xfull, lo, g.nranks);
    for (int i = 2 * n2 + blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const double val = v[i];
        for (int j = 0; j < g.nranks; ++j) g.xfull[j][lo + i] = val;
    }
    if (last_block(g.counter + 0) && threadIdx.x == 0)       /* fenced by last_block */
        for (int j = 0; j < g.nranks; ++j) st_relaxed_sys(g.vflag[j] + g.rank, e);
}

/* One launch per product for back-to-back products (b200_peer_exchange): report the
 * previous vector (epoch e - 1) as consumed -- the product that read it precedes this
 * kernel in the stream --, wait until every rank has done so, push the new slice,
 * publish epoch e, and let the last block wait for everybody's epoch e, so that the
 * kernel only retires once the whole vector is in this rank's buffer. */
__global__ void __launch_bounds__(kThreads)
peer_exchange_kernel(PeerDev g, const double *__restrict__ v, int n, long long lo, unsigned long long e)
{
    if (e > 1) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            __threadfence_system();
            for (int j = 0; j < g.nranks; ++j) st_relaxed_sys(g.rflag[j] + g.rank, e - 1);
        }
        wait_flags(g.rflag[g.rank], g.nranks, e - 1);
    }
    const bool vec_ok = ((lo & 1) == 0) && ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
    const int n2 = vec_ok ? n >> 1 : 0;
    const double2 *v2 = reinterpret_cast<const double2 *>(v);
    push_slice_vec(v2, n2, g.xfull, lo, g.nranks);
    for (int i = 2 * n2 + blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const double val = v[i];
        for (int j = 0; j < g.nranks; ++j) g.xfull[j][lo + i] = val;
    }
    if (last_block(g.counter + 0)) {
        if (threadIdx.x == 0)                                /* fenced by last_block */
            for (int j = 0; j < g.nranks; ++j) st_relaxed_sys(g.vflag[j] + g.rank, e);
        wait_flags(g.vflag[g.rank], g.nranks, e);
    }
}

/* b200_peer_post: the exchange half of a product that waits for its x slices itself
 * (b200_spmv_exec_sliced).  Epoch e goes to buffer e & 1, which the products of epoch
 * e - 2 read last: report e - 1 as consumed, wait for every rank's report of e - 2 --
 * one whole step old, so this hardly ever spins and a slow rank does not hold the
 * others back --, push the slice, publish e.  Nobody waits for arrivals here. */
__global__ void __launch_bounds__(kThreads)
peer_post_kernel(PeerDev g, const double *__restrict__ v, int n, long long lo, unsigned long long e)
{
    if (e > 1 && blockIdx.x == 0 && threadIdx.x == 0) {
        __threadfence_system();
        for (int j = 0; j < g.nranks; ++j) st_relaxed_sys(g.rflag[j] + g.rank, e - 1);
    }
    if (e > 2) wait_flags(g.rflag[g.rank], g.nranks, e - 2);
    double *const *dstv = (e & 1) ? g.xalt : g.xfull;
    const bool vec_ok = ((lo & 1) == 0) && ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
    const int n2 = vec_ok ? n >> 1 : 0;
    const double2 *v2 = reinterpret_cast<const double2 *>(v);
    /* four loads in flight per thread before the stores (see push_slice_vec) */
    push_slice_vec(v2, n2, dstv, lo, g.nranks);
    for (int i = 2 * n2 + blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const double val = v[i];
        for (int j = 0; j < g.nranks; ++j) dstv[j][lo + i] = val;
    }
    if (last_block(g.counter + 4) && threadIdx.x == 0)       /* fenced by last_block */
        for (int j = 0; j < g.nranks; ++j) st_relaxed_sys(g.vflag[j] + g.rank, e);
}

__global__ void peer_wait_vector_kernel(PeerDev g, unsigned long long e)
{
    wait_flags(g.vflag[g.rank], g.nranks, e);
}

/* tell every rank that this rank has finished reading vector epoch e */
__global__ void peer_consumed_kernel(PeerDev g, unsigned long long e)
{
    if (threadIdx.x == 0) {
        __threadfence_system();
        for (int j = 0; j < g.nranks; ++j) st_relaxed_sys(g.rflag[j] + g.rank, e);
    }
}

__global__ void __launch_bounds__(kThreads)
peer_dot_kernel(PeerDev g, const double *__restrict__ x, const double *__restrict__ y, int n, int mode,
                int slot, unsigned long long e)
{
    double acc = 0.0;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        if (mode == 0) acc += x[i] * y[i];
        else { const double d = x[i] - y[i]; acc += d * d; }
    }
    const double s = block_sum(acc);
    if (threadIdx.x == 0) g.partial[blockIdx.x] = s;
    publish_scalar(g, slot, e, g.counter + 1);
}

__global__ void __launch_bounds__(kThreads)
peer_update_zr_kernel(PeerDev g, double *z, double *r, const double *__restrict__ p,
                      const double *__restrict__ q, int n, int slot_rho, int slot_d,
                      unsigned long long e_d, int slot_out, unsigned long long e_out)
{
    wait_flags(g.sflag[g.rank] + slot_d * NR, g.nranks, e_d);
    const double alpha = slot_sum(g, slot_rho) / slot_sum(g, slot_d);          /* cg.f:581 */
    double acc = 0.0;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        z[i] = z[i] + alpha * p[i];                                            /* cg.f:593-596 */
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        acc += ri * ri;                                                        /* cg.f:602-604 */
    }
    const double s = block_sum(acc);
    if (threadIdx.x == 0) g.partial[blockIdx.x] = s;
    publish_scalar(g, slot_out, e_out, g.counter + 2);
}

__global__ void __launch_bounds__(kThreads)
peer_update_p_kernel(PeerDev g, double *p, const double *__restrict__ r, int n, long long lo,
                     int slot_new, unsigned long long e_new, int slot_old, unsigned long long e_vec)
{
    wait_flags(g.sflag[g.rank] + slot_new * NR, g.nranks, e_new);
    const double beta = slot_sum(g, slot_new) / slot_sum(g, slot_old);         /* cg.f:609 */
    const bool vec_ok = ((lo & 1) == 0) && (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(r)) & 15) == 0);
    const int n2 = vec_ok ? n >> 1 : 0;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n2; i += gridDim.x * kThreads) {
        const double2 rv = reinterpret_cast<const double2 *>(r)[i];
        double2 pv = reinterpret_cast<double2 *>(p)[i];
        pv.x = rv.x + beta * pv.x;                                             /* cg.f:614-616 */
        pv.y = rv.y + beta * pv.y;
        reinterpret_cast<double2 *>(p)[i] = pv;
        /* the exchange: every rank's copy of the full vector gets the elements now */
        for (int j = 0; j < g.nranks; ++j) reinterpret_cast<double2 *>(g.xfull[j] + lo)[i] = pv;
    }
    for (int i = 2 * n2 + blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const double pi = r[i] + beta * p[i];
        p[i] = pi;
        for (int j = 0; j < g.nranks; ++j) g.xfull[j][lo + i] = pi;
    }
    if (last_block(g.counter + 3) && threadIdx.x == 0)       /* fenced by last_block */
        for (int j = 0; j < g.nranks; ++j) st_relaxed_sys(g.vflag[j] + g.rank, e_vec);
}

__global__ void __launch_bounds__(kThreads)
peer_scale_kernel(PeerDev g, double *x, const double *__restrict__ z, int n, int slot,
                  unsigned long long e)
{
    wait_flags(g.sflag[g.rank] + slot * NR, g.nranks, e);
    const double s = 1.0 / sqrt(slot_sum(g, slot));                            /* cg.f:322, 344-346 */
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) x[i] = s * z[i];
}

struct SlotList { int slot[B200_PEER_SLOTS]; unsigned long long epoch[B200_PEER_SLOTS]; int count; };

__global__ void peer_read_slots_kernel(PeerDev g, SlotList l, double *out)
{
    for (int k = 0; k < l.count; ++k) {
        wait_flags(g.sflag[g.rank] + l.slot[k] * NR, g.nranks, l.epoch[k]);
        if (threadIdx.x == 0) out[k] = slot_sum(g, l.slot[k]);
        __syncthreads();
    }
}

}  // namespace

struct b200_peer_group {
    int rank, nranks;
    int64_t n_global;
    size_t bytes;
    char *local;                 /* my segment */
    char *peer[NR];              /* mapped segments (peer[rank] == local) */
    unsigned int *counter;
    double *partial;
    PeerDev dev;
    bool connected;
};

static size_t seg_x_bytes(int64_t n) { return (((size_t)n + 2) * sizeof(double) + 255) & ~(size_t)255; }
static size_t seg_scal_off(int64_t n) { return 2 * seg_x_bytes(n); }      /* two vector buffers */
static size_t seg_sflag_off(int64_t n) { return seg_scal_off(n) + B200_PEER_SLOTS * NR * sizeof(double); }
static size_t seg_vflag_off(int64_t n) { return seg_sflag_off(n) + B200_PEER_SLOTS * NR * sizeof(unsigned long long); }
static size_t seg_rflag_off(int64_t n) { return seg_vflag_off(n) + NR * sizeof(unsigned long long); }
static size_t seg_total(int64_t n) { return seg_rflag_off(n) + NR * sizeof(unsigned long long) + 256; }

static void fill_dev(b200_peer_group *g)
{
    PeerDev &d = g->dev;
    d.rank = g->rank;
    d.nranks = g->nranks;
    for (int j = 0; j < g->nranks; ++j) {
        d.xfull[j] = (double *)g->peer[j];
        d.xalt[j] = (double *)(g->peer[j] + seg_x_bytes(g->n_global));
        d.scal[j] = (double *)(g->peer[j] + seg_scal_off(g->n_global));
        d.sflag[j] = (unsigned long long *)(g->peer[j] + seg_sflag_off(g->n_global));
        d.vflag[j] = (unsigned long long *)(g->peer[j] + seg_vflag_off(g->n_global));
        d.rflag[j] = (unsigned long long *)(g->peer[j] + seg_rflag_off(g->n_global));
    }
    d.counter = g->counter;
    d.partial = g->partial;
}

extern "C" b200_peer_group *b200_peer_create(int rank, int nranks, int64_t n_global, void *ipc_handle_out)
{
    if (nranks < 1 || nranks > NR || rank < 0 || rank >= nranks) return nullptr;
    b200_peer_group *g = (b200_peer_group *)calloc(1, sizeof *g);
    g->rank = rank; g->nranks = nranks; g->n_global = n_global;
    g->bytes = seg_total(n_global);
    PEER_OK(cudaMalloc((void **)&g->local, g->bytes));
    PEER_OK(cudaMemset(g->local, 0, g->bytes));
    PEER_OK(cudaMalloc((void **)&g->counter, 8 * sizeof(unsigned int)));
    PEER_OK(cudaMemset(g->counter, 0, 8 * sizeof(unsigned int)));
    PEER_OK(cudaMalloc((void **)&g->partial, kBlocks * sizeof(double)));
    g->peer[rank] = g->local;
    if (ipc_handle_out) {
        cudaIpcMemHandle_t h;
        memset(&h, 0, sizeof h);
        if (nranks > 1) PEER_OK(cudaIpcGetMemHandle(&h, g->local));
        static_assert(sizeof(cudaIpcMemHandle_t) == B200_IPC_HANDLE_BYTES, "IPC handle size");
        memcpy(ipc_handle_out, &h, sizeof h);
    }
    if (nranks == 1) { fill_dev(g); g->connected = true; }
    return g;
}

extern "C" int b200_peer_connect(b200_peer_group *g, const void *handles)
{
    for (int j = 0; j < g->nranks; ++j) {
        if (j == g->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)j * B200_IPC_HANDLE_BYTES, sizeof h);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            fprintf(stderr, "libb200-spmv: cudaIpcOpenMemHandle(rank %d) failed: %s\n", j,
                    cudaGetErrorString(e));
            return -1;
        }
        g->peer[j] = (char *)p;
    }
    fill_dev(g);
    g->connected = true;
    return 0;
}

extern "C" void b200_peer_destroy(b200_peer_group *g)
{
    if (!g) return;
    cudaDeviceSynchronize();
    for (int j = 0; j < g->nranks; ++j)
        if (j != g->rank && g->peer[j]) cudaIpcCloseMemHandle(g->peer[j]);
    cudaFree(g->local); cudaFree(g->counter); cudaFree(g->partial);
    free(g);
}

extern "C" double *b200_peer_xfull(b200_peer_group *g) { return (double *)g->local; }

extern "C" double *b200_peer_xbuf(b200_peer_group *g, uint64_t e)
{
    return (double *)(g->local + ((e & 1) ? seg_x_bytes(g->n_global) : 0));
}

extern "C" const unsigned long long *b200_peer_vflags(b200_peer_group *g) { return g->dev.vflag[g->rank]; }

/* what a product kernel needs to do the push of epoch e itself (b200_spmv_exec_pushed) */
extern "C" int b200_peer_describe_push(b200_peer_group *g, const double *v, int n_local, int64_t lo,
                                       uint64_t e, b200::XPush *out, const double **xbuf,
                                       const unsigned long long **vflags, int *cols_per_rank)
{
    if (!g || !g->connected) return -1;
    /* 16-byte granules on both sides */
    if ((lo & 1) || (n_local & 1) || (reinterpret_cast<uintptr_t>(v) & 15)) return -1;
    out->src = v;
    out->bytes = (size_t)n_local * sizeof(double);
    out->offset = (size_t)lo * sizeof(double);
    out->rank = g->rank;
    out->nranks = g->nranks;
    out->epoch = e;
    for (int j = 0; j < g->nranks; ++j) {
        out->dst[j] = (e & 1) ? (void *)g->dev.xalt[j] : (void *)g->dev.xfull[j];
        out->vflag[j] = g->dev.vflag[j];
        out->rflag[j] = g->dev.rflag[j];
    }
    out->counter = g->counter + 5;
    *xbuf = (e & 1) ? g->dev.xalt[g->rank] : g->dev.xfull[g->rank];
    *vflags = g->dev.vflag[g->rank];
    if (cols_per_rank) *cols_per_rank = 0;
    return 0;
}

extern "C" void b200_peer_post(b200_peer_group *g, const double *v, int n_local, int64_t lo, uint64_t e,
                               void *stream)
{
    peer_post_kernel<<<kBlocks, kThreads, 0, (cudaStream_t)stream>>>(g->dev, v, n_local, (long long)lo, e);
}

extern "C" void b200_peer_push(b200_peer_group *g, const double *v, int n_local, int64_t lo, uint64_t e,
                               void *stream)
{
    peer_push_kernel<<<kBlocks, kThreads, 0, (cudaStream_t)stream>>>(g->dev, v, n_local, (long long)lo, e, 0);
}

extern "C" void b200_peer_push_after(b200_peer_group *g, const double *v, int n_local, int64_t lo,
                                     uint64_t e, uint64_t e_consumed, void *stream)
{
    peer_push_kernel<<<kBlocks, kThreads, 0, (cudaStream_t)stream>>>(g->dev, v, n_local, (long long)lo, e,
                                                                     e_consumed);
}

extern "C" void b200_peer_exchange(b200_peer_group *g, const double *v, int n_local, int64_t lo,
                                   uint64_t e, void *stream)
{
    peer_exchange_kernel<<<kBlocks, kThreads, 0, (cudaStream_t)stream>>>(g->dev, v, n_local, (long long)lo, e);
}

extern "C" void b200_peer_consumed(b200_peer_group *g, uint64_t e, void *stream)
{
    peer_consumed_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(g->dev, e);
}

extern "C" void b200_peer_wait_vector(b200_peer_group *g, uint64_t e, void *stream)
{
    peer_wait_vector_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(g->dev, e);
}

extern "C" void b200_peer_dot(b200_peer_group *g, const double *x, const double *y, int n_local, int mode,
                              int slot, uint64_t e, void *stream)
{
    peer_dot_kernel<<<kBlocks, kThreads, 0, (cudaStream_t)stream>>>(g->dev, x, y, n_local, mode, slot, e);
}

extern "C" void b200_peer_update_zr(b200_peer_group *g, double *z, double *r, const double *p,
                                    const double *q, int n_local, int slot_rho, int slot_d, uint64_t e_d,
                                    int slot_out, uint64_t e_out, void *stream)
{
    peer_update_zr_kernel<<<kBlocks, kThreads, 0, (cudaStream_t)stream>>>(
        g->dev, z, r, p, q, n_local, slot_rho, slot_d, e_d, slot_out, e_out);
}

extern "C" void b200_peer_update_p(b200_peer_group *g, double *p, const double *r, int n_local, int64_t lo,
                                   int slot_new, uint64_t e_new, int slot_old, uint64_t e_vec, void *stream)
{
    peer_update_p_kernel<<<kBlocks, kThreads, 0, (cudaStream_t)stream>>>(
        g->dev, p, r, n_local, (long long)lo, slot_new, e_new, slot_old, e_vec);
}

extern "C" void b200_peer_scale(b200_peer_group *g, double *x, const double *z, int n_local, int slot,
                                uint64_t e, void *stream)
{
    peer_scale_kernel<<<kBlocks, kThreads, 0, (cudaStream_t)stream>>>(g->dev, x, z, n_local, slot, e);
}

extern "C" void b200_peer_read_slots(b200_peer_group *g, const int *slots, const uint64_t *epochs,
                                     int count, double *out, void *stream)
{
    SlotList l;
    l.count = count > B200_PEER_SLOTS ? B200_PEER_SLOTS : count;
    for (int k = 0; k < l.count; ++k) { l.slot[k] = slots[k]; l.epoch[k] = epochs[k]; }
    peer_read_slots_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(g->dev, l, out);
}
