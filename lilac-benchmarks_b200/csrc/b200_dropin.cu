/*
 * b200_dropin.cu -- host layer of libb200-spmv.so, part 2: the libspmv ABI
 * (spmv_harness_ / f_spmv_harness_) on top of the resident-matrix layer of
 * b200_host.cu.
 *
 * Reference behaviour mirrored (and where it deliberately differs):
 *   libspmv/gpu.c:227-262  matrix kept resident, keyed by the host pointers;
 *                          here the key also carries rows and nnz, and several
 *                          matrices can be resident at once (LRU).
 *   libspmv/gpu.c:140-209  mprotect/SIGSEGV invalidation: ON by default like the
 *                          reference's (B200_SPMV_GUARD=0 switches it off), but
 *                          only pages lying entirely inside the arrays are
 *                          protected and a previously installed handler is
 *                          chained to.  Second line of defence, for what the
 *                          guard cannot see (arrays smaller than a page, memory
 *                          freed and mapped again): a content fingerprint
 *                          checked on every call (B200_SPMV_VALIDATE=0: off) and
 *                          b200_spmv_invalidate().
 *   libspmv/gpu.c:264,285  x H2D and y D2H on every call: same, but on one device with a
 *                          PANEL kernel x goes up WHILE the product runs -- first chunk by
 *                          a PCIe-reading copy kernel on the product's stream, the others
 *                          by the copy engine on a second stream, a flag per chunk, the
 *                          kernel waits per chunk (watchdog on the wait; run_call) --,
 *                          pageable vectors through pinned bounce buffers (non-temporal
 *                          stores, chunk by chunk under the same overlap), and y is stored
 *                          by the kernel straight into pinned memory.  (A pool of copy
 *                          threads for the two memcpy's was measured and dropped: waking
 *                          sleeping helpers costs more than the 75 us copy, NPB CG class C
 *                          went from 1.34 s to 2.06 s, profiles/r02_run3_bench_C.json.)
 *   single device          gpu.c drives one GPU.  With B200_SPMV_DEVICES=0,1,..
 *                          (or "all") the SAME two symbols drive several: the rows
 *                          are split into nnz-balanced blocks, one per device;
 *                          per call every device pulls 1/G of x over its own PCIe
 *                          link and stores it into every device's x buffer over
 *                          NVLink (one kernel: PCIe read, G peer stores), runs
 *                          its block's product and writes its y block back
 *                          (SURVEY.md 8e "ABI mode").  One process, peer access
 *                          between the devices, no collective library.
 */
#include "host_internal.h"

#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

using namespace b200;

/* ------------------------------------------------------------------------
 * state
 * ---------------------------------------------------------------------- */
static const int kMaxFlags = 16;       /* >= kMaxDevices */

struct Part {
    DevCtx *ctx;
    b200_matrix *m;            /* rows [row_lo, row_hi); NULL when the block is empty */
    int row_lo, row_hi;
    size_t x_lo, x_hi;         /* byte range of x this device pulls from the host */
    char *d_x, *d_y;           /* full-length x, this block's y */
    unsigned long long *d_flags;   /* [kMaxFlags] flag[j] = call number whose x slice j has landed here (slices: one per
                                    * device in ABI mode, else the chunks the copy engine delivers while the product runs) */
    unsigned int *d_counter;       /* ... and the last-block counter of this device's copy kernel */
};

struct CacheEntry {
    const void *a; const int *rowstr; const int *colidx;
    int rows; int64_t nnz; int dtype;
    uint64_t fingerprint;
    std::vector<uint64_t> win_hash;   /* hash of every window of a / colidx, taken at upload */
    size_t win_next;                  /* window re-hashed by the next call */
    uint64_t last_use;
    int guard_slot;            /* index into g_guards, -1 when unguarded */
    int nparts;
    Part part[kMaxDevices];
    int ncols;
    int cols_per_part;         /* several devices: columns of x each device pulls (uniform) */
    unsigned long long epoch;  /* ... and the call number the flags carry */
    void *h_x, *h_y;           /* pinned bounce buffers (portable, mapped) */
    void *h_x_alias, *h_y_alias;   /* ... as the devices address them */
    size_t x_bytes, y_bytes;
    int x_chunk_cols, x_nchunks;   /* one device: x goes up in this many chunks of this many columns */
    unsigned long long *h_epoch;   /* pinned: source of the flag copies when stream memory ops are unavailable */
};

struct PinnedRange { char *lo, *hi; };

static std::vector<CacheEntry> g_cache;
static std::vector<PinnedRange> g_pinned;
static uint64_t g_tick = 0;
static b200_spmv_stats g_stats;
static bool g_conf_ready = false;
static int g_validate = 1, g_cache_cap = 4, g_time_kernels = 0;
static int g_zero_copy = 1, g_auto_pin = 0, g_guard = 1;
/* one device: x uploaded in chunks by the copy engine WHILE the product runs (the PANEL kernels
 * walk the columns left to right and wait per chunk) instead of in full before it */
static int g_x_overlap = 1, g_x_chunks = 6, g_x_prelaunch = 0;
static int g_x_second_early = 1;   /* ... and the second one requested before the launch */
static int g_x_first_kernel = 1;   /* the first chunk by the PCIe-reading copy kernel on the product's stream */
static int g_nt_copy = 1;          /* non-temporal stores into the bounce buffer (copy_to_bounce) */
static int g_x_overlap_auto = 0;   /* experiment: overlap also for vectors this library registered */
static int g_x_test_stall = 0;     /* test hook: the last chunk's flag of the next overlapped call is never written */
static size_t g_x_overlap_min = 256u << 10;
static unsigned long long g_x_timeout_ns = 20ull * 1000 * 1000;    /* watchdog of the product's wait for a chunk */
typedef int (*StreamWriteValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
static StreamWriteValue32Fn g_write_value32 = nullptr;     /* cuStreamWriteValue32, through the runtime */

/* caller vectors seen by the drop-in path (B200_SPMV_PIN_HOST) */
struct AutoPin { char *lo, *hi; int seen; bool registered; bool failed; uint64_t last_use; };
static std::vector<AutoPin> g_auto;
/* what the probe kernels report about an auto-registered vector (pinned, mapped) */
struct PinProbe { int x_bad; int x_timed_out; unsigned long long y_val[4]; };
static PinProbe *g_probe = nullptr;
struct RegSpan { char *lo, *hi; };              /* page-aligned host ranges registered by maybe_auto_pin */
static std::vector<RegSpan> g_spans;
static uint64_t g_reg_events = 0;               /* cudaHostRegister / Unregister calls made by this library */
static int g_ndev = 0, g_devs[kMaxDevices];
static int64_t g_multi_min_nnz = 1 << 22;

/* ------------------------------------------------------------------------
 * Write guard (libspmv/gpu.c:140-209, ALIGN macro :204-209): the host pages of a
 * resident matrix are made read-only and a SIGSEGV handler invalidates the
 * cache entry when the caller writes to them, chaining to any handler that was
 * installed before (gpu.c:180-181).  Differences: only pages lying entirely
 * inside an array are protected (gpu.c rounds outwards, so a write to a
 * neighbouring variable on a shared page silently drops the protection -- in
 * NPB `x` follows `a` in COMMON), and several matrices can be guarded.
 * The handler only touches this fixed table (async-signal-safe).
 * ---------------------------------------------------------------------- */
struct GuardRange { char *lo, *hi; };
struct GuardSlot {
    volatile int used;         /* 1 while a cache entry owns the slot */
    volatile int tripped;      /* set by the handler: host copy was written */
    GuardRange r[3];
};
static const int kMaxGuards = 16;
static GuardSlot g_guards[kMaxGuards];
static struct sigaction g_old_segv;
static bool g_guard_installed = false;

static void guard_handler(int sig, siginfo_t *si, void *ctx)
{
    char *addr = (char *)si->si_addr;
    for (int k = 0; k < kMaxGuards; ++k) {
        GuardSlot &g = g_guards[k];
        if (!g.used) continue;
        for (int j = 0; j < 3; ++j) {
            if (g.r[j].lo && addr >= g.r[j].lo && addr < g.r[j].hi) {
                for (int q = 0; q < 3; ++q)
                    if (g.r[q].lo && g.r[q].hi > g.r[q].lo)
                        mprotect(g.r[q].lo, (size_t)(g.r[q].hi - g.r[q].lo), PROT_READ | PROT_WRITE);
                g.tripped = 1;
                return;                           /* the faulting store is retried */
            }
        }
    }
    /* not ours: chain (gpu.c:180-181) or fall back to the default action */
    if ((g_old_segv.sa_flags & SA_SIGINFO) && g_old_segv.sa_sigaction) {
        g_old_segv.sa_sigaction(sig, si, ctx);
    } else if (g_old_segv.sa_handler != SIG_DFL && g_old_segv.sa_handler != SIG_IGN &&
               g_old_segv.sa_handler) {
        g_old_segv.sa_handler(sig);
    } else {
        signal(SIGSEGV, SIG_DFL);
    }
}

static GuardRange inner_pages(const void *p, size_t bytes)
{
    const uintptr_t page = (uintptr_t)sysconf(_SC_PAGE_SIZE);
    uintptr_t lo = ((uintptr_t)p + page - 1) / page * page;
    uintptr_t hi = ((uintptr_t)p + bytes) / page * page;
    GuardRange r = {nullptr, nullptr};
    if (hi > lo) { r.lo = (char *)lo; r.hi = (char *)hi; }
    return r;
}

static int guard_arm(const void *a, size_t a_bytes, const int *rowstr, size_t r_bytes,
                     const int *colidx, size_t c_bytes)
{
    if (!g_guard) return -1;
    if (!g_guard_installed) {
        struct sigaction sa;
        memset(&sa, 0, sizeof sa);
        sa.sa_flags = SA_SIGINFO;
        sigemptyset(&sa.sa_mask);
        sa.sa_sigaction = guard_handler;
        if (sigaction(SIGSEGV, &sa, &g_old_segv) != 0) return -1;
        g_guard_installed = true;
    }
    for (int k = 0; k < kMaxGuards; ++k) {
        GuardSlot &g = g_guards[k];
        if (g.used) continue;
        g.r[0] = inner_pages(a, a_bytes);
        g.r[1] = inner_pages(rowstr, r_bytes);
        g.r[2] = inner_pages(colidx, c_bytes);
        g.tripped = 0;
        g.used = 1;
        for (int j = 0; j < 3; ++j)
            if (g.r[j].lo && mprotect(g.r[j].lo, (size_t)(g.r[j].hi - g.r[j].lo), PROT_READ) != 0)
                g.r[j].lo = g.r[j].hi = nullptr;          /* not protectable (e.g. a read-only mapping) */
        return k;
    }
    return -1;
}

static void guard_disarm(int slot)
{
    if (slot < 0) return;
    GuardSlot &g = g_guards[slot];
    for (int j = 0; j < 3; ++j)
        if (g.r[j].lo) mprotect(g.r[j].lo, (size_t)(g.r[j].hi - g.r[j].lo), PROT_READ | PROT_WRITE);
    g.used = 0;
}

/* ------------------------------------------------------------------------
 * configuration
 * ---------------------------------------------------------------------- */
static void dump_stats_at_exit(void)
{
    if (!env_int("B200_SPMV_STATS", 0)) return;
    fprintf(stderr,
            "libb200-spmv stats: calls=%llu uploads=%llu launches=%llu kernel_ms=%.3f "
            "e2e_ms=%.3f upload_ms=%.3f h2d_MB=%.3f d2h_MB=%.3f devices=%d\n",
            (unsigned long long)g_stats.calls, (unsigned long long)g_stats.uploads,
            (unsigned long long)g_stats.kernel_launches, g_stats.kernel_ms, g_stats.e2e_ms,
            g_stats.upload_ms, g_stats.h2d_bytes / 1e6, g_stats.d2h_bytes / 1e6, std::max(g_ndev, 1));
}

/* B200_SPMV_DEVICES = "all" | "0,1,2,3": the devices the drop-in symbols spread a
 * large matrix over (ABI mode).  Unset: the one default device. */
static void parse_devices_locked(void)
{
    g_ndev = 0;
    int count = 0;
    CUDA_OK(cudaGetDeviceCount(&count));
    const char *v = getenv("B200_SPMV_DEVICES");
    if (v && *v) {
        if (!strcmp(v, "all")) {
            for (int d = 0; d < count && g_ndev < kMaxDevices; ++d) g_devs[g_ndev++] = d;
        } else {
            const char *p = v;
            while (*p && g_ndev < kMaxDevices) {
                char *end = nullptr;
                const long d = strtol(p, &end, 10);
                if (end == p) die("B200_SPMV_DEVICES=%s: not a comma-separated device list", v);
                if (d < 0 || d >= count) die("B200_SPMV_DEVICES=%s: device %ld does not exist (%d visible)", v, d, count);
                for (int k = 0; k < g_ndev; ++k)
                    if (g_devs[k] == (int)d) die("B200_SPMV_DEVICES=%s: device %ld listed twice", v, d);
                g_devs[g_ndev++] = (int)d;
                p = *end == ',' ? end + 1 : end;
            }
        }
    }
    if (g_ndev == 0) g_devs[g_ndev++] = default_device_locked();
    if (g_ndev > 1) {
        /* one process: plain peer access, no IPC handles */
        for (int i = 0; i < g_ndev; ++i) {
            DeviceScope scope(g_devs[i]);
            for (int j = 0; j < g_ndev; ++j) {
                if (i == j) continue;
                int can = 0;
                CUDA_OK(cudaDeviceCanAccessPeer(&can, g_devs[i], g_devs[j]));
                if (!can) die("B200_SPMV_DEVICES: device %d cannot access device %d", g_devs[i], g_devs[j]);
                cudaError_t e = cudaDeviceEnablePeerAccess(g_devs[j], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else CUDA_OK(e);
            }
        }
    }
}

static void ensure_conf_locked(void)
{
    ensure_init_locked(-1);
    if (g_conf_ready) return;
    g_validate = env_int("B200_SPMV_VALIDATE", 1);
    g_cache_cap = std::max(1, env_int("B200_SPMV_CACHE", 4));
    g_time_kernels = env_int("B200_SPMV_TIME_KERNELS", 0);
    g_zero_copy = env_int("B200_SPMV_ZEROCOPY", 1);
    g_auto_pin = std::max(0, env_int("B200_SPMV_PIN_HOST", 0));
    CUDA_OK(cudaHostAlloc((void **)&g_probe, sizeof(PinProbe), cudaHostAllocPortable | cudaHostAllocMapped));
    memset(g_probe, 0, sizeof(PinProbe));
    g_guard = env_int("B200_SPMV_GUARD", 1);
    g_x_overlap = env_int("B200_SPMV_X_OVERLAP", 1);
    {
        /* the product spins on flags that copies issued AFTER its launch will set */
        const char *lb = getenv("CUDA_LAUNCH_BLOCKING");
        if (lb && atoi(lb) != 0) g_x_overlap = 0;
    }
    g_x_chunks = std::min(kMaxFlags, std::max(1, env_int("B200_SPMV_X_CHUNKS", 6)));
    g_x_timeout_ns = (unsigned long long)std::max(1, env_int("B200_SPMV_X_TIMEOUT_MS", 20)) * 1000000ull;
    g_x_prelaunch = env_int("B200_SPMV_X_PRELAUNCH", 0);
    g_x_test_stall = env_int("B200_SPMV_X_TEST_STALL", 0);
    g_x_overlap_auto = env_int("B200_SPMV_X_OVERLAP_AUTO", 0);
    g_x_first_kernel = env_int("B200_SPMV_X_FIRST_KERNEL", 1);
    g_x_second_early = env_int("B200_SPMV_X_SECOND_EARLY", 1);
    g_nt_copy = env_int("B200_SPMV_NT_COPY", 1);      /* 1: every chunk is issued before the launch */
    g_x_overlap_min = (size_t)std::max(0, env_int("B200_SPMV_X_OVERLAP_MIN_KB", 256)) << 10;
    if (g_x_overlap && env_int("B200_SPMV_FLAG_WRITE", 1)) {
        void *fn = nullptr;
        enum cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_write_value32 = (StreamWriteValue32Fn)fn;
        else
            cudaGetLastError();
    }
    {
        const char *v = getenv("B200_SPMV_MULTI_MIN_NNZ");
        if (v && *v) g_multi_min_nnz = atoll(v);
    }
    parse_devices_locked();
    memset(&g_stats, 0, sizeof g_stats);
    atexit(dump_stats_at_exit);
    g_conf_ready = true;
}

/* ------------------------------------------------------------------------
 * staleness checks
 * ---------------------------------------------------------------------- */
static inline uint64_t mix(uint64_t h, uint64_t v)
{
    h = (h ^ v) * 0x9E3779B97F4A7C15ull;
    return h ^ (h >> 29);
}

static uint64_t hash_bytes(const void *p, size_t bytes, uint64_t h)
{
    const unsigned char *c = (const unsigned char *)p;
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t v;
        memcpy(&v, c + i, 8);
        h = mix(h, v);
    }
    uint64_t tail = 0;
    if (i < bytes) memcpy(&tail, c + i, bytes - i);
    return mix(h, tail ^ (uint64_t)bytes);
}

/* checked on every call: all of rowstr / a / colidx when they are small (<= 64 KB in
 * total: the guard cannot protect arrays below a page), else 64 evenly spaced probes
 * of each array */
static uint64_t fingerprint(const void *a, const int *rowstr, const int *colidx,
                            int rows, int64_t nnz, size_t es)
{
    uint64_t h = 1469598103934665603ull;
    if (rows <= 0) return h;
    const int64_t base = (int64_t)rowstr[0] - 1;
    const size_t total = (size_t)nnz * (es + 4) + ((size_t)rows + 1) * 4;
    if (total <= (64u << 10)) {
        h = hash_bytes(rowstr, ((size_t)rows + 1) * 4, h);
        h = hash_bytes((const char *)a + (size_t)base * es, (size_t)nnz * es, h);
        return hash_bytes(colidx + base, (size_t)nnz * 4, h);
    }
    const int probes = 64;
    for (int k = 0; k < probes && nnz > 0; ++k) {
        const int64_t i = base + (nnz - 1) * k / (probes - 1);
        uint64_t v = 0;
        memcpy(&v, (const char *)a + (size_t)i * es, es);
        h = mix(h, v);
        h = mix(h, (uint64_t)(uint32_t)colidx[i]);
    }
    for (int k = 0; k < probes; ++k) {
        const int64_t i = (int64_t)rows * k / (probes - 1);
        h = mix(h, (uint64_t)(uint32_t)rowstr[i]);
    }
    return h;
}

/* rolling check: a / colidx are cut into 32 KB windows hashed at upload; every call
 * re-hashes one window (~2 us), so an in-place change the probes miss is found within
 * one sweep over the windows even when the guard is off */
static const size_t kWindow = 32u << 10;

static size_t window_count(int64_t nnz, size_t es)
{
    return ((size_t)nnz * es + kWindow - 1) / kWindow + ((size_t)nnz * 4 + kWindow - 1) / kWindow;
}

static uint64_t window_hash(const CacheEntry &e, size_t w)
{
    const size_t es = elem_size(e.dtype);
    const int64_t base = e.rows > 0 ? (int64_t)e.rowstr[0] - 1 : 0;
    const size_t wa = ((size_t)e.nnz * es + kWindow - 1) / kWindow;
    const char *p; size_t total;
    if (w < wa) { p = (const char *)e.a + (size_t)base * es; total = (size_t)e.nnz * es; }
    else { p = (const char *)(e.colidx + base); total = (size_t)e.nnz * 4; w -= wa; }
    const size_t lo = w * kWindow, hi = std::min(total, lo + kWindow);
    return hash_bytes(p + lo, hi - lo, 0x243F6A8885A308D3ull);
}

/* ------------------------------------------------------------------------
 * pinned caller vectors
 * ---------------------------------------------------------------------- */
/* Device-usable alias of a pinned (cudaHostAlloc'ed or registered) host range, or NULL
 * when any part of [p, p + bytes) is pageable.  First and last byte must both be pinned and
 * map to device addresses `bytes - 1` apart; and since two registrations can sit at the two
 * ends of a range whose middle is pageable (neighbouring vectors registered page by page), the
 * allocations / registrations behind the addresses are walked (cuPointerGetAttribute
 * RANGE_START_ADDR / RANGE_SIZE, obtained through the runtime) until they cover the range. */
typedef int (*PointerGetAttributeFn)(void *, int, unsigned long long);
static PointerGetAttributeFn g_ptr_attr = nullptr;
static bool g_get_range_tried = false;
static const int kAttrRangeStart = 11, kAttrRangeSize = 12;   /* CU_POINTER_ATTRIBUTE_RANGE_START_ADDR / _SIZE */

static void *pinned_device_alias(const void *p, size_t bytes)
{
    if (bytes == 0) return nullptr;
    cudaPointerAttributes a0, a1;
    if (cudaPointerGetAttributes(&a0, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a0.type != cudaMemoryTypeHost || !a0.devicePointer) return nullptr;
    if (cudaPointerGetAttributes(&a1, (const char *)p + bytes - 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a1.type != cudaMemoryTypeHost || !a1.devicePointer) return nullptr;
    if ((char *)a1.devicePointer - (char *)a0.devicePointer != (ptrdiff_t)(bytes - 1)) return nullptr;
    if (!g_get_range_tried) {
        g_get_range_tried = true;
        void *fn = nullptr;
        enum cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_ptr_attr = (PointerGetAttributeFn)fn;
        else
            cudaGetLastError();
    }
    if (g_ptr_attr) {
        const unsigned long long lo = (unsigned long long)(uintptr_t)p, hi = lo + bytes;
        unsigned long long at = lo;
        for (int hops = 0; at < hi; ++hops) {
            unsigned long long base = 0;
            size_t size = 0;
            const bool ok = g_ptr_attr(&base, kAttrRangeStart, at) == 0 && g_ptr_attr(&size, kAttrRangeSize, at) == 0 &&
                            size > 0 && base <= at;
            if (!ok && hops == 0) { g_ptr_attr = nullptr; break; }   /* no ranges for host memory here: ends only */
            if (!ok || hops > 64) return nullptr;                    /* a pageable hole in the middle */
            at = base + size;                                        /* the next range must start right here */
        }
    }
    return a0.devicePointer;
}

static int register_range_locked(void *p, size_t bytes)
{
    ++g_reg_events;
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) { cudaGetLastError(); return -1; }
    PinnedRange r = {(char *)p, (char *)p + bytes};
    g_pinned.push_back(r);
    return 0;
}

static bool spans_intersect(const void *p, size_t bytes)
{
    const char *lo = (const char *)p, *hi = lo + bytes;
    for (const RegSpan &sp : g_spans)
        if (lo < sp.hi && sp.lo < hi) return true;
    return false;
}

static bool spans_cover(const void *p, size_t bytes)
{
    const char *lo = (const char *)p, *hi = lo + bytes;
    for (const RegSpan &sp : g_spans)
        if (lo >= sp.lo && hi <= sp.hi) return true;
    return false;
}

/* B200_SPMV_PIN_HOST=N / b200_spmv_set_auto_pin(N) (default 0: never): a caller vector seen
 * N times at the same address with the same length is registered (cudaHostRegister), so that
 * later calls move it in place over PCIe instead of through the bounce buffer -- NPB's COMMON
 * vectors, pagerank's two std::vectors, SparseBench's static work arrays all come back call
 * after call; bfs, whose vectors are new on every call, never qualifies with N >= 2.  The hazard
 * of registering memory one does not own -- the owner frees it and the address range is mapped
 * again, while the GPU mapping still points at the old pages -- is CHECKED on every call that
 * uses such a range: four sample words of x are compared as the GPU sees them through the
 * mapping with what the host sees, and four words of y as the GPU wrote them with what the host
 * reads back; on a mismatch the range is unregistered for good and the call is redone through
 * the bounce buffer.  It stays opt-in all the same: a stale registration also misleads every
 * OTHER user of the CUDA runtime in the process (a cudaMemcpy from that address range would
 * read the old pages), which this library cannot check; a caller whose vectors live as long
 * as the process (the reference's Fortran and C callers) has nothing to fear.
 * Returns true when [p, p + bytes) lies in memory registered by this library -- by this call
 * or by an earlier one for ANY vector: a different vector that now lives where a registered one
 * used to be is exactly the stale case -- and must therefore be checked. */
static bool maybe_auto_pin(const void *p, size_t bytes)
{
    if (bytes == 0) return false;
    if (spans_cover(p, bytes)) return true;
    if (!g_auto_pin || spans_intersect(p, bytes)) return false;
    char *lo = (char *)p, *hi = lo + bytes;
    AutoPin *hit = nullptr;
    for (AutoPin &a : g_auto)
        if (a.lo == lo && a.hi == hi) { hit = &a; break; }
    if (!hit) {
        if (g_auto.size() >= 32) {                 /* forget (and release) the least recently used */
            size_t v = 0;
            for (size_t i = 1; i < g_auto.size(); ++i)
                if (g_auto[i].last_use < g_auto[v].last_use) v = i;
            g_auto[v] = g_auto.back();              /* its registration, if any, stays: spans are shared */
            g_auto.pop_back();
        }
        AutoPin a = {lo, hi, 0, false, false, 0};
        g_auto.push_back(a);
        hit = &g_auto.back();
    }
    hit->last_use = g_tick;
    if (hit->failed) return false;
    if (hit->registered) return true;
    if (++hit->seen < g_auto_pin) return false;
    if (pinned_device_alias(p, bytes)) return false;        /* pinned by its owner: trusted */
    /* Registration is page-granular and must not overlap an existing one, and callers keep
     * their vectors side by side (NPB: x, z, p, q, r are consecutive in one COMMON block, so
     * neighbours share a page): spans that touch are merged into one registration. */
    const uintptr_t page = (uintptr_t)sysconf(_SC_PAGE_SIZE);
    char *slo = (char *)((uintptr_t)lo / page * page);
    char *shi = (char *)(((uintptr_t)hi + page - 1) / page * page);
    for (size_t i = 0; i < g_spans.size();) {
        if (g_spans[i].lo <= shi && slo <= g_spans[i].hi) {            /* overlaps or abuts */
            ++g_reg_events;
            cudaHostUnregister(g_spans[i].lo);
            cudaGetLastError();
            slo = std::min(slo, g_spans[i].lo);
            shi = std::max(shi, g_spans[i].hi);
            g_spans[i] = g_spans.back();
            g_spans.pop_back();
            i = 0;                                                     /* the union may touch more */
        } else {
            ++i;
        }
    }
    ++g_reg_events;
    if (cudaHostRegister(slo, (size_t)(shi - slo), cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
        cudaGetLastError();
        /* e.g. the span touches memory registered by somebody else: nothing of it is ours now */
        for (AutoPin &a : g_auto)
            if (a.registered && a.lo < shi && slo < a.hi) { a.registered = false; a.failed = true; }
        hit->failed = true;
        return false;
    }
    RegSpan sp = {slo, shi};
    g_spans.push_back(sp);
    hit->registered = true;
    return true;
}

/* the mapping of an auto-registered range turned out stale: never use it (or what shares its
 * registration) again */
static void auto_pin_revoke(const void *p)
{
    const char *c = (const char *)p;
    for (size_t i = 0; i < g_spans.size(); ++i) {
        if (c < g_spans[i].lo || c >= g_spans[i].hi) continue;
        ++g_reg_events;
        cudaHostUnregister(g_spans[i].lo);
        cudaGetLastError();
        for (AutoPin &a : g_auto)
            if (a.lo < g_spans[i].hi && g_spans[i].lo < a.hi) { a.registered = false; a.failed = true; }
        g_spans[i] = g_spans.back();
        g_spans.pop_back();
        if (g_verbose) fprintf(stderr, "libb200-spmv: registered vector %p was remapped by its owner; "
                                       "back to the bounce buffer\n", p);
        return;
    }
}

/* Caller vector -> pinned bounce buffer.  The destination is read next by the GPU (DMA or a
 * PCIe-reading kernel), never by this CPU, so the stores bypass the cache: no read-for-ownership
 * of lines the device fetched last call, no eviction of the caller's working set -- glibc's memcpy
 * switches to such stores only far above these sizes.  Ends with a store fence: the data is
 * globally visible before the copy that reads it is issued.  (B200_SPMV_NT_COPY=0: plain memcpy.) */
static void copy_to_bounce(void *dst, const void *src, size_t n)
{
#if defined(__x86_64__)
    if (g_nt_copy && n >= 4096) {
        char *d = (char *)dst;
        const char *s = (const char *)src;
        const size_t head = (size_t)(-(intptr_t)d) & 15;          /* to a 16-byte destination boundary */
        if (head) { memcpy(d, s, head); d += head; s += head; n -= head; }
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            const __m128i v0 = _mm_loadu_si128((const __m128i *)(s + i));
            const __m128i v1 = _mm_loadu_si128((const __m128i *)(s + i + 16));
            const __m128i v2 = _mm_loadu_si128((const __m128i *)(s + i + 32));
            const __m128i v3 = _mm_loadu_si128((const __m128i *)(s + i + 48));
            _mm_stream_si128((__m128i *)(d + i), v0);
            _mm_stream_si128((__m128i *)(d + i + 16), v1);
            _mm_stream_si128((__m128i *)(d + i + 32), v2);
            _mm_stream_si128((__m128i *)(d + i + 48), v3);
        }
        _mm_sfence();
        if (i < n) memcpy(d + i, s + i, n - i);
        return;
    }
#endif
    memcpy(dst, src, n);
}

/* four sample positions (byte offsets, multiples of es) spread over [0, bytes) */
static void sample_offsets(size_t bytes, size_t es, size_t off[4])
{
    const size_t n = bytes / es;
    const size_t idx[4] = {0, n / 3, (2 * n) / 3, n ? n - 1 : 0};
    for (int k = 0; k < 4; ++k) off[k] = idx[k] * es;
}

/* ------------------------------------------------------------------------
 * cache
 * ---------------------------------------------------------------------- */
static void release_entry_locked(CacheEntry &e)
{
    guard_disarm(e.guard_slot);
    for (int p = 0; p < e.nparts; ++p) {
        Part &pt = e.part[p];
        DeviceScope scope(pt.ctx->device);
        cudaStreamSynchronize(pt.ctx->stream);
        cudaStreamSynchronize(pt.ctx->copy_stream);
        release_locked(pt.m);
        cudaFree(pt.d_x);
        cudaFree(pt.d_y);
        cudaFree(pt.d_flags);
        cudaFree(pt.d_counter);
    }
    if (e.h_x) cudaFreeHost(e.h_x);
    if (e.h_y) cudaFreeHost(e.h_y);
    if (e.h_epoch) cudaFreeHost(e.h_epoch);
}

static void build_entry_locked(CacheEntry &e)
{
    const size_t es = elem_size(e.dtype);
    const bool multi = g_ndev > 1 && e.nnz >= g_multi_min_nnz && e.rows >= g_ndev;
    e.nparts = multi ? g_ndev : 1;
    int bounds[kMaxDevices + 1];
    if (multi) b200_spmv_partition_rows(e.rowstr, e.rows, e.nparts, bounds);
    else { bounds[0] = 0; bounds[1] = e.rows; }
    e.ncols = 0;
    for (int p = 0; p < e.nparts; ++p) {
        Part &pt = e.part[p];
        pt.ctx = ctx_for_device_locked(multi ? g_devs[p] : g_devs[0]);
        pt.row_lo = bounds[p]; pt.row_hi = bounds[p + 1];
        pt.m = nullptr; pt.d_x = pt.d_y = nullptr;
        pt.d_flags = nullptr; pt.d_counter = nullptr;
        if (pt.row_hi > pt.row_lo || !multi) {
            /* a row block is addressed exactly as the ABI would: (a, rowstr + lo, colidx, hi - lo) */
            pt.m = upload_locked(pt.ctx, e.a, e.rowstr + pt.row_lo, e.colidx, pt.row_hi - pt.row_lo,
                                 e.dtype, B200_KERNEL_AUTO);
            e.ncols = std::max(e.ncols, pt.m->ncols);
        }
    }
    e.x_bytes = (size_t)std::max(e.ncols, 1) * es;
    e.y_bytes = (size_t)std::max(e.rows, 1) * es;
    const size_t x_used = (size_t)e.ncols * es;
    /* slice of x each device fetches: the same number of columns everywhere (the product
     * kernels map a column to the device that delivers it by one division), 16-byte granules */
    const int gran = (int)(16 / es);
    e.cols_per_part = std::max(gran, ((e.ncols + e.nparts - 1) / e.nparts + gran - 1) / gran * gran);
    e.epoch = 0;
    for (int p = 0; p < e.nparts; ++p) {
        Part &pt = e.part[p];
        DeviceScope scope(pt.ctx->device);
        pt.x_lo = std::min(x_used, (size_t)p * e.cols_per_part * es);
        pt.x_hi = std::min(x_used, (size_t)(p + 1) * e.cols_per_part * es);
        CUDA_OK(cudaMalloc((void **)&pt.d_x, std::max<size_t>(e.x_bytes, 16)));
        CUDA_OK(cudaMalloc((void **)&pt.d_y, std::max<size_t>((size_t)(pt.row_hi - pt.row_lo) * es, 16)));
        CUDA_OK(cudaMalloc((void **)&pt.d_flags, kMaxFlags * sizeof(unsigned long long)));
        CUDA_OK(cudaMemsetAsync(pt.d_flags, 0, kMaxFlags * sizeof(unsigned long long), pt.ctx->stream));
        if (multi) {
            CUDA_OK(cudaMalloc((void **)&pt.d_counter, sizeof(unsigned int)));
            CUDA_OK(cudaMemsetAsync(pt.d_counter, 0, sizeof(unsigned int), pt.ctx->stream));
        }
        /* the flags are written from other streams (and other devices) from the first call on */
        CUDA_OK(cudaStreamSynchronize(pt.ctx->stream));
    }
    CUDA_OK(cudaHostAlloc(&e.h_x, e.x_bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    CUDA_OK(cudaHostAlloc(&e.h_y, e.y_bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    CUDA_OK(cudaHostGetDevicePointer(&e.h_x_alias, e.h_x, 0));
    CUDA_OK(cudaHostGetDevicePointer(&e.h_y_alias, e.h_y, 0));
    CUDA_OK(cudaHostAlloc((void **)&e.h_epoch, kMaxFlags * sizeof(unsigned long long), cudaHostAllocPortable));
    memset(e.h_epoch, 0, kMaxFlags * sizeof(unsigned long long));
    /* chunks of the overlapped upload: a multiple of 2048 columns each (whole 16-byte granules
     * for the bulk copies of the kernels, and no chunk worth less than a copy's fixed cost) */
    {
        const int cg = 2048;
        const int per = (std::max(e.ncols, 1) + g_x_chunks - 1) / g_x_chunks;
        e.x_chunk_cols = std::max(cg, (per + cg - 1) / cg * cg);
        e.x_nchunks = (std::max(e.ncols, 1) + e.x_chunk_cols - 1) / e.x_chunk_cols;
    }
    if (g_verbose && multi) {
        fprintf(stderr, "libb200-spmv: matrix rows=%d nnz=%lld spread over %d devices:", e.rows,
                (long long)e.nnz, e.nparts);
        for (int p = 0; p < e.nparts; ++p)
            fprintf(stderr, " dev%d[rows %d..%d, %s]", e.part[p].ctx->device, e.part[p].row_lo, e.part[p].row_hi,
                    e.part[p].m ? b200_spmv_kernel_name(e.part[p].m) : "empty");
        fprintf(stderr, "\n");
    }
}

/* the content checks of a cache hit (probes + one rolling window): ~10 us of host time on a
 * large matrix, so the caller runs them AFTER the launches, while the GPU works, and redoes
 * the call when they fail */
static bool entry_content_stale(CacheEntry &e)
{
    const size_t es = elem_size(e.dtype);
    bool stale = fingerprint(e.a, e.rowstr, e.colidx, e.rows, e.nnz, es) != e.fingerprint;
    if (!stale && !e.win_hash.empty()) {
        stale = window_hash(e, e.win_next) != e.win_hash[e.win_next];
        e.win_next = (e.win_next + 1) % e.win_hash.size();
    }
    return stale;
}

static void drop_entry_locked(CacheEntry *ep)
{
    const size_t i = (size_t)(ep - g_cache.data());
    if (g_verbose) fprintf(stderr, "libb200-spmv: host matrix changed, re-uploading\n");
    release_entry_locked(g_cache[i]);
    g_cache[i] = g_cache.back();
    g_cache.pop_back();
}

/* *unchecked: the entry is a cache hit whose content checks are still to be run */
static CacheEntry *lookup_locked(const void *a, const int *rowstr, const int *colidx,
                                 int rows, int dtype, bool *unchecked)
{
    const int64_t nnz = rows > 0 ? (int64_t)rowstr[rows] - rowstr[0] : 0;
    const size_t es = elem_size(dtype);
    ++g_tick;
    *unchecked = false;
    for (size_t i = 0; i < g_cache.size(); ++i) {
        CacheEntry &e = g_cache[i];
        if (e.a == a && e.rowstr == rowstr && e.colidx == colidx && e.rows == rows &&
            e.nnz == nnz && e.dtype == dtype) {
            const bool stale = e.guard_slot >= 0 && g_guards[e.guard_slot].tripped;
            if (stale) {
                if (g_verbose) fprintf(stderr, "libb200-spmv: host matrix changed, re-uploading\n");
                release_entry_locked(e);
                g_cache[i] = g_cache.back();
                g_cache.pop_back();
                break;
            }
            e.last_use = g_tick;
            *unchecked = g_validate != 0;
            return &e;
        }
    }
    /* miss: upload (evict the least recently used entry beyond the cap) */
    const double t0 = now_ms();
    if ((int)g_cache.size() >= g_cache_cap) {
        size_t victim = 0;
        for (size_t i = 1; i < g_cache.size(); ++i)
            if (g_cache[i].last_use < g_cache[victim].last_use) victim = i;
        release_entry_locked(g_cache[victim]);
        g_cache[victim] = g_cache.back();
        g_cache.pop_back();
    }
    g_cache.emplace_back();
    CacheEntry &e = g_cache.back();
    e.a = a; e.rowstr = rowstr; e.colidx = colidx; e.rows = rows; e.nnz = nnz; e.dtype = dtype;
    e.last_use = g_tick;
    e.guard_slot = -1;
    e.win_next = 0;
    e.h_x = e.h_y = nullptr;
    e.h_epoch = nullptr;
    build_entry_locked(e);
    if (g_validate) {
        e.fingerprint = fingerprint(a, rowstr, colidx, rows, nnz, es);
        const size_t total = (size_t)nnz * (es + 4) + ((size_t)rows + 1) * 4;
        if (total > (64u << 10)) {
            e.win_hash.resize(window_count(nnz, es));
            for (size_t w = 0; w < e.win_hash.size(); ++w) e.win_hash[w] = window_hash(e, w);
        }
    } else {
        e.fingerprint = 0;
    }
    const size_t last = rows > 0 ? (size_t)rowstr[rows] - 1 : 0;          /* entries up to the last offset */
    e.guard_slot = guard_arm(a, last * es, rowstr, ((size_t)rows + 1) * sizeof(int), colidx, last * sizeof(int));
    g_stats.uploads++;
    g_stats.upload_ms += now_ms() - t0;
    return &e;
}

/* ------------------------------------------------------------------------
 * one call
 * ---------------------------------------------------------------------- */
enum CallResult {
    CALL_OK = 0,
    CALL_REDO_PLAIN,       /* the product gave up waiting for a chunk of x (overlap is off from now on): redo */
    CALL_REDO_BOUNCE,      /* an auto-registered vector turned out to be remapped (revoked): redo without it */
    CALL_STALE_MATRIX      /* the content checks failed: the resident copy is out of date, re-upload and redo */
};

/* "slice k of x has landed": the call number into d_flag, ordered after the copies issued on
 * `s` before it.  A stream memory operation when the driver has them, else an 8-byte copy. */
static void write_flag(cudaStream_t s, unsigned long long *d_flag, unsigned long long epoch,
                       unsigned long long *h_word)
{
    if (g_write_value32) {
        /* flags are 64-bit, little endian; call numbers stay below 2^32 (run_call resets them) */
        if (g_write_value32(s, (unsigned long long)(uintptr_t)d_flag, (unsigned int)epoch, 0u) == 0) return;
        g_write_value32 = nullptr;                  /* not on this driver / device: copies from now on */
    }
    *h_word = epoch;
    CUDA_OK(cudaMemcpyAsync(d_flag, h_word, sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
}

/* one product through the resident entry.  check_content: run the staleness checks of a cache
 * hit between the launches and the synchronisation */
static CallResult run_call(CacheEntry &e, void *ov, const void *iv, int n, int dtype, bool allow_auto,
                           bool check_content)
{
    const size_t es = elem_size(dtype);
    const size_t x_used = (size_t)e.ncols * es;
    static uint64_t reg_events_seen = 0;
    /* memory registered by this library is only ever used with the probes on; a range that
     * merely touches such memory (or a redo after a failed probe) takes the bounce buffer */
    const bool x_auto = allow_auto && x_used > 0 && maybe_auto_pin(iv, x_used);
    const bool y_auto = allow_auto && maybe_auto_pin(ov, (size_t)n * es);
    const bool x_avoid = !x_auto && x_used > 0 && spans_intersect(iv, x_used);
    const bool y_avoid = !y_auto && spans_intersect(ov, (size_t)n * es);
    size_t x_off[4] = {0, 0, 0, 0}, y_off[4] = {0, 0, 0, 0};
    unsigned long long x_val[4] = {0, 0, 0, 0};
    if (x_auto) {
        sample_offsets(x_used, es, x_off);
        for (int k = 0; k < 4; ++k) memcpy(&x_val[k], (const char *)iv + x_off[k], es);
        g_probe->x_bad = 0;
    }
    if (y_auto) sample_offsets((size_t)n * es, es, y_off);

    const bool multi = e.nparts > 1;
    /* One device and the paired PANEL kernel (it walks the columns left to right, and its
     * flagged instance has a watchdog on the wait): x goes up in chunks WHILE the product runs
     * -- the first by a copy kernel on the product's stream, the others through the copy engine
     * on a second stream --; the kernel waits per chunk just before the panels that need it
     * (XFlags).  The first two chunks are requested before the launch, the others after it,
     * so the product starts ~10 us into the call
     * instead of after the whole vector has crossed PCIe -- and for a pageable vector the
     * memcpy into the bounce buffer overlaps the product chunk by chunk as well.
     * (The other way round -- the product kernel fetching x itself over PCIe, chunk by chunk
     * behind flags -- was built and measured: a GPU-issued PCIe read takes ~10 us to come
     * back, a window of chunks in flight serialises on that, and without a window the chunks
     * do not land in column order: 208 us per call against 157 us, profiles/r02_run13_bench_C*.json.) */
    /* Not in a call that has just (un)registered host memory: the driver applies the change
     * to the GPU's address space with the next work it submits, and that did not get past a
     * product spinning on all SMs -- copies queued behind it never ran (bench.py's auto-pin
     * leg hung in the calls that registered a vector; profiles/r02_run29_bench_C.err,
     * r02_run31_e2e_variants.txt).  Such a call uploads x before the product, and its stream
     * synchronisation flushes the change.  Vectors this library registered itself stay with
     * the copy kernel altogether unless B200_SPMV_X_OVERLAP_AUTO=1.  Whatever ELSE may
     * serialise the two streams (a registration by another thread, a profiler) costs one
     * watchdog timeout, after which the overlap is off for the process. */
    const bool reg_changed = g_reg_events != reg_events_seen;
    reg_events_seen = g_reg_events;
    const bool overlap = !multi && g_x_overlap && x_used >= g_x_overlap_min && x_used > 0 && !reg_changed &&
                         (!x_auto || g_x_overlap_auto) &&
                         e.part[0].m && exec_takes_guarded_flags(e.part[0].m) && e.x_nchunks <= kMaxFlags;
    if (overlap) g_probe->x_timed_out = 0;

    /* x: host -> device (gpu.c:264).  Without the overlap, pinned caller memory is read in
     * place over PCIe by a copy kernel on the library's stream; pageable memory goes through
     * the pinned bounce buffer first. */
    const char *x_pinned = nullptr;       /* host pointer to pinned x, for the copy engine */
    const char *x_alias = nullptr;        /* its device alias, for the copy kernel */
    bool x_bounce = false;
    if (x_used > 0) {
        x_alias = x_avoid ? nullptr : (const char *)pinned_device_alias(iv, x_used);
        x_pinned = (const char *)iv;
        if (!x_alias) {
            x_bounce = true;
            if (!overlap) copy_to_bounce(e.h_x, iv, x_used);
            x_pinned = (const char *)e.h_x;
            x_alias = (const char *)e.h_x_alias;
        }
    }
    /* y: device -> host (gpu.c:285).  The PANEL and SMALL kernels store y coalesced, so they
     * write straight into pinned host memory (the caller's, else the bounce buffer). */
    char *y_alias = y_avoid ? nullptr : (char *)pinned_device_alias(ov, (size_t)n * es);
    const bool y_direct = y_alias != nullptr;
    char *y_host = y_direct ? (char *)ov : (char *)e.h_y;
    if (!y_direct) y_alias = (char *)e.h_y_alias;

    /* one cudaSetDevice per device and loop, restored once at the end: with eight devices
     * the host side of a call is a few dozen runtime calls, and it is the critical path
     * (the first version -- a device scope per step, an event per slice and 56 stream
     * waits -- spent 0.24 ms of a 0.51 ms call in the runtime) */
    int saved_dev = -1, cur_dev = -1;
    cudaGetDevice(&saved_dev);
    cur_dev = saved_dev;
    auto use = [&](int d) { if (d != cur_dev) { CUDA_OK(cudaSetDevice(d)); cur_dev = d; } };
    if (overlap && e.epoch >= 0xFFFFFFF0ull) {        /* 32-bit flag writes: start the call numbers over */
        use(e.part[0].ctx->device);
        CUDA_OK(cudaMemsetAsync(e.part[0].d_flags, 0, kMaxFlags * sizeof(unsigned long long), e.part[0].ctx->stream));
        CUDA_OK(cudaStreamSynchronize(e.part[0].ctx->stream));
        e.epoch = 0;
    }
    const unsigned long long epoch = ++e.epoch;
    /* chunk k of the overlapped upload: (memcpy into the bounce buffer,) copy, flag */
    /* The first chunk is on the critical path (the product cannot start without it): a copy
     * engine transfer takes ~20 us from the call to the flag, the PCIe-reading copy kernel
     * on the product's own stream -- no flag needed, stream order -- about half of that. */
    const bool first_by_kernel = overlap && g_x_first_kernel && g_zero_copy;
    int x_sent = 0;                                /* chunks requested so far */
    auto send_chunk = [&](int k) {
        Part &pt = e.part[0];
        const size_t lo = std::min(x_used, (size_t)k * e.x_chunk_cols * es);
        const size_t hi = std::min(x_used, (size_t)(k + 1) * e.x_chunk_cols * es);
        if (hi > lo && x_bounce) copy_to_bounce((char *)e.h_x + lo, (const char *)iv + lo, hi - lo);
        if (k == 0 && first_by_kernel) {
            void *dst[1] = {pt.d_x};
            if (hi > lo) launch_copy_in_multi(x_alias, dst, 1, hi - lo, pt.ctx->stream);
            return;
        }
        if (hi > lo)
            CUDA_OK(cudaMemcpyAsync(pt.d_x + lo, x_pinned + lo, hi - lo, cudaMemcpyHostToDevice, pt.ctx->copy_stream));
        if (g_x_test_stall && k == e.x_nchunks - 1) { g_x_test_stall = 0; return; }    /* the watchdog's test */
        write_flag(pt.ctx->copy_stream, pt.d_flags + k, epoch, e.h_epoch + k);
    };
    for (int p = 0; p < e.nparts; ++p) {
        Part &pt = e.part[p];
        use(pt.ctx->device);
        cudaStream_t s = pt.ctx->stream;
        const size_t bytes = pt.x_hi - pt.x_lo;
        if (p == 0 && x_auto)      /* what does the GPU see through the mapping we registered? */
            launch_probe_x(x_alias, x_off, x_val, (int)es, &g_probe->x_bad, s);
        if (multi) {
            /* one kernel: read the slice over this device's PCIe link, store it into every
             * device's x buffer (the peers' over NVLink), then publish the call number on
             * every device's flag for this slice -- the products wait on flags, not events */
            void *dst[kMaxDevices];
            unsigned long long *flag[kMaxDevices];
            for (int j = 0; j < e.nparts; ++j) {
                dst[j] = e.part[j].d_x + pt.x_lo;
                flag[j] = e.part[j].d_flags + p;
            }
            launch_copy_in_multi_flagged(x_alias ? x_alias + pt.x_lo : nullptr, dst, e.nparts, bytes, flag,
                                         epoch, pt.d_counter, s);
        } else if (overlap) {
            /* with the first chunk on the product's stream, the second one (copy engine: ~20 us
             * from here to its flag) is requested before the launch as well */
            x_sent = g_x_prelaunch ? e.x_nchunks : std::min(e.x_nchunks, first_by_kernel && g_x_second_early ? 2 : 1);
            for (int k = 0; k < x_sent; ++k) send_chunk(k);
        } else if (bytes > 0) {
            if (g_zero_copy) {
                void *dst[1] = {pt.d_x + pt.x_lo};
                launch_copy_in_multi(x_alias + pt.x_lo, dst, 1, bytes, s);
            } else {
                CUDA_OK(cudaMemcpyAsync(pt.d_x + pt.x_lo, x_pinned + pt.x_lo, bytes, cudaMemcpyHostToDevice, s));
            }
        }
    }
    int launched = 0;
    int timed_part = -1;
    bool y_probed = false;
    for (int p = 0; p < e.nparts; ++p) {
        Part &pt = e.part[p];
        const int prow = pt.row_hi - pt.row_lo;
        if (!pt.m || prow <= 0) continue;
        use(pt.ctx->device);
        cudaStream_t s = pt.ctx->stream;
        /* kernel timing: every call on one device; on several, the first block only */
        const bool timed = g_time_kernels && timed_part < 0;
        if (timed) { CUDA_OK(cudaEventRecord(pt.ctx->ev0, s)); timed_part = p; }
        char *y_target = nullptr;
        if (g_zero_copy && (pt.m->kernel == B200_KERNEL_PANEL || pt.m->kernel == B200_KERNEL_SMALL) && y_alias)
            y_target = y_alias + (size_t)pt.row_lo * es;
        if (overlap) {
            SliceFlags sf = {pt.d_flags, epoch, e.x_chunk_cols, e.x_nchunks, &g_probe->x_timed_out, g_x_timeout_ns,
                             first_by_kernel ? 1 : 0};
            launched += exec_locked(pt.m, pt.d_x, y_target ? y_target : pt.d_y, s, &sf);
        } else if (multi && exec_waits_in_kernel(pt.m)) {
            /* the ring kernel waits per slice, just before the panels that need it */
            SliceFlags sf = {pt.d_flags, epoch, e.cols_per_part, e.nparts, nullptr, 0ull, 0};
            launched += exec_locked(pt.m, pt.d_x, y_target ? y_target : pt.d_y, s, &sf);
        } else {
            if (multi) launch_wait_flags(pt.d_flags, e.nparts, epoch, s);
            launched += exec_locked(pt.m, pt.d_x, y_target ? y_target : pt.d_y, s, nullptr);
        }
        if (timed) CUDA_OK(cudaEventRecord(pt.ctx->ev1, s));
        if (!y_target)
            CUDA_OK(cudaMemcpyAsync(y_host + (size_t)pt.row_lo * es, pt.d_y, (size_t)prow * es,
                                    cudaMemcpyDeviceToHost, s));
        if (y_auto && y_direct && pt.row_lo == 0) {  /* ... and what went back through it? */
            launch_probe_y(y_alias, y_off, (int)es, (size_t)prow * es, g_probe->y_val, s);
            y_probed = true;
        }
    }
    if (overlap)                                   /* the rest of x follows the launch */
        for (int k = x_sent; k < e.x_nchunks; ++k) send_chunk(k);
    /* the GPU is busy: now the content checks of the cache hit */
    const bool matrix_stale = check_content && entry_content_stale(e);
    float kernel_ms = 0.f;
    for (int p = 0; p < e.nparts; ++p)
        CUDA_OK(cudaStreamSynchronize(e.part[p].ctx->stream));
    if (overlap) CUDA_OK(cudaStreamSynchronize(e.part[0].ctx->copy_stream));
    if (timed_part >= 0)
        CUDA_OK(cudaEventElapsedTime(&kernel_ms, e.part[timed_part].ctx->ev0, e.part[timed_part].ctx->ev1));
    use(saved_dev);
    if (overlap) {
        g_stats.x_overlapped_calls++;
        if (g_probe->x_timed_out) {
            g_stats.x_overlap_timeouts++;
            g_x_overlap = 0;
            fprintf(stderr, "libb200-spmv: the product waited %llu ms for a chunk of x that the copy stream "
                            "did not deliver (streams serialised?); uploading x before the product from now on\n",
                    g_x_timeout_ns / 1000000ull);
            return CALL_REDO_PLAIN;
        }
    }
    if (matrix_stale) return CALL_STALE_MATRIX;
    /* an auto-registered range whose owner has remapped it: the GPU read / wrote the OLD
     * pages.  Revoke the registration and have the call redone through the bounce buffer. */
    bool stale = false;
    if (x_auto && g_probe->x_bad) { auto_pin_revoke(iv); stale = true; }
    if (y_probed) {
        const size_t first_rows = (size_t)(e.part[0].row_hi - e.part[0].row_lo) * es;
        for (int k = 0; k < 4 && !stale; ++k) {
            if (y_off[k] >= first_rows) continue;       /* sampled inside the first block only */
            unsigned long long host_view = 0;
            memcpy(&host_view, (const char *)ov + y_off[k], es);
            if (host_view != g_probe->y_val[k]) { auto_pin_revoke(ov); stale = true; }
        }
    }
    if (stale) { g_stats.auto_pin_revoked++; return CALL_REDO_BOUNCE; }
    if (x_auto || y_auto) g_stats.auto_pinned_calls++;
    if (!y_direct) memcpy(ov, e.h_y, (size_t)n * es);
    g_stats.kernel_ms += kernel_ms;
    g_stats.kernel_launches += (uint64_t)launched;
    g_stats.h2d_bytes += x_used;
    g_stats.d2h_bytes += (size_t)n * es;
    return CALL_OK;
}

static void harness_common(void *ov, const void *a, const void *iv, const int *rowstr,
                           const int *colidx, const int *rows, int dtype)
{
    pthread_mutex_lock(&g_lock);
    ensure_conf_locked();
    const int n = *rows;
    bool unchecked = false;
    CacheEntry *ep = lookup_locked(a, rowstr, colidx, n, dtype, &unchecked);
    const double t0 = now_ms();
    if (n > 0) {
        bool allow_auto = true;
        for (int attempt = 0; ; ++attempt) {
            const CallResult r = run_call(*ep, ov, iv, n, dtype, allow_auto, unchecked);
            if (r == CALL_OK) break;
            if (attempt >= 3) die("the call could not be completed (result %d)", (int)r);
            if (r == CALL_REDO_BOUNCE) allow_auto = false;
            if (r == CALL_STALE_MATRIX) {               /* the product ran on the old copy: upload, redo */
                drop_entry_locked(ep);
                ep = lookup_locked(a, rowstr, colidx, n, dtype, &unchecked);
            }
        }
    } else if (unchecked && entry_content_stale(*ep)) {
        drop_entry_locked(ep);
        ep = lookup_locked(a, rowstr, colidx, n, dtype, &unchecked);
    }
    g_stats.calls++;
    g_stats.e2e_ms += now_ms() - t0;
    pthread_mutex_unlock(&g_lock);
}

extern "C" void *spmv_harness_(double *ov, double *a, double *iv, int *rowstr, int *colidx, int *rows)
{
    harness_common(ov, a, iv, rowstr, colidx, rows, B200_F64);
    return NULL;
}

extern "C" void *f_spmv_harness_(float *ov, float *a, float *iv, int *rowstr, int *colidx, int *rows)
{
    harness_common(ov, a, iv, rowstr, colidx, rows, B200_F32);
    return NULL;
}

extern "C" void b200_spmv_invalidate(void)
{
    pthread_mutex_lock(&g_lock);
    for (CacheEntry &e : g_cache) release_entry_locked(e);
    g_cache.clear();
    pthread_mutex_unlock(&g_lock);
}

extern "C" int b200_spmv_devices_in_use(void)
{
    pthread_mutex_lock(&g_lock);
    int n = 0;
    for (const CacheEntry &e : g_cache) n = std::max(n, e.nparts);
    pthread_mutex_unlock(&g_lock);
    return n;
}

extern "C" void b200_spmv_get_stats(b200_spmv_stats *out)
{
    pthread_mutex_lock(&g_lock);
    *out = g_stats;
    pthread_mutex_unlock(&g_lock);
}

extern "C" void b200_spmv_reset_stats(void)
{
    pthread_mutex_lock(&g_lock);
    memset(&g_stats, 0, sizeof g_stats);
    pthread_mutex_unlock(&g_lock);
}

/* the bounce-buffer copy by itself (host only; tests/test_abi_surface.py checks it against
 * memcpy for every alignment without a GPU) */
extern "C" void b200_spmv_host_copy(void *dst, const void *src, size_t bytes)
{
    copy_to_bounce(dst, src, bytes);
}

extern "C" void b200_spmv_set_time_kernels(int on)
{
    pthread_mutex_lock(&g_lock);
    ensure_conf_locked();
    g_time_kernels = on ? 1 : 0;
    pthread_mutex_unlock(&g_lock);
}

extern "C" void b200_spmv_set_auto_pin(int sightings)
{
    pthread_mutex_lock(&g_lock);
    ensure_conf_locked();
    g_auto_pin = sightings > 0 ? sightings : 0;
    if (g_auto_pin == 0) {                       /* switched off: give everything back */
        for (const RegSpan &sp : g_spans) { ++g_reg_events; cudaHostUnregister(sp.lo); cudaGetLastError(); }
        g_spans.clear();
        g_auto.clear();
    }
    pthread_mutex_unlock(&g_lock);
}

extern "C" int b200_spmv_pin_host(void *ptr, size_t bytes)
{
    pthread_mutex_lock(&g_lock);
    ensure_conf_locked();
    const int rc = register_range_locked(ptr, bytes);
    pthread_mutex_unlock(&g_lock);
    return rc;
}

extern "C" int b200_spmv_unpin_host(void *ptr)
{
    int rc = -1;
    pthread_mutex_lock(&g_lock);
    for (size_t i = 0; i < g_pinned.size(); ++i)
        if (g_pinned[i].lo == (char *)ptr) {
            ++g_reg_events;
            cudaHostUnregister(ptr);
            g_pinned[i] = g_pinned.back();
            g_pinned.pop_back();
            rc = 0;
            break;
        }
    pthread_mutex_unlock(&g_lock);
    return rc;
}
