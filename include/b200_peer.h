/*
 * b200_peer.h -- fused compute + exchange kernels over NVLink peer memory for
 * the row-block sharded, device-resident CG.
 *
 * The exchange steps of the sharded path are (1) re-assembling the full
 * vector from the ranks' slices before a product and (2) completing a dot
 * product.  With NCCL each is a separate collective launch after the kernel
 * that produced the data.  Here the producing kernel does the exchange itself:
 *
 *   p = r + beta p        stores every new p element into the x buffer of
 *                         EVERY rank (plain st.global on peer-mapped pointers,
 *                         NVLink 5 / NVSwitch), so the "allgather" overlaps the
 *                         update element by element and the next product only
 *                         waits on a flag;
 *   dot products          each rank reduces its slice, writes the partial into
 *                         slot [rank] of every peer, and the consumer kernel
 *                         sums the slots in rank order -- the same value, bit
 *                         for bit, on every rank, without an allreduce launch.
 *
 * One process per GPU; buffers are cudaMalloc'ed by this library and shared
 * with cudaIpc handles that the host side exchanges (torch.distributed is the
 * plumbing for that one-off exchange only).  Completion is signalled with
 * monotonically increasing epochs written with system-scope release stores.
 * The reference has no multi-device path at all (libspmv/gpu.c is single
 * device); the algebra is NPB3.3.1/CG/cg.f:447-644.
 */
#ifndef B200_PEER_H
#define B200_PEER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_PEER_MAX_RANKS 8
#define B200_PEER_SLOTS     8      /* scalar exchange slots */
#define B200_IPC_HANDLE_BYTES 64

typedef struct b200_peer_group b200_peer_group;

/* Step 1 (every rank): allocate the local symmetric segment -- a vector of
 * `n_global` doubles plus the scalar slots and flags -- and export its IPC
 * handle (B200_IPC_HANDLE_BYTES bytes). */
b200_peer_group *b200_peer_create(int rank, int nranks, int64_t n_global, void *ipc_handle_out);
/* Step 2: after the handles of all ranks have been exchanged (rank order),
 * map the peers.  handles = nranks * B200_IPC_HANDLE_BYTES bytes. */
int b200_peer_connect(b200_peer_group *g, const void *handles);
void b200_peer_destroy(b200_peer_group *g);

/* the local full-length vector that products read (device pointer) */
double *b200_peer_xfull(b200_peer_group *g);

/* --- kernels; all asynchronous on `stream` ------------------------------- */
/* copy the local slice v[0..n_local) to offset `lo` of every rank's x buffer,
 * then publish epoch `e` on the vector flag */
void b200_peer_push(b200_peer_group *g, const double *v, int n_local, int64_t lo, uint64_t e, void *stream);
/* the same, but first wait until every rank has reported (b200_peer_consumed)
 * that it finished reading vector epoch e_consumed -- for back-to-back
 * products without a scalar exchange in between */
void b200_peer_push_after(b200_peer_group *g, const double *v, int n_local, int64_t lo, uint64_t e,
                          uint64_t e_consumed, void *stream);
/* the whole exchange of one of a series of back-to-back products in ONE launch
 * (epochs e = 1, 2, 3, ... on the stream that also runs the products): report
 * epoch e - 1 as consumed, wait for every rank's report, push the slice, publish
 * epoch e and retire only when every rank's epoch e has arrived */
void b200_peer_exchange(b200_peer_group *g, const double *v, int n_local, int64_t lo, uint64_t e,
                        void *stream);
/* The exchange half of a product that waits for its x slices itself
 * (b200_spmv_exec_sliced): epochs e = 1, 2, 3, ... alternate between two vector buffers
 * (b200_peer_xbuf(g, e)); report epoch e - 1 as consumed, wait for every rank's report of
 * e - 2 (a whole step old: a slow rank does not hold the others back), push the slice,
 * publish epoch e on the vector flags (b200_peer_vflags).  Does not wait for arrivals. */
void b200_peer_post(b200_peer_group *g, const double *v, int n_local, int64_t lo, uint64_t e,
                    void *stream);
/* the local full-length vector of epoch e, and this rank's row of vector flags
 * (flag[r] >= e: the slice of rank r of epoch e has arrived); device pointers */
double *b200_peer_xbuf(b200_peer_group *g, uint64_t e);
const unsigned long long *b200_peer_vflags(b200_peer_group *g);
/* report that this rank's product has consumed vector epoch e */
void b200_peer_consumed(b200_peer_group *g, uint64_t e, void *stream);
/* block until every rank has published epoch >= e on the vector flag */
void b200_peer_wait_vector(b200_peer_group *g, uint64_t e, void *stream);
/* slot <- sum_i x[i]*y[i] (mode 0) or sum_i (x[i]-y[i])^2 (mode 1) of the local
 * slices, published to every rank with epoch e */
void b200_peer_dot(b200_peer_group *g, const double *x, const double *y, int n_local, int mode,
                   int slot, uint64_t e, void *stream);
/* waits for slot_d at epoch e_d; alpha = sum(slot_rho)/sum(slot_d);
 * z += alpha p; r -= alpha q; slot_out <- sum r*r, epoch e_out */
void b200_peer_update_zr(b200_peer_group *g, double *z, double *r, const double *p, const double *q,
                         int n_local, int slot_rho, int slot_d, uint64_t e_d, int slot_out,
                         uint64_t e_out, void *stream);
/* waits for slot_new at epoch e_new; beta = sum(slot_new)/sum(slot_old);
 * p = r + beta p, every new element also stored to offset lo of every rank's
 * x buffer; vector flag <- e_vec */
void b200_peer_update_p(b200_peer_group *g, double *p, const double *r, int n_local, int64_t lo,
                        int slot_new, uint64_t e_new, int slot_old, uint64_t e_vec, void *stream);
/* waits for slot at epoch e; x = z / sqrt(sum(slot)) */
void b200_peer_scale(b200_peer_group *g, double *x, const double *z, int n_local, int slot,
                     uint64_t e, void *stream);
/* waits for each listed slot at its epoch and writes the rank-ordered sums to
 * out[0..count) (device) */
void b200_peer_read_slots(b200_peer_group *g, const int *slots, const uint64_t *epochs, int count,
                          double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif
