"""Row-block sharded SpMV across the GPUs of one box.

The reference is single-process / single-device (libspmv/gpu.c); this is new
ground required by the north star: contiguous row blocks, one per rank, each
rank keeps its block resident and owns the matching slice of every vector.
The one exchange step of the path is re-assembling the full x from the ranks'
slices before a product -- an allgather -- after which the product itself is
rank-local.  One process per GPU; `torch.distributed` (NCCL over NVLink /
NVSwitch on the GPU box, gloo in the CPU tests) is the plumbing.

The local product is injected (`local_spmv`) so the same host logic runs on a
GPU (lilac_benchmarks_b200.libspmv.ResidentMatrix.exec) and in the world_size-2
gloo tests on a CPU box.
"""
from dataclasses import dataclass

import numpy as np


def equal_row_bounds(rows, parts):
    """Equal-count contiguous row blocks (the last ones may be one row shorter).
    NPB CG rows are near-uniform (row-length CV 0.21-0.24), so equal rows are
    within ~1 % of nnz balance and allow the plain equal-count allgather."""
    per = -(-rows // parts)
    b = np.minimum(np.arange(parts + 1, dtype=np.int64) * per, rows)
    return b.astype(np.int32)


def block_imbalance(rowstr, bounds):
    """max block nnz / mean block nnz."""
    per = np.diff(rowstr[np.asarray(bounds)].astype(np.int64))
    return float(per.max() / max(per.mean(), 1.0))


@dataclass
class ShardLayout:
    rows: int            # global rows (= global length of x and y; square operator)
    parts: int
    bounds: np.ndarray   # parts+1 row boundaries
    slot: int            # padded slice length used by the equal-count allgather

    @classmethod
    def build(cls, rows, parts, bounds=None):
        bounds = equal_row_bounds(rows, parts) if bounds is None else np.asarray(bounds, dtype=np.int32)
        assert bounds[0] == 0 and bounds[-1] == rows and len(bounds) == parts + 1
        slot = int(np.diff(bounds).max()) if parts else 0
        return cls(rows=rows, parts=parts, bounds=bounds, slot=slot)

    def local_range(self, rank):
        return int(self.bounds[rank]), int(self.bounds[rank + 1])

    @property
    def contiguous(self):
        """True when the padded slots tile x without gaps (every block but the
        last is exactly `slot` rows), so the gathered buffer IS x."""
        d = np.diff(self.bounds)
        return bool(np.all(d[:-1] == self.slot)) if self.parts > 1 else True


class ShardedSpmv:
    """y_local = A[rows of this rank, :] @ allgather(x_local)."""

    def __init__(self, layout, rank, local_spmv, dist=None, device=None, dtype=None):
        import torch
        self.torch = torch
        self.layout = layout
        self.rank = rank
        self.local_spmv = local_spmv          # (x_full_tensor, y_local_tensor) -> None
        self.dist = dist
        self.lo, self.hi = layout.local_range(rank)
        dtype = dtype or torch.float64
        device = device or "cpu"
        # gathered buffer: parts slots of `slot` elements; with a contiguous
        # layout its first `rows` elements are x itself
        self.x_gather = torch.zeros(layout.parts * layout.slot, dtype=dtype, device=device)
        self.x_full = (self.x_gather if layout.contiguous
                       else torch.zeros(layout.rows, dtype=dtype, device=device))
        self.x_slot = torch.zeros(layout.slot, dtype=dtype, device=device)
        self.y_local = torch.zeros(self.hi - self.lo, dtype=dtype, device=device)

    def assemble_x(self, x_local):
        """The exchange step: allgather of the ranks' x slices."""
        n_local = self.hi - self.lo
        self.x_slot[:n_local].copy_(x_local[:n_local])
        if self.dist is None or self.layout.parts == 1:
            self.x_gather[: self.layout.slot].copy_(self.x_slot)
        else:
            self.dist.all_gather_into_tensor(self.x_gather, self.x_slot)
        if not self.layout.contiguous:
            for p in range(self.layout.parts):
                lo, hi = self.layout.local_range(p)
                self.x_full[lo:hi].copy_(self.x_gather[p * self.layout.slot: p * self.layout.slot + hi - lo])
        return self.x_full

    def step(self, x_local):
        """One sharded product: exchange, then the rank-local kernel."""
        x_full = self.assemble_x(x_local)
        self.local_spmv(x_full, self.y_local)
        return self.y_local


# ---------------------------------------------------------------------------
# Device-resident NPB CG over row blocks (SURVEY.md section 8e "device-resident
# CG mode", section 8f row 1): every rank keeps its slices of x, z, p, q, r in
# HBM; per CG iteration the ranks allgather their p slices, run the local
# product, and complete the two dot products with an allreduce of one scalar.
# Follows NPB3.3.1/CG/cg.f:447-644 (conj_grad) and :299-349 (outer loop).
# ---------------------------------------------------------------------------
class TorchVectorOps:
    """Vector algebra of conj_grad with plain torch ops (CPU gloo tests)."""

    def dot(self, x, y):
        return (x * y).sum().reshape(1)

    def update_zr(self, z, r, p, q, rho, d):
        alpha = rho / d
        z.add_(alpha * p)
        r.sub_(alpha * q)
        return (r * r).sum().reshape(1)

    def update_p(self, p, r, rho_new, rho_old):
        beta = rho_new / rho_old
        p.mul_(beta).add_(r)


class B200VectorOps:
    """The same through the fused CUDA kernels of include/b200_cg.h (torch only
    owns the memory and the stream)."""

    def __init__(self, libspmv_module, device):
        import torch
        self.torch = torch
        self.L = libspmv_module.lib()
        nb = self.L.b200_cg_partials()
        self.partial = torch.zeros(nb, dtype=torch.float64, device=device)
        self.device = device

    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def _finish(self):
        out = self.torch.empty(1, dtype=self.torch.float64, device=self.device)
        self.L.b200_cg_finish(self.partial.data_ptr(), out.data_ptr(), self._stream())
        return out

    def dot(self, x, y):
        self.L.b200_cg_dot(x.data_ptr(), y.data_ptr(), x.numel(), self.partial.data_ptr(), self._stream())
        return self._finish()

    def update_zr(self, z, r, p, q, rho, d):
        self.L.b200_cg_update_zr(z.data_ptr(), r.data_ptr(), p.data_ptr(), q.data_ptr(), z.numel(),
                                 rho.data_ptr(), d.data_ptr(), self.partial.data_ptr(), self._stream())
        return self._finish()

    def update_p(self, p, r, rho_new, rho_old):
        self.L.b200_cg_update_p(p.data_ptr(), r.data_ptr(), p.numel(), rho_new.data_ptr(),
                                rho_old.data_ptr(), self._stream())


class ShardedNpbCg:
    """NPB CG with row-block sharded, device-resident vectors."""

    def __init__(self, sharded_spmv, ops, shift, cgitmax=25):
        self.sp = sharded_spmv
        self.ops = ops
        self.shift = shift
        self.cgitmax = cgitmax
        t = sharded_spmv.torch
        n_local = sharded_spmv.hi - sharded_spmv.lo
        dev, dt = sharded_spmv.y_local.device, sharded_spmv.y_local.dtype
        self.x, self.z, self.p, self.q, self.r = (t.zeros(n_local, dtype=dt, device=dev) for _ in range(5))
        self.spmv_count = 0
        self.collectives = 0

    def _allreduce(self, v):
        if self.sp.dist is not None and self.sp.layout.parts > 1:
            self.sp.dist.all_reduce(v)
            self.collectives += 1
        return v

    def _product(self, src, dst):
        dst.copy_(self.sp.step(src))
        self.spmv_count += 1
        if self.sp.layout.parts > 1:
            self.collectives += 1

    def conj_grad(self):
        """cg.f:447-644; returns ||x - A z|| as a 1-element tensor."""
        ops = self.ops
        self.q.zero_()
        self.z.zero_()
        self.r.copy_(self.x)
        self.p.copy_(self.r)
        rho = self._allreduce(ops.dot(self.r, self.r))
        for _ in range(self.cgitmax):
            self._product(self.p, self.q)                              # q = A p
            d = self._allreduce(ops.dot(self.p, self.q))
            rho_new = self._allreduce(ops.update_zr(self.z, self.r, self.p, self.q, rho, d))
            ops.update_p(self.p, self.r, rho_new, rho)
            rho = rho_new
        self._product(self.z, self.r)                                  # r = A z
        diff = self.x - self.r
        return self._allreduce(ops.dot(diff, diff)).sqrt()

    def run(self, niter, untimed_first=True, sync=None):
        """cg.f:216-352.  Returns (zeta history, rnorm history, seconds of the timed loop)."""
        import time
        t = self.sp.torch
        hist_z, hist_r = [], []

        def outer():
            rnorm = self.conj_grad()
            nt = t.cat([self.ops.dot(self.x, self.z), self.ops.dot(self.z, self.z)])
            nt = self._allreduce(nt)
            self.x.copy_(self.z / nt[1].sqrt())
            return rnorm, nt

        if untimed_first:
            self.x.fill_(1.0)
            outer()
        self.x.fill_(1.0)
        self.spmv_count = 0
        self.collectives = 0
        if sync:
            sync()
        t0 = time.perf_counter()
        for _ in range(niter):
            rnorm, nt = outer()
            vals = t.cat([rnorm, nt]).cpu()          # one small device->host read per outer iteration
            hist_r.append(float(vals[0]))
            hist_z.append(self.shift + 1.0 / float(vals[1]))
        if sync:
            sync()
        return hist_z, hist_r, time.perf_counter() - t0


# ---------------------------------------------------------------------------
# The same CG with the exchanges fused into the producing kernels over NVLink
# peer memory (include/b200_peer.h): no NCCL call and no host synchronisation
# inside conj_grad.  torch.distributed is used once, to swap the IPC handles.
# ---------------------------------------------------------------------------
class PeerNpbCg:
    RHO = (0, 1)       # ping-pong slots for rho
    D, XZ, ZZ, RES = 2, 3, 4, 5

    def __init__(self, libspmv_module, resident_matrix, layout, rank, shift, dist=None, device="cuda",
                 cgitmax=25):
        import ctypes as C
        import torch
        self.torch, self.C = torch, C
        self.L = libspmv_module.lib()
        self.rm = resident_matrix
        self.layout, self.rank, self.shift, self.cgitmax = layout, rank, shift, cgitmax
        self.lo, self.hi = layout.local_range(rank)
        self.n_local = self.hi - self.lo
        world = layout.parts
        handle = (C.c_ubyte * 64)()
        self.g = self.L.b200_peer_create(rank, world, layout.rows, handle)
        if not self.g:
            raise RuntimeError("b200_peer_create failed")
        if world > 1:
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
            allh = torch.empty(64 * world, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allh, mine)
            buf = (C.c_ubyte * (64 * world))(*allh.cpu().tolist())
            if self.L.b200_peer_connect(self.g, buf) != 0:
                raise RuntimeError("b200_peer_connect failed (no peer access between the GPUs?)")
            dist.barrier()
        self.xfull = self.L.b200_peer_xfull(self.g)
        mk = lambda: torch.zeros(self.n_local, dtype=torch.float64, device=device)   # noqa: E731
        self.x, self.z, self.p, self.q, self.r = mk(), mk(), mk(), mk(), mk()
        self.out = torch.zeros(4, dtype=torch.float64, device=device)
        self.epoch = 0
        self.spmv_count = 0
        self.dist = dist

    def calibrate(self, x_local, steps=20):
        """Time the forms this matrix supports on the actual GPUs (max over ranks, CUDA events)
        and keep the fastest; which one wins depends on the rank count and the block shape.
        Returns {form: ms per step}."""
        torch = self.torch
        forms = [("blocking", False, False)]
        if self.overlap:
            forms.append(("overlapped", True, False))
        if self.fused:
            forms.append(("fused", True, True))
        best, times = None, {}
        for name, ov, fu in forms:
            self.overlap, self.fused = ov, fu
            for _ in range(3):
                self.step(x_local)
            torch.cuda.synchronize()
            if self.dist is not None and self.layout.parts > 1:
                self.dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                self.step(x_local)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=self.y_local.device)
            if self.dist is not None and self.layout.parts > 1:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            times[name] = float(t.item())
            if best is None or times[name] < times[best[0]]:
                best = (name, ov, fu)
        self.overlap, self.fused = best[1], best[2]
        self.form = best[0]
        return times

    def close(self):
        if self.g:
            self.torch.cuda.synchronize()
            if self.dist is not None and self.layout.parts > 1:
                self.dist.barrier()
            self.L.b200_peer_destroy(self.g)
            self.g = None

    def _s(self):
        return self.torch.cuda.current_stream().cuda_stream

    def _next(self):
        self.epoch += 1
        return self.epoch

    def _spmv_from_xfull(self, e_vec, dst):
        self.L.b200_peer_wait_vector(self.g, e_vec, self._s())
        self.rm.exec_ptr(self.xfull, dst.data_ptr(), self._s())
        self.spmv_count += 1

    def outer_iteration(self):
        """conj_grad (cg.f:447-644) + the zeta / normalisation step (cg.f:315-346);
        leaves (x.z, residual sum) in self.out."""
        L, g, n, lo, s = self.L, self.g, self.n_local, self.lo, self._s
        x, z, p, q, r = self.x, self.z, self.p, self.q, self.r
        q.zero_(); z.zero_(); r.copy_(x); p.copy_(r)
        L.b200_peer_dot(g, r.data_ptr(), r.data_ptr(), n, 0, self.RHO[0], self._next(), s())
        e_vec = self._next()
        L.b200_peer_push(g, p.data_ptr(), n, lo, e_vec, s())
        for it in range(self.cgitmax):
            rho_old, rho_new = self.RHO[it & 1], self.RHO[(it + 1) & 1]
            self._spmv_from_xfull(e_vec, q)                                        # q = A p
            e_d = self._next()
            L.b200_peer_dot(g, p.data_ptr(), q.data_ptr(), n, 0, self.D, e_d, s())
            e_rho = self._next()
            L.b200_peer_update_zr(g, z.data_ptr(), r.data_ptr(), p.data_ptr(), q.data_ptr(), n,
                                  rho_old, self.D, e_d, rho_new, e_rho, s())
            e_vec = self._next()
            L.b200_peer_update_p(g, p.data_ptr(), r.data_ptr(), n, lo, rho_new, e_rho, rho_old, e_vec, s())
        e_vec = self._next()
        L.b200_peer_push(g, z.data_ptr(), n, lo, e_vec, s())
        self._spmv_from_xfull(e_vec, r)                                            # r = A z
        e_res, e_xz, e_zz = self._next(), self._next(), self._next()
        L.b200_peer_dot(g, x.data_ptr(), r.data_ptr(), n, 1, self.RES, e_res, s())
        L.b200_peer_dot(g, x.data_ptr(), z.data_ptr(), n, 0, self.XZ, e_xz, s())
        L.b200_peer_dot(g, z.data_ptr(), z.data_ptr(), n, 0, self.ZZ, e_zz, s())
        C = self.C
        slots = (C.c_int * 2)(self.XZ, self.RES)
        epochs = (C.c_uint64 * 2)(e_xz, e_res)
        L.b200_peer_read_slots(g, slots, epochs, 2, self.out.data_ptr(), s())
        L.b200_peer_scale(g, x.data_ptr(), z.data_ptr(), n, self.ZZ, e_zz, s())

    def run(self, niter, sync=None):
        import time
        hist_z, hist_r = [], []
        self.x.fill_(1.0)
        self.outer_iteration()                     # untimed (cg.f:233-272)
        self.x.fill_(1.0)
        self.spmv_count = 0
        if sync:
            sync()
        t0 = time.perf_counter()
        for _ in range(niter):
            self.outer_iteration()
            vals = self.out.cpu()                  # the only host read per outer iteration
            hist_z.append(self.shift + 1.0 / float(vals[0]))
            hist_r.append(float(vals[1]) ** 0.5)
        if sync:
            sync()
        return hist_z, hist_r, time.perf_counter() - t0


class PeerShardedSpmv:
    """y_local = A[rows of this rank, :] @ x with the exchange done by this rank's own
    kernels over NVLink peer memory (include/b200_peer.h) instead of an NCCL allgather.

    Overlapped form (the RING kernel): `b200_peer_post` pushes my slice into every
    rank's buffer of this epoch (two buffers alternate, so nobody waits for a slow
    rank's previous product) and publishes the epoch; the product then waits per slice,
    just before the column panels that need it (`b200_spmv_exec_sliced`), i.e. it starts
    on the slices that have arrived while the others are still in flight.
    Fused form (row blocks of at most one CTA per SM): the product kernel does the push
    itself in its prologue (`b200_spmv_exec_pushed`) -- one launch per step.
    Other kernels: one exchange kernel that also waits for every rank's slice, then the
    product."""

    def __init__(self, libspmv_module, resident_matrix, layout, rank, dist=None, device="cuda",
                 overlap=True, fused=True):
        import ctypes as C
        import torch
        self.torch = torch
        self.L = libspmv_module.lib()
        self.rm, self.layout, self.rank, self.dist = resident_matrix, layout, rank, dist
        self.lo, self.hi = layout.local_range(rank)
        world = layout.parts
        handle = (C.c_ubyte * 64)()
        self.g = self.L.b200_peer_create(rank, world, layout.rows, handle)
        if not self.g:
            raise RuntimeError("b200_peer_create failed")
        if world > 1:
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
            allh = torch.empty(64 * world, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allh, mine)
            buf = (C.c_ubyte * (64 * world))(*allh.cpu().tolist())
            if self.L.b200_peer_connect(self.g, buf) != 0:
                raise RuntimeError("b200_peer_connect failed")
            dist.barrier()
        self.xfull = self.L.b200_peer_xfull(self.g)
        self.vflags = self.L.b200_peer_vflags(self.g)
        self.y_local = torch.zeros(self.hi - self.lo, dtype=torch.float64, device=device)
        self.epoch = 0
        # every rank must take the same branch: the overlapped form needs a kernel that can
        # wait in-kernel on every rank and slices that are whole multiples of `slot` columns
        ok = 1 if (overlap and resident_matrix.waits_in_kernel and layout.contiguous) else 0
        if dist is not None and world > 1:
            t = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = int(t.item())
        self.overlap = bool(ok)
        # fused form: the product kernel pushes the slice itself (one launch per step); needs
        # every rank's row blocks co-resident (one CTA per SM) and 16-byte slice boundaries
        okf = 1 if (self.overlap and fused and resident_matrix.can_push and
                    all(int(b) % 2 == 0 for b in layout.bounds)) else 0
        if dist is not None and world > 1:
            t = torch.tensor([okf], dtype=torch.int32, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            okf = int(t.item())
        self.fused = bool(okf)
        self.form = "fused" if self.fused else "overlapped" if self.overlap else "blocking"

    def step(self, x_local):
        s = self.torch.cuda.current_stream().cuda_stream
        self.epoch += 1
        e = self.epoch
        if self.fused:
            rc = self.L.b200_spmv_exec_pushed(self.rm._h, self.y_local.data_ptr(), s, self.g,
                                              x_local.data_ptr(), self.hi - self.lo, self.lo, e,
                                              self.layout.slot)
            if rc < 0:
                raise RuntimeError("b200_spmv_exec_pushed refused a matrix that reported can_push")
        elif self.overlap:
            self.L.b200_peer_post(self.g, x_local.data_ptr(), self.hi - self.lo, self.lo, e, s)
            self.rm.exec_sliced_ptr(self.L.b200_peer_xbuf(self.g, e), self.y_local.data_ptr(), s,
                                    self.vflags, e, self.layout.slot, self.layout.parts)
        else:
            # one launch: consumed(previous) -> wait -> push -> publish -> wait for everybody
            self.L.b200_peer_exchange(self.g, x_local.data_ptr(), self.hi - self.lo, self.lo, e, s)
            self.rm.exec_ptr(self.xfull, self.y_local.data_ptr(), s)
        return self.y_local

    def calibrate(self, x_local, steps=20):
        """Time the forms this matrix supports on the actual GPUs (max over ranks, CUDA events)
        and keep the fastest; which one wins depends on the rank count and the block shape.
        Returns {form: ms per step}."""
        torch = self.torch
        forms = [("blocking", False, False)]
        if self.overlap:
            forms.append(("overlapped", True, False))
        if self.fused:
            forms.append(("fused", True, True))
        best, times = None, {}
        for name, ov, fu in forms:
            self.overlap, self.fused = ov, fu
            for _ in range(3):
                self.step(x_local)
            torch.cuda.synchronize()
            if self.dist is not None and self.layout.parts > 1:
                self.dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                self.step(x_local)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=self.y_local.device)
            if self.dist is not None and self.layout.parts > 1:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            times[name] = float(t.item())
            if best is None or times[name] < times[best[0]]:
                best = (name, ov, fu)
        self.overlap, self.fused = best[1], best[2]
        self.form = best[0]
        return times

    def close(self):
        if self.g:
            self.torch.cuda.synchronize()
            if self.dist is not None and self.layout.parts > 1:
                self.dist.barrier()
            self.L.b200_peer_destroy(self.g)
            self.g = None
