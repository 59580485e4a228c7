"""ctypes mirror of include/b200_spmv.h.

`spmv_harness` / `f_spmv_harness` take numpy arrays and call the drop-in
symbols exactly as the reference's callers do (libspmv/test.cpp:52,
NPB3.3.1/CG/cg.f:531-532): 1-based `rowstr` / `colidx`, `rows` by reference,
result written into `ov`.  `ResidentMatrix` wraps the device-pointer API used
by bench.py and the row-block (multi-GPU) path.

There is no fallback: if csrc/b200.so is missing this module raises, and on a
box without a CUDA device the library itself aborts.
"""
import ctypes as C
from ctypes import POINTER, c_double, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p, c_char_p

import numpy as np

from .build import B200_SO

KERNEL_AUTO, KERNEL_ORDERED, KERNEL_VECTOR, KERNEL_PANEL, KERNEL_MERGE = range(5)
KERNEL_IDS = {"auto": 0, "ordered": 1, "vector": 2, "panel": 3, "merge": 4, "sell": 5, "small": 6}
F64, F32 = 0, 1


class CgResult(C.Structure):
    _fields_ = [("zeta", c_double), ("rnorm", c_double), ("seconds", c_double), ("mops", c_double),
                ("spmv_launches", c_int), ("vector_launches", c_int)]


class NpbDeviceCsr(C.Structure):
    """b200_npb_csr of include/b200_npb.h."""
    _fields_ = [("rows", c_int), ("nnz", c_int64), ("d_rowstr", c_void_p), ("d_colidx", c_void_p),
                ("d_a", c_void_p)]


class Stats(C.Structure):
    _fields_ = [("calls", c_uint64), ("uploads", c_uint64), ("kernel_launches", c_uint64),
                ("kernel_ms", c_double), ("e2e_ms", c_double), ("upload_ms", c_double),
                ("h2d_bytes", c_uint64), ("d2h_bytes", c_uint64),
                ("auto_pinned_calls", c_uint64), ("auto_pin_revoked", c_uint64),
                ("x_overlapped_calls", c_uint64), ("x_overlap_timeouts", c_uint64)]


_lib = None


def lib():
    """Load csrc/b200.so (RTLD_GLOBAL not needed).  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not B200_SO.exists():
        raise RuntimeError(f"{B200_SO} is not built: run __graft_entry__.build() "
                           "(there is no CPU fallback for the b200 platform)")
    L = C.CDLL(str(B200_SO))
    dp, fp, ip = POINTER(c_double), POINTER(c_float), POINTER(c_int)
    L.spmv_harness_.argtypes = [dp, dp, dp, ip, ip, ip]
    L.spmv_harness_.restype = c_void_p
    L.f_spmv_harness_.argtypes = [fp, fp, fp, ip, ip, ip]
    L.f_spmv_harness_.restype = c_void_p
    L.b200_spmv_init.argtypes = [c_int]
    L.b200_spmv_init.restype = c_int
    L.b200_spmv_upload.argtypes = [c_void_p, ip, ip, c_int, c_int, c_int]
    L.b200_spmv_upload.restype = c_void_p
    L.b200_spmv_release.argtypes = [c_void_p]
    L.b200_spmv_release.restype = None
    L.b200_spmv_exec.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    L.b200_spmv_exec.restype = c_int
    L.b200_spmv_exec_sliced.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64,
                                        c_int, c_int]
    L.b200_spmv_exec_sliced.restype = c_int
    L.b200_spmv_exec_pushed.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64,
                                        c_uint64, c_int]
    L.b200_spmv_exec_pushed.restype = c_int
    L.b200_spmv_can_push.argtypes = [c_void_p]
    L.b200_spmv_can_push.restype = c_int
    L.b200_spmv_dot_partials.argtypes = [c_void_p]
    L.b200_spmv_dot_partials.restype = c_int
    L.b200_spmv_exec_dot.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.b200_spmv_exec_dot.restype = c_int
    L.b200_spmv_upload_device.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]
    L.b200_spmv_upload_device.restype = c_void_p
    L.b200_spmv_device.argtypes = [c_void_p]
    L.b200_spmv_device.restype = c_int
    L.b200_spmv_waits_in_kernel.argtypes = [c_void_p]
    L.b200_spmv_waits_in_kernel.restype = c_int
    L.b200_spmv_devices_in_use.argtypes = []
    L.b200_spmv_devices_in_use.restype = c_int
    for name, res in (("rows", c_int), ("ncols", c_int), ("nnz", c_int64), ("kernel", c_int),
                      ("kernel_name", c_char_p), ("launches_per_exec", c_int),
                      ("algorithmic_bytes", c_int64), ("resident_bytes", c_int64)):
        fn = getattr(L, f"b200_spmv_{name}")
        fn.argtypes = [c_void_p]
        fn.restype = res
    L.b200_spmv_row_histogram.argtypes = [c_void_p, POINTER(c_int64), ip, ip]
    L.b200_spmv_row_histogram.restype = None
    L.b200_spmv_partition_rows.argtypes = [ip, c_int, c_int, ip]
    L.b200_spmv_partition_rows.restype = None
    L.b200_spmv_invalidate.argtypes = []
    L.b200_spmv_invalidate.restype = None
    L.b200_spmv_get_stats.argtypes = [POINTER(Stats)]
    L.b200_spmv_get_stats.restype = None
    L.b200_spmv_reset_stats.argtypes = []
    L.b200_spmv_reset_stats.restype = None
    L.b200_spmv_set_auto_pin.argtypes = [c_int]
    L.b200_spmv_set_time_kernels.argtypes = [c_int]
    L.b200_spmv_set_time_kernels.restype = None
    L.b200_spmv_set_auto_pin.restype = None
    L.b200_spmv_pin_host.argtypes = [c_void_p, c_size_t]
    L.b200_spmv_pin_host.restype = c_int
    L.b200_spmv_unpin_host.argtypes = [c_void_p]
    L.b200_spmv_unpin_host.restype = c_int
    L.b200_spmv_version.argtypes = []
    L.b200_spmv_version.restype = c_char_p
    # include/b200_cg.h
    L.b200_cg_npb_run.argtypes = [c_void_p, c_int, c_int, c_double, c_int, POINTER(c_double),
                                  POINTER(c_double), POINTER(CgResult)]
    L.b200_cg_npb_run.restype = c_int
    L.b200_cg_partials.argtypes = []
    L.b200_cg_partials.restype = c_int
    L.b200_cg_dot.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    L.b200_cg_dot.restype = None
    L.b200_cg_update_zr.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p]
    L.b200_cg_update_zr.restype = None
    L.b200_cg_update_p.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    L.b200_cg_update_p.restype = None
    L.b200_cg_finish.argtypes = [c_void_p, c_void_p, c_void_p]
    L.b200_cg_finish.restype = None
    # include/b200_npb.h
    L.b200_npb_makea_device.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_double,
                                        c_double, c_int, c_int, POINTER(NpbDeviceCsr)]
    L.b200_npb_makea_device.restype = c_int
    L.b200_npb_csr_free.argtypes = [POINTER(NpbDeviceCsr)]
    L.b200_npb_csr_free.restype = None
    L.b200_npb_csr_to_host.argtypes = [POINTER(NpbDeviceCsr), ip, ip, dp]
    L.b200_npb_csr_to_host.restype = c_int
    # include/b200_peer.h
    u64 = C.c_uint64
    L.b200_peer_create.argtypes = [c_int, c_int, c_int64, c_void_p]
    L.b200_peer_create.restype = c_void_p
    L.b200_peer_connect.argtypes = [c_void_p, c_void_p]
    L.b200_peer_connect.restype = c_int
    L.b200_peer_destroy.argtypes = [c_void_p]
    L.b200_peer_destroy.restype = None
    L.b200_peer_xfull.argtypes = [c_void_p]
    L.b200_peer_xfull.restype = c_void_p
    L.b200_peer_push.argtypes = [c_void_p, c_void_p, c_int, c_int64, u64, c_void_p]
    L.b200_peer_push_after.argtypes = [c_void_p, c_void_p, c_int, c_int64, u64, u64, c_void_p]
    L.b200_peer_exchange.argtypes = [c_void_p, c_void_p, c_int, c_int64, u64, c_void_p]
    L.b200_peer_consumed.argtypes = [c_void_p, u64, c_void_p]
    L.b200_peer_wait_vector.argtypes = [c_void_p, u64, c_void_p]
    L.b200_peer_dot.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, u64, c_void_p]
    L.b200_peer_update_zr.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                      c_int, u64, c_int, u64, c_void_p]
    L.b200_peer_update_p.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, u64, c_int,
                                     u64, c_void_p]
    L.b200_peer_scale.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, u64, c_void_p]
    L.b200_peer_read_slots.argtypes = [c_void_p, POINTER(c_int), POINTER(u64), c_int, c_void_p, c_void_p]
    L.b200_peer_post.argtypes = [c_void_p, c_void_p, c_int, c_int64, u64, c_void_p]
    L.b200_peer_post.restype = None
    L.b200_peer_xbuf.argtypes = [c_void_p, u64]
    L.b200_peer_xbuf.restype = c_void_p
    L.b200_peer_vflags.argtypes = [c_void_p]
    L.b200_peer_vflags.restype = c_void_p
    for nm in ("push", "push_after", "consumed", "wait_vector", "dot", "update_zr", "update_p", "scale", "read_slots"):
        getattr(L, "b200_peer_" + nm).restype = None
    _lib = L
    return L


def _ptr(a, ctype):
    return a.ctypes.data_as(POINTER(ctype))


def _check(a, dtype, name):
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.c_contiguous:
        raise TypeError(f"{name} must be a C-contiguous numpy array of {np.dtype(dtype)}")


def spmv_harness(ov, a, iv, rowstr, colidx, rows):
    """y = A x through the drop-in fp64 symbol (libspmv/native.c:3-6)."""
    for arr, dt, nm in ((ov, np.float64, "ov"), (a, np.float64, "a"), (iv, np.float64, "iv"),
                        (rowstr, np.int32, "rowstr"), (colidx, np.int32, "colidx")):
        _check(arr, dt, nm)
    n = c_int(int(rows))
    lib().spmv_harness_(_ptr(ov, c_double), _ptr(a, c_double), _ptr(iv, c_double),
                        _ptr(rowstr, c_int), _ptr(colidx, c_int), C.byref(n))
    return ov


def f_spmv_harness(ov, a, iv, rowstr, colidx, rows):
    """y = A x through the drop-in fp32 symbol (libspmv/native.c:8-11)."""
    for arr, dt, nm in ((ov, np.float32, "ov"), (a, np.float32, "a"), (iv, np.float32, "iv"),
                        (rowstr, np.int32, "rowstr"), (colidx, np.int32, "colidx")):
        _check(arr, dt, nm)
    n = c_int(int(rows))
    lib().f_spmv_harness_(_ptr(ov, c_float), _ptr(a, c_float), _ptr(iv, c_float),
                          _ptr(rowstr, c_int), _ptr(colidx, c_int), C.byref(n))
    return ov


def harness_address(f32=False):
    """Raw address of the ABI symbol, for C callers that take a function pointer."""
    fn = lib().f_spmv_harness_ if f32 else lib().spmv_harness_
    return C.cast(fn, c_void_p).value


def exec_address():
    """Address of b200_spmv_exec (what a C caller of the resident-matrix API binds)."""
    return C.cast(lib().b200_spmv_exec, c_void_p).value


def stats():
    s = Stats()
    lib().b200_spmv_get_stats(C.byref(s))
    return {k: getattr(s, k) for k, _ in Stats._fields_}


def reset_stats():
    lib().b200_spmv_reset_stats()


def invalidate():
    lib().b200_spmv_invalidate()


def devices_in_use():
    """Most devices any matrix cached by the drop-in symbols is spread over."""
    return lib().b200_spmv_devices_in_use()


def partition_rows(rowstr, parts):
    """nnz-balanced contiguous row partition: parts+1 boundaries (0-based rows)."""
    _check(rowstr, np.int32, "rowstr")
    bounds = np.zeros(parts + 1, dtype=np.int32)
    lib().b200_spmv_partition_rows(_ptr(rowstr, c_int), len(rowstr) - 1, parts, _ptr(bounds, c_int))
    return bounds


class ResidentMatrix:
    """A CSR row block resident in HBM (b200_spmv_upload / _exec / _release)."""

    def __init__(self, a, rowstr, colidx, rows=None, kernel="auto", row_lo=0):
        dtype = a.dtype
        if dtype == np.float64:
            self.dtype, self.np_dtype = F64, np.float64
        elif dtype == np.float32:
            self.dtype, self.np_dtype = F32, np.float32
        else:
            raise TypeError("values must be float64 or float32")
        _check(a, dtype, "a")
        _check(rowstr, np.int32, "rowstr")
        _check(colidx, np.int32, "colidx")
        if rows is None:
            rows = len(rowstr) - 1 - row_lo
        self._keep = (a, rowstr, colidx)
        rs = rowstr[row_lo:]
        self._h = lib().b200_spmv_upload(a.ctypes.data_as(c_void_p), _ptr(rs, c_int),
                                         _ptr(colidx, c_int), int(rows), self.dtype,
                                         KERNEL_IDS[kernel] if isinstance(kernel, str) else int(kernel))
        if not self._h:
            raise RuntimeError("b200_spmv_upload failed")
        self._describe()

    @classmethod
    def from_device(cls, d_a, d_rowstr, d_colidx, rows, np_dtype=np.float64, kernel="auto", keep=None):
        """Upload from DEVICE arrays (raw pointers as ints) of the current device; same 1-based
        contents as the ABI.  `keep` holds whatever owns that memory until the upload is done."""
        self = cls.__new__(cls)
        self.dtype, self.np_dtype = (F64, np.float64) if np_dtype == np.float64 else (F32, np.float32)
        self._keep = keep
        self._h = lib().b200_spmv_upload_device(c_void_p(d_a), c_void_p(d_rowstr), c_void_p(d_colidx),
                                                int(rows), self.dtype,
                                                KERNEL_IDS[kernel] if isinstance(kernel, str) else int(kernel))
        if not self._h:
            raise RuntimeError("b200_spmv_upload_device failed")
        self._describe()
        self._keep = None
        return self

    def _describe(self):
        L = lib()
        self.rows = L.b200_spmv_rows(self._h)
        self.ncols = L.b200_spmv_ncols(self._h)
        self.nnz = L.b200_spmv_nnz(self._h)
        self.kernel_name = L.b200_spmv_kernel_name(self._h).decode()
        self.launches_per_exec = L.b200_spmv_launches_per_exec(self._h)
        self.algorithmic_bytes = L.b200_spmv_algorithmic_bytes(self._h)
        self.resident_bytes = L.b200_spmv_resident_bytes(self._h)
        self.device = L.b200_spmv_device(self._h)
        self.waits_in_kernel = bool(L.b200_spmv_waits_in_kernel(self._h))
        self.can_push = bool(L.b200_spmv_can_push(self._h))
        self.dot_partials = L.b200_spmv_dot_partials(self._h)

    def exec_sliced_ptr(self, d_x, d_y, stream, flags, epoch, cols_per_rank, nranks):
        """Product on an x that is still arriving slice by slice (include/b200_peer.h);
        returns -1 without launching if this matrix's kernel cannot wait in-kernel."""
        return lib().b200_spmv_exec_sliced(self._h, c_void_p(d_x), c_void_p(d_y), c_void_p(stream),
                                           c_void_p(flags), int(epoch), int(cols_per_rank), int(nranks))

    @property
    def handle(self):
        """The b200_matrix* as an int (for C callers of the resident-matrix API)."""
        return int(self._h)

    def exec_ptr(self, d_x, d_y, stream=0):
        """Launch on raw device pointers (ints) and a cudaStream_t handle (int)."""
        return lib().b200_spmv_exec(self._h, c_void_p(d_x), c_void_p(d_y), c_void_p(stream))

    def exec(self, x, y, stream=None):
        """Launch on torch CUDA tensors, on torch's current stream by default."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        assert x.is_cuda and y.is_cuda and x.is_contiguous() and y.is_contiguous()
        assert x.numel() >= self.ncols and y.numel() >= self.rows
        return self.exec_ptr(x.data_ptr(), y.data_ptr(), stream)

    def exec_dot(self, x, y, dotv, stream=None):
        """y = A x and, in the same launch, the per-row-block shares of dotv . y (their sum, in
        index order, is the dot product; fixed reduction order).  Returns the partials tensor,
        or None if this matrix's kernel has no fused epilogue (nothing is launched)."""
        import torch
        if self.dot_partials <= 0:
            return None
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        assert x.is_cuda and y.is_cuda and dotv.is_cuda and dotv.numel() >= self.rows
        part = torch.zeros(self.dot_partials, dtype=torch.float64, device=x.device)
        rc = lib().b200_spmv_exec_dot(self._h, c_void_p(x.data_ptr()), c_void_p(y.data_ptr()),
                                      c_void_p(dotv.data_ptr()), c_void_p(part.data_ptr()), c_void_p(stream))
        if rc != 1:
            raise RuntimeError("b200_spmv_exec_dot refused a matrix that reported dot partials")
        return part

    def npb_cg_device(self, nonzer, niter, shift, use_graph=True):
        """Whole NPB CG benchmark with every vector resident in HBM (include/b200_cg.h)."""
        zeta_hist = np.zeros(niter)
        rnorm_hist = np.zeros(niter)
        res = CgResult()
        rc = lib().b200_cg_npb_run(self._h, int(nonzer), int(niter), float(shift), int(bool(use_graph)),
                                   _ptr(zeta_hist, c_double), _ptr(rnorm_hist, c_double), C.byref(res))
        if rc != 0:
            raise RuntimeError(f"b200_cg_npb_run failed with {rc}")
        return {"zeta": res.zeta, "rnorm": res.rnorm, "seconds": res.seconds, "mops": res.mops,
                "spmv_launches": res.spmv_launches, "vector_launches": res.vector_launches,
                "zeta_hist": zeta_hist, "rnorm_hist": rnorm_hist}

    def row_histogram(self):
        bins = (c_int64 * 32)()
        mn, mx = c_int(0), c_int(0)
        lib().b200_spmv_row_histogram(self._h, bins, C.byref(mn), C.byref(mx))
        return list(bins), mn.value, mx.value

    def release(self):
        if self._h:
            lib().b200_spmv_release(self._h)
            self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
