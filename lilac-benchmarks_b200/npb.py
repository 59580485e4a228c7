"""ctypes binding of callers/libb200callers.so: the NPB CG matrix generator
and driver restated in C from NPB3.3.1/CG/cg.f (see callers/npb/*.c)."""
import ctypes as C
from ctypes import POINTER, c_char, c_double, c_int, c_int64, c_void_p

import numpy as np

from .build import CALLERS_SO

HARNESS_FN = C.CFUNCTYPE(c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_double),
                         POINTER(c_int), POINTER(c_int), POINTER(c_int))


class CgClass(C.Structure):
    _fields_ = [("cls", c_char), ("na", c_int), ("nonzer", c_int), ("niter", c_int),
                ("shift", c_double), ("rcond", c_double), ("zeta_verify", c_double)]


class Csr(C.Structure):
    _fields_ = [("n", c_int), ("nnz", c_int64), ("rowstr", POINTER(c_int)),
                ("colidx", POINTER(c_int)), ("a", POINTER(c_double))]


class CgResult(C.Structure):
    _fields_ = [("zeta", c_double), ("rnorm", c_double), ("err", c_double), ("verified", c_int),
                ("t_bench", c_double), ("t_init", c_double), ("mops", c_double),
                ("spmv_calls", c_int), ("zeta_hist", POINTER(c_double)),
                ("rnorm_hist", POINTER(c_double))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not CALLERS_SO.exists():
            raise RuntimeError(f"{CALLERS_SO} is not built: run __graft_entry__.build()")
        L = C.CDLL(str(CALLERS_SO))
        L.npb_cg_class_lookup.argtypes = [c_char, POINTER(CgClass)]
        L.npb_cg_class_lookup.restype = c_int
        L.npb_makea.argtypes = [POINTER(CgClass), POINTER(Csr)]
        L.npb_makea.restype = c_int
        L.npb_makea_rows.argtypes = [POINTER(CgClass), c_int, c_int, POINTER(Csr)]
        L.npb_makea_rows.restype = c_int
        L.npb_makea_release_cache.argtypes = []
        L.npb_makea_release_cache.restype = None
        L.npb_csr_free.argtypes = [POINTER(Csr)]
        L.npb_csr_free.restype = None
        L.npb_cg_run.argtypes = [POINTER(CgClass), POINTER(Csr), c_void_p, POINTER(CgResult), c_int]
        L.npb_cg_run.restype = c_int
        L.npb_time_spmv_calls.argtypes = [c_void_p, POINTER(c_double), POINTER(c_double),
                                          POINTER(POINTER(c_double)), c_int, POINTER(c_int),
                                          POINTER(c_int), c_int, c_int]
        L.npb_time_spmv_calls.restype = c_double
        L.npb_issue_exec_calls.argtypes = [c_void_p, c_void_p, POINTER(c_void_p), c_int, c_void_p, c_void_p, c_int]
        L.npb_issue_exec_calls.restype = None
        L.npb_vectors_get.argtypes = [POINTER(CgClass), POINTER(c_void_p), POINTER(c_void_p),
                                      POINTER(c_void_p), POINTER(c_void_p)]
        L.npb_vectors_get.restype = c_int
        L.npb_free.argtypes = [c_void_p]
        L.npb_free.restype = None
        L.npb_randlc.argtypes = [POINTER(c_double), c_double]
        L.npb_randlc.restype = c_double
        _lib = L
    return _lib


def cg_class(letter):
    c = CgClass()
    if lib().npb_cg_class_lookup(letter.encode()[:1], C.byref(c)) != 0:
        raise ValueError(f"unknown NPB class {letter!r}")
    return c


class NpbMatrix:
    """1-based CSR of one NPB CG class (or a row block of it) as numpy arrays."""

    def __init__(self, letter, row_lo=None, row_hi=None, pieces=1):
        """`pieces` > 1 builds the block in that many consecutive row ranges to
        bound the generator's temporary memory (class D/E shards)."""
        self.cls = cg_class(letter)
        if row_lo is None:
            row_lo, row_hi = 0, self.cls.na
        pieces = max(1, min(int(pieces), max(row_hi - row_lo, 1)))
        cuts = [row_lo + (row_hi - row_lo) * k // pieces for k in range(pieces + 1)]
        # upper bound of the block's nnz (cg.f: nz = na*(nonzer+1)**2): the pieces are
        # written straight into preallocated arrays, so no second copy is ever held
        bound = (row_hi - row_lo) * (self.cls.nonzer + 1) ** 2 if pieces > 1 else 0
        rowstr = np.empty(row_hi - row_lo + 1, dtype=np.int64)
        rowstr[0] = 1
        colidx = np.empty(bound, dtype=np.int32) if pieces > 1 else None
        a = np.empty(bound, dtype=np.float64) if pieces > 1 else None
        offset = 0
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            csr = Csr()
            rc = lib().npb_makea_rows(C.byref(self.cls), int(lo), int(hi), C.byref(csr))
            if rc != 0:
                raise RuntimeError(f"npb_makea failed with {rc}")
            rs = np.ctypeslib.as_array(csr.rowstr, shape=(csr.n + 1,))
            rowstr[lo - row_lo + 1: hi - row_lo + 1] = rs[1:].astype(np.int64) + offset
            c_view = np.ctypeslib.as_array(csr.colidx, shape=(max(csr.nnz, 1),))[:csr.nnz]
            a_view = np.ctypeslib.as_array(csr.a, shape=(max(csr.nnz, 1),))[:csr.nnz]
            if pieces > 1:
                colidx[offset: offset + csr.nnz] = c_view
                a[offset: offset + csr.nnz] = a_view
            else:
                colidx, a = c_view.copy(), a_view.copy()     # numpy-owned before the C arrays go
            offset += int(csr.nnz)
            lib().npb_csr_free(C.byref(csr))
        if pieces > 1:
            lib().npb_makea_release_cache()
        self.n = row_hi - row_lo
        self.nnz = offset
        if offset + 1 > 2 ** 31 - 1:
            raise RuntimeError("row block exceeds the int32 ABI; use more ranks")
        self.rowstr = rowstr.astype(np.int32)
        self.colidx = colidx[:offset]
        self.a = a[:offset]

    def as_csr_struct(self):
        csr = Csr()
        csr.n = self.n
        csr.nnz = self.nnz
        csr.rowstr = self.rowstr.ctypes.data_as(POINTER(c_int))
        csr.colidx = self.colidx.ctypes.data_as(POINTER(c_int))
        csr.a = self.a.ctypes.data_as(POINTER(c_double))
        return csr


class NpbDeviceMatrix:
    """Rows [row_lo, row_hi) of an NPB CG class assembled ON the current CUDA device
    (include/b200_npb.h): the host only draws the generating vectors (the sequential
    random stream of cg.f:709-718), the rows are built by the GPU, bit for bit what
    `NpbMatrix` builds on the host.  Device pointers are plain ints."""

    def __init__(self, letter, row_lo=None, row_hi=None, release_vectors=True):
        from . import libspmv
        self.cls = cg_class(letter)
        if row_lo is None:
            row_lo, row_hi = 0, self.cls.na
        arow, acol, aelt, size = c_void_p(), c_void_p(), c_void_p(), c_void_p()
        if lib().npb_vectors_get(C.byref(self.cls), C.byref(arow), C.byref(acol), C.byref(aelt),
                                 C.byref(size)) != 0:
            raise RuntimeError("npb_vectors_get failed")
        self._L = libspmv.lib()
        self._csr = libspmv.NpbDeviceCsr()
        try:
            rc = self._L.b200_npb_makea_device(self.cls.na, self.cls.nonzer + 1, arow, acol, aelt, size,
                                               self.cls.rcond, self.cls.shift, int(row_lo), int(row_hi),
                                               C.byref(self._csr))
        finally:
            lib().npb_free(size)
            if release_vectors:
                lib().npb_makea_release_cache()
        if rc == -1:
            raise RuntimeError("row block exceeds the int32 ABI; use more ranks")
        if rc != 0:
            raise RuntimeError(f"b200_npb_makea_device failed with {rc}")
        self.n = row_hi - row_lo
        self.row_lo, self.row_hi = row_lo, row_hi
        self.nnz = int(self._csr.nnz)

    def resident(self, kernel="auto"):
        """Upload device -> device and keep resident (b200_spmv_upload_device)."""
        from . import libspmv
        return libspmv.ResidentMatrix.from_device(self._csr.d_a, self._csr.d_rowstr, self._csr.d_colidx,
                                                  self.n, kernel=kernel, keep=self)

    def to_host(self):
        """(a, rowstr, colidx) as numpy arrays, for the CPU checker."""
        rowstr = np.empty(self.n + 1, dtype=np.int32)
        colidx = np.empty(max(self.nnz, 1), dtype=np.int32)
        a = np.empty(max(self.nnz, 1), dtype=np.float64)
        rc = self._L.b200_npb_csr_to_host(C.byref(self._csr), rowstr.ctypes.data_as(POINTER(c_int)),
                                          colidx.ctypes.data_as(POINTER(c_int)),
                                          a.ctypes.data_as(POINTER(c_double)))
        if rc != 0:
            raise RuntimeError(f"b200_npb_csr_to_host failed with {rc}")
        return a[:self.nnz], rowstr, colidx[:self.nnz]

    def free(self):
        if self._csr.d_rowstr:
            self._L.b200_npb_csr_free(C.byref(self._csr))

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def run_cg(matrix, harness_addr, verbose=False):
    """Whole NPB CG benchmark (cg.f:53-443) through the ABI symbol at `harness_addr`."""
    niter = matrix.cls.niter
    zeta_hist = np.zeros(niter)
    rnorm_hist = np.zeros(niter)
    res = CgResult()
    res.zeta_hist = zeta_hist.ctypes.data_as(POINTER(c_double))
    res.rnorm_hist = rnorm_hist.ctypes.data_as(POINTER(c_double))
    csr = matrix.as_csr_struct()
    rc = lib().npb_cg_run(C.byref(matrix.cls), C.byref(csr), c_void_p(harness_addr),
                          C.byref(res), 1 if verbose else 0)
    if rc != 0:
        raise RuntimeError(f"npb_cg_run failed with {rc}")
    return {"zeta": res.zeta, "rnorm": res.rnorm, "err": res.err, "verified": bool(res.verified),
            "t_bench": res.t_bench, "t_init": res.t_init, "mops": res.mops,
            "spmv_calls": res.spmv_calls, "zeta_hist": zeta_hist, "rnorm_hist": rnorm_hist}


def issue_exec_calls(exec_addr, matrix_handle, d_xs, d_y, stream, calls):
    """`calls` back-to-back launches of the resident product issued by the C loop
    npb_issue_exec_calls (callers/npb/cg.c): `exec_addr` is the address of b200_spmv_exec,
    `d_xs` device pointers (ints) the x argument rotates over.  No synchronisation."""
    arr = (c_void_p * len(d_xs))(*[c_void_p(int(p)) for p in d_xs])
    lib().npb_issue_exec_calls(c_void_p(exec_addr), c_void_p(matrix_handle), arr, len(d_xs), c_void_p(int(d_y)),
                               c_void_p(int(stream)), int(calls))


def time_spmv_calls(harness_addr, ov, a, xs, rowstr, colidx, rows, calls):
    """Seconds per ABI call measured by the C caller loop (callers/npb/cg.c,
    npb_time_spmv_calls): `calls` products y = A x through the function at
    `harness_addr`, x rotating over the numpy vectors `xs`."""
    dp = POINTER(c_double)
    xp = (dp * len(xs))(*[x.ctypes.data_as(dp) for x in xs])
    sec = lib().npb_time_spmv_calls(harness_addr, ov.ctypes.data_as(dp), a.ctypes.data_as(dp), xp, len(xs),
                                    rowstr.ctypes.data_as(POINTER(c_int)),
                                    colidx.ctypes.data_as(POINTER(c_int)), int(rows), int(calls))
    return sec / max(int(calls), 1)
