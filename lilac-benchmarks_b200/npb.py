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
        L.npb_csr_free.argtypes = [POINTER(Csr)]
        L.npb_csr_free.restype = None
        L.npb_cg_run.argtypes = [POINTER(CgClass), POINTER(Csr), c_void_p, POINTER(CgResult), c_int]
        L.npb_cg_run.restype = c_int
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

    def __init__(self, letter, row_lo=None, row_hi=None):
        self.cls = cg_class(letter)
        csr = Csr()
        if row_lo is None:
            rc = lib().npb_makea(C.byref(self.cls), C.byref(csr))
        else:
            rc = lib().npb_makea_rows(C.byref(self.cls), int(row_lo), int(row_hi), C.byref(csr))
        if rc != 0:
            raise RuntimeError(f"npb_makea failed with {rc}")
        self.n = csr.n
        self.nnz = csr.nnz
        # copy out into numpy-owned memory, then free the C arrays
        self.rowstr = np.ctypeslib.as_array(csr.rowstr, shape=(csr.n + 1,)).copy()
        self.colidx = np.ctypeslib.as_array(csr.colidx, shape=(max(csr.nnz, 1),))[:csr.nnz].copy()
        self.a = np.ctypeslib.as_array(csr.a, shape=(max(csr.nnz, 1),))[:csr.nnz].copy()
        lib().npb_csr_free(C.byref(csr))

    def as_csr_struct(self):
        csr = Csr()
        csr.n = self.n
        csr.nnz = self.nnz
        csr.rowstr = self.rowstr.ctypes.data_as(POINTER(c_int))
        csr.colidx = self.colidx.ctypes.data_as(POINTER(c_int))
        csr.a = self.a.ctypes.data_as(POINTER(c_double))
        return csr


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
