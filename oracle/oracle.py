"""ctypes access to the CHECKER: oracle/liboracle.so (the in-repo restatement
of libspmv/native-impl.c) and, when it has been built from the reference tree,
oracle/_ref/native.so (the reference's own native backend).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; nothing under
lilac-benchmarks_b200/ does.
"""
import ctypes as C
from ctypes import POINTER, c_double, c_float, c_int, c_void_p
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "liboracle.so"
REF_NATIVE_SO = HERE / "_ref" / "native.so"
REF_TEST_BIN = HERE / "_ref" / "test"

_dp, _fp, _ip = POINTER(c_double), POINTER(c_float), POINTER(c_int)
_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not ORACLE_SO.exists():
            raise RuntimeError(f"{ORACLE_SO} not built (make -C oracle)")
        L = C.CDLL(str(ORACLE_SO))
        for name in ("oracle_spmv_f64", "oracle_spmv_f64_omp", "spmv_harness_"):
            getattr(L, name).argtypes = [_dp, _dp, _dp, _ip, _ip, _ip]
        for name in ("oracle_spmv_f32", "oracle_spmv_f32_omp", "f_spmv_harness_"):
            getattr(L, name).argtypes = [_fp, _fp, _fp, _ip, _ip, _ip]
        L.spmv_harness_.restype = c_void_p
        L.f_spmv_harness_.restype = c_void_p
        L.oracle_spmv_f64_extended.argtypes = [_dp, _dp, _dp, _dp, _ip, _ip, _ip]
        L.oracle_max_colidx.argtypes = [_ip, _ip, _ip]
        L.oracle_max_colidx.restype = c_int
        _lib = L
    return _lib


def ref_available():
    return REF_NATIVE_SO.exists()


def ref():
    """The reference's own native.so (libspmv/native.c + native-impl.c)."""
    global _ref
    if _ref is None:
        L = C.CDLL(str(REF_NATIVE_SO))
        L.spmv_harness_.argtypes = [_dp, _dp, _dp, _ip, _ip, _ip]
        L.f_spmv_harness_.argtypes = [_fp, _fp, _fp, _ip, _ip, _ip]
        _ref = L
    return _ref


def _p(a, ct):
    return a.ctypes.data_as(POINTER(ct))


def spmv(a, x, rowstr, colidx, rows=None, omp=False, use_ref=False):
    """y = A x with the reference semantics; dtype follows `a`."""
    if rows is None:
        rows = len(rowstr) - 1
    n = c_int(int(rows))
    y = np.empty(rows, dtype=a.dtype)
    f32 = a.dtype == np.float32
    ct = c_float if f32 else c_double
    assert x.dtype == a.dtype and rowstr.dtype == np.int32 and colidx.dtype == np.int32
    if use_ref:
        fn = ref().f_spmv_harness_ if f32 else ref().spmv_harness_
    elif omp:
        fn = lib().oracle_spmv_f32_omp if f32 else lib().oracle_spmv_f64_omp
    else:
        fn = lib().oracle_spmv_f32 if f32 else lib().oracle_spmv_f64
    fn(_p(y, ct), _p(a, ct), _p(x, ct), _p(rowstr, c_int), _p(colidx, c_int), C.byref(n))
    return y


def spmv_extended(a, x, rowstr, colidx, rows=None):
    """(long-double row sums rounded to double, sum of |terms|) per row."""
    if rows is None:
        rows = len(rowstr) - 1
    n = c_int(int(rows))
    y = np.empty(rows)
    mag = np.empty(rows)
    lib().oracle_spmv_f64_extended(_p(y, c_double), _p(mag, c_double), _p(a, c_double),
                                   _p(x, c_double), _p(rowstr, c_int), _p(colidx, c_int), C.byref(n))
    return y, mag


def max_colidx(rowstr, colidx, rows=None):
    if rows is None:
        rows = len(rowstr) - 1
    n = c_int(int(rows))
    return lib().oracle_max_colidx(_p(rowstr, c_int), _p(colidx, c_int), C.byref(n))


def harness_address(f32=False, use_ref=False):
    L = ref() if use_ref else lib()
    fn = L.f_spmv_harness_ if f32 else L.spmv_harness_
    return C.cast(fn, c_void_p).value
