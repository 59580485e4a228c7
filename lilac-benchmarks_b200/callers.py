"""ctypes bindings of the other C callers in callers/libb200callers.so:
SparseBench's BiCG (callers/sparsebench) and the PageRank power iteration
(callers/pagerank).  NPB CG is bound in npb.py."""
import ctypes as C
from ctypes import POINTER, c_char_p, c_double, c_int, c_void_p

import numpy as np

from .npb import lib as _lib


class BicgResult(C.Structure):
    _fields_ = [("its", c_int), ("rnorm0", c_double), ("rnorm", c_double),
                ("t_iter", c_double), ("t_matprod", c_double), ("matprod_calls", c_int)]


_ready = False


def lib():
    global _ready
    L = _lib()
    if not _ready:
        L.sb_bicg.argtypes = [c_int, POINTER(c_double), POINTER(c_int), POINTER(c_int), c_void_p,
                              c_int, c_double, POINTER(c_double), POINTER(c_double), POINTER(BicgResult)]
        L.sb_bicg.restype = c_int
        L.sb_read_crs.argtypes = [c_char_p, POINTER(c_int), POINTER(c_int), POINTER(POINTER(c_int)),
                                  POINTER(POINTER(c_int)), POINTER(POINTER(c_double))]
        L.sb_read_crs.restype = c_int
        L.pr_power_iterations.argtypes = [c_int, POINTER(c_double), POINTER(c_int), POINTER(c_int),
                                          POINTER(c_double), POINTER(c_double), c_double, c_int,
                                          c_void_p, POINTER(c_double)]
        L.pr_power_iterations.restype = c_double
        L.pr_load_mtx.argtypes = [c_char_p, c_double, POINTER(c_int), POINTER(c_int),
                                  POINTER(POINTER(c_int)), POINTER(POINTER(c_int)),
                                  POINTER(POINTER(c_double))]
        L.pr_load_mtx.restype = c_int
        _ready = True
    return L


def _p(a, ct):
    return a.ctypes.data_as(POINTER(ct))


def bicg(a, rowstr, colidx, harness_addr, maxit=100, rtol=1e-6):
    """SparseBench BiCG (iter.f:18-104), x0 = 0, rhs = 1, through the ABI symbol."""
    n = len(rowstr) - 1
    x = np.zeros(n + 1)
    hist = np.zeros(maxit)
    res = BicgResult()
    rc = lib().sb_bicg(n, _p(a, c_double), _p(rowstr, c_int), _p(colidx, c_int),
                       c_void_p(harness_addr), maxit, rtol, _p(x, c_double), _p(hist, c_double),
                       C.byref(res))
    if rc != 0:
        raise RuntimeError(f"sb_bicg failed with {rc}")
    return {"its": res.its, "rnorm0": res.rnorm0, "rnorm": res.rnorm, "t_iter": res.t_iter,
            "t_matprod": res.t_matprod, "matprod_calls": res.matprod_calls,
            "hist": hist[:abs(res.its)], "x": x[:n]}


def write_crs(path, a, rowstr, colidx, extra_lines=None):
    """The CRS text file of SparseBench/big_gen.py:52-57."""
    with open(path, "w") as f:
        f.write("{:12}{:12}\n".format(len(rowstr) - 1, len(a)))
        for p in rowstr:
            f.write("{:12}\n".format(int(p)))
        for c, v in zip(colidx, a):
            f.write("{:12} {:20.17f}\n".format(int(c), float(v)))
        for c, v in (extra_lines or []):
            f.write("{:12} {:20.17f}\n".format(int(c), float(v)))


def read_crs(path):
    n, nnz = c_int(), c_int()
    ptr, idx, val = POINTER(c_int)(), POINTER(c_int)(), POINTER(c_double)()
    rc = lib().sb_read_crs(str(path).encode(), C.byref(n), C.byref(nnz), C.byref(ptr),
                           C.byref(idx), C.byref(val))
    if rc != 0:
        raise RuntimeError(f"sb_read_crs failed with {rc}")
    rowstr = np.ctypeslib.as_array(ptr, shape=(n.value + 1,)).copy()
    colidx = np.ctypeslib.as_array(idx, shape=(max(nnz.value, 1),))[:nnz.value].copy()
    a = np.ctypeslib.as_array(val, shape=(max(nnz.value, 1),))[:nnz.value].copy()
    return a, rowstr, colidx


def pagerank(a, rowstr, colidx, x0, harness_addr, iters=1024, d=0.85):
    """`iters` power iterations of pagerank/main.cpp:125-149; returns (x, error, seconds)."""
    n = len(rowstr) - 1
    x = np.array(x0, dtype=np.float64, copy=True)
    y = np.zeros(n)
    sec = c_double()
    err = lib().pr_power_iterations(n, _p(a, c_double), _p(rowstr, c_int), _p(colidx, c_int),
                                    _p(x, c_double), _p(y, c_double), d, iters,
                                    c_void_p(harness_addr), C.byref(sec))
    return x, err, sec.value


def load_mtx(path, d=0.85):
    n, nnz = c_int(), c_int()
    ptr, idx, val = POINTER(c_int)(), POINTER(c_int)(), POINTER(c_double)()
    rc = lib().pr_load_mtx(str(path).encode(), d, C.byref(n), C.byref(nnz), C.byref(ptr),
                           C.byref(idx), C.byref(val))
    if rc != 0:
        raise RuntimeError(f"pr_load_mtx failed with {rc}")
    rowstr = np.ctypeslib.as_array(ptr, shape=(n.value + 1,)).copy()
    colidx = np.ctypeslib.as_array(idx, shape=(max(nnz.value, 1),))[:nnz.value].copy()
    a = np.ctypeslib.as_array(val, shape=(max(nnz.value, 1),))[:nnz.value].copy()
    return a, rowstr, colidx
