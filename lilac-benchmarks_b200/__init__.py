"""lilac-benchmarks_b200 -- the `b200` libspmv platform for mob-group/lilac-benchmarks.

The product is a C-ABI shared library (csrc/b200.so, installed as
libb200-spmv.so) that exports the reference's two symbols, `spmv_harness_` and
`f_spmv_harness_` (libspmv/native.c:3-11), on top of hand-written sm_100a CSR
SpMV kernels.  This Python package is only the thin host-side mirror used by
the tests and bench.py: ctypes bindings (libspmv.py), the C callers of the ABI
(callers/, bound in npb.py) and the build helper (build.py).

The directory name carries a hyphen, so it is loaded under the module name
`lilac_benchmarks_b200` by __graft_entry__.load_package().
"""
from . import build  # noqa: F401

__all__ = ["build"]
