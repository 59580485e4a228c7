import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = Path(__file__).resolve().parent / "golden"

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Make sure every native artefact exists (no-op when already built)."""
    entry.build()
    return entry.load_package()


@pytest.fixture(scope="session")
def oracle(built):
    return entry.load_oracle()


@pytest.fixture(scope="session")
def libspmv(built):
    from lilac_benchmarks_b200 import libspmv as mod
    return mod


@pytest.fixture(scope="session")
def npb(built):
    from lilac_benchmarks_b200 import npb as mod
    return mod


@pytest.fixture(scope="session")
def kat():
    return json.loads((GOLDEN / "libspmv_test_kat.json").read_text())


@pytest.fixture(scope="session")
def native_vectors():
    z = np.load(GOLDEN / "native_ref_vectors.npz")
    names = sorted({k.split(".")[0] for k in z.files})
    return {n: {f: z[f"{n}.{f}"] for f in ("a", "colidx", "rowstr", "x", "y")} for n in names}


@pytest.fixture(scope="session")
def parboil():
    z = np.load(GOLDEN / "parboil_spmv.npz")
    return {n: {f: z[f"{n}.{f}"] for f in ("a", "colidx", "rowstr", "x", "y_golden")}
            for n in ("small", "medium")}


def parboil_compare(ref, got):
    """parboil/benchmarks/spmv/tools/compare-output:12-37: abs tol 1e-4 * max|ref| OR rel 0.2 %."""
    abstol = 1e-4 * np.abs(ref).max()
    diff = np.abs(ref.astype(np.float64) - got.astype(np.float64))
    return bool(np.all((diff <= abstol) | (diff < 0.002 * np.abs(ref))))


@pytest.fixture(scope="session")
def npb_history():
    return json.loads((GOLDEN / "npb_history.json").read_text())


def make_csr(rng, n, ncols, lens, dtype=np.float64, base=1, sort=True, positive=False):
    """1-based CSR with the given row lengths (helper shared by the tests)."""
    lens = np.asarray(lens, dtype=np.int64)
    rowstr = np.empty(n + 1, dtype=np.int32)
    rowstr[0] = base
    rowstr[1:] = base + np.cumsum(lens)
    nnz = int(lens.sum())
    colidx = rng.integers(1, ncols + 1, nnz).astype(np.int32)
    if sort and nnz:
        rows = np.repeat(np.arange(n), lens)
        order = np.lexsort((colidx, rows))
        colidx = colidx[order]
    lead = base - 1
    vals = rng.random(nnz) + 0.1 if positive else rng.standard_normal(nnz)
    a = np.concatenate([rng.standard_normal(lead), vals]).astype(dtype)
    c = np.concatenate([rng.integers(1, ncols + 1, lead).astype(np.int32), colidx])
    x = (rng.random(ncols) + 0.1 if positive else rng.standard_normal(ncols)).astype(dtype)
    return a, c, rowstr, x
