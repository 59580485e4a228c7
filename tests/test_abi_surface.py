"""CPU-side checks of the boundary: the shared library loads without a GPU,
exports every function include/*.h declares, and the host-only helpers work.
No compute entry point is called here."""
import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _declared_functions():
    names = []
    for header in (ROOT / "include").glob("*.h"):
        text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
        names += re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(n for n in names if not n.startswith("__")))


def test_library_exports_every_declared_symbol(libspmv):
    L = libspmv.lib()
    declared = _declared_functions()
    assert "spmv_harness_" in declared and "f_spmv_harness_" in declared
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"


def test_drop_in_symbols_have_default_visibility(libspmv):
    out = subprocess.run(["nm", "-D", "--defined-only", str(libspmv.B200_SO)],
                         stdout=subprocess.PIPE, text=True, check=True).stdout
    assert re.search(r" T spmv_harness_$", out, re.M)
    assert re.search(r" T f_spmv_harness_$", out, re.M)


def test_library_needs_cudart_but_no_vendor_sparse_library(libspmv):
    out = subprocess.run(["readelf", "-d", str(libspmv.B200_SO)],
                         stdout=subprocess.PIPE, text=True, check=True).stdout
    needed = re.findall(r"NEEDED.*\[(.*?)\]", out)
    assert any(n.startswith("libcudart") for n in needed)
    assert not any(("cusparse" in n) or ("cublas" in n) for n in needed)


def test_partition_rows_balances_nnz(libspmv, npb):
    m = npb.NpbMatrix("S")
    for parts in (1, 2, 3, 8):
        b = libspmv.partition_rows(m.rowstr, parts)
        assert b[0] == 0 and b[-1] == m.n and np.all(np.diff(b) >= 0)
        per = np.diff(m.rowstr[b].astype(np.int64))
        assert per.sum() == m.nnz
        assert per.max() - per.min() <= 2 * np.diff(m.rowstr).max()


def test_partition_rows_skewed_and_empty(libspmv):
    rowstr = np.array([1, 1, 1, 1001, 1001, 1002], dtype=np.int32)
    b = libspmv.partition_rows(rowstr, 4)
    assert b[0] == 0 and b[-1] == 5 and np.all(np.diff(b) >= 0)
    empty = np.array([1], dtype=np.int32)
    assert list(libspmv.partition_rows(empty, 2)) == [0, 0, 0]


def test_bounce_buffer_copy_matches_memcpy_for_every_alignment(libspmv):
    """b200_spmv_host_copy (cache-bypassing stores, head / 64-byte body / tail): every
    source and destination alignment, sizes around the thresholds; bytes next to the
    destination stay untouched."""
    L = libspmv.lib()
    L.b200_spmv_host_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    L.b200_spmv_host_copy.restype = None
    rng = np.random.default_rng(2)
    src = rng.integers(0, 256, 1 << 17, dtype=np.uint8)
    for n in (0, 1, 63, 64, 4095, 4096, 4097, 4096 + 63, 8192 + 17, 100000):
        for so in (0, 1, 7, 8, 15, 16, 33):
            for do in (0, 1, 5, 8, 15, 16, 47):
                dst = np.full(n + 128, 0xAB, dtype=np.uint8)
                L.b200_spmv_host_copy(dst.ctypes.data + do, src.ctypes.data + so, n)
                assert np.array_equal(dst[do:do + n], src[so:so + n]), (n, so, do)
                assert np.all(dst[:do] == 0xAB) and np.all(dst[do + n:] == 0xAB), (n, so, do)


def test_version_string(libspmv):
    assert b"sm_100a" in libspmv.lib().b200_spmv_version()


def test_product_does_not_reference_the_oracle():
    pkg = ROOT / "lilac-benchmarks_b200"
    for path in pkg.rglob("*"):
        if path.suffix in {".py", ".c", ".cu", ".cuh", ".h", ".cpp"} or path.name == "Makefile":
            text = path.read_text()
            if path.name == "build.py":
                continue          # builds the checker, never loads it
            assert "liboracle" not in text and "oracle/" not in text, path


def test_bench_arms_name_the_workload_identically():
    """bench.py: `config` must be the same dict in the b200 and the --impl reference arm (the
    driver compares them), and hold only what names the workload."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    cfg = bench.workload_config(bench.workload_label("A"), 14000, 1853104, 14000, 1)
    assert set(cfg) == {"workload", "rows", "nnz", "ncols", "algorithmic_bytes_per_step", "l2_policy"}
    assert cfg["algorithmic_bytes_per_step"] == 12 * 1853104 + 4 * 14001 + 16 * 14000
    proc = subprocess.run(["python", str(ROOT / "bench.py"), "--impl", "reference", "--workload", "A",
                           "--steps", "2", "--warmup", "1"], stdout=subprocess.PIPE, text=True, check=True,
                          timeout=600)
    line = json.loads(proc.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["config"] == cfg
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert bench.workload_label("crsmat170u") == "sparsebench-crsmat170u"
    assert bench.workload_label("pl22") == "pagerank-powerlaw-2^22"
