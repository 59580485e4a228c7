"""Regenerate the golden fixtures in tests/golden/ from the REFERENCE itself.

Run in the build container (needs /root/reference and `make -C oracle all`):
    python tests/golden/make_golden.py

  native_ref_vectors.npz   seeded CSR inputs and the outputs of the reference's
                           own libspmv/native.c build (oracle/_ref/native.so),
                           fp64 and fp32, with the edge cases the ABI allows
                           (empty rows, rowstr[0] != 1, unsorted and repeated
                           columns, one very long row, colidx == ncols).
  npb_history.json         per-iteration ||r|| and zeta printed by the
                           reference's C twin of cg.f (SNU_NPB/NPB3.3-OMP-C/CG,
                           built by `make -C oracle snu CLASS=X`, 1 thread) for
                           classes S, W, A, plus the zeta constants of
                           NPB3.3.1/CG/cg.f:122-166.
  libspmv_test_kat.json    the known-answer vector of libspmv/test.cpp:44-49.
  parboil_spmv.npz         the fp32 CSR + x that parboil's CPU caller feeds
                           f_spmv_harness_ for its small / medium datasets,
                           recorded from the reference program, with the
                           reference's golden outputs; plus the small .mtx,
                           vector.bin, .out and bfs/input.mtx as data files for
                           the relinked-binary tests.
"""
import json
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_oracle  # noqa: E402

oracle = load_oracle()


def random_csr(rng, n, ncols, mean_len, dtype, base=1, sort=True, dup=False, long_row=None):
    lens = rng.poisson(mean_len, n).astype(np.int64)
    lens[rng.random(n) < 0.1] = 0                       # empty rows
    if long_row is not None:
        lens[n // 2] = long_row
    rowstr = np.empty(n + 1, dtype=np.int32)
    rowstr[0] = base
    rowstr[1:] = base + np.cumsum(lens)
    nnz = int(lens.sum())
    colidx = rng.integers(1, ncols + 1, nnz).astype(np.int32)
    if not dup:
        pass
    if sort:
        for r in range(n):
            s, e = rowstr[r] - base, rowstr[r + 1] - base
            colidx[s:e].sort()
    if nnz:
        colidx[nnz - 1] = ncols                         # colidx == ncols occurs
    a = rng.standard_normal(nnz).astype(dtype)
    lead = base - 1                                      # entries before the block
    a_full = np.concatenate([rng.standard_normal(lead).astype(dtype), a])
    c_full = np.concatenate([rng.integers(1, ncols + 1, lead).astype(np.int32), colidx])
    x = rng.standard_normal(ncols).astype(dtype)
    return a_full, c_full, rowstr, x


def make_native_vectors():
    assert oracle.ref_available(), "build oracle/_ref first (make -C oracle all)"
    rng = np.random.default_rng(20261018)
    cases = {}
    specs = [
        ("f64_small_sorted", dict(n=97, ncols=113, mean_len=9, dtype=np.float64)),
        ("f64_unsorted_dups", dict(n=300, ncols=64, mean_len=40, dtype=np.float64, sort=False, dup=True)),
        ("f64_base_offset", dict(n=128, ncols=200, mean_len=17, dtype=np.float64, base=58)),
        ("f64_long_row", dict(n=41, ncols=5000, mean_len=6, dtype=np.float64, long_row=9001)),
        ("f64_short_rows", dict(n=5000, ncols=5000, mean_len=3, dtype=np.float64)),
        ("f32_small_sorted", dict(n=97, ncols=113, mean_len=9, dtype=np.float32)),
        ("f32_unsorted_dups", dict(n=300, ncols=64, mean_len=40, dtype=np.float32, sort=False, dup=True)),
        ("f32_long_row", dict(n=41, ncols=5000, mean_len=6, dtype=np.float32, long_row=17001)),
    ]
    for name, kw in specs:
        a, c, rowstr, x = random_csr(rng, **kw)
        y = oracle.spmv(a, x, rowstr, c, use_ref=True)
        cases[name] = (a, c, rowstr, x, y)
    out = {}
    for name, (a, c, rowstr, x, y) in cases.items():
        out[f"{name}.a"] = a
        out[f"{name}.colidx"] = c
        out[f"{name}.rowstr"] = rowstr
        out[f"{name}.x"] = x
        out[f"{name}.y"] = y
    np.savez_compressed(HERE / "native_ref_vectors.npz", **out)
    print("wrote native_ref_vectors.npz:", ", ".join(cases))


ZETA = {"S": 8.5971775078648, "W": 10.362595087124, "A": 17.130235054029,
        "B": 22.712745482631, "C": 28.973605592845, "D": 52.514532105794,
        "E": 77.522164599383}
NNZ = {"S": 78148, "W": 508402, "A": 1853104, "B": 13708072, "C": 36121058}


def make_npb_history():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    env.pop("CC", None)
    hist = {"zeta_verify": ZETA, "nnz": NNZ, "classes": {}}
    for cls in "SWA":
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "snu", f"CLASS={cls}"], check=True,
                       env=env, stdout=subprocess.DEVNULL)
        out = subprocess.run([str(ROOT / "oracle" / "_ref" / f"snu_cg.{cls}")], env=env, check=True,
                             stdout=subprocess.PIPE, text=True).stdout
        rows = re.findall(r"^\s+(\d+)\s+([0-9.]+E[-+]\d+)\s+([0-9.]+)\s*$", out, re.M)
        assert "VERIFICATION SUCCESSFUL" in out and rows
        hist["classes"][cls] = {"rnorm": [r[1] for r in rows], "zeta": [r[2] for r in rows]}
    (HERE / "npb_history.json").write_text(json.dumps(hist, indent=1))
    print("wrote npb_history.json")


def make_kat():
    kat = {  # libspmv/test.cpp:44-49
        "a": [5.0, 8.0, 3.0, 6.0],
        "rowstr": [1, 2, 2, 3, 3, 3, 4, 4, 4, 4, 4, 4, 5],
        "colidx": [1, 4, 2, 12],
        "x": [1.0] * 6 + [2.0] * 6,
        "y": [5.0, 0.0, 8.0, 0.0, 0.0, 3.0, 0.0, 0.0, 0.0, 0.0, 0.0, 12.0],
        "rows": 12,
    }
    (HERE / "libspmv_test_kat.json").write_text(json.dumps(kat, indent=1))
    print("wrote libspmv_test_kat.json")


def read_record(path):
    """File written by oracle/record_harness.c."""
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw, dtype=np.int64, count=5)
    assert hdr[0] == 0x53504D56
    f32, rows, nnz, ncols = bool(hdr[1]), int(hdr[2]), int(hdr[3]), int(hdr[4])
    off = 40
    rowstr = np.frombuffer(raw, dtype=np.int32, count=rows + 1, offset=off); off += 4 * (rows + 1)
    colidx = np.frombuffer(raw, dtype=np.int32, count=nnz, offset=off); off += 4 * nnz
    dt = np.float32 if f32 else np.float64
    a = np.frombuffer(raw, dtype=dt, count=nnz, offset=off); off += a.itemsize * nnz
    x = np.frombuffer(raw, dtype=dt, count=ncols, offset=off)
    return a.copy(), colidx.copy(), rowstr.copy(), x.copy()


def make_parboil():
    """The fp32 caller: run the reference's parboil spmv CPU program (built into
    oracle/_ref by oracle/Makefile) against a recording backend to capture the
    exact CSR it feeds f_spmv_harness_ (main.c:80-95), and keep the reference's
    golden outputs (datasets/spmv/*/output/*.out) beside it.  The small
    MatrixMarket input itself is copied so the relinked binary can run on the
    GPU box."""
    import shutil
    import tempfile
    ds = Path("/root/reference/parboil/datasets/spmv")
    out = {}
    for name, mtx in (("small", "1138_bus.mtx"), ("medium", "bcsstk18.mtx")):
        with tempfile.TemporaryDirectory() as td:
            rec = os.path.join(td, "rec.bin")
            env = dict(os.environ, SPMV_RECORD_PATH=rec)
            subprocess.run([str(ROOT / "oracle" / "_ref" / "parboil_spmv.record"), "-i",
                            f"{ds / name / 'input' / mtx},{ds / name / 'input' / 'vector.bin'}",
                            "-o", os.path.join(td, "y.out")], env=env, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            a, c, rowstr, x = read_record(rec)
        gold = np.fromfile(ds / name / "output" / f"{mtx}.out", dtype=np.uint8)
        n = int(np.frombuffer(gold[:4].tobytes(), dtype=np.uint32)[0])
        y = np.frombuffer(gold[4:].tobytes(), dtype=np.float32, count=n)
        assert n == len(rowstr) - 1
        for k, v in (("a", a), ("colidx", c), ("rowstr", rowstr), ("x", x), ("y_golden", y)):
            out[f"{name}.{k}"] = v
    np.savez_compressed(HERE / "parboil_spmv.npz", **out)
    shutil.copy(ds / "small" / "input" / "1138_bus.mtx", HERE / "parboil_small_1138_bus.mtx")
    shutil.copy(ds / "small" / "input" / "vector.bin", HERE / "parboil_small_vector.bin")
    shutil.copy(ds / "small" / "output" / "1138_bus.mtx.out", HERE / "parboil_small_1138_bus.mtx.out")
    shutil.copy("/root/reference/bfs/input.mtx", HERE / "bfs_input.mtx")
    for f in ("parboil_small_1138_bus.mtx", "parboil_small_vector.bin",
              "parboil_small_1138_bus.mtx.out", "bfs_input.mtx"):
        os.chmod(HERE / f, 0o644)
    print("wrote parboil_spmv.npz (+ small inputs, bfs input)")


if __name__ == "__main__":
    make_kat()
    make_native_vectors()
    make_npb_history()
    make_parboil()
