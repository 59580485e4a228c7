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


if __name__ == "__main__":
    make_kat()
    make_native_vectors()
    make_npb_history()
