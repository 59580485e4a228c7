"""Seeded synthetic inputs of the shapes BASELINE.json names.

crsmat(size)      SparseBench/big_gen.py:59-83 restated with numpy and a fixed
                  seed: n = size^3 rows, row count = max(1, int(gauss(5, 4))),
                  columns uniform in [1, n] plus the diagonal, sorted, values
                  |gauss(0, 2)|.  Like big_gen.py the header nnz counts the
                  sampled entries only while the diagonal is appended to the
                  stream when absent (big_gen.py:75-76), and like the Fortran
                  reader (SRC/reference/gen_crs.f:757-789) the first nnz stream
                  entries are consumed with the row pointers as written -- so
                  rows are windows over the stream: unsorted now and then, with
                  the occasional repeated column.  That is what the reference's
                  SparseBench run feeds the ABI, so that is what is generated.

powerlaw_graph()  pagerank input (pagerank/main.cpp:100-111): a square link
                  matrix with heavy-tailed row lengths (in-degree ~ Zipf),
                  column-normalised and scaled by d = 0.85, 1-based CSR.
"""
import numpy as np


def crsmat(size, seed=170):
    rng = np.random.default_rng(seed)
    n = size ** 3
    counts = np.trunc(rng.normal(5.0, 4.0, n)).astype(np.int64)   # int() truncates toward zero
    counts = np.clip(counts, 1, n)
    row_ptr = np.empty(n + 1, dtype=np.int64)
    row_ptr[0] = 1
    np.cumsum(counts, out=row_ptr[1:])
    row_ptr[1:] += 1
    nnz = int(row_ptr[-1] - 1)
    # sampled columns (big_gen samples without replacement; with n >= 10^4 a repeat inside a
    # row is a ~1e-6 event and the reader tolerates it anyway)
    rows_of = np.repeat(np.arange(n, dtype=np.int64), counts)
    cols = rng.integers(1, n + 1, rows_of.size, dtype=np.int64)
    has_diag = np.zeros(n, dtype=bool)
    has_diag[rows_of[cols == rows_of + 1]] = True
    need = np.flatnonzero(~has_diag)
    rows_all = np.concatenate([rows_of, need])
    cols_all = np.concatenate([cols, need + 1])
    order = np.lexsort((cols_all, rows_all))                # sorted(sample) per row
    stream_cols = cols_all[order]
    stream_vals = np.abs(rng.normal(0.0, 2.0, stream_cols.size))
    a = np.ascontiguousarray(stream_vals[:nnz])
    colidx = np.ascontiguousarray(stream_cols[:nnz].astype(np.int32))
    return a, colidx, row_ptr.astype(np.int32), n


def powerlaw_graph(n=1 << 22, mean_deg=16.0, alpha=2.1, seed=22, cap=None):
    """1-based CSR of 0.85 * column-normalised adjacency with Zipf-like row lengths."""
    rng = np.random.default_rng(seed)
    cap = cap or max(n // 64, 4)
    raw = (rng.pareto(alpha - 1.0, n) + 1.0)
    raw = np.minimum(raw, cap)
    lens = np.maximum(1, np.floor(raw * (mean_deg / raw.mean()))).astype(np.int64)
    lens = np.minimum(lens, cap)
    rowstr = np.empty(n + 1, dtype=np.int64)
    rowstr[0] = 1
    np.cumsum(lens, out=rowstr[1:])
    rowstr[1:] += 1
    nnz = int(lens.sum())
    if nnz >= 2 ** 31 - 1:
        raise ValueError("graph too large for the int32 ABI")
    rows_of = np.repeat(np.arange(n, dtype=np.int64), lens)
    cols = rng.integers(0, n, nnz, dtype=np.int64)
    order = np.lexsort((cols, rows_of))
    cols = cols[order]
    outdeg = np.bincount(cols, minlength=n).astype(np.float64)
    vals = 0.85 / outdeg[cols]                              # normalise() then scale(d)
    x0 = rng.random(n)
    x0 /= x0.sum()                                          # pagerank/main.cpp:82-98
    return vals, (cols + 1).astype(np.int32), rowstr.astype(np.int32), x0
