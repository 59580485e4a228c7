"""The C restatement of NPB3.3.1/CG/cg.f (callers/npb) on the CPU: matrix
sizes, zeta verification (cg.f:122-166, 363-392) and the per-iteration history
printed by the reference's own C twin (SNU_NPB, golden fixture)."""
import numpy as np
import pytest


def _fmt_rnorm(v):
    # SNU prints %20.14E, cg.f prints e20.14; compare on the SNU text
    return f"{v:.14E}"


@pytest.mark.parametrize("cls", ["S", "W", "A"])
def test_makea_sizes_and_structure(npb, npb_history, cls):
    m = npb.NpbMatrix(cls)
    assert m.nnz == npb_history["nnz"][cls]
    assert m.rowstr[0] == 1 and m.rowstr[-1] == m.nnz + 1
    assert m.colidx.min() >= 1 and m.colidx.max() <= m.n
    # columns strictly increasing inside every row (cg.f:838-850 keeps them ordered)
    d = np.diff(m.colidx.astype(np.int64))
    row_start = np.zeros(m.nnz, dtype=bool)
    row_start[m.rowstr[1:-1] - 1] = True
    assert np.all(d[~row_start[1:]] > 0)


@pytest.mark.parametrize("cls", ["S", "W", "A"])
def test_cg_history_matches_reference_c_twin(npb, oracle, npb_history, cls):
    m = npb.NpbMatrix(cls)
    res = npb.run_cg(m, oracle.harness_address())
    assert res["verified"] and res["err"] <= 1e-10
    assert res["spmv_calls"] == (m.cls.niter + 1) * 26
    gold = npb_history["classes"][cls]
    assert [f"{z:.13f}" for z in res["zeta_hist"]] == gold["zeta"]
    assert [_fmt_rnorm(r) for r in res["rnorm_hist"]] == gold["rnorm"]


def test_makea_row_blocks_tile_the_matrix(npb):
    full = npb.NpbMatrix("S")
    cuts = [0, 301, 302, 1000, full.n]
    a, c = [], []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        part = npb.NpbMatrix("S", lo, hi)
        assert part.n == hi - lo and part.rowstr[0] == 1
        assert np.array_equal(np.diff(part.rowstr), np.diff(full.rowstr[lo:hi + 1]))
        a.append(part.a)
        c.append(part.colidx)
    assert np.array_equal(np.concatenate(a), full.a)
    assert np.array_equal(np.concatenate(c), full.colidx)


def test_randlc_first_values(npb):
    import ctypes as C
    x = C.c_double(314159265.0)
    v = npb.lib().npb_randlc(C.byref(x), 1220703125.0)
    # x1 = a*x0 mod 2^46
    expect = (314159265 * 1220703125) % (1 << 46)
    assert x.value == float(expect) and v == expect / float(1 << 46)


def test_makea_in_pieces_is_identical(npb):
    whole = npb.NpbMatrix("W", 1000, 6000)
    parts = npb.NpbMatrix("W", 1000, 6000, pieces=7)
    assert parts.n == whole.n and parts.nnz == whole.nnz
    assert np.array_equal(parts.rowstr, whole.rowstr)
    assert np.array_equal(parts.colidx, whole.colidx) and np.array_equal(parts.a, whole.a)


def test_c_caller_loop_issues_the_calls(npb, oracle):
    """npb_time_spmv_calls (the loop bench.py times the ABI with): every call
    goes through the given function pointer with x rotating over the caller
    vectors, y lands in ov, and the time per call is positive."""
    m = npb.NpbMatrix("S")
    rng = np.random.default_rng(3)
    xs = [rng.standard_normal(m.n + 2) for _ in range(3)]
    y = np.full(m.n, np.nan)
    for calls in (1, 2, 3, 7):
        sec = npb.time_spmv_calls(oracle.harness_address(), y, m.a, xs, m.rowstr, m.colidx, m.n, calls)
        assert sec > 0.0
        # the last call used xs[(calls - 1) % 3]
        assert np.array_equal(y, oracle.spmv(m.a, xs[(calls - 1) % 3], m.rowstr, m.colidx))


def test_generating_vectors_for_the_device_generator(npb):
    """npb_vectors_get: the sequential part of makea (cg.f:709-718, :876) that
    include/b200_npb.h consumes -- nonzer or nonzer + 1 entries per vector, positions inside
    the matrix, the vector's own index forced to 0.5 (vecset, cg.f:991-1019), and the size_i
    sequence accumulated by repeated multiplication."""
    import ctypes as C
    from ctypes import POINTER, c_double, c_int, c_void_p
    cls = npb.cg_class("S")
    arow, acol, aelt, size = c_void_p(), c_void_p(), c_void_p(), c_void_p()
    assert npb.lib().npb_vectors_get(C.byref(cls), C.byref(arow), C.byref(acol), C.byref(aelt), C.byref(size)) == 0
    n, ld = cls.na, cls.nonzer + 1
    ar = np.ctypeslib.as_array(C.cast(arow, POINTER(c_int)), shape=(n,))
    ac = np.ctypeslib.as_array(C.cast(acol, POINTER(c_int)), shape=(n, ld))
    ae = np.ctypeslib.as_array(C.cast(aelt, POINTER(c_double)), shape=(n, ld))
    sz = np.ctypeslib.as_array(C.cast(size, POINTER(c_double)), shape=(n,)).copy()
    npb.lib().npb_free(size)
    assert set(np.unique(ar)) <= {cls.nonzer, cls.nonzer + 1}
    for i in (0, 1, n // 2, n - 1):
        k = ar[i]
        assert np.all((ac[i, :k] >= 1) & (ac[i, :k] <= n)) and len(set(ac[i, :k])) == k
        own = np.flatnonzero(ac[i, :k] == i + 1)
        assert len(own) == 1 and ae[i, own[0]] == 0.5
    ratio = cls.rcond ** (1.0 / n)
    assert sz[0] == 1.0 and sz[1] == ratio and sz[2] == ratio * ratio
    assert abs(sz[-1] * ratio - cls.rcond) < 1e-12
    # the triples these vectors generate are the matrix: nnz self-check through the host path
    m = npb.NpbMatrix("S")
    assert m.nnz == 78148
