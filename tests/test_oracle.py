"""The oracle is pinned before anything trusts it: against the reference's
own known-answer test (libspmv/test.cpp:44-49), against outputs of the
reference's native.so committed as fixtures (tests/golden/make_golden.py), and
-- when oracle/_ref/native.so is present -- against that library live."""
import numpy as np
import pytest

from conftest import make_csr


def test_oracle_matches_reference_kat(oracle, kat):
    for dt in (np.float64, np.float32):
        a = np.array(kat["a"], dtype=dt)
        x = np.array(kat["x"], dtype=dt)
        rowstr = np.array(kat["rowstr"], dtype=np.int32)
        colidx = np.array(kat["colidx"], dtype=np.int32)
        y = oracle.spmv(a, x, rowstr, colidx, rows=kat["rows"])
        assert np.array_equal(y, np.array(kat["y"], dtype=dt))


def test_oracle_bit_exact_on_reference_outputs(oracle, native_vectors):
    assert len(native_vectors) >= 8
    for name, v in native_vectors.items():
        y = oracle.spmv(v["a"], v["x"], v["rowstr"], v["colidx"])
        assert y.dtype == v["y"].dtype
        assert np.array_equal(y, v["y"]), name
        y_omp = oracle.spmv(v["a"], v["x"], v["rowstr"], v["colidx"], omp=True)
        assert np.array_equal(y_omp, v["y"]), name + " (omp)"


def test_oracle_equals_live_reference_build(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/native.so not built (reference tree absent)")
    rng = np.random.default_rng(3)
    for dt in (np.float64, np.float32):
        for n, ncols, mean in ((1, 1, 1), (50, 7, 3), (2000, 3000, 60)):
            lens = rng.poisson(mean, n)
            a, c, rowstr, x = make_csr(rng, n, ncols, lens, dtype=dt, sort=False)
            y0 = oracle.spmv(a, x, rowstr, c, use_ref=True)
            y1 = oracle.spmv(a, x, rowstr, c)
            assert np.array_equal(y0, y1)


def test_oracle_empty_and_degenerate(oracle):
    a = np.zeros(0)
    c = np.zeros(0, dtype=np.int32)
    x = np.ones(3)
    rowstr = np.array([1, 1, 1, 1], dtype=np.int32)
    assert np.array_equal(oracle.spmv(a, x, rowstr, c), np.zeros(3))
    assert oracle.spmv(a, x, np.array([1], dtype=np.int32), c, rows=0).shape == (0,)
    assert oracle.max_colidx(rowstr, c) == 0


def test_oracle_sum_is_sequential_not_pairwise(oracle):
    # 1 + 2^-53 + 2^-53 : left-to-right gives 1.0, a tree (or FMA-free pairwise) gives 1+2^-52
    a = np.array([1.0, 2.0 ** -53, 2.0 ** -53])
    c = np.array([1, 1, 1], dtype=np.int32)
    rowstr = np.array([1, 4], dtype=np.int32)
    y = oracle.spmv(a, np.ones(1), rowstr, c)
    assert y[0] == 1.0


def test_extended_metric(oracle):
    rng = np.random.default_rng(0)
    a, c, rowstr, x = make_csr(rng, 200, 300, rng.poisson(30, 200))
    y = oracle.spmv(a, x, rowstr, c)
    y_ld, mag = oracle.spmv_extended(a, x, rowstr, c)
    nz = mag > 0
    assert np.all(np.abs(y - y_ld)[nz] <= 1e-14 * mag[nz])


def test_oracle_fp32_against_parboil_golden_outputs(oracle, parboil):
    """fp32 path pinned on the reference's parboil golden files: the small
    dataset is reproduced bit for bit, the medium one within the reference's
    own checker tolerance (tools/compare-output)."""
    from conftest import parboil_compare
    s = parboil["small"]
    y = oracle.spmv(s["a"], s["x"], s["rowstr"], s["colidx"])
    assert np.array_equal(y, s["y_golden"])
    m = parboil["medium"]
    y = oracle.spmv(m["a"], m["x"], m["rowstr"], m["colidx"])
    assert parboil_compare(m["y_golden"], y)
